/*
 * oracle/anchor_check.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The drop-in proof of the anchoring path at the reference's own call site (alignment/Blasr.cpp:2282-2296): a genome indexed
 * by the reference's own DNASuffixArray (Larsson-Sadakane on the ThreeBit text + BuildLookupTable, alignment/SAWriter.cpp:160-225),
 * reads as SMRTSequence objects, every read and its MakeRC mapped twice --
 *   (1) by the reference's MapReadToGenome (common/algorithms/anchoring/MapBySuffixArray.h:209-309) into vector<ChainedMatchPos>,
 *   (2) by blasr_gpu::AnchorBatch (include/blasr_gpu_adapter.hpp) on the GPU, loaded from the same objects, stored into the
 *       same vector type --
 * and compares the two lists element by element.  Exit code 0 = identical.  Built by oracle/Makefile into
 * oracle/_ref/anchor_check; run by tests/test_gpu_adapter.py.
 */
#define _GLIBCXX_USE_CXX11_ABI 0
#include "algorithms/anchoring/MapBySuffixArray.h"
#include "datastructures/suffixarray/SuffixArrayTypes.h"
#include "SMRTSequence.h"

#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>
#include "blasr_gpu_adapter.hpp"

int main(int argc, char **argv) {
  const int nReads = argc > 1 ? atoi(argv[1]) : 24;
  const int genomeLen = argc > 2 ? atoi(argv[2]) : 400000;
  std::mt19937 rng(20261018);
  static const char B[] = "ACGT";
  std::string g;
  for (int i = 0; i < genomeLen - 1; i++) g.push_back(B[rng() & 3]);
  for (int k = 0; k < 4; k++) {                       /* repeat copies: positions with several anchors */
    const int ln = 2000 + rng() % 3000, src = rng() % (genomeLen - ln - 1), dst = rng() % (genomeLen - ln - 1);
    for (int i = 0; i < ln; i++) g[dst + i] = (rng() % 100 < 3) ? B[rng() & 3] : g[src + i];
  }
  g.push_back('N');                                   /* FASTAReader.h:130 */
  DNASequence genome;
  genome.seq = (Nucleotide *)g.data(); genome.length = g.size(); genome.deleteOnExit = false;

  DNASuffixArray sa;
  {
    DNASequence coded; coded.Copy(genome); coded.ToThreeBit();       /* SAWriter.cpp:160 */
    vector<int> alphabet; sa.InitThreeBitDNAAlphabet(alphabet);
    sa.LarssonBuildSuffixArray(coded.seq, coded.length, alphabet);
    sa.BuildLookupTable(genome.seq, genome.length, 8);               /* Blasr.cpp:4419 */
  }

  AnchorParameters ap;                                /* MappingParameters.h:243-309 */
  ap.minMatchLength = 12; ap.stopMappingOnceUnique = true; ap.maxAnchorsPerPosition = 1000; ap.useLookupTable = true;

  std::vector<std::string> bases;
  std::uniform_real_distribution<double> U(0, 1);
  for (int i = 0; i < nReads; i++) {
    const int len = 500 + rng() % 6000, at = rng() % (genomeLen - len);
    std::string q;
    for (int k = 0; k < len; k++) {
      const double r = U(rng);
      if (r < 0.15 * 0.55) { q.push_back(B[rng() & 3]); q.push_back(g[at + k]); }
      else if (r < 0.15 * 0.90) { }
      else if (r < 0.15) q.push_back(B[rng() & 3]);
      else q.push_back(g[at + k]);
    }
    bases.push_back(q);
  }
  std::vector<SMRTSequence> reads(2 * nReads);
  for (int i = 0; i < nReads; i++) {
    SMRTSequence &r = reads[2 * i];
    r.seq = (Nucleotide *)bases[i].data(); r.length = bases[i].size(); r.deleteOnExit = false;
    r.subreadStart = 0; r.subreadEnd = r.length;
    if (i % 5 == 4) { r.subreadStart = r.length / 4; r.subreadEnd = r.length - r.length / 5; }
    r.MakeRC(reads[2 * i + 1]);                       /* Blasr.cpp:3337: readRC */
    reads[2 * i + 1].subreadStart = r.length - r.subreadEnd; reads[2 * i + 1].subreadEnd = r.length - r.subreadStart;
  }

  blasr_gpu::Context ctx(0);
  blasr_gpu::AnchorBatch::LoadIndex(ctx, sa, genome);
  blasr_gpu::AnchorBatch batch;
  for (size_t i = 0; i < reads.size(); i++) batch.Add(reads[i]);
  batch.Run(ctx, 8 /* params.lookupTableLength */, ap);

  int bad = 0; size_t total = 0;
  for (size_t i = 0; i < reads.size(); i++) {
    vector<ChainedMatchPos> want, got;
    const int nw = MapReadToGenome(genome, sa, reads[i], 8, want, ap);
    const int ng = batch.Store((uint32_t)i, got);
    bool same = nw == ng && want.size() == got.size();
    for (size_t k = 0; same && k < want.size(); k++) same = want[k].t == got[k].t && want[k].q == got[k].q && want[k].l == got[k].l;
    if (!same) { printf("read %zu: device match list differs from MapReadToGenome's (%zu vs %zu)\n", i, got.size(), want.size()); bad++; }
    total += want.size();
  }
  if (bad == 0) printf("anchor_check: MapReadToGenome x%zu read strands (%zu anchors) through blasr_gpu::AnchorBatch: identical to the reference call site\n", reads.size(), total);
  return bad ? 1 : 0;
}

/*
 * oracle/ref_anchor.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The UNMODIFIED reference anchoring templates behind a flat C interface:
 *   MapReadToGenome / LocateAnchorBoundsInSuffixArray   common/algorithms/anchoring/MapBySuffixArray.h:24-314
 *   SuffixArray::StoreLCPBounds / SearchLeftBound / SearchRightBound / BuildLookupTable / LarssonBuildSuffixArray
 *                                                       common/datastructures/suffixarray/SuffixArray.h:193-269,736-822,928-1067
 * instantiated the way blasr does (DNASuffixArray, DNASequence genome, SMRTSequence read, ChainedMatchPos;
 * alignment/Blasr.cpp:2282-2296), and the suffix array built the way sawriter does (alignment/SAWriter.cpp:160,202,225:
 * ToThreeBit, Larsson-Sadakane, BuildLookupTable).  The headers are #include'd from where they lie; nothing is copied.
 * Output: oracle/_ref/libblasr_ref_anchor.so (git-ignored, travels to the GPU box).
 */
#define _GLIBCXX_USE_CXX11_ABI 0
#include "algorithms/anchoring/MapBySuffixArray.h"
#include "datastructures/suffixarray/SuffixArrayTypes.h"
#include "SMRTSequence.h"
#include <stdint.h>
#include <string.h>
#include <thread>
#include <atomic>
#include <vector>

/* params[]: 0 minPrefixMatchLength (blasr: params.lookupTableLength), 1 minMatchLength, 2 expand, 3 useLookupTable,
 * 4 maxAnchorsPerPosition, 5 advanceExactMatches, 6 maxLCPLength, 7 stopMappingOnceUnique, 8 removeEncompassedMatches */
static void FillParams(const int32_t *p, AnchorParameters &ap) {
  ap.minMatchLength = (DNALength)p[1];
  ap.expand = p[2];
  ap.useLookupTable = p[3] != 0;
  ap.maxAnchorsPerPosition = p[4];
  ap.advanceExactMatches = p[5];
  ap.maxLCPLength = p[6];
  ap.stopMappingOnceUnique = p[7] != 0;
  ap.removeEncompassedMatches = p[8] != 0;
}

/* sawriter's default construction: genome -> ThreeBit codes -> Larsson-Sadakane.  index[n] out. */
extern "C" int ref_sa_build(const uint8_t *genome, uint32_t n, uint32_t *index) {
  DNASequence seq;
  seq.seq = new Nucleotide[n + 1];
  memcpy(seq.seq, genome, n);
  seq.length = n;
  seq.deleteOnExit = false;
  seq.ToThreeBit();
  DNASuffixArray sa;
  vector<int> alphabet;
  sa.InitThreeBitDNAAlphabet(alphabet);
  sa.LarssonBuildSuffixArray(seq.seq, seq.length, alphabet);
  memcpy(index, sa.index, sizeof(uint32_t) * (size_t)n);
  delete[] seq.seq;
  seq.seq = NULL;
  return 0;
}

/* SuffixArray::BuildLookupTable on the ASCII genome (what blasr does when the .sa file carries no table, Blasr.cpp:4419;
 * sawriter calls it on the ThreeBit-coded text, same tuples).  startPos / endPos: 4^prefixLength entries each. */
extern "C" int ref_sa_lookup_table(const uint8_t *genome, uint32_t n, const uint32_t *index, int prefixLength,
                                   uint32_t *startPos, uint32_t *endPos) {
  DNASuffixArray sa;
  sa.index = const_cast<uint32_t *>(index);
  sa.length = n;
  sa.BuildLookupTable((Nucleotide *)genome, n, prefixLength);
  memcpy(startPos, sa.startPosTable, sizeof(uint32_t) * sa.lookupTableLength);
  memcpy(endPos, sa.endPosTable, sizeof(uint32_t) * sa.lookupTableLength);
  sa.index = NULL;          /* borrowed */
  return 0;
}

struct SAView {
  DNASuffixArray sa;
  DNASequence genome;
  SAView(const uint8_t *g, uint32_t n, const uint32_t *index, const uint32_t *startPos, const uint32_t *endPos, int prefixLength) {
    sa.index = const_cast<uint32_t *>(index);
    sa.length = n;
    sa.deleteStructures = false;
    if (startPos) {
      sa.startPosTable = const_cast<uint32_t *>(startPos);
      sa.endPosTable = const_cast<uint32_t *>(endPos);
      sa.lookupPrefixLength = prefixLength;
      sa.lookupTableLength = 1u << (2 * prefixLength);
      sa.tm.Initialize(prefixLength);
    }
    genome.seq = (Nucleotide *)g;
    genome.length = n;
    genome.deleteOnExit = false;
  }
};

static int MapOne(SAView &v, const uint8_t *read, uint32_t readLen, uint32_t subStart, uint32_t subEnd, const int32_t *params,
                  vector<ChainedMatchPos> &out) {
  SMRTSequence r;
  r.seq = (Nucleotide *)read;
  r.length = readLen;
  r.deleteOnExit = false;
  r.subreadStart = subStart;
  r.subreadEnd = subEnd;
  AnchorParameters ap;
  FillParams(params, ap);
  out.clear();
  int n = MapReadToGenome(v.genome, v.sa, r, (unsigned int)params[0], out, ap);
  r.seq = NULL;
  return n;
}

/* One MapReadToGenome call.  matches[3 * i ..] = (t, q, l) of matchPosList[i]; returns the list's size (may exceed cap:
 * only the first cap entries are stored).  The genome buffer must be readable one byte past n (see DESIGN N3). */
extern "C" int64_t ref_map_read(const uint8_t *genome, uint32_t n, const uint32_t *index, const uint32_t *startPos,
                                const uint32_t *endPos, int prefixLength, const uint8_t *read, uint32_t readLen,
                                uint32_t subStart, uint32_t subEnd, const int32_t *params, uint32_t *matches, uint64_t cap) {
  SAView v(genome, n, index, startPos, endPos, prefixLength);
  vector<ChainedMatchPos> out;
  MapOne(v, read, readLen, subStart, subEnd, params, out);
  for (size_t i = 0; i < out.size() && i < cap; i++) {
    matches[3 * i] = out[i].t; matches[3 * i + 1] = out[i].q; matches[3 * i + 2] = out[i].l;
  }
  return (int64_t)out.size();
}

/* Many reads on nThreads threads (the CPU baseline of the bench and the bulk parity check): read i = reads[readOff[i] ..
 * readOff[i + 1]), whole-read subread.  counts[i] = matchPosList.size(); when matches != NULL the lists are written at
 * matchOff[i] (the caller sized them from an earlier counts pass).  Returns the total number of matches. */
extern "C" int64_t ref_map_reads(const uint8_t *genome, uint32_t n, const uint32_t *index, const uint32_t *startPos,
                                 const uint32_t *endPos, int prefixLength, const uint8_t *reads, const uint64_t *readOff,
                                 uint32_t nReads, const int32_t *params, int nThreads, uint64_t *counts, uint32_t *matches,
                                 const uint64_t *matchOff) {
  std::atomic<uint32_t> next(0);
  std::atomic<int64_t> total(0);
  auto work = [&]() {
    SAView v(genome, n, index, startPos, endPos, prefixLength);
    vector<ChainedMatchPos> out;
    for (;;) {
      uint32_t i = next.fetch_add(1);
      if (i >= nReads) break;
      const uint32_t len = (uint32_t)(readOff[i + 1] - readOff[i]);
      MapOne(v, reads + readOff[i], len, 0, len, params, out);
      counts[i] = out.size();
      total += (int64_t)out.size();
      if (matches) {
        uint32_t *m = matches + 3 * matchOff[i];
        for (size_t k = 0; k < out.size(); k++) { m[3 * k] = out[k].t; m[3 * k + 1] = out[k].q; m[3 * k + 2] = out[k].l; }
      }
    }
  };
  if (nThreads < 1) nThreads = 1;
  std::vector<std::thread> th;
  for (int t = 1; t < nThreads; t++) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
  return total.load();
}

/*
 * oracle/adapter_check.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The drop-in proof at the reference's own call site: builds candidates with the reference's SDPAlign the way
 * AlignIntervals does (alignment/Blasr.cpp:1716-1722), then refines every candidate twice --
 *   (1) with the reference's own AffineGuidedAlign/GuidedAlign + ComputeAlignmentStats (alignment/Blasr.cpp:863-878),
 *   (2) with blasr_gpu::RefineBatch (include/blasr_gpu_adapter.hpp) on the GPU, storing into the reference's real
 *       T_AlignmentCandidate --
 * and compares blocks, gaps, qPos, tPos, nCells, score and the stats fields.  Exit code 0 = identical.
 *
 * Built by oracle/Makefile into oracle/_ref/adapter_check (it contains compiled reference templates, so it lives
 * beside libblasr_ref.so, git-ignored, and travels to the GPU box); run by tests/test_gpu_adapter.py.
 */
#define _GLIBCXX_USE_CXX11_ABI 0
#include "algorithms/alignment.h"
#include "algorithms/alignment/GuidedAlign.h"
#include "algorithms/alignment/AffineGuidedAlign.h"
#include "algorithms/alignment/SDPAlign.h"
#include "algorithms/alignment/AffineKBandAlign.h"
#include "algorithms/alignment/DistanceMatrixScoreFunction.h"
#include "datastructures/alignment/AlignmentCandidate.h"
#include "FASTQSequence.h"
#include "algorithms/alignment/printers/SAMPrinter.h"

#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>
#include "blasr_gpu_adapter.hpp"

typedef DistanceMatrixScoreFunction<DNASequence, FASTQSequence> DistFn;

struct Pair { std::string q, t; };

static Pair MakePair(std::mt19937 &rng, int len, double err) {
  static const char B[] = "ACGT";
  Pair p;
  std::uniform_real_distribution<double> U(0, 1);
  for (int i = 0; i < len; i++) p.t.push_back(B[rng() & 3]);
  for (int i = 0; i < len; i++) {
    const bool prot = i < 20 || i >= len - 20;
    const double r = prot ? 1.0 : U(rng);
    if (r < err * 0.55) { p.q.push_back(B[rng() & 3]); p.q.push_back(p.t[i]); }
    else if (r < err * 0.90) { /* deletion */ }
    else if (r < err) p.q.push_back(B[(std::string(B).find(p.t[i]) + 1 + rng() % 3) & 3]);
    else p.q.push_back(p.t[i]);
  }
  return p;
}

template <typename A, typename B>
static int Diff(const A &a, const B &b, int job, const char *what) {
  int bad = 0;
#define CHK(f) if (a.f != b.f) { printf("job %d %s: %s differs (%ld vs %ld)\n", job, what, #f, (long)a.f, (long)b.f); bad++; }
  CHK(qPos) CHK(tPos) CHK(nCells) CHK(score) CHK(nMatch) CHK(nMismatch) CHK(nIns) CHK(nDel)
#undef CHK
  if (a.pctSimilarity != b.pctSimilarity) { printf("job %d %s: pctSimilarity differs\n", job, what); bad++; }
  if (a.blocks.size() != b.blocks.size()) { printf("job %d %s: %zu vs %zu blocks\n", job, what, a.blocks.size(), b.blocks.size()); return bad + 1; }
  for (size_t i = 0; i < a.blocks.size(); i++)
    if (a.blocks[i].qPos != b.blocks[i].qPos || a.blocks[i].tPos != b.blocks[i].tPos || a.blocks[i].length != b.blocks[i].length) { printf("job %d %s: block %zu differs\n", job, what, i); return bad + 1; }
  if (a.gaps.size() != b.gaps.size()) { printf("job %d %s: %zu vs %zu gap lists\n", job, what, a.gaps.size(), b.gaps.size()); return bad + 1; }
  for (size_t i = 0; i < a.gaps.size(); i++) {
    if (a.gaps[i].size() != b.gaps[i].size()) { printf("job %d %s: gap list %zu size differs\n", job, what, i); return bad + 1; }
    for (size_t j = 0; j < a.gaps[i].size(); j++)
      if (a.gaps[i][j].seq != b.gaps[i][j].seq || a.gaps[i][j].length != b.gaps[i][j].length) { printf("job %d %s: gap %zu/%zu differs\n", job, what, i, j); return bad + 1; }
  }
  return bad;
}

/* What the reference prints from a refined candidate: the SAM CIGAR core (printers/SAMPrinter.h:203-293, '=' / 'X' / 'I' / 'D'
 * ops) and the three m5 strings (AlignmentUtils.h:390-533, printed by printers/CompareSequencesAlignmentPrinter.h:17-89),
 * produced by the reference's own printers from whatever the aligner stored into the candidate. */
static std::string PrintedForm(T_AlignmentCandidate &a, FASTQSequence &q, DNASequence &t) {
  a.qAlignedSeq.ReferenceSubstring(q, 0, q.length); a.tAlignedSeq.ReferenceSubstring(t, 0, t.length);
  a.qAlignedSeqPos = 0; a.tAlignedSeqPos = 0; a.qLength = q.length; a.tLength = t.length;
  std::string out;
  if (a.blocks.size()) {
    vector<int> opSize; vector<char> opChar; std::string cigar;
    SAMOutput::CreateNoClippingCigarOps(a, a.qPos + a.blocks[0].qPos, a.tPos + a.blocks[0].tPos, opSize, opChar);
    SAMOutput::CigarOpsToString(opSize, opChar, cigar);
    out += cigar;
  }
  std::string qs, as, ts;
  CreateAlignmentStrings(a, q.seq, t.seq, ts, as, qs, q.length, t.length);
  return out + "|" + qs + "|" + as + "|" + ts;
}

/* What KBandAlign / SWAlign / AffineKBandAlign leave in the alignment: blocks, gaps, qPos / tPos and the return value. */
template <typename A, typename B>
static int DiffDense(const A &a, int sa, const B &b, int sb, int job, const char *what, bool pos) {
  int bad = 0;
  if (sa != sb) { printf("job %d %s: return value differs (%d vs %d)\n", job, what, sa, sb); bad++; }
  if (pos && (a.qPos != b.qPos || a.tPos != b.tPos)) { printf("job %d %s: qPos/tPos differ\n", job, what); bad++; }
  if (a.blocks.size() != b.blocks.size()) { printf("job %d %s: %zu vs %zu blocks\n", job, what, a.blocks.size(), b.blocks.size()); return bad + 1; }
  for (size_t i = 0; i < a.blocks.size(); i++)
    if (a.blocks[i].qPos != b.blocks[i].qPos || a.blocks[i].tPos != b.blocks[i].tPos || a.blocks[i].length != b.blocks[i].length) { printf("job %d %s: block %zu differs\n", job, what, i); return bad + 1; }
  if (a.gaps.size() != b.gaps.size()) { printf("job %d %s: %zu vs %zu gap lists\n", job, what, a.gaps.size(), b.gaps.size()); return bad + 1; }
  for (size_t i = 0; i < a.gaps.size(); i++) {
    if (a.gaps[i].size() != b.gaps[i].size()) { printf("job %d %s: gap list %zu size differs\n", job, what, i); return bad + 1; }
    for (size_t j = 0; j < a.gaps[i].size(); j++)
      if (a.gaps[i][j].seq != b.gaps[i][j].seq || a.gaps[i][j].length != b.gaps[i][j].length) { printf("job %d %s: gap %zu/%zu differs\n", job, what, i, j); return bad + 1; }
  }
  return bad;
}

/* The guide-less call sites through blasr_gpu::DenseBatch against the reference's own templates, called the way blasr calls
 * them: KBandAlign Global (-global, Blasr.cpp:820-824) and Fit (PairwiseLocalAlign, :717-730), SWAlign Global on short gap
 * fragments (SDPAlign.h:440,503,563), AffineKBandAlign Global with AlignSubstring's parameter pattern (Blasr.cpp:1067-1076). */
static int CheckDense(std::vector<Pair> &pairs, DistFn &fn) {
  int bad = 0;
  std::mt19937 rng(99);
  blasr_gpu::Context ctx(0);
  const int n = (int)std::min<size_t>(pairs.size(), 24);
  std::vector<FASTQSequence> qs(n); std::vector<DNASequence> ts(n);
  std::vector<std::string> qstr(n), tstr(n);
  int mat[5][5];
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) mat[i][j] = SMRTDistanceMatrix[i][j];
  for (int mode = 0; mode < 4; mode++) {
    const char *what = mode == 0 ? "KBandAlign Global" : mode == 1 ? "KBandAlign Fit" : mode == 2 ? "SWAlign Global" : "AffineKBandAlign Global";
    blasr_gpu::DenseBatch batch;
    for (int i = 0; i < n; i++) {
      /* short fragments for the gap-fill shapes, longer slices for the k-band ones */
      const size_t len = mode >= 2 ? 4 + rng() % 28 : 200 + rng() % 700;
      qstr[i] = pairs[i].q.substr(0, std::min(pairs[i].q.size(), len));
      tstr[i] = pairs[i].t.substr(0, std::min(pairs[i].t.size(), mode >= 2 ? 4 + rng() % 28 : len));
      qs[i].seq = (Nucleotide *)qstr[i].data(); qs[i].length = qstr[i].size();
      ts[i].seq = (Nucleotide *)tstr[i].data(); ts[i].length = tstr[i].size();
      batch.Add(qs[i].seq, qs[i].length, ts[i].seq, ts[i].length);
    }
    const int k = mode == 3 ? 6 : 15, indel = 5;
    if (mode == 0) batch.RunKBand(ctx, fn, indel, indel, k, BGPU_GLOBAL);
    else if (mode == 1) batch.RunKBand(ctx, fn, indel, indel, k, BGPU_FIT);
    else if (mode == 2) batch.RunSW(ctx, fn, BGPU_GLOBAL);
    else batch.RunAffineKBand(ctx, mat, indel + 2, indel - 3, indel + 2, indel - 1, indel, k, BGPU_GLOBAL);
    int checked = 0;
    for (int i = 0; i < n; i++) {
      if (batch.Result(i).status != BGPU_JOB_OK) continue;   /* shapes on which the reference itself is undefined */
      T_AlignmentCandidate ref, gpu;
      vector<int> scoreMat, hpS, insS; vector<Arrow> pathMat, hpP, insP;
      int sr;
      if (mode == 0) sr = KBandAlign(qs[i], ts[i], mat, indel, indel, k, scoreMat, pathMat, ref, Global, fn, false);
      else if (mode == 1) sr = KBandAlign(qs[i], ts[i], mat, indel, indel, k, scoreMat, pathMat, ref, Fit, fn, false);
      else if (mode == 2) sr = SWAlign(qs[i], ts[i], scoreMat, pathMat, ref, fn, Global);
      else sr = AffineKBandAlign(qs[i], ts[i], mat, indel + 2, indel - 3, indel + 2, indel - 1, indel, k, scoreMat, pathMat, hpS, hpP, insS, insP, ref, Global);
      const int sg = batch.Store(i, gpu);
      bad += DiffDense(ref, sr, gpu, sg, i, what, mode != 3);
      checked++;
    }
    printf("adapter_check: %s x%d jobs through blasr_gpu::DenseBatch: %s\n", what, checked, bad ? "MISMATCH" : "identical to the reference call site");
  }
  return bad;
}

int main(int argc, char **argv) {
  const int nJobs = argc > 1 ? atoi(argv[1]) : 48;
  const int maxLen = argc > 2 ? atoi(argv[2]) : 6000;
  std::mt19937 rng(20261017);
  DistFn fn;
  fn.InitializeScoreMatrix(SMRTDistanceMatrix);
  fn.ins = 5; fn.del = 5; fn.affineOpen = 50; fn.affineExtend = 0;          /* MappingParameters.h:339-342 */

  std::vector<Pair> pairs;
  std::vector<T_AlignmentCandidate> cands(nJobs);
  for (int i = 0; i < nJobs; i++) {
    pairs.push_back(MakePair(rng, 300 + (int)(rng() % (maxLen - 300)), 0.15));
    FASTQSequence q; DNASequence t;
    q.seq = (Nucleotide *)pairs[i].q.data(); q.length = pairs[i].q.size();
    t.seq = (Nucleotide *)pairs[i].t.data(); t.length = pairs[i].t.size();
    /* the candidate AlignIntervals hands to RefineAlignment: SDPAlign(Local, detailed), Blasr.cpp:1716-1722 */
    SDPAlign(q, t, fn, 11, 5, 10, 0.30f, cands[i], Local, true, false, 50, 2, 1000);
  }

  if (argc > 3 && std::string(argv[3]) == "dense") return CheckDense(pairs, fn) ? 1 : 0;   /* the guide-less call sites */
  if (argc > 3 && std::string(argv[3]) == "sdp") {     /* the candidate-producing call itself: SDPAlign on the device */
    int badS = 0;
    blasr_gpu::Context ctx(0);
    blasr_gpu::SdpBatch batch;
    for (int i = 0; i < nJobs; i++) batch.Add((const uint8_t *)pairs[i].q.data(), pairs[i].q.size(), (const uint8_t *)pairs[i].t.data(), pairs[i].t.size());
    /* the argument list of Blasr.cpp:1716-1722 with MappingParameters' defaults, as above */
    batch.Run(ctx, fn, 11, 5, 10, 0.30f, Local, true, false, 50, 2, 1000, 0);
    size_t nBlocks = 0;
    for (int i = 0; i < nJobs; i++) {
      T_AlignmentCandidate g;
      batch.Store(i, g);
      const T_AlignmentCandidate &r = cands[i];
      bool same = g.qPos == r.qPos && g.tPos == r.tPos && g.blocks.size() == r.blocks.size();
      for (size_t k = 0; same && k < r.blocks.size(); k++)
        same = g.blocks[k].qPos == r.blocks[k].qPos && g.blocks[k].tPos == r.blocks[k].tPos && g.blocks[k].length == r.blocks[k].length;
      if (!same) { printf("pair %d: device SDPAlign differs from the reference's (blocks %zu vs %zu)\n", i, g.blocks.size(), r.blocks.size()); badS++; }
      nBlocks += r.blocks.size();
    }
    /* the Global pattern of AlignSubstring (Blasr.cpp:1080-1090): front / tail extension on, no recursion over 10000 cells */
    std::vector<T_AlignmentCandidate> glob(nJobs);
    for (int i = 0; i < nJobs; i++) {
      FASTQSequence q; DNASequence t;
      q.seq = (Nucleotide *)pairs[i].q.data(); q.length = pairs[i].q.size();
      t.seq = (Nucleotide *)pairs[i].t.data(); t.length = pairs[i].t.size();
      SDPAlign(q, t, fn, 11, 5, 10, 0.25f, glob[i], Global, true, true, 50, 2, 1000, 0);
    }
    batch.Run(ctx, fn, 11, 5, 10, 0.25f, Global, true, true, 50, 2, 1000, 0);
    for (int i = 0; i < nJobs; i++) {
      T_AlignmentCandidate g;
      batch.Store(i, g);
      const T_AlignmentCandidate &r = glob[i];
      bool same = g.qPos == r.qPos && g.tPos == r.tPos && g.blocks.size() == r.blocks.size();
      for (size_t k = 0; same && k < r.blocks.size(); k++)
        same = g.blocks[k].qPos == r.blocks[k].qPos && g.blocks[k].tPos == r.blocks[k].tPos && g.blocks[k].length == r.blocks[k].length;
      if (!same) { printf("pair %d: device SDPAlign (Global) differs from the reference's (blocks %zu vs %zu)\n", i, g.blocks.size(), r.blocks.size()); badS++; }
      nBlocks += r.blocks.size();
    }
    printf("adapter_check: SDPAlign x%d pairs (Local and Global patterns, %zu blocks) through blasr_gpu::SdpBatch: %s\n", nJobs, nBlocks,
           badS ? "MISMATCH" : "identical to the reference call site");
    return badS ? 1 : 0;
  }

  int bad = 0;
  size_t printedBytes = 0;
  for (int affine = 1; affine >= 0; affine--) {
    const int band = affine ? 16 : 10;                                       /* bandSize / guidedAlignBandSize */
    blasr_gpu::Context ctx(0);
    blasr_gpu::RefineBatch batch;
    std::vector<FASTQSequence> qs(nJobs); std::vector<DNASequence> ts(nJobs);
    std::vector<int> jobOf;
    for (int i = 0; i < nJobs; i++) {
      T_AlignmentCandidate &c = cands[i];
      if (c.blocks.size() == 0) continue;
      const int last = c.blocks.size() - 1;
      /* the slices of Blasr.cpp:850-859 */
      ts[i].seq = (Nucleotide *)pairs[i].t.data() + c.tPos; ts[i].length = c.blocks[last].tPos + c.blocks[last].length;
      qs[i].seq = (Nucleotide *)pairs[i].q.data() + c.qPos; qs[i].length = c.blocks[last].qPos + c.blocks[last].length;
      batch.Add(qs[i].seq, qs[i].length, ts[i].seq, ts[i].length, c.blocks);
      jobOf.push_back(i);
    }
    batch.Run(ctx, fn, band, affine != 0);
    for (size_t j = 0; j < jobOf.size(); j++) {
      const int i = jobOf[j];
      T_AlignmentCandidate refRefined, gpuRefined;
      vector<int> scoreMat; vector<Arrow> pathMat; vector<double> pm, opm; vector<float> a, b, c, d;
      if (affine) AffineGuidedAlign(qs[i], ts[i], cands[i], fn, band, refRefined, scoreMat, pathMat, pm, opm, a, b, c, d, Global, false);
      else GuidedAlign(qs[i], ts[i], cands[i], fn, band, refRefined, scoreMat, pathMat, pm, opm, a, b, c, d, Global, false);
      ComputeAlignmentStats(refRefined, qs[i].seq, ts[i].seq, fn, affine != 0);
      batch.Store(j, gpuRefined);
      bad += Diff(refRefined, gpuRefined, i, affine ? "AffineGuidedAlign" : "GuidedAlign");
      const std::string pr = PrintedForm(refRefined, qs[i], ts[i]), pg = PrintedForm(gpuRefined, qs[i], ts[i]);
      if (pr != pg) { printf("job %d: printed CIGAR / m5 strings differ\n", i); bad++; }
      /* the CIGAR built on the device (bgpu_cigar) against the reference printer's text */
      if (batch.Cigar(j) != pr.substr(0, pr.find('|'))) { printf("job %d: device CIGAR differs from the reference printer's\n", i); bad++; }
      printedBytes += pr.size();
      /* StoreMapQVs' rescoring (Blasr.cpp:2768-2780): the run's gap costs with SMRTLogProbMatrix, the Alignment overload of
       * ComputeAlignmentScore, on the refined candidate */
      DistFn logProb;
      logProb.InitializeScoreMatrix(SMRTLogProbMatrix);
      logProb.ins = fn.ins; logProb.del = fn.del; logProb.affineOpen = fn.affineOpen; logProb.affineExtend = fn.affineExtend;
      const int wantProb = ComputeAlignmentScore(refRefined, qs[i], ts[i], logProb, affine != 0);
      if (batch.Rescore(j, logProb, affine != 0) != wantProb) { printf("job %d: device rescoring differs from ComputeAlignmentScore (%d vs %d)\n", i, batch.Rescore(j, logProb, affine != 0), wantProb); bad++; }
    }
    printf("adapter_check: StoreMapQVs rescoring (SMRTLogProbMatrix) x%zu candidates on the device: %s\n", jobOf.size(), bad ? "MISMATCH" : "identical");
    printf("adapter_check: %zu bytes of SAM CIGAR + m5 alignment strings printed by the reference's printers: %s\n", printedBytes,
           bad ? "MISMATCH" : "identical");
    printf("adapter_check: %s x%zu candidates through blasr_gpu::RefineBatch: %s\n", affine ? "AffineGuidedAlign" : "GuidedAlign",
           jobOf.size(), bad ? "MISMATCH" : "identical to the reference call site");
  }
  return bad ? 1 : 0;
}

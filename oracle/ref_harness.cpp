/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Instantiates the UNMODIFIED reference templates behind the flat interface of
 * oracle/orc_align.h.  The reference headers are #include'd from where they lie
 * (-I /root/reference/common, see oracle/Makefile); no reference source is copied
 * into this repository.  Output: oracle/_ref/libblasr_ref.so (git-ignored).
 *
 * Instantiations:
 *   GuidedAlign        common/algorithms/alignment/GuidedAlign.h:278
 *   AffineGuidedAlign  common/algorithms/alignment/AffineGuidedAlign.h:31
 *   KBandAlign         common/algorithms/alignment/KBandAlign.h:75
 *   SWAlign            common/algorithms/alignment/SWAlign.h:18
 *   AffineKBandAlign   common/algorithms/alignment/AffineKBandAlign.h:12 (Global / QueryFit; takes matchMat, no score function)
 *   ComputeAlignmentStats  common/algorithms/alignment/AlignmentUtils.h:535
 *   SDPAlign           common/algorithms/alignment/SDPAlign.h (guide producer, test inputs only)
 * each with DistanceMatrixScoreFunction<DNASequence,FASTQSequence>,
 * QualityValueScoreFunction<DNASequence,FASTQSequence> and (all but SWAlign, whose transposed position
 * arguments make it read out of bounds, SWAlign.h:166-167) IDSScoreFunction<DNASequence,FASTQSequence>.
 */
#define _GLIBCXX_USE_CXX11_ABI 0
#include "algorithms/alignment.h"
#include "algorithms/alignment/GuidedAlign.h"
#include "algorithms/alignment/AffineGuidedAlign.h"
#include "algorithms/alignment/SDPAlign.h"
#include "algorithms/alignment/AffineKBandAlign.h"
#include "algorithms/alignment/DistanceMatrixScoreFunction.h"
#include "algorithms/alignment/QualityValueScoreFunction.h"
#include "algorithms/alignment/IDSScoreFunction.h"
#include "datastructures/alignment/AlignmentCandidate.h"
#include "FASTQSequence.h"
#include "algorithms/alignment/printers/SAMPrinter.h"

#include <thread>
#include <atomic>
#include <vector>
#include "orc_align.h"

typedef DistanceMatrixScoreFunction<DNASequence, FASTQSequence> DistFn;

/* GuidedAlign references NormalizedMatch/Insertion/Deletion even with computeProb=false
 * (GuidedAlign.h:577-583); QualityValueScoreFunction lacks them, so add a no-op shim.
 * The integer path never evaluates them. */
class QVFn : public QualityValueScoreFunction<DNASequence, FASTQSequence> {
 public:
  float NormalizedMatch(DNASequence &, DNALength, FASTQSequence &, DNALength) { return 0; }
  float NormalizedInsertion(DNASequence &, DNALength, FASTQSequence &, DNALength) { return 0; }
  float NormalizedDeletion(DNASequence &, DNALength, FASTQSequence &, DNALength) { return 0; }
};

static void FillDist(const orc_scorefn *fn, DistFn &f) {
  int m[5][5];
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) m[i][j] = fn->M[i * 5 + j];
  f.InitializeScoreMatrix(m);
  f.ins = fn->ins; f.del = fn->del;
  f.affineOpen = fn->affineOpen; f.affineExtend = fn->affineExtend;
}
static void FillQV(const orc_scorefn *fn, QVFn &f) {
  f.ins = fn->ins; f.del = fn->del;
  f.affineOpen = fn->affineOpen; f.affineExtend = fn->affineExtend;
}

typedef IDSScoreFunction<DNASequence, FASTQSequence> IDSFn;
static void FillIDS(const orc_scorefn *fn, IDSFn &f) {
  f.ins = fn->ins; f.del = fn->del;
  f.substitutionPrior = fn->substitutionPrior; f.globalDeletionPrior = fn->globalDeletionPrior;
  f.affineOpen = fn->affineOpen; f.affineExtend = fn->affineExtend;
}

struct Scratch {
  vector<int> scoreMat; vector<Arrow> pathMat; vector<double> probMat, optPathProbMat;
  vector<int> hpInsScoreMat, insScoreMat; vector<Arrow> hpInsPathMat, insPathMat;
  vector<float> a, b, c, d;
};

template <typename T_Fn>
static int RunAligner(const orc_job *job, T_Fn &f, const orc_scorefn *fn, FASTQSequence &q, DNASequence &t,
                      Alignment &aln, Scratch &s) {
  int m[5][5];
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) m[i][j] = fn->M[i * 5 + j];
  AlignmentType at = (AlignmentType)job->alignType;
  switch (job->algo) {
    case ORC_GUIDED:
    case ORC_AFFINE_GUIDED: {
      Alignment guide;
      guide.blocks.resize(job->nGuide);
      for (uint32_t i = 0; i < job->nGuide; i++) {
        guide.blocks[i].qPos = job->guide[3 * i];
        guide.blocks[i].tPos = job->guide[3 * i + 1];
        guide.blocks[i].length = job->guide[3 * i + 2];
      }
      if (job->algo == ORC_GUIDED)
        return GuidedAlign(q, t, guide, f, job->band, aln, s.scoreMat, s.pathMat, s.probMat, s.optPathProbMat,
                           s.a, s.b, s.c, s.d, at, false);
      return AffineGuidedAlign(q, t, guide, f, job->band, aln, s.scoreMat, s.pathMat, s.probMat,
                               s.optPathProbMat, s.a, s.b, s.c, s.d, at, false);
    }
    case ORC_KBAND:
      return KBandAlign(q, t, m, job->bndIns, job->bndDel, (int)job->band, s.scoreMat, s.pathMat, aln, at, f,
                        false);
    case ORC_SW:
      return SWAlign(q, t, s.scoreMat, s.pathMat, aln, f, at);
    case ORC_AFFINE_KBAND:   /* argument order of Blasr.cpp:1067-1076 */
      return AffineKBandAlign(q, t, m, job->hpInsOpen, job->hpInsExtend, job->insOpen, job->insExtend, job->bndDel,
                              (int)job->band, s.scoreMat, s.pathMat, s.hpInsScoreMat, s.hpInsPathMat, s.insScoreMat,
                              s.insPathMat, aln, at);
  }
  return 0;
}

static int RunOne(const orc_scorefn *fn, const orc_job *job, orc_result *res, Scratch &s, Alignment &aln) {
  memset(res, 0, sizeof(*res));
  FASTQSequence q; DNASequence t;
  q.seq = (Nucleotide *)job->q; q.length = job->qLen;
  t.seq = (Nucleotide *)job->t; t.length = job->tLen;
  if (job->qual) q.qual.data = (QualityValue *)job->qual;
  DistFn df; FillDist(fn, df);
  int score;
  if (fn->kind == ORC_FN_QUALITY) {
    if (!job->qual) { q.qual.data = NULL; res->status = ORC_BAD_INPUT; return 0; }
    QVFn qf; FillQV(fn, qf);
    score = RunAligner(job, qf, fn, q, t, aln, s);
  } else if (fn->kind == ORC_FN_IDS) {
    if (!job->insQV || !job->subQV || !job->subTag || job->algo == ORC_SW) { q.qual.data = NULL; res->status = ORC_BAD_INPUT; return 0; }
    q.insertionQV.data = (QualityValue *)job->insQV; q.substitutionQV.data = (QualityValue *)job->subQV;
    q.substitutionTag = (Nucleotide *)job->subTag;
    if (job->delQV && job->delTag) { q.deletionQV.data = (QualityValue *)job->delQV; q.deletionTag = (Nucleotide *)job->delTag; }
    IDSFn idf; FillIDS(fn, idf);
    score = RunAligner(job, idf, fn, q, t, aln, s);
    q.insertionQV.data = NULL; q.substitutionQV.data = NULL; q.deletionQV.data = NULL;   /* borrowed */
    q.substitutionTag = NULL; q.deletionTag = NULL;
  } else {
    score = RunAligner(job, df, fn, q, t, aln, s);
  }
  res->score = score;
  res->alnScore = aln.score;
  res->qPos = aln.qPos; res->tPos = aln.tPos; res->nCells = aln.nCells;
  if ((job->algo == ORC_GUIDED || job->algo == ORC_AFFINE_GUIDED) && job->nGuide == 0)
    res->status = ORC_EMPTY_GUIDE;
  if (job->doStats) {
    /* blasr always rescoring with the distance-matrix function (Blasr.cpp:875-878) */
    ComputeAlignmentStats(aln, q.seq, t.seq, df, job->statsAffine != 0);
    res->nMatch = aln.nMatch; res->nMismatch = aln.nMismatch; res->nIns = aln.nIns; res->nDel = aln.nDel;
    res->pctSimilarity = aln.pctSimilarity; res->statsScore = aln.score;
  }
  q.qual.data = NULL; /* borrowed */
  return 0;
}

extern "C" int ref_align(const orc_scorefn *fn, const orc_job *job, orc_result *res, uint32_t *blocks,
                         uint32_t capBlocks, uint32_t *gapCounts, uint32_t capGapLists, int32_t *gaps,
                         uint32_t capGaps) {
  Scratch s; Alignment aln;
  RunOne(fn, job, res, s, aln);
  res->nBlocks = aln.blocks.size();
  res->nGapLists = aln.gaps.size();
  uint32_t ng = 0;
  for (size_t i = 0; i < aln.gaps.size(); i++) ng += aln.gaps[i].size();
  res->nGaps = ng;
  if (res->nBlocks > capBlocks || res->nGapLists > capGapLists || ng > capGaps) return ORC_OVERFLOW;
  for (size_t i = 0; i < aln.blocks.size(); i++) {
    blocks[3 * i] = aln.blocks[i].qPos; blocks[3 * i + 1] = aln.blocks[i].tPos;
    blocks[3 * i + 2] = aln.blocks[i].length;
  }
  uint32_t g = 0;
  for (size_t i = 0; i < aln.gaps.size(); i++) {
    gapCounts[i] = aln.gaps[i].size();
    for (size_t j = 0; j < aln.gaps[i].size(); j++) {
      gaps[2 * g] = (int)aln.gaps[i][j].seq; gaps[2 * g + 1] = aln.gaps[i][j].length; g++;
    }
  }
  return 0;
}

extern "C" int ref_cigar(const orc_scorefn *fn, const orc_job *job, uint32_t *ops, uint32_t capOps) {
  Scratch s; T_AlignmentCandidate cand; orc_result res;
  RunOne(fn, job, &res, s, cand);
  if (res.status != ORC_OK || cand.blocks.size() == 0) return 0;
  /* the printer reads the aligned sequences through the candidate (SAMPrinter.h:146-147) */
  DNASequence t; t.seq = (Nucleotide *)job->t; t.length = job->tLen;
  FASTQSequence q; q.seq = (Nucleotide *)job->q; q.length = job->qLen;
  cand.qAlignedSeq.ReferenceSubstring(q, 0, q.length); cand.tAlignedSeq.ReferenceSubstring(t, 0, t.length);
  vector<int> opSize; vector<char> opChar;
  SAMOutput::CreateNoClippingCigarOps(cand, cand.qPos + cand.blocks[0].qPos, cand.tPos + cand.blocks[0].tPos, opSize, opChar);
  if (opSize.size() > capOps) return -1;
  for (size_t i = 0; i < opSize.size(); i++) {
    const char c = opChar[i];
    const uint32_t code = c == '=' ? 7 : c == 'X' ? 8 : c == 'I' ? 1 : c == 'D' ? 2 : 15;
    ops[i] = ((uint32_t)opSize[i] << 4) | code;
  }
  return (int)opSize.size();
}

/* The whole CreateCIGARString (SAMPrinter.h:345-400) on the job's alignment placed inside a longer read: the candidate's
 * qAlignedSeqPos = qSeqPos, read.length = readLength, read.lowQualityPrefix / Suffix, tStrand.  clipping: 0 hard, 1 soft,
 * 2 subread, 3 none.  Writes the CIGAR text and the four clip lengths (hard prefix, soft prefix, soft suffix, hard suffix). */
extern "C" int ref_cigar_string(const orc_scorefn *fn, const orc_job *job, int clipping, int tStrand, uint32_t qSeqPos,
                                uint32_t readLength, uint32_t lowQPrefix, uint32_t lowQSuffix, char *out, uint32_t capOut,
                                uint32_t *clips) {
  Scratch s; T_AlignmentCandidate cand; orc_result res;
  RunOne(fn, job, &res, s, cand);
  if (res.status != ORC_OK || cand.blocks.size() == 0) return 0;
  DNASequence t; t.seq = (Nucleotide *)job->t; t.length = job->tLen;
  FASTQSequence q; q.seq = (Nucleotide *)job->q; q.length = job->qLen;
  cand.qAlignedSeq.ReferenceSubstring(q, 0, q.length); cand.tAlignedSeq.ReferenceSubstring(t, 0, t.length);
  cand.qAlignedSeqPos = qSeqPos; cand.tStrand = tStrand;
  SMRTSequence read;
  read.length = readLength; read.lowQualityPrefix = lowQPrefix; read.lowQualitySuffix = lowQSuffix;
  read.subreadStart = 0; read.subreadEnd = readLength;
  std::string cigar;
  DNALength pS = 0, sS = 0, pH = 0, sH = 0;
  SAMOutput::CreateCIGARString(cand, read, cigar, (SAMOutput::Clipping)clipping, pS, sS, pH, sH);
  clips[0] = pH; clips[1] = pS; clips[2] = sS; clips[3] = sH;
  if (cigar.size() + 1 > capOut) return -1;
  memcpy(out, cigar.data(), cigar.size()); out[cigar.size()] = 0;
  return (int)cigar.size();
}

extern "C" int ref_alignment_strings(const orc_scorefn *fn, const orc_job *job, char *textStr, char *alignStr, char *queryStr,
                                     uint32_t capOut) {
  Scratch s; Alignment aln; orc_result res;
  RunOne(fn, job, &res, s, aln);
  if (res.status != ORC_OK) return 0;
  std::string ts, as, qs;
  Nucleotide *q = (Nucleotide *)job->q, *t = (Nucleotide *)job->t;
  CreateAlignmentStrings(aln, q, t, ts, as, qs, job->qLen, job->tLen);
  if (ts.size() > capOut) return -1;
  memcpy(textStr, ts.data(), ts.size()); memcpy(alignStr, as.data(), as.size()); memcpy(queryStr, qs.data(), qs.size());
  return (int)ts.size();
}

/* StoreMapQVs' rescoring (alignment/Blasr.cpp:2768-2780): the job aligned under fn, then ComputeAlignmentScore(alignment, query,
 * text, fn2, useAffinePenalty) -- the Alignment overload, AlignmentUtils.h:127-169 -- under a second distance-matrix function. */
extern "C" int ref_rescore(const orc_scorefn *fn, const orc_job *job, const orc_scorefn *fn2, int useAffine, int32_t *score) {
  Scratch s; Alignment aln; orc_result res;
  RunOne(fn, job, &res, s, aln);
  *score = 0;
  if (res.status != ORC_OK) return 0;
  FASTQSequence q; DNASequence t;
  q.seq = (Nucleotide *)job->q; q.length = job->qLen;
  t.seq = (Nucleotide *)job->t; t.length = job->tLen;
  DistFn f2; FillDist(fn2, f2);
  *score = ComputeAlignmentScore(aln, q, t, f2, useAffine != 0);
  q.seq = NULL; t.seq = NULL;
  return 1;
}

extern "C" int ref_block_strings(const uint8_t *qs, uint32_t qLen, const uint8_t *ts, uint32_t tLen, const uint32_t *blocks,
                                 uint32_t nBlocks, char *textStr, char *alignStr, char *queryStr, uint32_t capOut) {
  Alignment aln;                                   /* blocks only, no gap lists: the form SDPAlign returns */
  aln.qPos = 0; aln.tPos = 0;
  aln.blocks.resize(nBlocks);
  for (uint32_t i = 0; i < nBlocks; i++) { aln.blocks[i].qPos = blocks[3 * i]; aln.blocks[i].tPos = blocks[3 * i + 1]; aln.blocks[i].length = blocks[3 * i + 2]; }
  std::string t3, a3, q3;
  Nucleotide *q = (Nucleotide *)qs, *t = (Nucleotide *)ts;
  CreateAlignmentStrings(aln, q, t, t3, a3, q3, qLen, tLen);
  if (t3.size() > capOut) return -1;
  memcpy(textStr, t3.data(), t3.size()); memcpy(alignStr, a3.data(), a3.size()); memcpy(queryStr, q3.data(), q3.size());
  return (int)t3.size();
}

extern "C" int ref_sdp_chain(const uint32_t *frags, uint32_t n, uint32_t queryLength, uint32_t fragmentLength,
                             int insertion, int deletion, int match, int alignType, int32_t *chain, uint32_t capChain) {
  vector<Fragment> fragmentSet;
  for (uint32_t i = 0; i < n; i++) {
    Fragment f(frags[4 * i], frags[4 * i + 1], (int)frags[4 * i + 3]);
    f.length = frags[4 * i + 2];
    fragmentSet.push_back(f);
  }
  vector<int> maxFragmentChain;
  SDPLongestCommonSubsequence(queryLength, fragmentSet, fragmentLength, insertion, deletion, match, maxFragmentChain,
                              (AlignmentType)alignType);
  if (maxFragmentChain.size() > capChain) return -1;
  for (size_t i = 0; i < maxFragmentChain.size(); i++) chain[i] = maxFragmentChain[i];
  return (int)maxFragmentChain.size();
}

extern "C" int ref_sdp_fragments(const uint8_t *qs, uint32_t qLen, const uint8_t *ts, uint32_t tLen, const orc_scorefn *fn,
                                 int wordSize, int sdpIns, int sdpDel, int alignType, uint32_t *frags, uint32_t capFrags,
                                 int32_t *chain, uint32_t capChain, int32_t *nChain) {
  FASTQSequence q; DNASequence t;
  q.seq = (Nucleotide *)qs; q.length = qLen; t.seq = (Nucleotide *)ts; t.length = tLen;
  DistFn df; FillDist(fn, df);
  Alignment sdp;
  vector<Fragment> fragmentSet, prefixFragmentSet, suffixFragmentSet;
  TupleList<PositionDNATuple> targetTupleList, targetPrefixTupleList, targetSuffixTupleList;
  vector<int> maxFragmentChain;
  /* detailedAlignment = false, extendFrontByLocalAlignment = false, recurse = 0, noRecurseUnder = 0: nothing after the
   * chain touches the buffers again (the recursive calls of SDPAlign.h:445-456,505-520,565-580 reuse them) */
  SDPAlign(q, t, df, wordSize, sdpIns, sdpDel, 0.30f, sdp, fragmentSet, prefixFragmentSet, suffixFragmentSet,
           targetTupleList, targetPrefixTupleList, targetSuffixTupleList, maxFragmentChain,
           (AlignmentType)alignType, false, false, 50, 0, 0);
  *nChain = (int32_t)maxFragmentChain.size();
  if (fragmentSet.size() > capFrags || maxFragmentChain.size() > capChain) return -1;
  for (size_t i = 0; i < fragmentSet.size(); i++) {
    frags[4 * i] = fragmentSet[i].x; frags[4 * i + 1] = fragmentSet[i].y;
    frags[4 * i + 2] = fragmentSet[i].length; frags[4 * i + 3] = fragmentSet[i].weight;
  }
  for (size_t i = 0; i < maxFragmentChain.size(); i++) chain[i] = maxFragmentChain[i];
  return (int)fragmentSet.size();
}

extern "C" int ref_guide_rows(const uint32_t *guide, uint32_t nGuide, int band, int32_t *rows, uint32_t capRows,
                              int64_t *nCells) {
  Alignment a; a.blocks.resize(nGuide);
  for (uint32_t i = 0; i < nGuide; i++) {
    a.blocks[i].qPos = guide[3 * i]; a.blocks[i].tPos = guide[3 * i + 1]; a.blocks[i].length = guide[3 * i + 2];
  }
  Guide g;
  AlignmentToGuide(a, g, band);
  if (nCells) *nCells = ComputeMatrixNElem(g);
  if (g.size() > capRows) return -1;
  for (size_t i = 0; i < g.size(); i++) {
    rows[4 * i] = g[i].q; rows[4 * i + 1] = g[i].t; rows[4 * i + 2] = g[i].tPre; rows[4 * i + 3] = g[i].tPost;
  }
  return (int)g.size();
}

extern "C" int ref_sdp_guide(const uint8_t *qs, uint32_t qLen, const uint8_t *ts, uint32_t tLen,
                             const orc_scorefn *fn, int tupleSize, int sdpIns, int sdpDel, float indelRate,
                             uint32_t *blocks, uint32_t capBlocks) {
  FASTQSequence q; DNASequence t;
  q.seq = (Nucleotide *)qs; q.length = qLen; t.seq = (Nucleotide *)ts; t.length = tLen;
  DistFn df; FillDist(fn, df);
  Alignment sdp;
  /* argument pattern of Blasr.cpp:1716-1722 with MappingParameters defaults
   * (detailedSDPAlignment=true, extendFrontAlignment=false, sdpPrefix=50, recurse=2, recurseOver=1000) */
  SDPAlign(q, t, df, tupleSize, sdpIns, sdpDel, indelRate, sdp, Local, true, false, 50, 2, 1000);
  if (sdp.blocks.size() > capBlocks) return -1;
  for (size_t i = 0; i < sdp.blocks.size(); i++) {
    blocks[3 * i] = sdp.blocks[i].qPos + sdp.qPos;
    blocks[3 * i + 1] = sdp.blocks[i].tPos + sdp.tPos;
    blocks[3 * i + 2] = sdp.blocks[i].length;
  }
  return (int)sdp.blocks.size();
}

extern "C" int64_t ref_replay(const orc_scorefn *fn, const orc_job *jobs, uint32_t n, int nThreads,
                              int64_t *scoreSum) {
  std::atomic<uint32_t> next(0);
  std::atomic<long long> cells(0), ssum(0);
  if (nThreads < 1) nThreads = 1;
  auto worker = [&]() {
    Scratch s;
    long long myCells = 0, mySum = 0;
    for (;;) {
      uint32_t i = next.fetch_add(1);
      if (i >= n) break;
      Alignment aln; orc_result r;
      RunOne(fn, &jobs[i], &r, s, aln);
      myCells += r.nCells; mySum += r.score + (long long)aln.blocks.size();
    }
    cells += myCells; ssum += mySum;
  };
  std::vector<std::thread> th;
  for (int i = 0; i < nThreads; i++) th.emplace_back(worker);
  for (auto &x : th) x.join();
  if (scoreSum) *scoreSum = ssum;
  return cells;
}

/* The reference's own base -> code table (NucConversion.h:48-84), entry by entry. */
extern "C" int ref_three_bit(int c) { return ThreeBit[c & 255]; }

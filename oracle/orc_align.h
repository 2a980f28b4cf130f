/*
 * oracle/orc_align.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Flat C interface shared by the two CPU checkers of the refinement hot path:
 *
 *   ref_align()  (oracle/ref_harness.cpp -> oracle/_ref/libblasr_ref.so)
 *       the UNMODIFIED reference templates, #include'd from /root/reference/common
 *       at build time and instantiated behind this interface;
 *   orc_align()  (oracle/orc_align.c    -> oracle/liborc.so)
 *       a plain-C restatement of the same algorithms, each function citing the
 *       reference file:line it follows.
 *
 * Nothing under oracle/ is product code: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load these libraries.
 * The product path is include/blasr_gpu.h + blasr_b200/csrc (CUDA only, no CPU fallback).
 */
#ifndef ORC_ALIGN_H_
#define ORC_ALIGN_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* algorithms (mirrors bgpu_algo in include/blasr_gpu.h) */
enum { ORC_GUIDED = 0, ORC_AFFINE_GUIDED = 1, ORC_KBAND = 2, ORC_SW = 3, ORC_AFFINE_KBAND = 4 };
/* score function kinds */
enum { ORC_FN_DISTANCE = 0, ORC_FN_QUALITY = 1, ORC_FN_IDS = 2 };
/* AlignmentType ordinals, common/algorithms/alignment/AlignmentUtils.h:14-58 */
enum { ORC_LOCAL = 0, ORC_GLOBAL = 1, ORC_QUERYFIT = 2, ORC_TARGETFIT = 3, ORC_OVERLAP = 4,
       ORC_FRONTANCHORED = 5, ORC_ENDANCHORED = 6, ORC_FIT = 7, ORC_TSUFFIXQPREFIX = 8,
       ORC_TPREFIXQSUFFIX = 9 };
/* status */
enum { ORC_OK = 0, ORC_EMPTY_GUIDE = 1, ORC_PATH_AWRY = 2, ORC_BAD_INPUT = 3, ORC_OVERFLOW = 4 };

typedef struct {
  int32_t M[25];          /* scoreMatrix[5][5] row-major (DistanceMatrixScoreFunction.h:24) */
  int32_t ins, del;       /* BaseScoreFunction.h:6-7 */
  int32_t affineOpen, affineExtend; /* BaseScoreFunction.h:10-11 */
  int32_t kind;           /* ORC_FN_* */
  int32_t substitutionPrior, globalDeletionPrior; /* BaseScoreFunction.h:8-9 (IDSScoreFunction defaults 20 / 13) */
} orc_scorefn;

typedef struct {
  int32_t  algo;          /* ORC_GUIDED ... */
  int32_t  alignType;     /* AlignmentType ordinal */
  int32_t  band;          /* bandSize (guided) or k (kband); unused for SW */
  int32_t  bndIns, bndDel;/* KBandAlign's int ins/del parameters (KBandAlign.h:116,121) */
  int32_t  doStats;       /* run ComputeAlignmentStats afterwards */
  int32_t  statsAffine;   /* its useAffineScore argument */
  const uint8_t *q; uint32_t qLen;   /* ASCII */
  const uint8_t *t; uint32_t tLen;
  const uint8_t *qual;    /* qLen QVs or NULL (required for ORC_FN_QUALITY) */
  const uint32_t *guide;  /* nGuide x {qPos,tPos,length} */
  uint32_t nGuide;
  /* rich QV tracks of FASTQSequence (FASTQSequence.h:19-26), qLen bytes each, used by ORC_FN_IDS:
   * insQV, subQV, subTag are required; delQV + delTag are optional as a pair (IDSScoreFunction.h:85) */
  const uint8_t *insQV, *delQV, *subQV, *delTag, *subTag;
  /* AffineKBandAlign's int parameters (AffineKBandAlign.h:14-15); its `del` is bndDel, its matchMat is fn->M */
  int32_t  hpInsOpen, hpInsExtend, insOpen, insExtend;
} orc_job;

typedef struct {
  int32_t  status;
  int32_t  score;         /* value returned by the aligner */
  int32_t  alnScore;      /* alignment.score after the aligner (KBandAlign leaves it 0) */
  uint32_t qPos, tPos;
  int32_t  nCells;
  int32_t  nMatch, nMismatch, nIns, nDel;
  float    pctSimilarity;
  int32_t  statsScore;    /* alignment.score after ComputeAlignmentStats */
  uint32_t nBlocks, nGapLists, nGaps;
} orc_result;

/*
 * blocks:    capBlocks x {qPos,tPos,length}
 * gapCounts: one entry per GapList (alignment.gaps[i].size())
 * gaps:      flattened {seq(0=Gap::Query,1=Gap::Target), length} pairs
 * Returns 0, or ORC_OVERFLOW if a capacity was too small (counts are still set).
 */
typedef int (*orc_align_fn)(const orc_scorefn *fn, const orc_job *job, orc_result *res,
                            uint32_t *blocks, uint32_t capBlocks,
                            uint32_t *gapCounts, uint32_t capGapLists,
                            int32_t *gaps, uint32_t capGaps);

int orc_align(const orc_scorefn *fn, const orc_job *job, orc_result *res,
              uint32_t *blocks, uint32_t capBlocks, uint32_t *gapCounts, uint32_t capGapLists,
              int32_t *gaps, uint32_t capGaps);
int ref_align(const orc_scorefn *fn, const orc_job *job, orc_result *res,
              uint32_t *blocks, uint32_t capBlocks, uint32_t *gapCounts, uint32_t capGapLists,
              int32_t *gaps, uint32_t capGaps);

/* SAM CIGAR core (printers/SAMPrinter.h:203-293, CreateNoClippingCigarOps with AddGaps :120-137 and
 * AddUngappedOperations :138-166), BAM-packed: length << 4 | code, '=' 7, 'X' 8, 'I' 1, 'D' 2.
 *   ref_cigar       runs the job's aligner and then the reference's own printer code on the result;
 *   orc_cigar_from  the C restatement, from a stored alignment (blocks / gap lists as ref_align / orc_align return them).
 * Both return the number of ops (0 for an alignment without blocks), -1 when capOps is too small. */
int ref_cigar(const orc_scorefn *fn, const orc_job *job, uint32_t *ops, uint32_t capOps);
int orc_cigar_from(const uint8_t *q, const uint8_t *t, uint32_t qPos, uint32_t tPos,
                   const uint32_t *blocks, uint32_t nBlocks, const uint32_t *gapCounts, uint32_t nGapLists,
                   const int32_t *gaps, uint32_t *ops, uint32_t capOps);

/* CPU model of the next fill kernel's number format (oracle/orc_s16.c): the linear GuidedAlign sweep with 16-bit slots
 * relative to an offset re-based every 64 anti-diagonals, next to the same sweep in 32 bits.  out[9]: cells compared,
 * arrow mismatches, score mismatches, end score (32-bit), end score (16-bit + offset), min / max legit relative value
 * (shifted by 2), max unreachable value, re-bases.  Returns 0, -1 on unsupported input. */
int orc_guided_s16_model(const orc_scorefn *fn, const orc_job *job, int64_t *out);

/* The three strings the m5 / stick printers print (CreateAlignmentStrings, AlignmentUtils.h:390-533): text, match pattern,
 * query, each capOut bytes at most; returns their common length (0 without blocks), -1 on overflow.
 *   ref_alignment_strings runs the job's aligner first and prints from its result;
 *   orc_alignment_strings prints from a stored alignment (nGapLists == 0: the block-only form SDPAlign returns). */
int ref_alignment_strings(const orc_scorefn *fn, const orc_job *job, char *textStr, char *alignStr, char *queryStr, uint32_t capOut);
int ref_block_strings(const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, const uint32_t *blocks, uint32_t nBlocks,
                      char *textStr, char *alignStr, char *queryStr, uint32_t capOut);   /* a block-only alignment at qPos = tPos = 0 */
int orc_alignment_strings(const uint8_t *query, const uint8_t *text, uint32_t qPos, uint32_t tPos,
                          const uint32_t *blocks, uint32_t nBlocks, const uint32_t *gapCounts, uint32_t nGapLists,
                          const int32_t *gaps, char *textStr, char *alignStr, char *queryStr, uint32_t capOut);

/* The chaining step of SDPAlign (next scope row): SDPLongestCommonSubsequence
 * (sdp/SparseDynamicProgramming.h:71-322) over a fragment set with unique (x, y).
 * frags: n x {x, y, length, weight}.  chain: indices into the set sorted by (x, y), first fragment first.
 * Returns the chain length, -1 when capChain is too small.  alignType: ORC_GLOBAL or ORC_LOCAL. */
int orc_sdp_chain(const uint32_t *frags, uint32_t n, uint32_t queryLength, uint32_t fragmentLength,
                  int insertion, int deletion, int match, int alignType, int32_t *chain, uint32_t capChain);
int ref_sdp_chain(const uint32_t *frags, uint32_t n, uint32_t queryLength, uint32_t fragmentLength,
                  int insertion, int deletion, int match, int alignType, int32_t *chain, uint32_t capChain);

/* The fragment set SDPAlign hands to the chain (SDPAlign.h:133-262): prefix / middle / suffix k-mer matches, sorted by
 * (x, y), de-duplicated.  frags: n x {x, y, length, weight}.  Returns n, -1 when capFrags is too small, -2 when the
 * restated std::sort would have to fall back to heapsort (not restated: result unpinned for that input).
 *   ref_sdp_fragments runs the reference's SDPAlign (no detailed alignment, no recursion) and returns the fragment set
 *   it left in its buffers, plus its chain (indices into that set). */
int orc_sdp_fragments(const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, int wordSize, int sdpPrefixLength,
                      uint32_t *frags, uint32_t capFrags);
int ref_sdp_fragments(const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, const orc_scorefn *fn, int wordSize,
                      int sdpIns, int sdpDel, int alignType, uint32_t *frags, uint32_t capFrags,
                      int32_t *chain, uint32_t capChain, int32_t *nChain);

/* SDPAlign as a whole (SDPAlign.h:95-637), the call that produces the guide of the refinement (Blasr.cpp:1716-1722 passes
 * Local, detailed, no front extension, prefix 50, recurse 2, noRecurseUnder 1000): blocks relative to (*qPos, *tPos).
 * Returns nBlocks, -1 when capBlocks is too small.  The reference counterpart is ref_sdp_guide below. */
int orc_sdp_align(const orc_scorefn *fn, const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, int wordSize,
                  int sdpIns, int sdpDel, float indelRate, int alignType, int detailed, int extendFront, int sdpPrefixLength,
                  int recurse, int noRecurseUnder, uint32_t *blocks, uint32_t capBlocks, uint32_t *qPos, uint32_t *tPos);

/* Guide rows exactly as AlignmentToGuide builds them (GuidedAlign.h:104-259):
 * rows[i] = {q, t, tPre, tPost}; returns number of rows (0 for an empty guide),
 * -1 if capRows is too small. nCells = sum(tPre+tPost+1) (GuidedAlign.h:83-92). */
int orc_guide_rows(const uint32_t *guide, uint32_t nGuide, int band,
                   int32_t *rows, uint32_t capRows, int64_t *nCells);
int ref_guide_rows(const uint32_t *guide, uint32_t nGuide, int band,
                   int32_t *rows, uint32_t capRows, int64_t *nCells);

/* reference-only: SDPAlign exactly as blasr's AlignIntervals calls it (Blasr.cpp:1716-1722),
 * returning absolute blocks (qPos/tPos folded in) usable as a guide. Returns nBlocks or -1. */
int ref_sdp_guide(const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen,
                  const orc_scorefn *fn, int tupleSize, int sdpIns, int sdpDel, float indelRate,
                  uint32_t *blocks, uint32_t capBlocks);

/* multi-threaded replay used for the CPU baseline: runs jobs[0..n) on nThreads, returns
 * total nCells; per-job results are discarded except score sum (anti-DCE). */
int64_t orc_replay(const orc_scorefn *fn, const orc_job *jobs, uint32_t n, int nThreads, int64_t *scoreSum);
int64_t ref_replay(const orc_scorefn *fn, const orc_job *jobs, uint32_t n, int nThreads, int64_t *scoreSum);

#ifdef __cplusplus
}
#endif
#endif

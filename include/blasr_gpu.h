/*
 * blasr_gpu.h -- C ABI of the B200-native refinement DP (drop-in boundary).
 *
 * The reference (mchaisso/blasr) exposes no FFI; its "boundary" for this path is a set of
 * C++ template call sites.  Each entry point below replaces the per-candidate CPU call named
 * beside it with a batched, device-side one that returns the same score, coordinates, blocks,
 * gaps and statistics:
 *
 *   AffineGuidedAlign(...)   common/algorithms/alignment/AffineGuidedAlign.h:31   called at alignment/Blasr.cpp:863
 *   GuidedAlign(...)         common/algorithms/alignment/GuidedAlign.h:278        called at alignment/Blasr.cpp:869
 *   KBandAlign(...)          common/algorithms/alignment/KBandAlign.h:75          called at alignment/Blasr.cpp:717,820
 *   SWAlign(...)             common/algorithms/alignment/SWAlign.h:18             called at common/algorithms/alignment/SDPAlign.h:440,503,563
 *   AffineKBandAlign(...)    common/algorithms/alignment/AffineKBandAlign.h:12    called at alignment/Blasr.cpp:695,1067
 *   ComputeAlignmentStats    common/algorithms/alignment/AlignmentUtils.h:535     called at alignment/Blasr.cpp:740,875
 *
 * Plain pointers and sizes only; no C++/torch types.  All work runs in hand-written sm_100a
 * kernels; there is no CPU fallback -- every call fails with BGPU_E_NO_DEVICE without a GPU.
 */
#ifndef BLASR_GPU_H_
#define BLASR_GPU_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BGPU_VERSION 107 /* 0.1.7: + bgpu_set_suffix_array / bgpu_map_reads (0.1.6: bgpu_set_reference / bgpu_batch.tRefOff; 0.1.5: bgpu_sdp_align; 0.1.4: compact results, packed guides, ...) */

/* ---- return codes (API level) ---- */
enum {
  BGPU_OK = 0,
  BGPU_E_NO_DEVICE = -1,   /* no CUDA device / driver: the product path never falls back to CPU */
  BGPU_E_CUDA = -2,        /* a CUDA call failed; see bgpu_last_error() */
  BGPU_E_INVALID = -3,     /* bad argument */
  BGPU_E_OOM = -4,         /* device or pinned allocation failed */
  BGPU_E_BUSY = -5         /* ticket not in the state the call needs */
};

/* ---- per-job status (bgpu_result.status) ---- */
enum {
  BGPU_JOB_OK = 0,
  BGPU_JOB_EMPTY_GUIDE = 1,   /* guided aligners: no blocks -> score 0, empty alignment (GuidedAlign.h:388-392) */
  BGPU_JOB_PATH_AWRY = 2,     /* traceback met NoArrow: the reference prints and exit(1)s (GuidedAlign.h:637-651) */
  BGPU_JOB_BAD_INPUT = 3,     /* base outside ThreeBit 0..4, guide not monotone / outside the sequences, k<0 ... */
  BGPU_JOB_REF_UNDEFINED = 4, /* the reference itself is undefined here (e.g. KBandAlign TargetFit with k > tLen) */
  BGPU_JOB_TOO_WIDE = 5,      /* band wider than this build's widest kernel (BGPU_MAX_DIAGONALS) */
  BGPU_JOB_RANGE = 6          /* scores could exceed the 26-bit range the kernels carry */
};

/* ---- enums mirrored from the reference ---- */
typedef enum { BGPU_GUIDED = 0, BGPU_AFFINE_GUIDED = 1, BGPU_KBAND = 2, BGPU_SW = 3,
               BGPU_AFFINE_KBAND = 4   /* Global and QueryFit; see the note at bgpu_params */ } bgpu_algo;
/* AlignmentType ordinals, AlignmentUtils.h:14-58 */
typedef enum {
  BGPU_LOCAL = 0, BGPU_GLOBAL = 1, BGPU_QUERYFIT = 2, BGPU_TARGETFIT = 3, BGPU_OVERLAP = 4,
  BGPU_FRONTANCHORED = 5, BGPU_ENDANCHORED = 6, BGPU_FIT = 7, BGPU_TSUFFIXQPREFIX = 8, BGPU_TPREFIXQSUFFIX = 9
} bgpu_align_type;
typedef enum { BGPU_FN_DISTANCE = 0, /* DistanceMatrixScoreFunction.h:11 */
               BGPU_FN_QUALITY = 1,  /* QualityValueScoreFunction.h:9    */
               BGPU_FN_IDS = 2       /* IDSScoreFunction.h:21 (rich QV tracks: GuidedAlign, AffineGuidedAlign, KBandAlign;
                                        SWAlign x IDS reads out of bounds in the reference, SWAlign.h:166-167 -> BGPU_E_INVALID) */
             } bgpu_fn_kind;

/* ---- PODs mirroring the reference's containers ---- */
typedef struct { uint32_t qPos, tPos, length; } bgpu_block;   /* Block, datastructures/alignment/AlignmentBlock.h:17 */
typedef struct { int32_t seq;  /* 0 = Gap::Query (deletion), 1 = Gap::Target (insertion) */
                 int32_t length; } bgpu_gap;                  /* Gap, datastructures/alignment/AlignmentGapList.h:9-24 */

typedef struct {
  int32_t M[25];            /* scoreMatrix[5][5], row = query code (DistanceMatrixScoreFunction.h:100-105) */
  int32_t ins, del;         /* BaseScoreFunction.h:6-7 */
  int32_t affineOpen, affineExtend; /* BaseScoreFunction.h:10-11 */
  int32_t kind;             /* bgpu_fn_kind */
  int32_t substitutionPrior, globalDeletionPrior; /* BaseScoreFunction.h:8-9; IDSScoreFunction defaults 20 / 13 (:29-32) */
} bgpu_scorefn;

typedef struct {
  int32_t algo;             /* bgpu_algo */
  int32_t alignType;        /* bgpu_align_type */
  int32_t band;             /* bandSize (guided) / k (KBandAlign) when batch.band == NULL */
  int32_t bndIns, bndDel;   /* KBandAlign's int ins/del *parameters* (boundary costs, KBandAlign.h:116,121) */
  int32_t doStats;          /* also run ComputeAlignmentStats (fills nMatch.. statsScore) */
  int32_t statsAffine;      /* its useAffineScore argument (blasr: params.affineAlign) */
  /* AffineKBandAlign's int parameters (AffineKBandAlign.h:14-15; blasr passes indel+2, indel-3, indel+2, indel-1 and
   * del = indel, Blasr.cpp:1067-1071).  Its `del` travels as bndDel, its matchMat is scorefn.M (row = query), k is band.
   * Global and QueryFit are implemented; TargetFit -> BGPU_JOB_REF_UNDEFINED (the reference's end search reads mirrored
   * band columns, :322-330, and its traceback then spins on NoArrow cells).  It returns blocks / gaps / score only:
   * qPos, tPos and nCells are never set by the reference and come back 0. */
  int32_t hpInsOpen, hpInsExtend, insOpen, insExtend;
  /* Guided aligners: return every alignment as its run-length path instead of Block / Gap arrays (a third of the bytes over
   * PCIe).  arena.runs then holds, for job i, nBlocks + nGaps words from runs[blockOff + gapOff]: type << 30 | length in
   * path order from the first block to the last, type 0 = a block (advance q and t), 1 = Gap::Target (an insertion: advance
   * q), 2 = Gap::Query (a deletion: advance t); arena.blocks / gapCounts / gaps are NULL.  blasr_gpu::RefineBatch::Store
   * expands them into Alignment::blocks / gaps (block positions relative to qPos / tPos, gaps[0] and the last list empty). */
  int32_t compactResults;
} bgpu_params;

/* A batch in structure-of-arrays form.  Sequences are ASCII exactly as DNASequence::seq holds
 * them; job i uses qBases[qOff[i] .. qOff[i+1]) etc.  For the guided aligners the caller has
 * already sliced q/t the way RefineAlignment does (Blasr.cpp:850-859): guide blocks are used
 * raw, the alignment runs from guide.front() to guide.back(). */
typedef struct {
  uint32_t nJobs;
  const uint8_t  *qBases; const uint64_t *qOff;     /* nJobs+1 offsets */
  const uint8_t  *tBases; const uint64_t *tOff;     /* nJobs+1 offsets */
  const uint8_t  *qual;                             /* QVs parallel to qBases, or NULL (needed for BGPU_FN_QUALITY) */
  const bgpu_block *guide; const uint64_t *guideOff;/* nJobs+1 offsets; guided algos only, else NULL */
  const int32_t  *band;                             /* per-job band / k, or NULL -> params.band */
  /* FASTQSequence's rich QV tracks (FASTQSequence.h:19-26), parallel to qBases; read by BGPU_FN_IDS only:
   * insertionQV, substitutionQV, substitutionTag are required there, deletionQV + deletionTag are optional as a pair
   * (without them Deletion() is the constant del, IDSScoreFunction.h:85-101).  mergeQV is never read (`if (false)`, :82). */
  const uint8_t  *insQV, *delQV, *subQV, *delTag, *subTag;
  /* Optional packed form of the guides (a quarter of the bytes over PCIe); when guidePacked != NULL, `guide` is ignored
   * (guideOff, the per-job block offsets, is still required).  Block g of the batch is the three bytes guidePacked[3g..3g+2]:
   * qPos - (end of the previous block of the job in q), tPos - (end of the previous block in t) -- both absolute for a job's
   * first block -- and length.  A block with a value >= 255 stores 255,255,255 there and its three numbers in guideWide:
   * nGuideWide entries {g, dq, dt, length} sorted by g.  (blasr_gpu::RefineBatch::Add writes this form.) */
  const uint8_t  *guidePacked; const uint32_t *guideWide; uint64_t nGuideWide;
  /* Targets taken from the reference resident on the device (bgpu_set_reference) instead of tBases: when tRefOff != NULL,
   * job i's target is reference[tRefOff[i] .. tRefOff[i] + (tOff[i+1] - tOff[i])), reverse-complemented when tRefRc != NULL
   * and tRefRc[i] != 0 (the window is then read backwards from tRefOff[i] + length - 1; ACGT / acgt are complemented as
   * ReverseComplementNuc does, NucConversion.h:337-352, N / n and every other byte are kept -- the reference's table turns
   * the latter into 127, an invalid base); tBases is ignored (tOff still gives the lengths).  Guided aligners only.
   * blasr's targets are windows of the genome it holds in RAM for the whole run (Blasr.cpp:850-853 copies them out per
   * candidate): with the genome on the device only 8 bytes per job cross PCIe instead of the window. */
  const uint64_t *tRefOff; const uint8_t *tRefRc;
} bgpu_batch;

/* One job in pointer form (what a per-candidate call site has in hand). */
typedef struct {
  const uint8_t *q; uint32_t qLen;
  const uint8_t *t; uint32_t tLen;
  const uint8_t *qual;                  /* nullable */
  const bgpu_block *guide; uint32_t nGuide;
  int32_t band;
  const uint8_t *insQV, *delQV, *subQV, *delTag, *subTag;   /* nullable, qLen each (BGPU_FN_IDS) */
} bgpu_job;

typedef struct {
  int32_t  status;          /* BGPU_JOB_* */
  int32_t  score;           /* the aligner's return value */
  uint32_t qPos, tPos;      /* alignment.qPos / tPos */
  int32_t  nCells;          /* alignment.nCells (0 for SWAlign, which never sets it) */
  int32_t  nMatch, nMismatch, nIns, nDel;   /* ComputeAlignmentStats, when params.doStats */
  float    pctSimilarity;
  int32_t  statsScore;      /* alignment.score after ComputeAlignmentStats (what SAM AS:i / m4 / m5 print) */
  uint32_t nBlocks;   uint64_t blockOff;    /* arena.blocks[blockOff .. +nBlocks) */
  uint32_t nGapLists; uint64_t gapListOff;  /* arena.gapCounts[gapListOff .. +nGapLists) = alignment.gaps[i].size() */
  uint32_t nGaps;     uint64_t gapOff;      /* arena.gaps[gapOff .. +nGaps), lists concatenated in order */
} bgpu_result;

typedef struct {
  const bgpu_block *blocks;    uint64_t nBlocks;
  const uint32_t   *gapCounts; uint64_t nGapLists;
  const bgpu_gap   *gaps;      uint64_t nGaps;
  const uint32_t   *runs;      uint64_t nRuns;   /* params.compactResults: the run-length paths, else NULL */
} bgpu_arena;   /* pinned host memory owned by the library until bgpu_release() */

typedef struct {
  double   msPrep, msFill, msTrace, msEmit; /* CUDA-event times of the last run of this ticket, per stage */
  double   msTotal;                         /* first kernel start -> last kernel end */
  uint64_t cells;                           /* sum of nCells definitions of SURVEY 8(d) over the batch */
  uint64_t fillCells;                       /* lane-steps executed by the fill kernels (incl. idle lanes) */
  uint32_t kernelLaunches;                  /* kernels launched by the last run */
  uint64_t h2dBytes, d2hBytes;              /* bytes copied by submit / collect */
  double   msHostSubmit, msHostCollect;     /* wall time spent inside bgpu_submit / bgpu_collect */
  uint32_t devAllocs, pinAllocs;            /* cudaMalloc / cudaHostAlloc calls made by the context so far (allocation-cache misses) */
} bgpu_timing;

typedef struct bgpu_ctx bgpu_ctx;
typedef struct bgpu_ticket_s *bgpu_ticket;

/* ---- lifecycle ---- */
int  bgpu_create(bgpu_ctx **ctx, int device);       /* binds to one GPU; one ctx per host thread or shared (calls are serialised) */
void bgpu_destroy(bgpu_ctx *ctx);
const char *bgpu_last_error(const bgpu_ctx *ctx);
/* A context caches the device / pinned slabs of released tickets for the next ones; bgpu_trim gives the idle ones back to the
 * driver (a failing allocation also trims the idle slabs of the other contexts on the same device before it gives up). */
int  bgpu_trim(bgpu_ctx *ctx);
int  bgpu_version(void);
/* The base -> code table every kernel uses: ThreeBit[] of common/NucConversion.h:48-84 (0..3 ACGT, 4 = N / IUPAC,
 * 5 = '$', 255 = not a base -> BGPU_JOB_BAD_INPUT).  Host-callable so that the table can be pinned without a GPU. */
int  bgpu_base_code(int asciiByte);

/* The reference (genome) the targets of later batches may be windows of (bgpu_batch.tRefOff): copied to the device once,
 * shared by every context on that device, replaced by the next call, freed with n == 0.  Synchronous. */
int  bgpu_set_reference(bgpu_ctx *ctx, const uint8_t *bases, uint64_t n);

/* ---- asynchronous batch API ---- */
/* Copies the batch into pinned staging, enqueues H2D + all kernels on the ctx stream, returns at once. */
int  bgpu_submit(bgpu_ctx *ctx, const bgpu_scorefn *fn, const bgpu_params *p, const bgpu_batch *b, bgpu_ticket *out);
/* Pointer-form convenience: gathers jobs[] into a batch, then bgpu_submit. */
int  bgpu_submit_jobs(bgpu_ctx *ctx, const bgpu_scorefn *fn, const bgpu_params *p, const bgpu_job *jobs,
                      uint32_t nJobs, bgpu_ticket *out);
/* Blocks until the ticket's kernels are done, copies results to the host. results[nJobs]. */
int  bgpu_collect(bgpu_ctx *ctx, bgpu_ticket t, bgpu_result *results, bgpu_arena *arena);
/* 1 when bgpu_collect(t) would not wait for the device any more, 0 while the ticket's work is still in flight (a host
 * that multiplexes many readers over few threads polls this instead of blocking in bgpu_collect), < 0 on error. */
int  bgpu_query(bgpu_ctx *ctx, bgpu_ticket t);
int  bgpu_release(bgpu_ctx *ctx, bgpu_ticket t);
/* Re-executes every kernel of an already submitted ticket on its device-resident inputs
 * (benchmarking: inputs stay in HBM); synchronous. */
int  bgpu_rerun(bgpu_ctx *ctx, bgpu_ticket t);
int  bgpu_timing_of(bgpu_ctx *ctx, bgpu_ticket t, bgpu_timing *out);

/* ---- SAM CIGAR core of every alignment of a collected GuidedAlign / AffineGuidedAlign ticket, built on the device.
 * Replaces SAMOutput::CreateNoClippingCigarOps (common/algorithms/alignment/printers/SAMPrinter.h:203-293, with AddGaps
 * :120-137 and AddUngappedOperations :138-166): per block, maximal runs of unequal / equal RAW bytes as 'X' / '=', every
 * stored Gap as 'D' (Gap::Query) or 'I' (Gap::Target), nothing merged; clipping ops and the strand reversal stay with
 * the caller (CreateCIGARString :330-400 wraps this core).  ops are BAM-packed (length << 4 | code; '=' 7, 'X' 8, 'I' 1,
 * 'D' 2); job i owns ops[cigarOff[i] .. cigarOff[i+1]) (empty for jobs without blocks).  Both arrays are pinned host
 * memory owned by the library until bgpu_release(). */
int  bgpu_cigar(bgpu_ctx *ctx, bgpu_ticket t, const uint32_t **ops, const uint64_t **cigarOff);
/* The whole CreateCIGARString (SAMPrinter.h:345-400): the core above wrapped in the clipping ops and reversed for reverse-strand
 * alignments.  clips[4 * i .. 4 * i + 3] = hard-clipped prefix, soft-clipped prefix, soft-clipped suffix, hard-clipped suffix
 * of job i (the values SetHardClip :311-327 / SetSoftClip :295-309 compute from the read; 0 = no such op; NULL = no clipping,
 * `-clipping none`), tStrand[i] = alignment.tStrand (NULL = all forward).  Ops: 'H' 5, 'S' 4 and the core's codes; the
 * arrays are new pinned memory owned by the library until bgpu_release(). */
int  bgpu_cigar_clipped(bgpu_ctx *ctx, bgpu_ticket t, const uint32_t *clips, const uint8_t *tStrand, const uint32_t **ops,
                        const uint64_t **cigarOff);
/* The three strings of CreateAlignmentStrings (common/algorithms/alignment/AlignmentUtils.h:390-533, what
 * PrintCompareSequencesAlignment, printers/CompareSequencesAlignmentPrinter.h:17-89, prints for -m 5 and what StoreMapQVs
 * rescoring walks) for every alignment of a collected GuidedAlign / AffineGuidedAlign ticket, built on the device: text = target
 * bases / '-', align = '|' (TwoBit-equal) / '*' / ' ' (gap), query = query bases / '-'.  Job i owns [strOff[i], strOff[i+1]) of
 * each of the three arrays (no terminators); pinned host memory owned by the library until bgpu_release(). */
int  bgpu_strings(bgpu_ctx *ctx, bgpu_ticket t, const char **text, const char **align, const char **query, const uint64_t **strOff);

/* ---- The rescoring step of StoreMapQVs (alignment/Blasr.cpp:2768-2780): ComputeAlignmentScore(alignment, qAlignedSeq,
 * tAlignedSeq, scoreFn, useAffinePenalty) -- the Alignment overload, common/algorithms/alignment/AlignmentUtils.h:127-169: Match()
 * over every block plus one cost per Gap of the lists between blocks (affineOpen + length * affineExtend when that is smaller
 * and useAffinePenalty) -- of every alignment of a collected GuidedAlign / AffineGuidedAlign ticket under ANOTHER score function
 * (blasr: the caller's ins / del / affine costs with SMRTLogProbMatrix; probScore = -scores[i] / 10.0).  scores[nJobs] is the
 * caller's; jobs without an alignment get 0.  The partitioning of overlapping alignments and the phred arithmetic on these
 * scores (Blasr.cpp:2781-2920) stay with the caller: O(candidates^2) scalar work per read. */
int  bgpu_rescore(bgpu_ctx *ctx, bgpu_ticket t, const bgpu_scorefn *fn, int useAffinePenalty, int32_t *scores);

/* ---- SDPAlign (common/algorithms/alignment/SDPAlign.h:95-637), the step that produces the guide the refinement consumes
 * (called at alignment/Blasr.cpp:1716-1722 and :1080-1090): fragment set (k-mer matches of prefix / whole /
 * suffix), sparse-DP chain, chain -> blocks, and -- when `detailed` -- the SWAlign / recursive SDPAlign fills of the boxes
 * between chained blocks.  Field names are the reference's parameter names (SDPAlign.h:95-107).  Jobs are (query, target)
 * pairs of `b` (guide / band / qual ignored); results[i] carries status, qPos, tPos and the blocks (arena->blocks[blockOff ..
 * +nBlocks), positions relative to qPos / tPos as Alignment::blocks holds them; the other result fields are 0).  Synchronous;
 * arena is pinned memory owned by the library until the next bgpu_sdp_align on ctx or bgpu_destroy.  A job whose fragment
 * set outgrows its scratch slice (room for 2 (|q| + |t|) + 4096 fragments) comes back BGPU_JOB_RANGE. */
typedef struct {
  int32_t wordSize;                 /* blasr: params.sdpTupleSize (11) */
  int32_t sdpIns, sdpDel;           /* 5, 10 */
  float   indelRate;                /* params.indelRate * 3 */
  int32_t alignType;                /* BGPU_LOCAL (Blasr.cpp:1719) or BGPU_GLOBAL (AlignSubstring, :1075) */
  int32_t detailed;                 /* params.detailedSDPAlignment (true) */
  int32_t extendFront;              /* params.extendFrontAlignment (false) */
  int32_t sdpPrefix;                /* params.sdpPrefix (50) */
  int32_t recurse, noRecurseUnder;  /* params.recurse (2), params.recurseOver (1000) */
  int32_t maxMatches;               /* params.sdpMaxAnchorsPerPosition (0 = any number) */
} bgpu_sdp_params;
int  bgpu_sdp_align(bgpu_ctx *ctx, const bgpu_scorefn *fn, const bgpu_sdp_params *p, const bgpu_batch *b,
                    bgpu_result *results, bgpu_arena *arena);

/* ---- Suffix-array anchoring: MapReadToGenome (common/algorithms/anchoring/MapBySuffixArray.h:209-309, called
 * for the read and its reverse complement at alignment/Blasr.cpp:2282-2296) with LocateAnchorBoundsInSuffixArray (:24-207) and
 * SuffixArray::StoreLCPBounds / SearchLeftBound / SearchRightBound (common/datastructures/suffixarray/SuffixArray.h:928-1067,
 * 736-822) on the device, over an index that stays resident in HBM (8 B per base on the device -- position + the next ten bases, so that a probe of
 * the binary searches is one load: 25 GB for a human genome).
 *
 * bgpu_set_suffix_array copies the members of the reference's SuffixArray object the search reads -- index[n], and
 * startPosTable / endPosTable[4^lookupPrefixLength] with lookupPrefixLength (NULL / 0: no table, like a SuffixArray whose
 * startPosTable is NULL) -- to the device of ctx; the genome is the one bgpu_set_reference gave (same n; set it FIRST: the device index
 * entries carry its bases, and a later bgpu_set_reference invalidates them).
 * Shared by every context on that device, replaced by the next call, freed with n == 0.  Synchronous. */
int  bgpu_set_suffix_array(bgpu_ctx *ctx, const uint32_t *index, uint64_t n, const uint32_t *startPosTable,
                           const uint32_t *endPosTable, uint32_t lookupPrefixLength);
/* SuffixArray::BuildLookupTable (SuffixArray.h:193-250; blasr builds the table at start-up when the .sa file has none,
 * alignment/Blasr.cpp:4415-4419): startPosTable / endPosTable[4^lookupPrefixLength] from the genome and its suffix array, with
 * the reference's own edge behaviour (see bgpu_anchor.cu).  Index preparation, host code: needs no device. */
int  bgpu_build_lookup_table(const uint8_t *genome, uint64_t n, const uint32_t *index, uint32_t lookupPrefixLength,
                             uint32_t *startPosTable, uint32_t *endPosTable);
typedef struct {                      /* MapReadToGenome's scalar arguments; AnchorParameters.h:10-27 member names */
  uint32_t minPrefixMatchLength;      /* 4th argument; blasr passes params.lookupTableLength (8) */
  uint32_t minMatchLength;            /* anchorParameters.minMatchLength (12) */
  int32_t  expand;                    /* 0 .. 14 */
  int32_t  useLookupTable;            /* 1 */
  int32_t  maxAnchorsPerPosition;     /* 1000 */
  int32_t  advanceExactMatches;       /* 0 */
  int32_t  maxLCPLength;              /* 0 = no limit */
  int32_t  stopMappingOnceUnique;     /* 1 (MappingParameters.h:309) */
  int32_t  removeEncompassedMatches;  /* must be 0: the reference reads its vectors out of bounds there (:247-251) */
} bgpu_anchor_params;
typedef struct { uint32_t t, q, l; } bgpu_match;   /* MatchPos::t, q, l (datastructures/anchoring/MatchPos.h:10-21) */
/* Read i = reads[readOff[i] .. readOff[i + 1]) (ASCII as DNASequence::seq; pass the reverse complement as a read of its own,
 * as blasr does); subreadStart / subreadEnd per read, or NULL = whole reads.  matchOff[nReads + 1] (caller's) receives the
 * CSR offsets; *matches = every read's matchPosList in the reference's order (position ascending, then suffix-array order),
 * pinned memory owned by the library until the next bgpu_map_reads on ctx or bgpu_destroy.  Synchronous.  Bytes of the genome
 * at and beyond n read as 'N'.  BGPU_E_INVALID where the reference itself asserts or reads out of bounds (lookupPrefixLength >
 * minPrefixMatchLength with the table in use; minPrefixMatchLength > max(minMatchLength, lookupPrefixLength) + 2, :279). */
int  bgpu_map_reads(bgpu_ctx *ctx, const bgpu_anchor_params *p, const uint8_t *reads, const uint64_t *readOff, uint32_t nReads,
                    const uint32_t *subreadStart, const uint32_t *subreadEnd, uint64_t *matchOff, const bgpu_match **matches);
/* Device time of the kernels of the last bgpu_map_reads on ctx (CUDA events), ms: [0] locate, [1] count + scan + emit;
 * positions searched and H2D / D2H bytes. */
int  bgpu_map_timing(bgpu_ctx *ctx, double ms[2], uint64_t *positions, uint64_t *h2dBytes, uint64_t *d2hBytes);
/* Re-executes the kernels of the last bgpu_map_reads on its device-resident reads (benchmarking); synchronous. */
int  bgpu_map_rerun(bgpu_ctx *ctx);

/* ---- synchronous one-shot: submit + collect; arena valid until the next call on ctx ---- */
int  bgpu_align(bgpu_ctx *ctx, const bgpu_scorefn *fn, const bgpu_params *p, const bgpu_batch *b,
                bgpu_result *results, bgpu_arena *arena);

/* ---- device introspection used by bench.py ---- */
int  bgpu_device_count(void);
/* Measures the int32 ALU peak of the bound device with a dependent IADD3/VIMNMX chain kernel:
 * returns lane-ops per second (the integer roofline denominator). */
int  bgpu_measure_int_peak(bgpu_ctx *ctx, double *opsPerSec, double *smClockMHz);
/* Per-instruction-mix rates of the last bgpu_measure_int_peak call, lane-ops/s:
 * [0] add (ptxas splits it over IADD3/alu and IMAD.IADD/fma), [1] min (VIMNMX, alu only),
 * [2] mad (IMAD, fma only), [3] add+mad interleaved. */
int  bgpu_int_peak_modes(double out[4]);

#ifdef __cplusplus
}
#endif
#endif

// bgpu_common.cuh -- device-side data layout shared by the kernels of the refinement DP.
//
// Coordinates used by every guided kernel (see DESIGN.md "HBM layout"):
//   q' = q - qStart + 1  in [0, Qn]   (row 0 = the reference's boundary row, GuidedAlign.h:126-135)
//   t' = t - tStart + 1  in [0, Tn]   (column 0 = the boundary column tStart-1)
//   d  = q' + t'                      anti-diagonal, processed in d-blocks of 64
//   cd = t' - q' + C0                 diagonal index (C0 even >= Qn, so cd >= 0 and cd == d mod 2)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/blasr_gpu.h"

namespace bgpu {

constexpr int SH_LIN = 2;              // linear kernels carry (score << 2) | arrow
constexpr int SH_AFF = 5;              // affine kernels carry (score << 5) | arrow | open flags
constexpr int BIG = 1 << 30;           // "invalid neighbour" in the shifted domain (INF_INT stand-in)
constexpr int SCORE_LIMIT_AFF = 1 << 24;  // |score| bounds that keep (score << SH) clear of BIG
constexpr int SCORE_LIMIT_LIN = 1 << 27;
constexpr int DBLK = 64;               // anti-diagonals per d-block
constexpr int KRING = 6;               // most diagonal groups per lane in the register-ring kernels (6: the ring state still
                                       // fits 5 linear / 4 affine resident CTAs per SM; 8 and 5 measured slower)
constexpr int KWIDE = 128;             // groups per lane of the wide fallback kernel (32 lanes x 2 x 128 = 8192 diagonals)

// Job classes: how many lanes of a warp sweep one job.  A lane owns 2k consecutive diagonals ("lane-major"
// slots), k = groups active in the current d-block, so a class covers windows of up to 2 * LPJ * KRING diagonals.
// CLS_L8N is CLS_L8 restricted to k <= 4 (its kernel needs half the registers).
enum { CLS_L8N = 0, CLS_L8 = 1, CLS_L16 = 2, CLS_L32 = 3, CLS_WIDE = 4, N_CLS = 5 };
__host__ __device__ inline int cls_lpj(int cls) { return cls <= CLS_L8 ? 8 : (cls == CLS_L16 ? 16 : 32); }

// traceback codes written by the fill kernels
//   linear (2 bits / step, 16 steps per 32-bit word): 0 Diagonal, 1 Left, 2 Up, 3 NoArrow / out of band
//   affine (8 bits / step, 4 steps per word):
//     bits 0-2: 0 Diagonal, 1 Left, 2 Up, 3 AffineInsClose, 4 AffineDelClose, 7 NoArrow
//     bit 3   : affine-ins matrix arrow is AffineInsOpen (else AffineInsUp)
//     bit 4   : affine-del matrix arrow is AffineDelOpen (else AffineDelLeft)
// word layout per job: [d-block][row = (d & 63) / stepsPerWord][sigma = slot / 2], a row holding k * LPJ words.
enum { TB_DIAG = 0, TB_LEFT = 1, TB_UP = 2, TB_ICLOSE = 3, TB_DCLOSE = 4, TB_NONE = 7, TB_IOPEN = 8, TB_DOPEN = 16 };
enum { TL_DIAG = 0, TL_LEFT = 1, TL_UP = 2, TL_NONE = 3 };

struct alignas(8) RowInfo {  // 8 B per guide row in HBM, consumed as is by the fill kernels
  int32_t lo8;            // 8 * (diagonal of the row's first in-band cell), diagonal = t' - q' + C0
  int32_t nhi8;           // -8 * (diagonal of the row's last in-band cell)
  // (x 8: a lane turns the two into the byte offset of its row of the in-band mask table, see bgpu_fill.cu)
};
constexpr int DEAD_LO8 = 1 << 29;      // {DEAD_LO8, -DEAD_LO8}: a row no slot can be inside of
// The ring kernels stage a d-block's rows / target codes / query codes with ONE bulk copy (TMA) each, 16-byte aligned and
// without bounds checks: every job's band table carries ROW_PAD dead rows in front of row 0 and behind row Qn (a window
// reaches at most 6 * 32 rows past the live ones, plus the 32 of a block), and the per-base arrays (tc, qc, qual) sit
// BYTE_PAD bytes inside their allocations.
constexpr int ROW_PAD = 256;
constexpr int BYTE_PAD = 512;

struct DBlock {           // 16 B per d-block
  int32_t wbase;          // even: diagonal held by slot 0 of the job's window in this block
  int32_t k;              // prep: groups this job needs; fill: groups its warp actually used (>= needed)
  uint32_t arrowUnit;     // fill: offset of this block's arrows, in units of (64 / stepsPerWord) * LPJ words
  int32_t pad;
};

struct JobGeom {          // per job, written by prep, extended by fill / trace
  int32_t status;
  int32_t qStart, tStart; // first guide block (absolute positions inside the job's q / t)
  int32_t Qn, Tn;         // guide rows / target columns covered (without the boundary row/column)
  int32_t C0;
  int32_t nDB;            // number of d-blocks
  int32_t kmax;           // max k over the job's d-blocks (for its class)
  int32_t band;
  int32_t nCells;         // ComputeMatrixNElem (GuidedAlign.h:83-92)
  int32_t hi0;            // last in-band column of row 0
  int32_t score;          // fill result: S[qEnd-1][tEnd-1]
  int32_t cls;            // CLS_*
  int32_t ksum;           // sum of k over the d-blocks
  int32_t minW;           // cells of the job's narrowest band row (the ring kernels need rows >= a lane's 2k slots)
  uint64_t rowOff;        // RowInfo index of row 0
  uint64_t dblkOff;       // DBlock index of d-block 0
  uint64_t arrowBytes;    // unused by the guided path (the host bounds it per warp group)
  uint64_t runOff;        // u32 index into the run scratch (capacity Qn+Tn+2)
  uint32_t nRuns, nBlocks, nGaps, nGapLists;   // traceback results
  uint32_t qPos, tPos;    // alignment.qPos/tPos after RemoveAlignmentPrefixGaps
  int32_t startR, startC; // dense aligners (KBandAlign/SWAlign): traceback start cell chosen by the fill
  uint64_t rowBufOff;     // dense aligners: int index of the job's two score rows
};

struct ScoreParams {      // kernel argument (by value)
  int32_t M[25];
  int32_t ins, del, open, ext;
  int32_t kind, alignType, affine, pad;
  int32_t subPrior, delPrior;   // IDSScoreFunction: substitutionPrior, globalDeletionPrior (BaseScoreFunction.h:8-9)
};

struct BatchDev {         // device pointers of one submitted batch
  uint32_t nJobs;
  uint8_t *q; const uint64_t *qOff;
  const uint8_t *t; const uint64_t *tOff; // raw bytes as the caller passed them (emit / cigar compare these)
  uint8_t *tc;                            // the target as the fill kernels read it, written by prep: base codes 0..4
                                          // (raw bytes for BGPU_FN_IDS), indexed like t
  uint8_t *qc;                            // the query as the fill kernels read it, written by prep: base code * 20 (the
                                          // byte offset of the base's row in the 5x5 score table), indexed like q
  const uint8_t *qual;
  const uint8_t *insQV, *delQV, *subQV, *delTag, *subTag;   // IDSScoreFunction tracks (parallel to q), else NULL
  const bgpu_block *guide; const uint64_t *guideOff;
  const int32_t *band;
  JobGeom *geom;
  RowInfo *rows;
  DBlock *dblk;
  int32_t *dmin, *dmax;   // per d-block live diagonal range (prep scratch)
  uint8_t *arrows;
  const uint64_t *arrowOff; // per job: byte offset in the arrow pool (assigned by the host per wave)
  uint32_t *runs;
  int32_t *rowBuf;        // dense aligners: ping-pong score rows
  const uint32_t *order;  // job order for dynamic scheduling (longest first)
  uint32_t *counters;     // [0] fill job counter, [1] trace job counter, ...
  unsigned long long *cellSlots;  // cell slots executed by the guided fill kernels (lane-steps x groups), or NULL
};

// Device-resident schedule of one wave of a guided ticket: written by the planner kernels (bgpu_plan.cu) on the asynchronous
// path, uploaded by the host planner on the multi-wave path.  order[] holds, class after class, the warp groups (32 / LPJ
// job slots each, NOJOB pads a partial group) in dispatch order, then the traceback list.
struct PlanHead {
  uint32_t nGroups[N_CLS];       // warp groups per job class
  uint32_t orderBegin[N_CLS];    // first slot of the class in order[]
  uint32_t traceBegin, traceCount;
  uint32_t clsCount[N_CLS];      // jobs with status OK per class
  uint32_t nOk, nGroupsTotal, nSlots;
  uint32_t overflow;             // bit 0: the traceback pool is too small for this wave, bit 1: the result arena is
  unsigned long long arrowBytes; // traceback bytes reserved
  unsigned long long cells;      // sum of nCells (ComputeMatrixNElem) over the OK jobs
  unsigned long long laneSteps;  // lower bound of the cell slots the fill warps execute
  unsigned long long totals[3];  // blocks, gap lists, gaps of the whole ticket (scan_counts_kernel)
  unsigned long long caps[3];    // capacities of the result arena the emit kernel may write into
};
enum { PLAN_OVF_ARROWS = 1, PLAN_OVF_ARENA = 2 };

struct DenseArgs {
  int algo;            // BGPU_KBAND / BGPU_SW / BGPU_AFFINE_KBAND
  int defaultBand, bndIns, bndDel;
  int hpInsOpen, hpInsExtend, insOpen, insExtend;   // AffineKBandAlign.h:14 (its `del` is bndDel)
  const uint64_t *arrowOff;   // per job byte offset into B.arrows
};

__host__ __device__ __forceinline__ uint8_t base_code(uint8_t c) {
  // NucConversion.h:48-84 (ThreeBit), restated as arithmetic: ACGT/acgt and raw 0..3 -> 0..3, raw 4, the IUPAC
  // ambiguity letters in either case, 'x', 'y' (but not 'X') and '_' -> 4, '$' -> 5, everything else 255.
  // tests/test_host_cpu.py::test_base_code_table pins all 256 entries against the reference's own table.
  if (c <= 4) return c;
  const uint8_t u = c & 0xDF;  // fold case
  switch (u) {
    case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
    case 'B': case 'D': case 'H': case 'K': case 'M': case 'N': case 'R': case 'S':
    case 'U': case 'V': case 'W': case 'Y': return 4;
    default: break;
  }
  if (c == 'x' || c == '_') return 4;   // the table maps 'x' (but not 'X') and '_' to N
  if (c == '$') return 5;
  return 255;
}

}  // namespace bgpu

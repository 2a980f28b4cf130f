// bgpu_fill.cu -- the guided banded DP fill.
//
// Reference semantics restated (not translated):
//   GuidedAlign        common/algorithms/alignment/GuidedAlign.h:474-624   (linear gaps)
//   AffineGuidedAlign  common/algorithms/alignment/AffineGuidedAlign.h:241-375
//
// B200 mapping.  The band is swept by anti-diagonals d = q'+t' in blocks of 64.  A job is swept by LPJ lanes
// (8, 16 or 32: narrow bands put 4 or 2 jobs in one warp so the lanes stay busy); lane `sl` of the job owns the
// 2k consecutive diagonals [2k*sl, 2k*sl + 2k) of the job's window ("slots", k = groups active in this block).
// On an even step every lane computes the cells on its k even slots, on an odd step the k odd ones, so
//     even step, group g: left = own odd slot g-1 (g = 0: lane sl-1's top odd slot, one __shfl), up = own odd slot g
//     odd  step, group g: left = own even slot g, up = own even slot g+1 (g = k-1: lane sl+1's first even slot)
// i.e. ONE shuffle per lane-step however wide the band is.  Scores live in registers as (score << SH) | tag: the
// tie-ordered candidates carry their arrow code in the low bits, so one VIADDMNMX chain (DPX) yields both the
// minimum and the reference's first-match-wins arrow (Diagonal > Left > Up > AffineInsClose > AffineDelClose);
// the affine open/extend decisions land in two more tag bits the same way.  Cells outside the guide are held at
// BIG, which reproduces the reference's INF_INT-for-missing-neighbour rule.
//
// Staging.  The band table rows (RowInfo, 8 B, written by prep in the form the inner loop consumes) and the target
// codes of block b+1 are copied global -> shared with cp.async while block b is computed (double buffer), so the
// DP loop never waits on HBM.
//
// Inner loop (run_block_ring): the k rows and k columns a lane touches slide by one per two steps, so they are kept
// in register rings (the loop is unrolled by k, every ring index is static) and refilled with one shared-memory
// load per row / column; per cell that leaves one LDS (the substitution score), two VIADDMNMX, the in-band test
// and the tag split; one funnel shift appends the arrow to the lane's traceback word.  Words are
// stored [d-block][16-step row][slot pair], 2 bits per cell for the linear aligner (8 for the affine one).
// Blocks that touch the boundary row, a job's last block, the IDS score function and very wide windows take the
// generic path (run_block_gen), which reads its rows and columns from shared memory per cell.
#include "bgpu_common.cuh"

namespace bgpu {

constexpr uint32_t NOJOB = 0xffffffffu;
constexpr int DEAD_CD8 = 1 << 30;   // a row no slot can be inside of

struct FillConsts {
  int delT, insT;        // (del<<SH)|LEFT, (ins<<SH)|UP
  int extT3, extT4;      // (ext<<SH)|TB_ICLOSE, |TB_DCLOSE
  int ext, openI, openD; // ext<<SH, (open<<SH)|TB_IOPEN, (open<<SH)|TB_DOPEN
  int open;              // open<<SH
  int del0;              // row-0 step: (Global ? del : 0) << SH
  int k256, kacc;        // 256 and 1 << BITS held in registers the compiler cannot fold: keeps these
                         // multiply-adds on the FMA pipe (IMAD) instead of the busier ALU pipe
  int subPrior, delPrior, del;   // IDSScoreFunction: unshifted substitutionPrior / globalDeletionPrior / del
};

template <int LPJ, int KM, int FN>
struct SubSmem {          // staging of one job: two d-blocks (current, next)
  int2 rows[2][KM * LPJ + 32];           // RowInfo as prep wrote it: {cd8, (width << 8) | qcode * 20}
  uint32_t colw[2][(KM * LPJ + 44) / 4]; // target codes (bytes), 4-byte chunks from an aligned-down address
  int shift[2 * KM * LPJ];               // window re-mapping scratch
  // per-row score data of the current block (QualityValueScoreFunction needs none: the QV rides in RowInfo::cd8);
  // IDSScoreFunction: two words per row,
  //   [r]      query byte | substitutionTag << 8 | substitutionQV << 16 | insertionQV << 24
  //   [NR + r] deletionTag | deletionQV << 8 | (deletion tracks present) << 16
  static constexpr int NR = KM * LPJ + 32;
  int rowq[FN == 2 ? 2 * NR : 2];
};

template <bool AFFINE> struct Fmt {
  static constexpr int SHv = AFFINE ? SH_AFF : SH_LIN;
  static constexpr int BITS = AFFINE ? 8 : 2;
  static constexpr int SPW = 32 / BITS;                    // steps per traceback word
  static constexpr int TAGMASK = (1 << SHv) - 1;
  static constexpr int NONE = AFFINE ? (int)TB_NONE : (int)TL_NONE;
};

__device__ __forceinline__ int lds32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The min chain of one DP cell.
template <bool AFFINE>
__device__ __forceinline__ int dp_core(const int S, const int leftS, const int leftAD, const int upS, const int upAI,
                                       const int m, const int delT, const int insT, const FillConsts &c, int &ai, int &ad) {
  int cnd = S + m;                                         // Diagonal (tag 0)
  cnd = __viaddmin_s32(leftS, delT, cnd);                  // Left
  cnd = __viaddmin_s32(upS, insT, cnd);                    // Up
  if (AFFINE) {
    cnd = __viaddmin_s32(upAI, c.extT3, cnd);              // AffineInsClose
    cnd = __viaddmin_s32(leftAD, c.extT4, cnd);            // AffineDelClose
    // strict '<' in the reference: a tie extends (AffineGuidedAlign.h:357-373) -> the open candidate carries a
    // flag bit, so it loses ties.
    const int s0 = cnd & ~31;
    ai = __viaddmin_s32(s0, c.openI, upAI + c.ext);
    ad = __viaddmin_s32(s0, c.openD, leftAD + c.ext);
    cnd |= (ai | ad) & 24;                                  // tag | affine flags, still below bit 5
  }
  return cnd;
}

template <int LPJ>
__device__ __forceinline__ int sub_up(int v) { return __shfl_up_sync(0xffffffffu, v, 1, LPJ); }
template <int LPJ>
__device__ __forceinline__ int sub_dn(int v) { return __shfl_down_sync(0xffffffffu, v, 1, LPJ); }

// Per-block view of the staged data.  Row index r <-> q' = qlo + r, column index c <-> t' = tlo + c.
struct BlockView {
  const int2 *rows;       // staged RowInfo
  const uint8_t *cols;    // staged target codes, cols[c]
  int wbase8;             // window base diagonal << 8
  int qlo, tlo;
};

// ---------------------------------------------------------------------------------------------------------------
// Fast path: KA groups per lane (compile time), rows / columns in register rings, no boundary row, full 64 steps.
template <int LPJ, int KM, int KA, bool AFFINE, int FN>
__device__ __forceinline__ void run_block_ring(int (&Se)[KM], int (&So)[KM], int (&AIe)[KM], int (&AIo)[KM],
                                               int (&ADe)[KM], int (&ADo)[KM], const BlockView &bv,
                                               const int mtabAddr, const FillConsts &c, const int sl,
                                               uint32_t *aw, const bool live) {
  typedef Fmt<AFFINE> F;
  const int kL = KA * sl;
  const int2 *rp = bv.rows + KA * (LPJ - 1 - sl);           // rp[i + m]: row of group KA-1-m at step pair i
  const uint8_t *cp = bv.cols + kL;                          // cp[i + g]: column of group g on the even step of pair i
  const int s08 = bv.wbase8 + ((2 * kL) << 8);
  int rX[KA], rY[KA], rQ[KA], cT[KA];
  int rV[FN == 1 ? KA : 1];                                  // QualityValueScoreFunction: the rows' QVs
  uint32_t acc[KA];
#pragma unroll
  for (int m = 0; m < KA; m++) {
    const int2 v = rp[m];
    if (FN == 1) { rV[m] = v.x & 0xff; rX[m] = s08 - v.x + rV[m]; } else rX[m] = s08 - v.x;
    rY[m] = v.y; rQ[m] = v.y & 0xff;
    cT[m] = (int)cp[m] * 4 + mtabAddr;
    acc[m] = 0;
  }
  aw += kL;
#pragma unroll 1
  for (int i0 = 0; i0 < 32; i0 += KA) {
#pragma unroll
    for (int j = 0; j < KA; j++) {
      if ((32 % KA) != 0 && i0 + j >= 32) break;
      // ---- even step: cells on slots 2g
      {
        const int left0 = sub_up<LPJ>(So[KA - 1]);
        const int leftA0 = AFFINE ? sub_up<LPJ>(ADo[KA - 1]) : 0;
#pragma unroll
        for (int g = 0; g < KA; g++) {
          const int p = (j + KA - 1 - g) % KA, cs = (j + g) % KA;
          const int leftS = g == 0 ? left0 : So[g - 1];
          const int leftAD = AFFINE ? (g == 0 ? leftA0 : ADo[g - 1]) : 0;
          int m = lds32((uint32_t)(rQ[p] + cT[cs]));
          if (FN == 1) m *= rV[p];                            // +-(1<<SH) * QV  (QualityValueScoreFunction.h:78-83)
          int ai = 0, ad = 0;
          int cnd = dp_core<AFFINE>(Se[g], leftS, leftAD, So[g], AFFINE ? AIo[g] : 0, m, c.delT, c.insT, c, ai, ad);
          const bool inb = (unsigned)(c.k256 * (2 * g) + rX[p]) <= (unsigned)rY[p];
          cnd = inb ? cnd : (BIG | F::NONE);
          Se[g] = cnd & ~F::TAGMASK;
          if (AFFINE) { AIe[g] = inb ? (ai & ~31) : BIG; ADe[g] = inb ? (ad & ~31) : BIG; }
          acc[g] = __funnelshift_r(acc[g], (uint32_t)cnd, F::BITS);   // first step of a word ends up in its lowest field
        }
      }
      cT[j] = (int)cp[i0 + j + KA] * 4 + mtabAddr;          // column kL + i + KA replaces kL + i
      // ---- odd step: cells on slots 2g+1
      {
        const int up0 = sub_dn<LPJ>(Se[0]);
        const int upA0 = AFFINE ? sub_dn<LPJ>(AIe[0]) : 0;
#pragma unroll
        for (int g = 0; g < KA; g++) {
          const int p = (j + KA - 1 - g) % KA, cs = (j + 1 + g) % KA;
          const int upS = g == KA - 1 ? up0 : Se[g + 1];
          const int upAI = AFFINE ? (g == KA - 1 ? upA0 : AIe[g + 1]) : 0;
          int m = lds32((uint32_t)(rQ[p] + cT[cs]));
          if (FN == 1) m *= rV[p];
          int ai = 0, ad = 0;
          int cnd = dp_core<AFFINE>(So[g], Se[g], AFFINE ? ADe[g] : 0, upS, upAI, m, c.delT, c.insT, c, ai, ad);
          const bool inb = (unsigned)(c.k256 * (2 * g + 1) + rX[p]) <= (unsigned)rY[p];
          cnd = inb ? cnd : (BIG | F::NONE);
          So[g] = cnd & ~F::TAGMASK;
          if (AFFINE) { AIo[g] = inb ? (ai & ~31) : BIG; ADo[g] = inb ? (ad & ~31) : BIG; }
          acc[g] = __funnelshift_r(acc[g], (uint32_t)cnd, F::BITS);   // first step of a word ends up in its lowest field
        }
      }
      {                                                       // row baseR + i + KA replaces row baseR + i
        const int2 v = rp[i0 + j + KA];
        if (FN == 1) { rV[j] = v.x & 0xff; rX[j] = s08 - v.x + rV[j]; } else rX[j] = s08 - v.x;
        rY[j] = v.y; rQ[j] = v.y & 0xff;
      }
      if (((i0 + j) & (F::SPW / 2 - 1)) == F::SPW / 2 - 1) {  // a traceback word is complete
        if (live) {
#pragma unroll
          for (int g = 0; g < KA; g++) aw[g] = acc[g];
        }
        aw += KA * LPJ;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Generic path: run-time k, rows / columns read from shared memory per cell, boundary row (first) and the early
// stop of a job's last block (eLast) handled per cell.
template <int LPJ, int KM, bool AFFINE, int FN>
__device__ __forceinline__ void run_block_gen(int (&Se)[KM], int (&So)[KM], int (&AIe)[KM], int (&AIo)[KM],
                                              int (&ADe)[KM], int (&ADo)[KM], const BlockView &bv, const int *rowq,
                                              const int mtabAddr, const FillConsts &c, const int k, const int sl,
                                              const bool first, const int eLast, uint32_t *aw, const bool live) {
  typedef Fmt<AFFINE> F;
  constexpr int UG = KM <= KRING ? KM : 1;
  const int kL = k * sl, baseR = k * (LPJ - 1 - sl);
  uint32_t acc[KM];
#pragma unroll(UG)
  for (int g = 0; g < KM; g++) acc[g] = 0;
  aw += kL;
  auto cell = [&](int &S, int &AI, int &AD, const int leftS, const int leftAD, const int upS, const int upAI,
                  const int ridx, const int cidx, const int slot, const int e, uint32_t &a) {
    const int2 rv = bv.rows[ridx];
    int m, delT = c.delT, insT = c.insT;
    if (FN == 2) {
      // IDSScoreFunction.h:80-139 on the raw bytes: the row's tracks against the column's target byte
      constexpr int NR = KM * LPJ + 32;
      const int ra = rowq[ridx], rb = rowq[NR + ridx], tb = (int)bv.cols[cidx];
      const int stag = (ra >> 8) & 0xff, dtag = rb & 0xff;
      m = ((ra & 0xff) == tb) ? 0 : ((stag == tb ? ((ra >> 16) & 0xff) : c.subPrior) << F::SHv);
      insT = (int)(((unsigned)ra >> 24) << F::SHv) | TB_UP;
      const int dc = (rb >> 16) ? ((dtag != 'N' && dtag == tb) ? ((rb >> 8) & 0xff) : c.delPrior) : c.del;
      delT = (dc << F::SHv) | TB_LEFT;
    } else {
      m = lds32((uint32_t)(mtabAddr + (rv.y & 0xff) + (int)bv.cols[cidx] * 4));
      if (FN == 1) m *= rv.x & 0xff;                        // +-(1<<SH) * QV  (QualityValueScoreFunction.h:78-83)
    }
    int ai = 0, ad = 0;
    int cnd = dp_core<AFFINE>(S, leftS, leftAD, upS, upAI, m, delT, insT, c, ai, ad);
    if (first && bv.qlo + ridx == 0) {                      // boundary row (GuidedAlign.h:415-442)
      cnd = ((bv.tlo + cidx) * c.del0) | (AFFINE ? (TB_LEFT | TB_IOPEN | TB_DOPEN) : TL_LEFT);
      ai = c.open; ad = c.open;
    }
    const bool inb = (unsigned)(bv.wbase8 + (slot << 8) - (FN == 1 ? (rv.x & ~0xff) : rv.x)) <= (unsigned)rv.y;
    cnd = inb ? cnd : (BIG | F::NONE);
    if (e <= eLast) {
      S = cnd & ~F::TAGMASK;
      if (AFFINE) { AI = inb ? (ai & ~31) : BIG; AD = inb ? (ad & ~31) : BIG; }
    } else cnd = F::NONE;
    a = __funnelshift_r(a, (uint32_t)cnd, F::BITS);
  };
  auto pick = [&](int (&X)[KM], const int idx) {
    int v = BIG;
    if (KM <= KRING) {
#pragma unroll
      for (int g = 0; g < KM; g++) if (g == idx) v = X[g];
    } else v = X[idx];
    return v;
  };
#pragma unroll 1
  for (int i = 0; i < 32; i++) {
    {
      const int left0 = sub_up<LPJ>(pick(So, k - 1));
      const int leftA0 = AFFINE ? sub_up<LPJ>(pick(ADo, k - 1)) : 0;
      int prevS = left0, prevA = leftA0;
#pragma unroll(UG)
      for (int g = 0; g < KM; g++) {
        if (g < k) {
          const int so = So[g], ado = AFFINE ? ADo[g] : 0;
          cell(Se[g], AIe[g], ADe[g], prevS, prevA, so, AFFINE ? AIo[g] : 0, baseR + i + (k - 1 - g), kL + g + i,
               2 * (kL + g), 2 * i, acc[g]);
          prevS = so; prevA = ado;
        }
      }
    }
    {
      const int up0 = sub_dn<LPJ>(Se[0]);
      const int upA0 = AFFINE ? sub_dn<LPJ>(AIe[0]) : 0;
#pragma unroll(UG)
      for (int g = 0; g < KM; g++) {
        if (g < k) {
          int upS = up0, upAI = upA0;
          if (g + 1 < k) { upS = Se[(g + 1) % KM]; if (AFFINE) upAI = AIe[(g + 1) % KM]; }
          cell(So[g], AIo[g], ADo[g], Se[g], AFFINE ? ADe[g] : 0, upS, upAI, baseR + i + (k - 1 - g), kL + g + i + 1,
               2 * (kL + g) + 1, 2 * i + 1, acc[g]);
        }
      }
    }
    if ((i & (F::SPW / 2 - 1)) == F::SPW / 2 - 1) {
      if (live) {
#pragma unroll(UG)
        for (int g = 0; g < KM; g++) if (g < k) aw[g] = acc[g];
      }
      aw += k * LPJ;
    }
  }
}

template <int LPJ, int KM, bool AFFINE, int FN, int KA>
struct RingDispatch {
  static __device__ __forceinline__ void run(const int k, int (&Se)[KM], int (&So)[KM], int (&AIe)[KM], int (&AIo)[KM],
                                             int (&ADe)[KM], int (&ADo)[KM], const BlockView &bv, const int mtabAddr,
                                             const FillConsts &c, const int sl, uint32_t *aw, const bool live) {
    if (k == KA) run_block_ring<LPJ, KM, KA, AFFINE, FN>(Se, So, AIe, AIo, ADe, ADo, bv, mtabAddr, c, sl, aw, live);
    else RingDispatch<LPJ, KM, AFFINE, FN, KA + 1>::run(k, Se, So, AIe, AIo, ADe, ADo, bv, mtabAddr, c, sl, aw, live);
  }
};
template <int LPJ, int KM, bool AFFINE, int FN>
struct RingDispatch<LPJ, KM, AFFINE, FN, KM + 1> {
  static __device__ __forceinline__ void run(const int, int (&)[KM], int (&)[KM], int (&)[KM], int (&)[KM], int (&)[KM],
                                             int (&)[KM], const BlockView &, const int, const FillConsts &, const int,
                                             uint32_t *, const bool) {}
};

// order[] holds warp groups: 32 / LPJ job indices each (NOJOB pads the last group); a warp sweeps its jobs in
// lockstep from d-block 0, with k = the widest member's need per block.
template <int LPJ, int KM, bool AFFINE, int FN>
__global__ void __launch_bounds__(KM <= KRING ? 128 : 32, KM <= 4 ? (AFFINE ? 4 : 5) : (KM <= KRING ? (AFFINE ? 4 : 5) : 1))
fill_guided_kernel(BatchDev B, ScoreParams P, const uint32_t *orderBase, const PlanHead *plan, int cls, uint32_t *counter) {
  typedef Fmt<AFFINE> F;
  typedef SubSmem<LPJ, KM, FN> Smem;
  constexpr int NJ = 32 / LPJ;
  constexpr bool RING = KM <= KRING && FN <= 1;
  constexpr int UG = KM <= KRING ? KM : 1;
  constexpr int UNITW = (64 / F::SPW) * LPJ;                // words per arrow unit
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ int Mtab[25];
  // the schedule lives on the device (planner kernels or the host planner's upload): this class's warp groups
  if (plan->overflow & PLAN_OVF_ARROWS) return;                // the traceback pool cannot hold this wave: the host re-plans
  const uint32_t nGroups = plan->nGroups[cls];
  const uint32_t *order = orderBase + plan->orderBegin[cls];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPJ, sl = lane % LPJ;
  Smem &sm = reinterpret_cast<Smem *>(smemRaw)[warp * NJ + sub];
  if (threadIdx.x < 25) {
    if (FN == 1) { const int r = threadIdx.x / 5, cc = threadIdx.x % 5; Mtab[threadIdx.x] = ((r == cc && r < 4) ? -1 : 1) << F::SHv; }  // ScoreMatrices.h:4-10
    else Mtab[threadIdx.x] = P.M[threadIdx.x] << F::SHv;
  }
  __syncthreads();
  const int mtabAddr = (int)__cvta_generic_to_shared(Mtab);
  FillConsts c;
  c.delT = (P.del << F::SHv) | TB_LEFT; c.insT = (P.ins << F::SHv) | TB_UP;
  c.extT3 = (P.ext << F::SHv) | TB_ICLOSE; c.extT4 = (P.ext << F::SHv) | TB_DCLOSE;
  c.ext = P.ext << F::SHv; c.open = P.open << F::SHv;
  c.openI = (P.open << F::SHv) | TB_IOPEN; c.openD = (P.open << F::SHv) | TB_DOPEN;
  c.del0 = (P.alignType == BGPU_GLOBAL ? P.del : 0) << F::SHv;
  c.k256 = 256 + P.pad; c.kacc = (1 << F::BITS) + P.pad;    // P.pad is always 0
  c.subPrior = P.subPrior; c.delPrior = P.delPrior; c.del = P.del;

  for (;;) {
    uint32_t grp = 0;
    if (lane == 0) grp = atomicAdd(counter, 1u);
    grp = __shfl_sync(0xffffffffu, grp, 0);
    if (grp >= nGroups) break;
    const uint32_t job = order[(size_t)grp * NJ + sub];
    bool have = job != NOJOB;
    JobGeom *G = have ? &B.geom[job] : nullptr;
    if (have && G->status != BGPU_JOB_OK) have = false;
    int Qn = 0, Tn = 0, C0 = 0, nDB = 0, hi0 = 0;
    const RowInfo *rows = nullptr; DBlock *dblk = nullptr; const uint8_t *tcodes = nullptr, *tJobLo = nullptr, *tJobHi = nullptr;
    size_t trackOff = 0;                                     // IDS: track index of row q' is trackOff + q'
    uint32_t *arrowsJob = nullptr;
    if (have) {
      Qn = G->Qn; Tn = G->Tn; C0 = G->C0; nDB = G->nDB; hi0 = G->hi0;
      rows = B.rows + G->rowOff; dblk = B.dblk + G->dblkOff;
      tJobLo = B.tc + B.tOff[job]; tJobHi = B.tc + B.tOff[job + 1];
      tcodes = tJobLo + G->tStart - 1;                       // tcodes[t'] for t' in [1,Tn]
      if (FN == 2) trackOff = (size_t)B.qOff[job] + (size_t)G->qStart - 1;
      arrowsJob = reinterpret_cast<uint32_t *>(B.arrows + B.arrowOff[job]);
    }
    const int nDBw = __reduce_max_sync(0xffffffffu, nDB);
    const int nD = Qn + Tn + 1;

    // issue the cp.async copies of one block's rows and columns into buffer `buf`
    auto stage = [&](const int buf, const int b, const int wbase, const int k, const bool liveB, int &colShift) {
      const int cq = (C0 - wbase) >> 1;
      const int qlo = 32 * b + cq - (k * LPJ - 1), tlo = 32 * b - cq;
      for (int r = sl; r < k * LPJ + 32; r += LPJ) {
        const int qp = qlo + r;
        if (liveB && qp >= 0 && qp <= Qn) cp_async8(&sm.rows[buf][r], rows + qp);
        else sm.rows[buf][r] = make_int2(DEAD_CD8, 0);
      }
      // columns: bytes [tcodes + tlo, + k*LPJ + 33), copied as aligned 4-byte chunks clamped to the job's own bytes
      const uint8_t *p0 = tcodes + tlo;
      const int a = liveB ? (int)((uintptr_t)p0 & 3u) : 0;
      colShift = a;
      if (liveB) {
        const uint8_t *lo4 = (const uint8_t *)((uintptr_t)tJobLo & ~(uintptr_t)3);
        const uint8_t *hi4 = (const uint8_t *)(((uintptr_t)tJobHi + 3) & ~(uintptr_t)3);
        const int nch = (k * LPJ + 33 + a + 3) >> 2;
        for (int ch = sl; ch < nch; ch += LPJ) {
          const uint8_t *src = p0 - a + 4 * ch;
          if (src >= lo4 && src < hi4) cp_async4(&sm.colw[buf][ch], src);
          else sm.colw[buf][ch] = 0;
        }
      } else {
        for (int ch = sl; ch < (k * LPJ + 36) >> 2; ch += LPJ) sm.colw[buf][ch] = 0;
      }
      cp_async_commit();
    };

    int Se[KM], So[KM], AIe[KM], AIo[KM], ADe[KM], ADo[KM];
#pragma unroll(UG)
    for (int g = 0; g < KM; g++) { Se[g] = So[g] = AIe[g] = AIo[g] = ADe[g] = ADo[g] = BIG; }
    int wprev = 0, kprev = 0;
    uint32_t unit = 0;

    // block 0 is staged up front; the DBlock of block b+1 is always in registers one block ahead
    int wbase = 0, kown = 1;
    if (have && nDB > 0) { const DBlock db = dblk[0]; wbase = db.wbase; kown = db.k; }
    int k = __reduce_max_sync(0xffffffffu, kown);
    int colShift0 = 0, colShift1 = 0;
    __syncwarp();
    stage(0, 0, wbase, k, have && nDB > 0, colShift0);
    int wnext = wbase, knextOwn = 1;
    if (have && 1 < nDB) { const DBlock db = dblk[1]; wnext = db.wbase; knextOwn = db.k; }

    for (int b = 0; b < nDBw; b++) {
      const bool live = have && b < nDB;
      const int buf = b & 1;
      // ---- next block: its window is known, start its copies, fetch the DBlock after it
      const int knext = __reduce_max_sync(0xffffffffu, knextOwn);
      const bool liveN = have && b + 1 < nDB;
      if (b + 1 < nDBw) { if (buf) stage(0, b + 1, wnext, knext, liveN, colShift0); else stage(1, b + 1, wnext, knext, liveN, colShift1); }
      int wnext2 = wnext, knext2 = 1;
      if (have && b + 2 < nDB) { const DBlock db = dblk[b + 2]; wnext2 = db.wbase; knext2 = db.k; }
      // ---- this block's data has landed
      if (b + 1 < nDBw) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncwarp();
      // ---- re-map the register window (old: kprev groups from diagonal wprev; new: k groups from wbase)
      if (b > 0 && __any_sync(0xffffffffu, wbase != wprev || k != kprev)) {
        const int delta = wbase - wprev;
        auto remap = [&](int (&Xe)[KM], int (&Xo)[KM]) {
          __syncwarp();
#pragma unroll(UG)
          for (int g = 0; g < KM; g++)
            if (g < kprev) { sm.shift[2 * (kprev * sl + g)] = Xe[g]; sm.shift[2 * (kprev * sl + g) + 1] = Xo[g]; }
          __syncwarp();
#pragma unroll(UG)
          for (int g = 0; g < KM; g++) {
            if (g < k) {
              const int s = 2 * (k * sl + g) + delta;
              const bool ok = s >= 0 && s < 2 * kprev * LPJ;
              Xe[g] = ok ? sm.shift[s] : BIG;
              Xo[g] = ok ? sm.shift[s + 1] : BIG;
            }
          }
        };
        remap(Se, So);
        if (AFFINE) { remap(AIe, AIo); remap(ADe, ADo); }
      }
      wprev = wbase; kprev = k;
      const int cq = (C0 - wbase) >> 1;
      BlockView bv;
      bv.rows = sm.rows[buf];
      bv.cols = reinterpret_cast<const uint8_t *>(sm.colw[buf]) + (buf ? colShift1 : colShift0);
      bv.wbase8 = wbase << 8;
      bv.qlo = 32 * b + cq - (k * LPJ - 1); bv.tlo = 32 * b - cq;
      if (FN == 2) {
        for (int r = sl; r < k * LPJ + 32; r += LPJ) {
          const int qp = bv.qlo + r;
          int ra = 0, rb = 0;
          if (live && qp >= 1 && qp <= Qn) {
            const size_t x = trackOff + (size_t)qp;
            ra = (int)B.q[x] | ((int)B.subTag[x] << 8) | ((int)B.subQV[x] << 16) | ((int)B.insQV[x] << 24);
            if (B.delQV) rb = (int)B.delTag[x] | ((int)B.delQV[x] << 8) | (1 << 16);
          }
          sm.rowq[r] = ra; sm.rowq[Smem::NR + r] = rb;
        }
        __syncwarp();
      }
      uint32_t *aw = arrowsJob + (size_t)unit * UNITW;
      const bool first = live && (b << 6) <= hi0;          // row 0 holds cells on d <= hi0
      const bool last = live && (b == nDB - 1);
      const int eLast = last ? ((nD - 1) & 63) : 63;
      if (RING && !__any_sync(0xffffffffu, first || last))
        RingDispatch<LPJ, KM, AFFINE, FN, 1>::run(k, Se, So, AIe, AIo, ADe, ADo, bv, mtabAddr, c, sl, aw, live);
      else
        run_block_gen<LPJ, KM, AFFINE, FN>(Se, So, AIe, AIo, ADe, ADo, bv, sm.rowq, mtabAddr, c, k, sl, first, eLast, aw, live);
      if (live && sl == 0) { dblk[b].k = k; dblk[b].arrowUnit = unit; }
      unit += (uint32_t)k;
      // ---- the end cell (Qn, Tn) sits on diagonal Tn-Qn+C0 and is the last cell written to its slot
      if (__any_sync(0xffffffffu, last)) {
        const int s = Tn - Qn + C0 - wbase;
        const int sg = s >> 1, owner = last ? sg / k : 0, g = last ? sg % k : 0;
        int v = BIG;
        if (KM <= KRING) {
#pragma unroll
          for (int gg = 0; gg < KM; gg++) if (gg == g) v = (s & 1) ? So[gg] : Se[gg];
        } else v = (s & 1) ? So[g] : Se[g];
        v = __shfl_sync(0xffffffffu, v, owner, LPJ);
        if (last && sl == 0) G->score = v >> F::SHv;
      }
      __syncwarp();                                         // everyone is done with buffer `buf` before it is refilled
      wbase = wnext; k = knext; wnext = wnext2; knextOwn = knext2;
    }
    if (lane == 0 && B.cellSlots) atomicAdd(B.cellSlots, (unsigned long long)unit * 2048ull);   // 64 steps x 32 lanes x k cells
  }
}

template <int LPJ, int KM, bool AFFINE, int FN>
static void launch_one(const BatchDev &B, const ScoreParams &P, const uint32_t *order, const PlanHead *plan, int cls, uint32_t nGroups,
                       uint32_t *counter, int nSM, cudaStream_t s) {
  constexpr int WPC = KM <= KRING ? 4 : 1;          // warps per CTA
  const size_t smem = sizeof(SubSmem<LPJ, KM, FN>) * (32 / LPJ) * WPC;
  auto kern = fill_guided_kernel<LPJ, KM, AFFINE, FN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int perSM = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, WPC * 32, smem);
  if (perSM < 1) perSM = 1;
  unsigned grid = (unsigned)(nSM * perSM);
  const unsigned need = (nGroups + WPC - 1) / WPC;
  if (grid > need) grid = need;
  if (grid) kern<<<grid, WPC * 32, smem, s>>>(B, P, order, plan, cls, counter);
}

template <bool AFFINE, int FN>
static void launch_cls(const BatchDev &B, const ScoreParams &P, int cls, const uint32_t *order, const PlanHead *plan, uint32_t nGroups,
                       uint32_t *counter, int nSM, cudaStream_t s) {
  if (cls == CLS_L8N) launch_one<8, 4, AFFINE, FN>(B, P, order, plan, cls, nGroups, counter, nSM, s);
  else if (cls == CLS_L8) launch_one<8, KRING, AFFINE, FN>(B, P, order, plan, cls, nGroups, counter, nSM, s);
  else if (cls == CLS_L16) launch_one<16, KRING, AFFINE, FN>(B, P, order, plan, cls, nGroups, counter, nSM, s);
  else if (cls == CLS_L32) launch_one<32, KRING, AFFINE, FN>(B, P, order, plan, cls, nGroups, counter, nSM, s);
  else launch_one<32, KWIDE, AFFINE, FN>(B, P, order, plan, cls, nGroups, counter, nSM, s);
}

// cls: CLS_*; the class's warp groups (32 / cls_lpj(cls) job slots each) are plan->nGroups[cls] groups from
// order[plan->orderBegin[cls]]; nGroupsBound (an upper bound known to the host) only sizes the grid.
void launch_fill_guided(const BatchDev &B, const ScoreParams &P, int cls, const uint32_t *order, const PlanHead *plan,
                        uint32_t nGroupsBound, uint32_t *counter, int nSM, cudaStream_t s) {
  const bool aff = P.affine != 0;
  if (P.kind == BGPU_FN_IDS) {
    if (aff) launch_cls<true, 2>(B, P, cls, order, plan, nGroupsBound, counter, nSM, s);
    else launch_cls<false, 2>(B, P, cls, order, plan, nGroupsBound, counter, nSM, s);
  } else if (P.kind == BGPU_FN_QUALITY) {
    if (aff) launch_cls<true, 1>(B, P, cls, order, plan, nGroupsBound, counter, nSM, s);
    else launch_cls<false, 1>(B, P, cls, order, plan, nGroupsBound, counter, nSM, s);
  } else {
    if (aff) launch_cls<true, 0>(B, P, cls, order, plan, nGroupsBound, counter, nSM, s);
    else launch_cls<false, 0>(B, P, cls, order, plan, nGroupsBound, counter, nSM, s);
  }
}

}  // namespace bgpu

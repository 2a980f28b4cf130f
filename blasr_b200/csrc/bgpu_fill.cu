// bgpu_fill.cu -- the guided banded DP fill, one warp per job.
//
// Reference semantics restated (not translated):
//   GuidedAlign        common/algorithms/alignment/GuidedAlign.h:474-624   (linear gaps)
//   AffineGuidedAlign  common/algorithms/alignment/AffineGuidedAlign.h:241-375
//
// B200 mapping.  The band is swept by anti-diagonals d = q'+t'.  Lane j of the warp owns the two
// adjacent diagonals (slots) 2j and 2j+1 of a 64-diagonal register window (KMAX such windows for
// wide bands); on an even step it computes the cell on its even slot, on an odd step the one on its
// odd slot, so every lane has exactly one cell per step and all three DP neighbours are either its
// own registers or one __shfl away:
//     even step: left = lane j-1's odd slot (shfl), up = own odd slot, diag = own even slot
//     odd  step: left = own even slot, up = lane j+1's even slot (shfl), diag = own odd slot
// Scores live in registers as (score << 5) | tag: the five tie-ordered candidates carry their arrow
// code in the low bits, so one VIADDMNMX chain yields both the minimum and the reference's
// first-match-wins arrow (Diagonal > Left > Up > AffineInsClose > AffineDelClose); the affine
// open/extend decisions land in bits 3/4 the same way.  Cells outside the guide are held at BIG,
// which reproduces the reference's INF_INT-for-missing-neighbour rule.  Per d-block of 64 steps the
// warp stages the band table and target codes of the rows/columns it will touch into shared memory,
// and writes one traceback byte per cell as coalesced 128 B stores ([4 steps][32 lanes]).
#include "bgpu_common.cuh"

namespace bgpu {

struct FillConsts {
  int delT, insT;        // (del<<SH)|TB_LEFT, (ins<<SH)|TB_UP
  int extT3, extT4;      // (ext<<SH)|TB_ICLOSE, |TB_DCLOSE
  int ext, openI, openD; // ext<<SH, (open<<SH)|TB_IOPEN, (open<<SH)|TB_DOPEN
  int open;              // open<<SH
  int del0;              // row-0 step: (Global ? del : 0) << SH
};

template <bool AFFINE, bool QV, bool FIRST>
__device__ __forceinline__ uint32_t dp_cell(int &S, int &AI, int &AD, const int leftS, const int leftAD,
                                            const int upS, const int upAI, const int4 ri, const int tent,
                                            const int tprime, const uint32_t mtabAddr, const FillConsts &c) {
  const bool inb = (unsigned)(tprime - ri.x) <= (unsigned)ri.y;
  int m;
  asm("ld.shared.s32 %0, [%1];" : "=r"(m) : "r"(mtabAddr + (uint32_t)(ri.z + tent)));
  if (QV) m *= (FIRST ? (ri.w & 0xff) : ri.w);           // +-(1<<SH) * QV  (QualityValueScoreFunction.h:78-83)
  int cnd = S + m;                                         // Diagonal (tag 0)
  cnd = __viaddmin_s32(leftS, c.delT, cnd);                // Left
  cnd = __viaddmin_s32(upS, c.insT, cnd);                  // Up
  if (AFFINE) {
    cnd = __viaddmin_s32(upAI, c.extT3, cnd);              // AffineInsClose
    cnd = __viaddmin_s32(leftAD, c.extT4, cnd);            // AffineDelClose
  }
  int ai = 0, ad = 0;
  if (AFFINE) {
    // strict '<' in the reference: a tie extends (AffineGuidedAlign.h:357-373) -> the open
    // candidate carries a flag bit, so it loses ties.
    const int s0 = cnd & ~31;
    ai = __viaddmin_s32(s0, c.openI, upAI + c.ext);
    ad = __viaddmin_s32(s0, c.openD, leftAD + c.ext);
    cnd |= (ai | ad) & 24;                                  // tag | affine flags, still below bit 5
  }
  if (FIRST) {
    if (ri.w < 0) {                                        // boundary row (GuidedAlign.h:415-442)
      cnd = (tprime * c.del0) | (TB_LEFT | TB_IOPEN | TB_DOPEN); ai = c.open; ad = c.open;
    }
  }
  // cells outside the guide: BIG everywhere, arrow NoArrow
  cnd = inb ? cnd : (BIG | TB_NONE);
  S = cnd & ~31;
  if (AFFINE) { AI = inb ? (ai & ~31) : BIG; AD = inb ? (ad & ~31) : BIG; }
  return (uint32_t)cnd & 31u;
}

__device__ __forceinline__ int rot_up(int v, int lane) { return __shfl_sync(0xffffffffu, v, (lane + 31) & 31); }
__device__ __forceinline__ int rot_dn(int v, int lane) { return __shfl_sync(0xffffffffu, v, (lane + 1) & 31); }

template <int KMAX>
struct WarpSmem {
  int4 rows[32 * KMAX + 32];
  int tcol[32 * KMAX + 32];
  int shift[64 * KMAX];
};

// KACT > 0: exactly KACT groups are active (compile time, registers, no per-group branches);
// KACT == 0: the active count k is a run-time value (first/last blocks and the wide kernel).
template <int KMAX, int KACT, bool AFFINE, bool QV, bool FIRST, bool LAST>
__device__ __forceinline__ void run_block(int (&Se)[KMAX], int (&So)[KMAX], int (&AIe)[KMAX], int (&AIo)[KMAX],
                                          int (&ADe)[KMAX], int (&ADo)[KMAX], const WarpSmem<KMAX> &sm,
                                          const uint32_t mtabAddr, const FillConsts &c, const int k, const int lane,
                                          const int tlo, const int eLast, uint32_t *arrowWords) {
  constexpr int UG = KMAX <= 4 ? KMAX : 1;
  const int gEnd = KACT > 0 ? KACT : (KMAX <= 4 ? KMAX : k);
  const int kk = KACT > 0 ? KACT : k;
#define BGPU_ACTIVE(g) (KACT > 0 || (g) < k)
  // per-lane bases: row index (e>>1) + 32k-1-j-32g, column index ((e+1)>>1) + j + 32g
  const int4 *rp = sm.rows + (32 * kk - 1 - lane);
  const int *cp = sm.tcol + lane;
  int tcur = tlo + lane;
  uint32_t *aw = arrowWords + lane;
#pragma unroll 1
  for (int e4 = 0; e4 < 16; e4++, rp += 2, cp += 2, tcur += 2, aw += 32 * kk) {
    if (LAST && (e4 << 2) > eLast) break;
    uint32_t acc[KMAX];
#pragma unroll(UG)
    for (int g = 0; g < gEnd; g++) acc[g] = 0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (LAST && (e4 << 2) + u > eLast) break;
      constexpr int dummy = 0; (void)dummy;
      const int rI = u >> 1;            // (e>>1) - 2*e4
      const int cI = (u + 1) >> 1;      // ((e+1)>>1) - 2*e4
      int rS[KMAX], rA[KMAX];
      if ((u & 1) == 0) {
        // even step: left = odd slot of the lane below (ring over lanes and groups), up = own odd slot
#pragma unroll(UG)
        for (int g = 0; g < gEnd; g++)
          if (BGPU_ACTIVE(g)) { rS[g] = rot_up(So[g], lane); if (AFFINE) rA[g] = rot_up(ADo[g], lane); }
#pragma unroll(UG)
        for (int g = 0; g < gEnd; g++) {
          if (BGPU_ACTIVE(g)) {
            int leftS = rS[g], leftAD = AFFINE ? rA[g] : 0;
            if (KMAX > 1 && lane == 0) {
              const int pg = g > 0 ? g - 1 : KMAX - 1;     // the slot below slot 64g is the top slot of group g-1
              leftS = pg < kk ? rS[pg] : BIG;
              if (AFFINE) leftAD = pg < kk ? rA[pg] : BIG;
            }
            const int4 ri = rp[rI - 32 * g];
            const int tent = cp[cI + 32 * g];
            const uint32_t b = dp_cell<AFFINE, QV, FIRST>(Se[g], AIe[g], ADe[g], leftS, leftAD, So[g], AIo[g], ri,
                                                          tent, tcur + cI + 32 * g, mtabAddr, c);
            acc[g] += b << (8 * u);
          }
        }
      } else {
#pragma unroll(UG)
        for (int g = 0; g < gEnd; g++)
          if (BGPU_ACTIVE(g)) { rS[g] = rot_dn(Se[g], lane); if (AFFINE) rA[g] = rot_dn(AIe[g], lane); }
#pragma unroll(UG)
        for (int g = 0; g < gEnd; g++) {
          if (BGPU_ACTIVE(g)) {
            int upS = rS[g], upAI = AFFINE ? rA[g] : 0;
            if (KMAX > 1 && lane == 31) {
              const int ng = g + 1 < KMAX ? g + 1 : 0;
              upS = ng < kk ? rS[ng] : BIG;
              if (AFFINE) upAI = ng < kk ? rA[ng] : BIG;
            }
            const int4 ri = rp[rI - 32 * g];
            const int tent = cp[cI + 32 * g];
            const uint32_t b = dp_cell<AFFINE, QV, FIRST>(So[g], AIo[g], ADo[g], Se[g], ADe[g], upS, upAI, ri, tent,
                                                          tcur + cI + 32 * g, mtabAddr, c);
            acc[g] += b << (8 * u);
          }
        }
      }
    }
#pragma unroll(UG)
    for (int g = 0; g < gEnd; g++)
      if (BGPU_ACTIVE(g)) aw[g << 5] = acc[g];
  }
#undef BGPU_ACTIVE
}

template <int KMAX, bool AFFINE, bool QV>
__global__ void __launch_bounds__(KMAX <= 4 ? 128 : 32) fill_guided_kernel(BatchDev B, ScoreParams P, const uint32_t *order,
                                                          uint32_t nOrder, uint32_t *counter) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ int Mtab[25];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpSmem<KMAX> &sm = reinterpret_cast<WarpSmem<KMAX> *>(smemRaw)[warp];
  if (threadIdx.x < 25) {
    if (QV) { const int r = threadIdx.x / 5, cc = threadIdx.x % 5; Mtab[threadIdx.x] = ((r == cc && r < 4) ? -1 : 1) << SH; }  // ScoreMatrices.h:4-10
    else Mtab[threadIdx.x] = P.M[threadIdx.x] << SH;
  }
  __syncthreads();
  const uint32_t mtabAddr = (uint32_t)__cvta_generic_to_shared(Mtab);
  FillConsts c;
  c.delT = (P.del << SH) | TB_LEFT; c.insT = (P.ins << SH) | TB_UP;
  c.extT3 = (P.ext << SH) | TB_ICLOSE; c.extT4 = (P.ext << SH) | TB_DCLOSE;
  c.ext = P.ext << SH; c.open = P.open << SH;
  c.openI = (P.open << SH) | TB_IOPEN; c.openD = (P.open << SH) | TB_DOPEN;
  c.del0 = (P.alignType == BGPU_GLOBAL ? P.del : 0) << SH;

  for (;;) {
    uint32_t idx = 0;
    if (lane == 0) idx = atomicAdd(counter, 1u);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= nOrder) break;
    const uint32_t job = order[idx];
    JobGeom &G = B.geom[job];
    if (G.status != BGPU_JOB_OK) continue;
    const int Qn = G.Qn, Tn = G.Tn, C0 = G.C0, nDB = G.nDB, hi0 = G.hi0, tStart = G.tStart;
    const RowInfo *rows = B.rows + G.rowOff;
    const DBlock *dblk = B.dblk + G.dblkOff;
    const uint8_t *tcodes = B.t + B.tOff[job] + tStart - 1;   // tcodes[t'] for t' in [1,Tn]
    uint32_t *arrowsJob = reinterpret_cast<uint32_t *>(B.arrows + B.arrowOff[job]);
    const int nD = Qn + Tn + 1;

    constexpr int UG = KMAX <= 4 ? KMAX : 1;
    int Se[KMAX], So[KMAX], AIe[KMAX], AIo[KMAX], ADe[KMAX], ADo[KMAX];
#pragma unroll(UG)
    for (int g = 0; g < KMAX; g++) { Se[g] = So[g] = AIe[g] = AIo[g] = ADe[g] = ADo[g] = BIG; }
    int wprev = 0, kprev = 0;

    for (int b = 0; b < nDB; b++) {
      const DBlock db = dblk[b];
      const int wbase = db.wbase, k = db.k;
      // ---- slide the register window to this block's diagonals
      if (b > 0 && (wbase != wprev || k != kprev)) {
        const int delta = wbase - wprev;
        // invariant: groups >= kprev hold BIG.  Old slot s' = s + delta feeds new slot s.
        const int gW = KMAX <= 4 ? KMAX : kprev, gR = KMAX <= 4 ? KMAX : max(k, kprev);
        auto slide = [&](int (&Xe)[KMAX], int (&Xo)[KMAX]) {
          __syncwarp();
#pragma unroll(UG)
          for (int g = 0; g < gW; g++) { sm.shift[64 * g + 2 * lane] = Xe[g]; sm.shift[64 * g + 2 * lane + 1] = Xo[g]; }
          __syncwarp();
#pragma unroll(UG)
          for (int g = 0; g < gR; g++) {
            const int s = 64 * g + 2 * lane + delta;
            const bool ok = (g < k) && s >= 0 && s < 64 * (KMAX <= 4 ? KMAX : kprev);
            Xe[g] = ok ? sm.shift[s] : BIG;
            Xo[g] = ok ? sm.shift[s + 1] : BIG;
          }
        };
        slide(Se, So);
        if (AFFINE) { slide(AIe, AIo); slide(ADe, ADo); }
      }
      wprev = wbase; kprev = k;
      // ---- stage rows [qlo, qlo+32k+31) and columns [tlo, tlo+32k+32)
      const int cq = (C0 - wbase) >> 1;
      const int qlo = 32 * b + cq - 32 * k + 1, tlo = 32 * b - cq;
      __syncwarp();
      for (int r = lane; r < 32 * k + 31; r += 32) {
        const int qp = qlo + r;
        int4 v = make_int4(INT_MAX / 2, 0, 0, 0);
        if (qp >= 0 && qp <= Qn) {
          const RowInfo ri = rows[qp];
          v.x = ri.lo; v.y = (int)(ri.packed & ((1u << ROW_W_BITS) - 1));
          v.z = (int)((ri.packed >> 20) & 7u) * 20;
          v.w = (int)((ri.packed >> 23) & 0xffu) | (qp == 0 ? (int)0x80000000 : 0);
        }
        sm.rows[r] = v;
      }
      for (int cI = lane; cI < 32 * k + 32; cI += 32) {
        const int tp = tlo + cI;
        sm.tcol[cI] = (tp >= 1 && tp <= Tn) ? (int)tcodes[tp] * 4 : 0;
      }
      __syncwarp();
      uint32_t *aw = arrowsJob + (size_t)db.arrowUnit * 512u;
      const bool first = (b << 6) <= hi0;               // row 0 holds cells on d <= hi0
      const bool last = (b == nDB - 1);
      const int eLast = (nD - 1) & 63;
      if (first && last) run_block<KMAX, 0, AFFINE, QV, true, true>(Se, So, AIe, AIo, ADe, ADo, sm, mtabAddr, c, k, lane, tlo, eLast, aw);
      else if (first) run_block<KMAX, 0, AFFINE, QV, true, false>(Se, So, AIe, AIo, ADe, ADo, sm, mtabAddr, c, k, lane, tlo, eLast, aw);
      else if (last) run_block<KMAX, 0, AFFINE, QV, false, true>(Se, So, AIe, AIo, ADe, ADo, sm, mtabAddr, c, k, lane, tlo, eLast, aw);
      else if (KMAX > 4) run_block<KMAX, 0, AFFINE, QV, false, false>(Se, So, AIe, AIo, ADe, ADo, sm, mtabAddr, c, k, lane, tlo, eLast, aw);
      else if (k == 1) run_block<KMAX, 1, AFFINE, QV, false, false>(Se, So, AIe, AIo, ADe, ADo, sm, mtabAddr, c, k, lane, tlo, eLast, aw);
      else if (KMAX >= 2 && k == 2) run_block<KMAX, (KMAX >= 2 ? 2 : 1), AFFINE, QV, false, false>(Se, So, AIe, AIo, ADe, ADo, sm, mtabAddr, c, k, lane, tlo, eLast, aw);
      else if (KMAX >= 4 && k == 3) run_block<KMAX, (KMAX >= 4 ? 3 : 1), AFFINE, QV, false, false>(Se, So, AIe, AIo, ADe, ADo, sm, mtabAddr, c, k, lane, tlo, eLast, aw);
      else run_block<KMAX, (KMAX >= 4 ? 4 : 1), AFFINE, QV, false, false>(Se, So, AIe, AIo, ADe, ADo, sm, mtabAddr, c, k, lane, tlo, eLast, aw);
    }
    // ---- the end cell (Qn, Tn) sits on diagonal Tn-Qn+C0 and is the last cell written to its slot
    {
      const int s = Tn - Qn + C0 - wprev;
      int v = BIG;
#pragma unroll(UG)
      for (int g = 0; g < KMAX; g++) if ((s >> 6) == g) v = (s & 1) ? So[g] : Se[g];
      v = __shfl_sync(0xffffffffu, v, (s & 63) >> 1);
      if (lane == 0) G.score = v >> SH;
    }
  }
}

template <int KMAX, bool AFFINE, bool QV>
static void launch_one(const BatchDev &B, const ScoreParams &P, const uint32_t *order, uint32_t nOrder,
                       uint32_t *counter, int nSM, cudaStream_t s) {
  constexpr int WPC = KMAX <= 4 ? 4 : 1;           // warps per CTA
  const size_t smem = sizeof(WarpSmem<KMAX>) * WPC;
  auto kern = fill_guided_kernel<KMAX, AFFINE, QV>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int perSM = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, WPC * 32, smem);
  if (perSM < 1) perSM = 1;
  unsigned grid = (unsigned)(nSM * perSM);
  const unsigned need = (nOrder + WPC - 1) / WPC;
  if (grid > need) grid = need;
  if (grid) kern<<<grid, WPC * 32, smem, s>>>(B, P, order, nOrder, counter);
}

// kclass: 1 -> KMAX=1, 2 -> KMAX=2, 4 -> KMAX=4, anything larger -> the wide kernel (KMAX_BUILD groups)
void launch_fill_guided(const BatchDev &B, const ScoreParams &P, int kclass, const uint32_t *order, uint32_t nOrder,
                        uint32_t *counter, int nSM, cudaStream_t s) {
  const bool aff = P.affine != 0, qv = P.kind == BGPU_FN_QUALITY;
#define BGPU_DISPATCH(K)                                                                   \
  do {                                                                                     \
    if (aff && qv) launch_one<K, true, true>(B, P, order, nOrder, counter, nSM, s);        \
    else if (aff) launch_one<K, true, false>(B, P, order, nOrder, counter, nSM, s);        \
    else if (qv) launch_one<K, false, true>(B, P, order, nOrder, counter, nSM, s);         \
    else launch_one<K, false, false>(B, P, order, nOrder, counter, nSM, s);                \
  } while (0)
  if (kclass == 1) BGPU_DISPATCH(1);
  else if (kclass == 2) BGPU_DISPATCH(2);
  else if (kclass == 4) BGPU_DISPATCH(4);
  else BGPU_DISPATCH(KMAX_BUILD);
#undef BGPU_DISPATCH
}

}  // namespace bgpu

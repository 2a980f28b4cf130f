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
// in register rings refilled with one shared-memory load per row / column.  The loop body is a chunk of 8 (linear) / 4
// (affine) step pairs unrolled with every ring index static; between chunks the rings are rotated by (chunk mod k)
// registers, so traceback words complete at fixed places of the code for every k.
// Instruction budget.  The integer ALU pipe issues a warp instruction every second clock (VIADDMNMX, LOP3, ISETP, SEL,
// SHF all live there), the FMA pipe (IMAD) is otherwise idle and shared memory delivers one 32-lane load per clock and
// SM -- so the loop keeps on the ALU pipe only what has no other home: per cell the two VIADDMNMX of the min chain, one
// ISETP for the in-band test ((unsigned)(lane's diagonal - row's first diagonal) <= row width), the tag split and the
// funnel shift that appends the arrow to the lane's traceback word.  The diagonal offset of the test and the
// out-of-band overwrite are written as IMADs with a run-time 1 (FillConsts::one) so that they issue on the FMA pipe
// (the overwrite predicated on the test), S + m likewise; the substitution score is ONE LDS per cell (an in-band mask
// table in shared memory was measured: 2.4 shared-memory wavefronts per cell make the kernel LSU-bound and slower).
// Words are stored [d-block][16-step row][slot pair], 2 bits per cell for the linear aligner (8 for the affine one).
// Blocks that touch the boundary row, a job's last block, the IDS score function and very wide windows take the
// generic path (run_block_gen), which reads its rows and columns from shared memory per cell.
#include <type_traits>
#include "bgpu_common.cuh"

namespace bgpu {

constexpr uint32_t NOJOB = 0xffffffffu;

struct FillConsts {
  int delT, insT;        // (del<<SH)|LEFT, (ins<<SH)|UP
  int extT3, extT4;      // (ext<<SH)|TB_ICLOSE, |TB_DCLOSE
  int ext, openI, openD; // ext<<SH, (open<<SH)|TB_IOPEN, (open<<SH)|TB_DOPEN
  int open;              // open<<SH
  int del0;              // row-0 step: (Global ? del : 0) << SH
  int one, zero;         // 1 and 0 held in registers the compiler cannot fold (see "Instruction budget")
  int subPrior, delPrior, del;   // IDSScoreFunction: unshifted substitutionPrior / globalDeletionPrior / del
};

template <int LPJ, int KM, int FN>
struct alignas(16) SubSmem {   // staging of one job: two d-blocks (current, next)
  static constexpr int NR = KM * LPJ + 32;            // rows a d-block can touch
  static constexpr int NRS = (NR + 3) & ~1;           // ... staged from an even (16-byte aligned) row index
  static constexpr int NBW = (NR + 1 + 15 + 15) / 16 * 4;   // words of a byte window staged from a 16-byte aligned address
  int2 rows[2][NRS];                                  // RowInfo as prep wrote it: {lo8, nhi8}
  uint32_t colw[2][NBW];                              // target codes (bytes)
  uint32_t qcw[2][NBW];                               // query row offsets (code * 20, bytes)
  uint32_t qvw[FN == 1 ? 2 : 1][FN == 1 ? NBW : 4];   // QualityValueScoreFunction: the rows' QVs
  unsigned long long mbar[2];                         // one mbarrier per buffer: the bulk copies of a block complete on it
  int shift[2 * KM * LPJ];                            // window re-mapping scratch
  // IDSScoreFunction: two words per row of the current block,
  //   [r]      query byte | substitutionTag << 8 | substitutionQV << 16 | insertionQV << 24
  //   [NR + r] deletionTag | deletionQV << 8 | (deletion tracks present) << 16
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
// TMA (bulk copy) + mbarrier: one lane arms the barrier with the byte count and issues the copies, everyone waits on the phase
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}"
               ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// In-band test and out-of-band overwrite of the ring kernels.  x = rX + IMM is the cell's diagonal minus the row's first
// diagonal (x 8), rY the row's width: outside <=> (unsigned)x > rY.  The add and the overwrites are IMADs by a run-time 1,
// i.e. FMA-pipe instructions; only the compare issues on the ALU pipe.
template <int IMM>
__device__ __forceinline__ void mask_lin(int &cnd, const int rX, const int rY, const FillConsts &c) {
  // the overwrite is x * 0 + (BIG | NoArrow) with a run-time 0: depending on x keeps it a predicated IMAD of its own
  asm("{\n\t.reg .pred p;\n\t.reg .s32 x;\n\tmad.lo.s32 x, %1, %3, %5;\n\tsetp.gt.u32 p, x, %2;\n\t@p mad.lo.s32 %0, x, %4, %6;\n\t}"
      : "+r"(cnd) : "r"(rX), "r"(rY), "r"(c.one), "r"(c.zero), "n"(IMM), "n"(BIG | TL_NONE));
}
// (three different stand-ins for "invalid" -- BIG, BIG + 32, BIG + 64, any value >= BIG with clear tag bits does -- so
// that the three overwrites stay three predicated IMADs instead of being merged into one IMAD and SELs)
template <int IMM>
__device__ __forceinline__ void mask_aff(int &cnd, int &s0, int &ai, int &ad, const int rX, const int rY, const FillConsts &c) {
  asm("{\n\t.reg .pred p;\n\t.reg .s32 x;\n\tmad.lo.s32 x, %4, %6, %8;\n\tsetp.gt.u32 p, x, %5;\n\t"
      "@p mad.lo.s32 %0, x, %7, %9;\n\t@p mad.lo.s32 %1, x, %7, %10;\n\t@p mad.lo.s32 %2, x, %7, %11;\n\t@p mad.lo.s32 %3, x, %7, %12;\n\t}"
      : "+r"(cnd), "+r"(s0), "+r"(ai), "+r"(ad) : "r"(rX), "r"(rY), "r"(c.one), "r"(c.zero), "n"(IMM),
        "n"(BIG | TB_NONE), "n"(BIG), "n"(BIG + 32), "n"(BIG + 64));
}

// The min chain of one DP cell.  Returns the arrow word (tag | affine flags in its low bits); s0 / ai / ad = the
// scores with the tag bits cleared.  The caller masks out-of-band cells.
template <bool AFFINE>
__device__ __forceinline__ int dp_core(const int S, const int leftS, const int leftAD, const int upS, const int upAI,
                                       const int m, const int delT, const int insT, const FillConsts &c,
                                       int &s0, int &ai, int &ad) {
  int cnd = S + m;                                         // Diagonal (tag 0)
  cnd = __viaddmin_s32(leftS, delT, cnd);                  // Left
  cnd = __viaddmin_s32(upS, insT, cnd);                    // Up
  if (AFFINE) {
    cnd = __viaddmin_s32(upAI, c.extT3, cnd);              // AffineInsClose
    cnd = __viaddmin_s32(leftAD, c.extT4, cnd);            // AffineDelClose
    // strict '<' in the reference: a tie extends (AffineGuidedAlign.h:357-373) -> the open candidate carries a
    // flag bit, so it loses ties.
    s0 = cnd & ~31;
    ai = __viaddmin_s32(s0, c.openI, upAI + c.ext);
    ad = __viaddmin_s32(s0, c.openD, leftAD + c.ext);
    cnd = cnd | ai | ad;                                    // tag | the two open flags (ai: 0 / 8, ad: 0 / 16 below bit 5);
                                                            // the traceback reads bits 0-4 of the byte only
    ai &= ~31; ad &= ~31;
  } else {
    s0 = cnd & ~3;
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
  const uint8_t *qrows;   // staged query row offsets (code * 20), qrows[r]
  const uint8_t *qvs;     // staged QVs (QualityValueScoreFunction), qvs[r]
  int wbase;              // window base diagonal
  int qlo, tlo;
};

// compile-time loop: f(integral_constant<int, I>) for I = 0 .. N-1 (the asm immediates of the ring kernels need constant
// expressions, which a #pragma-unrolled loop variable is not)
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F &&f) {
  if constexpr (I < N) { f(std::integral_constant<int, I>{}); static_for<I + 1, N>(f); }
}

template <int KA, int ROT, typename T>
__device__ __forceinline__ void ring_rotate(T (&a)[KA]) {   // a[s] <- a[(s + ROT) % KA]
  if (ROT != 0) {
    T tmp[KA];
#pragma unroll
    for (int s = 0; s < KA; s++) tmp[s] = a[(s + ROT) % KA];
#pragma unroll
    for (int s = 0; s < KA; s++) a[s] = tmp[s];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fast path: KA groups per lane (compile time), rows / columns in register rings, no boundary row, full 64 steps.
template <int LPJ, int KM, int KA, bool AFFINE, int FN>
__device__ __forceinline__ void run_block_ring(int (&Se)[KM], int (&So)[KM], int (&AIe)[KM], int (&AIo)[KM],
                                               int (&ADe)[KM], int (&ADo)[KM], const BlockView &bv,
                                               const int mtabAddr, const FillConsts &c, const int sl,
                                               uint32_t *aw, const bool live) {
  typedef Fmt<AFFINE> F;
  constexpr int PPW = F::SPW / 2;                            // step pairs per traceback word
  // step pairs per unrolled chunk: the chunk is the loop body, and five warps per scheduler at different places of a
  // body that outgrows the instruction caches (L0 ~6 KB, L1.5 32 KB, ~190 B of code per linear cell) stall on fetches
  constexpr int CH = (AFFINE ? 2 : 4) * (KA <= 2 ? 2 : 1);
  constexpr int ROT = CH % KA;
  const int kL = KA * sl;
  const int rbase = KA * (LPJ - 1 - sl);
  const int2 *rp = bv.rows + rbase;                          // rp[i + m]: row of group KA-1-m at step pair i
  const uint8_t *qp = bv.qrows + rbase;
  const uint8_t *vp = bv.qvs + rbase;
  const uint8_t *cp = bv.cols + kL;                          // cp[i + g]: column of group g on the even step of pair i
  const int lb8 = (bv.wbase + 2 * kL) << 3;                 // 8 * (diagonal of the lane's slot 0)
  int rQ[KA], rX[KA], rY[KA], cT[KA];
  int rV[FN == 1 ? KA : 1];                                  // QualityValueScoreFunction: the rows' QVs
  uint32_t acc[KA];
  auto load_row = [&](const int slot, const int ridx) {
    const int2 v = rp[ridx];
    rX[slot] = lb8 - v.x;                                    // 8 * (slot 0's diagonal - the row's first diagonal)
    rY[slot] = -v.y - v.x;                                   // 8 * (cells in the row - 1)
    rQ[slot] = qp[ridx];
    if (FN == 1) rV[slot] = vp[ridx];
  };
#pragma unroll
  for (int m = 0; m < KA; m++) {
    load_row(m, m);
    cT[m] = (int)cp[m] * 4 + mtabAddr;
    acc[m] = 0;
  }
  aw += kL;
#pragma unroll 1
  for (int i0 = 0; i0 < 32; i0 += CH) {
    static_for<0, CH>([&](auto JJ) {
      constexpr int jj = decltype(JJ)::value, j = jj % KA;
      // ---- even step: cells on slots 2g
      {
        const int left0 = sub_up<LPJ>(So[KA - 1]);
        const int leftA0 = AFFINE ? sub_up<LPJ>(ADo[KA - 1]) : 0;
        static_for<0, KA>([&](auto GG) {
          constexpr int g = decltype(GG)::value, p = (jj + KA - 1 - g) % KA, cs = (jj + g) % KA;
          const int leftS = g == 0 ? left0 : So[g == 0 ? 0 : g - 1];
          const int leftAD = AFFINE ? (g == 0 ? leftA0 : ADo[g == 0 ? 0 : g - 1]) : 0;
          int m = lds32((uint32_t)(rQ[p] + cT[cs]));
          if (FN == 1) m *= rV[FN == 1 ? p : 0];            // +-(1<<SH) * QV  (QualityValueScoreFunction.h:78-83)
          int s0, ai = 0, ad = 0;
          int cnd = dp_core<AFFINE>(Se[g], leftS, leftAD, So[g], AFFINE ? AIo[g] : 0, m, c.delT, c.insT, c, s0, ai, ad);
          if (AFFINE) mask_aff<16 * g>(cnd, s0, ai, ad, rX[p], rY[p], c);
          else { mask_lin<16 * g>(cnd, rX[p], rY[p], c); s0 = cnd & ~3; }
          Se[g] = s0;
          if (AFFINE) { AIe[g] = ai; ADe[g] = ad; }
          acc[g] = __funnelshift_r(acc[g], (uint32_t)cnd, F::BITS);   // first step of a word ends up in its lowest field
        });
      }
      cT[j] = (int)cp[i0 + jj + KA] * 4 + mtabAddr;         // column kL + i + KA replaces kL + i
      // ---- odd step: cells on slots 2g+1
      {
        const int up0 = sub_dn<LPJ>(Se[0]);
        const int upA0 = AFFINE ? sub_dn<LPJ>(AIe[0]) : 0;
        static_for<0, KA>([&](auto GG) {
          constexpr int g = decltype(GG)::value, p = (jj + KA - 1 - g) % KA, cs = (jj + 1 + g) % KA;
          const int upS = g == KA - 1 ? up0 : Se[g == KA - 1 ? 0 : g + 1];
          const int upAI = AFFINE ? (g == KA - 1 ? upA0 : AIe[g == KA - 1 ? 0 : g + 1]) : 0;
          int m = lds32((uint32_t)(rQ[p] + cT[cs]));
          if (FN == 1) m *= rV[FN == 1 ? p : 0];
          int s0, ai = 0, ad = 0;
          int cnd = dp_core<AFFINE>(So[g], Se[g], AFFINE ? ADe[g] : 0, upS, upAI, m, c.delT, c.insT, c, s0, ai, ad);
          if (AFFINE) mask_aff<16 * g + 8>(cnd, s0, ai, ad, rX[p], rY[p], c);
          else { mask_lin<16 * g + 8>(cnd, rX[p], rY[p], c); s0 = cnd & ~3; }
          So[g] = s0;
          if (AFFINE) { AIo[g] = ai; ADo[g] = ad; }
          acc[g] = __funnelshift_r(acc[g], (uint32_t)cnd, F::BITS);
        });
      }
      load_row(j, i0 + jj + KA);                              // row baseR + i + KA replaces row baseR + i
      if (CH >= PPW && (jj % PPW) == PPW - 1) {               // a traceback word is complete
        if (live) {
#pragma unroll
          for (int g = 0; g < KA; g++) aw[g] = acc[g];
        }
        aw += KA * LPJ;
      }
    });
    if (CH < PPW && ((i0 + CH) & (PPW - 1)) == 0) {           // (linear, k >= 3: a word is two chunks)
      if (live) {
#pragma unroll
        for (int g = 0; g < KA; g++) aw[g] = acc[g];
      }
      aw += KA * LPJ;
    }
    ring_rotate<KA, ROT>(rQ); ring_rotate<KA, ROT>(rX); ring_rotate<KA, ROT>(rY); ring_rotate<KA, ROT>(cT);
    if (FN == 1) ring_rotate<(FN == 1 ? KA : 1), (FN == 1 ? ROT : 0)>(rV);
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
      m = lds32((uint32_t)(mtabAddr + (int)bv.qrows[ridx] + (int)bv.cols[cidx] * 4));
      if (FN == 1) m *= (int)bv.qvs[ridx];                  // +-(1<<SH) * QV  (QualityValueScoreFunction.h:78-83)
    }
    int s0, ai = 0, ad = 0;
    int cnd = dp_core<AFFINE>(S, leftS, leftAD, upS, upAI, m, delT, insT, c, s0, ai, ad);
    if (first && bv.qlo + ridx == 0) {                      // boundary row (GuidedAlign.h:415-442)
      cnd = ((bv.tlo + cidx) * c.del0) | (AFFINE ? (TB_LEFT | TB_IOPEN | TB_DOPEN) : TL_LEFT);
      s0 = cnd & ~F::TAGMASK; ai = c.open; ad = c.open;
    }
    const int d8 = (bv.wbase + slot) << 3;
    const bool inb = d8 >= rv.x && -d8 >= rv.y;
    cnd = inb ? cnd : (BIG | F::NONE);
    if (e <= eLast) {
      S = inb ? s0 : BIG;
      if (AFFINE) { AI = inb ? ai : BIG; AD = inb ? ad : BIG; }
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
__global__ void __launch_bounds__(KM <= KRING ? 128 : 32, KM <= KRING ? ((AFFINE || FN == 1) ? 4 : 5) : 1)
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
  if (RING && sl == 0) { mbar_init(&sm.mbar[0], 1); mbar_init(&sm.mbar[1], 1); }
  if (RING) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int mtabAddr = (int)__cvta_generic_to_shared(Mtab);
  FillConsts c;
  c.delT = (P.del << F::SHv) | TB_LEFT; c.insT = (P.ins << F::SHv) | TB_UP;
  c.extT3 = (P.ext << F::SHv) | TB_ICLOSE; c.extT4 = (P.ext << F::SHv) | TB_DCLOSE;
  c.ext = P.ext << F::SHv; c.open = P.open << F::SHv;
  c.openI = (P.open << F::SHv) | TB_IOPEN; c.openD = (P.open << F::SHv) | TB_DOPEN;
  c.del0 = (P.alignType == BGPU_GLOBAL ? P.del : 0) << F::SHv;
  c.one = 1 + P.pad; c.zero = P.pad;   // P.pad is always 0
  c.subPrior = P.subPrior; c.delPrior = P.delPrior; c.del = P.del;

  uint32_t mbarPhase = 0;                                    // bit b: parity the next wait on this job slot's mbar[b] looks for
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
    const RowInfo *rows = nullptr; DBlock *dblk = nullptr;
    const uint8_t *tcodes = nullptr, *qcodes = nullptr, *qvals = nullptr;
    int tLoOff = 0, tHiOff = 0, qLoOff = 0, qHiOff = 0;      // the job's own bytes: [tcodes + tLoOff, tcodes + tHiOff) etc.
    size_t trackOff = 0;                                     // IDS: track index of row q' is trackOff + q'
    uint32_t *arrowsJob = nullptr;
    if (have) {
      Qn = G->Qn; Tn = G->Tn; C0 = G->C0; nDB = G->nDB; hi0 = G->hi0;
      rows = B.rows + G->rowOff; dblk = B.dblk + G->dblkOff;
      const uint64_t t0 = B.tOff[job], q0 = B.qOff[job];
      tcodes = B.tc + t0 + G->tStart - 1;                    // tcodes[t'] for t' in [1,Tn]
      tLoOff = 1 - G->tStart; tHiOff = tLoOff + (int)(B.tOff[job + 1] - t0);
      qcodes = B.qc + q0 + G->qStart - 1;                    // qcodes[q'] for q' in [1,Qn]
      qLoOff = 1 - G->qStart; qHiOff = qLoOff + (int)(B.qOff[job + 1] - q0);
      if (FN == 1) qvals = B.qual + q0 + G->qStart - 1;
      if (FN == 2) trackOff = (size_t)B.qOff[job] + (size_t)G->qStart - 1;
      arrowsJob = reinterpret_cast<uint32_t *>(B.arrows + B.arrowOff[job]);
    }
    const int nDBw = __reduce_max_sync(0xffffffffu, nDB);
    const int nD = Qn + Tn + 1;

    // bytes [p0, p0 + n) of a per-base array, copied as aligned 4-byte chunks clamped to the job's own bytes [lo, hi);
    // returns the offset of p0 inside the first chunk
    auto stage_bytes = [&](uint32_t *dstw, const uint8_t *base, const int off, const int lo, const int hi, const int n, const bool liveB) {
      const int a = liveB ? (int)((uintptr_t)(base + off) & 3u) : 0;
      const int nch = (n + a + 3) >> 2;
      if (liveB) {
        // chunk ch covers bytes [off - a + 4 ch, + 4): whole aligned words that overlap [lo, hi) (the arrays are padded)
        for (int ch = sl; ch < nch; ch += LPJ) {
          const int o = off - a + 4 * ch;
          if (o + 3 >= lo && o < hi) cp_async4(&dstw[ch], base + o);
          else dstw[ch] = 0;
        }
      } else {
        for (int ch = sl; ch < nch; ch += LPJ) dstw[ch] = 0;
      }
      return a;
    };
    // Stage one block's rows, target codes and query codes into buffer `buf`.  Ring kernels: lane 0 of the job arms the
    // buffer's mbarrier and issues one bulk copy (TMA) per array, from 16-byte aligned addresses, no bounds checks (the
    // arrays are padded, see ROW_PAD / BYTE_PAD); desc = row offset | column offset << 4 | query offset << 8 of the block's first
    // element inside the staged data | (the block has to be waited for on the mbarrier) << 12 | (blanked) << 13.  Otherwise (wide kernel, IDS, a window that
    // would leave the pads): per-lane cp.async copies with the out-of-range elements filled in.
    auto stage = [&](const int buf, const int b, const int wbase, const int k, const bool liveB, int &desc) {
      const int cq = (C0 - wbase) >> 1;
      const int qlo = 32 * b + cq - (k * LPJ - 1), tlo = 32 * b - cq;
      const int nr = k * LPJ + 32;
      if (RING) {
        // a finished / absent job keeps computing (nothing is stored): its buffers only have to hold aligned table offsets,
        // so they are blanked once (bit 13) and then left alone
        if (!liveB) {
          if (desc & (1 << 13)) return;
          for (int r = sl; r < Smem::NRS; r += LPJ) sm.rows[buf][r] = make_int2(DEAD_LO8, -DEAD_LO8);   // the whole buffer: k may grow
          for (int w = sl; w < Smem::NBW; w += LPJ) { sm.colw[buf][w] = 0; sm.qcw[buf][w] = 0; if (FN == 1) sm.qvw[FN == 1 ? buf : 0][w] = 0; }
          desc = 1 << 13;
          return;
        }                                  // a finished / absent job computes on stale data, nothing is stored
        const bool inPads = qlo >= -ROW_PAD + 1 && qlo + nr + 2 <= Qn + ROW_PAD &&
                            tlo - 16 >= tLoOff - BYTE_PAD && tlo + nr + 1 + 32 <= tHiOff + BYTE_PAD &&
                            qlo - 16 >= qLoOff - BYTE_PAD && qlo + nr + 32 <= qHiOff + BYTE_PAD;
        if (liveB && inPads) {
          const RowInfo *rsrc = rows + qlo;
          const int ra = (int)(((uintptr_t)rsrc >> 3) & 1u);  // rows: 8 B each, the copy starts at an even one
          const uint8_t *csrc = tcodes + tlo, *qsrc = qcodes + qlo;
          const int ca = (int)((uintptr_t)csrc & 15u), qa = (int)((uintptr_t)qsrc & 15u);
          desc = ra | (ca << 4) | (qa << 8) | (1 << 12);
          if (sl == 0) {
            const uint32_t rb = (uint32_t)((nr + ra + 1) & ~1) * 8u, cb = (uint32_t)(nr + 1 + ca + 15) & ~15u, qb = (uint32_t)(nr + qa + 15) & ~15u;
            mbar_expect_tx(&sm.mbar[buf], rb + cb + qb + (FN == 1 ? qb : 0u));
            bulk_g2s(sm.rows[buf], rsrc - ra, rb, &sm.mbar[buf]);
            bulk_g2s(sm.colw[buf], csrc - ca, cb, &sm.mbar[buf]);
            bulk_g2s(sm.qcw[buf], qsrc - qa, qb, &sm.mbar[buf]);
            if (FN == 1) bulk_g2s(sm.qvw[buf], qvals + qlo - qa, qb, &sm.mbar[buf]);   // q and qual share their offsets
          }
          return;
        }
      }
      for (int r = sl; r < nr; r += LPJ) {
        const int qp = qlo + r;
        if (liveB && qp >= 0 && qp <= Qn) cp_async8(&sm.rows[buf][r], rows + qp);
        else sm.rows[buf][r] = make_int2(DEAD_LO8, -DEAD_LO8);
      }
      desc = (stage_bytes(sm.colw[buf], tcodes, tlo, tLoOff, tHiOff, nr + 1, liveB) << 4) | (liveB ? 0 : 1 << 13);
      if (FN != 2) desc |= stage_bytes(sm.qcw[buf], qcodes, qlo, qLoOff, qHiOff, nr, liveB) << 8;
      if (FN == 1) stage_bytes(sm.qvw[buf], qvals, qlo, qLoOff, qHiOff, nr, liveB);
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
    int desc0 = 0, desc1 = 0;
    __syncwarp();
    stage(0, 0, wbase, k, have && nDB > 0, desc0);
    int wnext = wbase, knextOwn = 1;
    if (have && 1 < nDB) { const DBlock db = dblk[1]; wnext = db.wbase; knextOwn = db.k; }

    for (int b = 0; b < nDBw; b++) {
      const bool live = have && b < nDB;
      const int buf = b & 1;
      // ---- next block: its window is known, start its copies, fetch the DBlock after it
      const int knext = __reduce_max_sync(0xffffffffu, knextOwn);
      const bool liveN = have && b + 1 < nDB;
      if (b + 1 < nDBw) { if (buf) stage(0, b + 1, wnext, knext, liveN, desc0); else stage(1, b + 1, wnext, knext, liveN, desc1); }
      int wnext2 = wnext, knext2 = 1;
      if (have && b + 2 < nDB) { const DBlock db = dblk[b + 2]; wnext2 = db.wbase; knext2 = db.k; }
      // ---- this block's data has landed
      if (b + 1 < nDBw) cp_async_wait<1>(); else cp_async_wait<0>();
      const int desc = buf ? desc1 : desc0;
      if (RING && (desc & (1 << 12))) { mbar_wait(&sm.mbar[buf], (mbarPhase >> buf) & 1u); mbarPhase ^= 1u << buf; }
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
      bv.rows = sm.rows[buf] + (desc & 1);
      bv.cols = reinterpret_cast<const uint8_t *>(sm.colw[buf]) + ((desc >> 4) & 15);
      bv.qrows = reinterpret_cast<const uint8_t *>(sm.qcw[buf]) + ((desc >> 8) & 15);
      bv.qvs = reinterpret_cast<const uint8_t *>(sm.qvw[FN == 1 ? buf : 0]) + ((desc >> 8) & 15);   // q and qual share their offsets
      bv.wbase = wbase;
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
      const bool slow = first || last;
      bool ring = false;
      if constexpr (RING) {
        ring = !__any_sync(0xffffffffu, slow);
        if (ring) RingDispatch<LPJ, KM, AFFINE, FN, 1>::run(k, Se, So, AIe, AIo, ADe, ADo, bv, mtabAddr, c, sl, aw, live);
      }
      if (!ring) run_block_gen<LPJ, KM, AFFINE, FN>(Se, So, AIe, AIo, ADe, ADo, bv, sm.rowq, mtabAddr, c, k, sl, first, eLast, aw, live);
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

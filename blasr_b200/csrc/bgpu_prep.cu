// bgpu_prep.cu -- guide construction on the device.
//
// One warp per job turns the candidate's block list into the per-row band table the fill kernels
// stage into shared memory, exactly as AlignmentToGuide builds its GuideRow list
// (common/algorithms/alignment/GuidedAlign.h:104-259), but without the row-to-row dependency:
// the reference's   tPre_i = min(cap_i, t_i - (t_{i-1} - tPre_{i-1}))   is the prefix maximum
//     L_i = max(L_{i-1}, t_i - cap_i),   L_0 = tStart - 1
// of the left edges (cap = bandSize inside a block, 250 on gap rows, unbounded on the first row
// of a block), so a warp scan over rows reproduces it.  The same pass measures, per block of 64
// anti-diagonals, which diagonals hold in-band cells; that fixes the register window and the
// traceback layout of the fill kernel.
#include "bgpu_common.cuh"

namespace bgpu {

__device__ __forceinline__ int warp_incl_max(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = max(v, u); }
  return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_or(int v) { return __reduce_or_sync(0xffffffffu, (unsigned)v); }

constexpr int MAX_BAND_SIZE = 250;  // GuidedAlign.h:29

// rowOffIn / dblkOffIn / runOffIn are host-computed exclusive prefix sums of the per-job capacities.
constexpr int PREP_WARPS = 4, PREP_WIN = 96;   // warps per CTA; guide blocks staged per warp
__global__ void __launch_bounds__(PREP_WARPS * 32) prep_guided_kernel(BatchDev B, ScoreParams P, int defaultBand,
                                                          const uint64_t *rowOffIn, const uint64_t *dblkOffIn,
                                                          const uint64_t *runOffIn) {
  __shared__ uint8_t lut[256];
  __shared__ bgpu_block sWin[PREP_WARPS][PREP_WIN];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = base_code((uint8_t)i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t job = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (job >= B.nJobs) return;
  JobGeom &G = B.geom[job];
  const uint64_t g0 = B.guideOff[job];
  const int nB = (int)(B.guideOff[job + 1] - g0);
  const bgpu_block *blk = B.guide + g0;
  const uint64_t qo = B.qOff[job], to = B.tOff[job];
  const uint32_t qLen = (uint32_t)(B.qOff[job + 1] - qo), tLen = (uint32_t)(B.tOff[job + 1] - to);
  const int band = B.band ? B.band[job] : defaultBand;

  int status = BGPU_JOB_OK;
  if (lane == 0) {
    G.rowOff = rowOffIn[job] + ROW_PAD; G.dblkOff = dblkOffIn[job]; G.runOff = runOffIn[job];
    G.arrowBytes = 0; G.nRuns = G.nBlocks = G.nGaps = G.nGapLists = 0;
    G.qPos = G.tPos = 0; G.score = 0; G.nCells = 0; G.kmax = 0; G.nDB = 0; G.band = band;
    G.Qn = G.Tn = 0; G.qStart = G.tStart = 0; G.C0 = 0; G.hi0 = 0; G.cls = 0; G.ksum = 0; G.minW = 0;
  }
  if (nB == 0) { if (lane == 0) G.status = BGPU_JOB_EMPTY_GUIDE; return; }   // GuidedAlign.h:388-392
  if (band < 0) { if (lane == 0) G.status = BGPU_JOB_BAD_INPUT; return; }

  // ---- validate the block chain: ordered, non-overlapping, non-empty, inside the sequences
  int bad = 0;
  for (int b = lane; b < nB; b += 32) {
    bgpu_block c = blk[b];
    if (c.length == 0 || c.length > 0x3fffffffu || c.qPos > 0x3fffffffu || c.tPos > 0x3fffffffu) bad = 1;
    if (b + 1 < nB) {
      bgpu_block n = blk[b + 1];
      if (n.qPos < c.qPos + c.length || n.tPos < c.tPos + c.length) bad = 1;
    }
  }
  const bgpu_block first = blk[0], last = blk[nB - 1];
  const int qStart = (int)first.qPos, tStart = (int)first.tPos;
  const long long qEndL = (long long)last.qPos + last.length, tEndL = (long long)last.tPos + last.length;
  if (qEndL > qLen || tEndL > tLen) bad = 1;
  if (warp_or(bad)) { if (lane == 0) G.status = BGPU_JOB_BAD_INPUT; return; }
  const int qEnd = (int)qEndL, tEnd = (int)tEndL;
  const int Qn = qEnd - qStart, Tn = tEnd - tStart;
  const int C0 = Qn + (Qn & 1);
  const int nD = Qn + Tn + 1, nDB = (nD + DBLK - 1) / DBLK;
  if (nD >= (1 << 21)) { if (lane == 0) G.status = BGPU_JOB_RANGE; return; }   // diagonals are carried * 8
  // score range the shifted-domain kernels can carry.  Every in-band cell is reachable by Diagonal / Left / Up moves
  // alone (they stay candidates of the affine 5-way min), so |S| <= steps * max(|M|, |ins|, |del|, |ext|); the affine
  // matrices sit at most one open above S.  QualityValueScoreFunction: |Match| <= the largest QV of this job's rows
  // (reduced below), IDSScoreFunction: <= 255 or a prior.
  int mxStep = max(max(abs(P.ins), abs(P.del)), abs(P.ext));
  if (P.kind == BGPU_FN_IDS) mxStep = max(max(mxStep, 255), max(abs(P.subPrior), abs(P.delPrior)));
  else if (P.kind == BGPU_FN_DISTANCE) for (int i = 0; i < 25; i++) mxStep = max(mxStep, abs(P.M[i]));
  if (P.open < 0) mxStep = max(mxStep, abs(P.open) + abs(P.ext));   // a negative open could be collected once per step
  if (mxStep >= (1 << 15) || abs(P.open) >= (1 << 20)) { if (lane == 0) G.status = BGPU_JOB_RANGE; return; }

  // ---- encode + validate the target window (codes 0..4 into B.tc), and check the query bases
  const uint8_t *tb = B.t + to;
  uint8_t *tcb = B.tc + to;
  const uint8_t *qb = B.q + qo;
  const bool keepRaw = P.kind == BGPU_FN_IDS;           // IDSScoreFunction compares raw bytes (IDSScoreFunction.h:129-132)
  {
    // t and tc share their offsets, so one word grid is aligned for both: bytes up to the first 4-byte boundary, whole
    // words (four table look-ups per load / store), then the tail bytes
    const int head = min((int)((4u - (unsigned)((uintptr_t)(tb + tStart) & 3u)) & 3u), Tn);
    const int nW = (Tn - head) >> 2, tail0 = tStart + head + 4 * nW;
    auto one = [&](int i) { const uint8_t r = tb[i], c = lut[r]; if (c > 4) bad = 1; tcb[i] = keepRaw ? r : c; };
    if (lane < head) one(tStart + lane);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(tb + tStart + head);
    uint32_t *dst = reinterpret_cast<uint32_t *>(tcb + tStart + head);
    for (int w = lane; w < nW; w += 32) {
      const uint32_t v = src[w];
      const uint32_t c0 = lut[v & 0xff], c1 = lut[(v >> 8) & 0xff], c2 = lut[(v >> 16) & 0xff], c3 = lut[v >> 24];
      if (max(max(c0, c1), max(c2, c3)) > 4) bad = 1;
      dst[w] = keepRaw ? v : (c0 | (c1 << 8) | (c2 << 16) | (c3 << 24));
    }
    if (tail0 + lane < tEnd) one(tail0 + lane);
  }
  if (warp_or(bad)) { if (lane == 0) G.status = BGPU_JOB_BAD_INPUT; return; }

  // ---- live-diagonal range per d-block.  The warp owns these arrays: contributions are reduced across the
  //      lanes with REDUX and lane 0 does a plain read-modify-write, no atomics.
  int32_t *dmin = B.dmin + dblkOffIn[job], *dmax = B.dmax + dblkOffIn[job];
  for (int b = lane; b < nDB; b += 32) { dmin[b] = INT_MAX; dmax[b] = INT_MIN; }
  __syncwarp();

  RowInfo *rows = B.rows + rowOffIn[job] + ROW_PAD;      // row 0; ROW_PAD dead rows lie on either side of rows [0, Qn]
  for (int r = lane; r < ROW_PAD; r += 32) {
    RowInfo dead; dead.lo8 = DEAD_LO8; dead.nhi8 = -DEAD_LO8;
    rows[r - ROW_PAD] = dead; rows[Qn + 1 + r] = dead;
  }
  const int drift0 = abs(tStart - qStart);                       // GuidedAlign.h:128
  const int tPost0 = drift0 > band ? drift0 : band;              // :129-134
  long long cells = 0;
  int wide = 0;

  // every lane passes its row (or an empty range lo > hi); row i covers t' in [lo,hi]
  auto add_rows = [&](int i, int lo, int hi) {
    const bool has = lo <= hi;
    const int bLo = has ? (i + lo) >> 6 : INT_MAX, bHi = has ? (i + hi) >> 6 : INT_MIN;
    const int w0 = __reduce_min_sync(0xffffffffu, bLo), w1 = __reduce_max_sync(0xffffffffu, bHi);
    for (int b = w0; b <= w1; b++) {
      int mn = INT_MAX, mx = INT_MIN;
      if (has && b >= bLo && b <= bHi) {
        const int tl = max(lo, (b << 6) - i), th = min(hi, (b << 6) + 63 - i);
        mn = tl - i + C0; mx = th - i + C0;
      }
      mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
      if (lane == 0) { dmin[b] = min(dmin[b], mn); dmax[b] = max(dmax[b], mx); }
    }
  };

  // row 0: boundary row, t' in [0, min(tPost0, Tn)]
  const int hi0 = min(tPost0, Tn);
  if (lane == 0) {
    rows[0].lo8 = C0 * 8; rows[0].nhi8 = -(C0 + hi0) * 8;          // boundary row: columns [0, hi0], no base
    cells += (long long)tPost0 + 1;                    // tPre=0
  }
  int minW = hi0 + 1;
  add_rows(0, lane == 0 ? 0 : 1, lane == 0 ? hi0 : 0);

  int carryL = tStart - 1;                              // L_0
  int bcur = 0;                                         // guide block of the chunk's first row
  // A window of the block list lives in shared memory (PREP_WIN blocks per warp, refilled as the rows advance): the
  // 32 rows of a chunk find their blocks without searching -- lane j looks at block bcur+1+j, the starts that fall
  // inside the chunk are OR-reduced into a 32-bit mask (blocks are ordered and hold >= 1 row each, so there are at
  // most 32 of them) and row r lies in block bcur + popc(mask & bits[0..r]).
  bgpu_block *win = sWin[threadIdx.x >> 5];
  int wb = 0, we = 0;                                   // window = blocks [wb, we)
  uint8_t qchNext = Qn >= 1 + lane ? qb[qStart + lane] : 0;
  // QualityValueScoreFunction: the fill kernels read the rows' QVs from B.qual themselves; only their maximum is needed here
  uint8_t *qcb = B.qc + qo;
  const uint8_t *qvb = (P.kind == BGPU_FN_QUALITY && B.qual) ? B.qual + qo : nullptr;
  uint8_t qvNext = (qvb && Qn >= 1 + lane) ? qvb[qStart + lane] : 0;
  int maxQV = 0;
  for (int base = 1; base <= Qn; base += 32) {
    const int i = base + lane;
    const bool act = i <= Qn;
    if (min(nB, bcur + 34) > we) {
      __syncwarp();
      wb = max(bcur - 1, 0); we = min(nB, wb + PREP_WIN);
      const uint32_t *src = reinterpret_cast<const uint32_t *>(blk + wb);
      uint32_t *dst = reinterpret_cast<uint32_t *>(win);
      for (int w = lane; w < 3 * (we - wb); w += 32) dst[w] = src[w];
      __syncwarp();
    }
    int t = 0, cap = 0, tPost = 0, x = INT_MIN;
    const uint8_t qch = qchNext, qv = qvNext;
    maxQV = max(maxQV, (int)qv);
    if (i + 32 <= Qn) { qchNext = qb[qStart + i + 31]; if (qvb) qvNext = qvb[qStart + i + 31]; }
    const uint32_t q0 = (uint32_t)(qStart + base - 1);
    const int cand = bcur + 1 + lane;
    const uint32_t rel = cand < nB ? win[cand - wb].qPos - q0 : 0xffffffffu;
    const uint32_t mask = __reduce_or_sync(0xffffffffu, rel < 32u ? 1u << rel : 0u);
    const int lo = bcur + __popc(mask & (0xffffffffu >> (31 - lane)));
    if (act) {
      const uint32_t q = q0 + (uint32_t)lane;
      const bgpu_block c = win[lo - wb];
      const uint32_t off = q - c.qPos;
      if (off < c.length) {
        t = (int)(c.tPos + off);
        if (off == 0) {                                 // first base of a block: :161-165
          int drift;
          if (lo == 0) drift = drift0;
          else { const bgpu_block p = win[lo - 1 - wb]; drift = (int)(c.tPos - (p.tPos + p.length)) - (int)(c.qPos - (p.qPos + p.length)); }
          cap = INT_MAX; tPost = band + abs(drift);
        } else { cap = band; tPost = min(MAX_BAND_SIZE, band); }   // :170-174
      } else {                                          // gap rows after block lo: :213-250
        const bgpu_block n = win[lo + 1 - wb];
        const int g = (int)(off - c.length);
        const int qGap = (int)(n.qPos - (c.qPos + c.length)), tGap = (int)(n.tPos - (c.tPos + c.length));
        const int diag = min(qGap, tGap);
        t = (int)(c.tPos + c.length) + min(g, diag);
        cap = MAX_BAND_SIZE; tPost = min(MAX_BAND_SIZE, band + abs(tGap - qGap));
      }
      x = (cap == INT_MAX) ? INT_MIN : t - cap;
    }
    bcur += __popc(mask);
    int L = warp_incl_max(x, lane);
    L = max(L, carryL);
    carryL = __shfl_sync(0xffffffffu, L, 31);
    int lop = 1, hip = 0;
    if (act) {
      const int tPre = t - L;                           // >= 0 for ordered blocks
      cells += (long long)tPre + tPost + 1;
      const int hi = min(t + tPost, tEnd - 1);
      lop = L - tStart + 1; hip = hi - tStart + 1;
      if (tPre < 0 || hip < lop) bad = 1;
      minW = min(minW, hip - lop + 1);
      const uint32_t qc = lut[qch];
      if (qc > 4) bad = 1;
      RowInfo r; r.lo8 = (lop - i + C0) * 8; r.nhi8 = -(hip - i + C0) * 8;
      rows[i] = r;
      qcb[qStart + i - 1] = (uint8_t)((qc & 7u) * 20u);
      if (bad || wide) { lop = 1; hip = 0; }
    }
    add_rows(i, lop, hip);
  }
  cells = warp_sum_ll(cells);
  if (warp_or(bad) || cells > INT_MAX) { if (lane == 0) G.status = BGPU_JOB_BAD_INPUT; return; }
  if (P.kind == BGPU_FN_QUALITY) mxStep = max(mxStep, __reduce_max_sync(0xffffffffu, maxQV));
  if ((long long)mxStep * (Qn + Tn + 2) + abs(P.open) >= (P.affine ? SCORE_LIMIT_AFF : SCORE_LIMIT_LIN)) {
    if (lane == 0) G.status = BGPU_JOB_RANGE;
    return;
  }
  __threadfence(); __syncwarp();

  // ---- per d-block window [mn-1, mx+1] aligned down to an even diagonal (the edge slots stay dead, see
  //      bgpu_fill.cu); the widest window picks the job's class, i.e. how many lanes sweep it
  DBlock *db = B.dblk + dblkOffIn[job];
  int maxSpan = 0;
  for (int b = lane; b < nDB; b += 32) {
    const int mn = __ldcg(&dmin[b]), mx = __ldcg(&dmax[b]);
    int wbase = (mn - 1) & ~1, span = mx + 1 - wbase;
    if (mx < mn) { wbase = 0; span = 0; }
    db[b].wbase = wbase;
    maxSpan = max(maxSpan, span);
  }
  maxSpan = __reduce_max_sync(0xffffffffu, maxSpan);
  int cls = CLS_L8N;
  if (maxSpan >= 2 * 8 * 4) cls = CLS_L8;
  if (maxSpan >= 2 * 8 * KRING) cls = CLS_L16;
  if (maxSpan >= 2 * 16 * KRING) cls = CLS_L32;
  if (maxSpan >= 2 * 32 * KRING) cls = CLS_WIDE;
  if (warp_or(wide) || maxSpan >= 2 * 32 * KWIDE) status = BGPU_JOB_TOO_WIDE;
  const int gw = 2 * cls_lpj(cls);
  int kmax = 0, ksum = 0;
  for (int b = lane; b < nDB; b += 32) {
    const int mn = __ldcg(&dmin[b]), mx = __ldcg(&dmax[b]);
    const int span = mx < mn ? 0 : mx + 1 - db[b].wbase;
    const int k = span / gw + 1;
    db[b].k = k; db[b].arrowUnit = 0; db[b].pad = 0;
    kmax = max(kmax, k); ksum += k;
  }
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  ksum = __reduce_add_sync(0xffffffffu, ksum);
  minW = __reduce_min_sync(0xffffffffu, minW);
  if (lane == 0) {
    G.status = status; G.qStart = qStart; G.tStart = tStart; G.Qn = Qn; G.Tn = Tn; G.C0 = C0;
    G.nDB = nDB; G.kmax = kmax; G.nCells = (int)cells; G.hi0 = hi0; G.cls = cls; G.ksum = ksum; G.minW = minW;
    G.arrowBytes = 0;
  }
}

// Packed guides (bgpu_batch::guidePacked): three bytes per block (gap before the block in q, in t, length), a side list
// for the rare block that does not fit a byte.  One warp per job turns them back into the {qPos, tPos, length} blocks
// prep reads: the positions are prefix sums of the advances.
__global__ void __launch_bounds__(128) unpack_guide_kernel(uint32_t nJobs, const uint64_t *guideOff, const uint8_t *packed,
                                                           const uint32_t *wide, uint64_t nWide, bgpu_block *guide) {
  const int lane = threadIdx.x & 31;
  const uint32_t job = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (job >= nJobs) return;
  const uint64_t g0 = guideOff[job], nb = guideOff[job + 1] - g0;
  uint32_t carryQ = 0, carryT = 0;
  for (uint64_t base = 0; base < nb; base += 32) {
    const uint64_t g = g0 + base + lane;
    const bool act = base + lane < nb;
    uint32_t dq = 0, dt = 0, len = 0;
    if (act) {
      dq = packed[3 * g]; dt = packed[3 * g + 1]; len = packed[3 * g + 2];
      if (dq == 255 && dt == 255 && len == 255) {          // escaped: binary search of the side list
        uint64_t lo = 0, hi = nWide;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (wide[4 * mid] < (uint32_t)g) lo = mid + 1; else hi = mid; }
        if (lo < nWide && wide[4 * lo] == (uint32_t)g) { dq = wide[4 * lo + 1]; dt = wide[4 * lo + 2]; len = wide[4 * lo + 3]; }
        else { dq = dt = 0; len = 0; }                      // malformed: a zero-length block makes prep reject the job
      }
    }
    uint32_t aq = dq + len, at = dt + len, xq = aq, xt = at;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t uq = __shfl_up_sync(0xffffffffu, xq, o), ut = __shfl_up_sync(0xffffffffu, xt, o);
      if (lane >= o) { xq += uq; xt += ut; }
    }
    if (act) { bgpu_block b; b.qPos = carryQ + xq - len; b.tPos = carryT + xt - len; b.length = len; guide[g] = b; }
    carryQ += __shfl_sync(0xffffffffu, xq, 31); carryT += __shfl_sync(0xffffffffu, xt, 31);
  }
}

// Targets as windows of the reference resident on the device (bgpu_batch.tRefOff): one CTA per job copies its window into
// the ticket's contiguous target array, reverse-complemented on request (ReverseComplementNuc, NucConversion.h:337-352: ACGT <->
// TGCA in either case, N kept; every other byte is kept too, where the reference's table would produce 127).
__global__ void __launch_bounds__(256) gather_reference_kernel(uint32_t nJobs, const uint8_t *ref, uint64_t refLen, const uint64_t *refOff,
                                                               const uint8_t *rc, const uint64_t *tOff, uint8_t *t) {
  const uint32_t job = blockIdx.x;
  if (job >= nJobs) return;
  const uint64_t o = tOff[job], len = tOff[job + 1] - o, r0 = refOff[job];
  const bool rev = rc && rc[job];
  for (uint64_t i = threadIdx.x; i < len; i += blockDim.x) {
    const uint64_t src = rev ? r0 + len - 1 - i : r0 + i;
    uint8_t c = src < refLen ? ref[src] : (uint8_t)'N';        // a window past the end of the reference reads as N
    if (rev) {
      switch (c) {
        case 'A': c = 'T'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break; case 'T': c = 'A'; break;
        case 'a': c = 't'; break; case 'c': c = 'g'; break; case 'g': c = 'c'; break; case 't': c = 'a'; break;
        default: break;
      }
    }
    t[o + i] = c;
  }
}
void launch_gather_reference(uint32_t nJobs, const uint8_t *ref, uint64_t refLen, const uint64_t *refOff, const uint8_t *rc,
                             const uint64_t *tOff, uint8_t *t, cudaStream_t s) {
  if (nJobs) gather_reference_kernel<<<nJobs, 256, 0, s>>>(nJobs, ref, refLen, refOff, rc, tOff, t);
}

void launch_unpack_guide(uint32_t nJobs, const uint64_t *guideOff, const uint8_t *packed, const uint32_t *wide, uint64_t nWide,
                         bgpu_block *guide, cudaStream_t s) {
  const unsigned grid = (nJobs + 3) / 4;
  if (grid) unpack_guide_kernel<<<grid, 128, 0, s>>>(nJobs, guideOff, packed, wide, nWide, guide);
}

void launch_prep_guided(const BatchDev &B, const ScoreParams &P, int defaultBand, const uint64_t *rowOff,
                        const uint64_t *dblkOff, const uint64_t *runOff, cudaStream_t s) {
  const int warpsPerBlock = PREP_WARPS;
  const unsigned grid = (B.nJobs + warpsPerBlock - 1) / warpsPerBlock;
  if (grid) prep_guided_kernel<<<grid, warpsPerBlock * 32, 0, s>>>(B, P, defaultBand, rowOff, dblkOff, runOff);
}

}  // namespace bgpu

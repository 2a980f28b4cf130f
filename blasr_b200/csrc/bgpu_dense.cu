// bgpu_dense.cu -- KBandAlign, SWAlign and AffineKBandAlign on the device (8f N1).
//
// Reference semantics restated:
//   KBandAlign  common/algorithms/alignment/KBandAlign.h:75-403  (+ SetKBoundedLengths :36-56)
//   SWAlign     common/algorithms/alignment/SWAlign.h:18-389
//   AffineKBandAlign  common/algorithms/alignment/AffineKBandAlign.h:12-401 (second half of this file)
//
// The in-row dependency  S[t] = min(A[t], S[t-1] + del_t)  is the min-plus prefix scan
//     S[t] = min_{j<=t} (A[j] - D[j]) + D[t],   D = prefix sum of the deletion costs of the row
// (D[t] = t * del for the DistanceMatrix / QualityValue score functions; IDSScoreFunction's deletion cost depends on
// the row's deletion tag and the column's base, so D comes from a warp prefix sum), which a warp evaluates for 32
// columns with five shuffles: one warp per job sweeps the matrix row by
// row, 32 columns per step, all lanes busy.  The previous row lives in an L1/L2-resident ping-pong
// buffer indexed by absolute column; one traceback byte per cell is written row-major with the layout
// the reference uses (SW: (|q|+1) x (|t|+1); k-band: (qLen+1) x (2k+1)), boundary cells included, so the
// traceback kernel needs no special cases.  The reference's quirks that change results are kept:
// k-band boundary costs come from the ins/del *parameters* while the fill uses the score function's;
// row 0 is only initialised for t < tLen; TargetFit/Fit write band column 0; Fit/TargetFit start the
// traceback in the corner's band column; SW's local minimum is remembered by its 0-based loop indices.
#include "bgpu_common.cuh"

namespace bgpu {

constexpr int DBIG = 1 << 29;
enum { DN_DIAG = 0, DN_UP = 1, DN_LEFT = 2, DN_NONE = 7 };   // Path.h:4-19 (only these occur here)
enum { RUN_D = 0, RUN_U = 1, RUN_L = 2 };


__device__ __forceinline__ int warp_prefix_min(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = min(v, u); }
  return v;
}

// SetKBoundedLengths, KBandAlign.h:36-56
__host__ __device__ inline void kbounded(uint32_t tLength, uint32_t qLength, uint32_t k, uint32_t &tLen, uint32_t &qLen) {
  if (tLength < qLength) { tLen = tLength; qLen = qLength < tLength + k ? qLength : tLength + k; }
  else if (qLength < tLength) { qLen = qLength; tLen = tLength < qLength + k ? tLength : qLength + k; }
  else { qLen = qLength; tLen = tLength; }
}

// warp per job: validate + encode bases in place, fix the job's extents.  Offsets come from the host.
__global__ void __launch_bounds__(128) dense_prep_kernel(BatchDev B, ScoreParams P, DenseArgs A, const uint64_t *rowBufOff,
                                                         const uint64_t *runOff) {
  __shared__ uint8_t lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = base_code((uint8_t)i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t job = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (job >= B.nJobs) return;
  JobGeom &G = B.geom[job];
  const uint64_t qo = B.qOff[job], to = B.tOff[job];
  const uint32_t qLength = (uint32_t)(B.qOff[job + 1] - qo), tLength = (uint32_t)(B.tOff[job + 1] - to);
  const int k = A.algo != BGPU_SW ? (B.band ? B.band[job] : A.defaultBand) : 0;
  const int at = P.alignType;
  uint32_t qLen = qLength, tLen = tLength;
  int status = BGPU_JOB_OK;
  long long nCells = 0;
  if (A.algo == BGPU_AFFINE_KBAND) {
    // AffineKBandAlign.h:12-401.  Global and QueryFit; TargetFit searches mirrored band columns (:322-330) and its
    // traceback can spin on a NoArrow cell; other types leave the end cell outside the matrix.
    if (k < 0 || (at != BGPU_GLOBAL && at != BGPU_QUERYFIT && at != BGPU_TARGETFIT)) status = BGPU_JOB_BAD_INPUT;
    else if (at == BGPU_TARGETFIT) status = BGPU_JOB_REF_UNDEFINED;
    else {
      kbounded(tLength, qLength, (uint32_t)k, tLen, qLen);
      nCells = ((long long)qLen + 1) * (2ll * k + 1);
      if (nCells > INT_MAX) status = BGPU_JOB_BAD_INPUT;
      else if (at == BGPU_QUERYFIT && (tLen == 0 || qLen == 0)) status = BGPU_JOB_REF_UNDEFINED;   // end search reads unwritten cells
      // INF_SCORE = INT_MAX - 1000 (:30): larger costs overflow int in the reference
      else if (max(max(abs(A.hpInsOpen), abs(A.hpInsExtend)), max(max(abs(A.insOpen), abs(A.insExtend)), abs(A.bndDel))) >= 1000) status = BGPU_JOB_RANGE;
      nCells = 0;                                          // the reference never sets alignment.nCells here
    }
  } else if (A.algo == BGPU_KBAND) {
    if (k < 0 || at < 0 || at > BGPU_TPREFIXQSUFFIX) status = BGPU_JOB_BAD_INPUT;
    else {
      kbounded(tLength, qLength, (uint32_t)k, tLen, qLen);
      nCells = ((long long)qLen + 1) * (2ll * k + 1);                                 // KBandAlign.h:96-98
      if (nCells > INT_MAX) status = BGPU_JOB_BAD_INPUT;
      // the reference compares q2 >= tLen - k in unsigned arithmetic and then reads an uninitialised index (:286,:307)
      else if ((at == BGPU_TARGETFIT || at == BGPU_FIT) && ((uint32_t)k > tLen || qLen == 0)) status = BGPU_JOB_REF_UNDEFINED;
    }
  } else {
    if (at < 0 || at == BGPU_FIT || at > BGPU_TPREFIXQSUFFIX) status = BGPU_JOB_BAD_INPUT;         // SWAlign.h has no Fit
    if (((long long)qLength + 1) * ((long long)tLength + 1) > INT_MAX) status = BGPU_JOB_BAD_INPUT;
  }
  if (P.kind == BGPU_FN_QUALITY && !B.qual) status = BGPU_JOB_BAD_INPUT;
  if (P.kind == BGPU_FN_IDS && (A.algo != BGPU_KBAND || !B.insQV || !B.subQV || !B.subTag)) status = BGPU_JOB_BAD_INPUT;
  {
    int mx = max(max(abs(P.ins), abs(P.del)), max(abs(A.bndIns), abs(A.bndDel)));
    if (P.kind == BGPU_FN_QUALITY) mx = max(mx, 255);
    else if (P.kind == BGPU_FN_IDS) mx = max(max(mx, 255), max(abs(P.subPrior), abs(P.delPrior)));
    else for (int i = 0; i < 25; i++) mx = max(mx, abs(P.M[i]));
    if ((long long)mx * ((long long)qLength + tLength + 2) >= (1 << 28) || mx >= (1 << 15)) status = status ? status : BGPU_JOB_RANGE;
  }
  int bad = 0;
  const uint8_t *tb = B.t + to; uint8_t *tcb = B.tc + to; uint8_t *qb = B.q + qo;
  const bool keepRaw = P.kind == BGPU_FN_IDS;            // IDSScoreFunction compares raw bytes; base_code() is idempotent on codes
  for (uint32_t i = lane; i < tLen; i += 32) { const uint8_t r = tb[i], c = lut[r]; if (c > 4) bad = 1; tcb[i] = keepRaw ? r : c; }
  for (uint32_t i = lane; i < qLen; i += 32) { const uint8_t c = lut[qb[i]]; if (c > 4) bad = 1; }
  bad = __reduce_or_sync(0xffffffffu, (unsigned)bad);
  if (bad && status == BGPU_JOB_OK) status = BGPU_JOB_BAD_INPUT;
  if (lane == 0) {
    G.status = status; G.Qn = (int)qLen; G.Tn = (int)tLen; G.band = k; G.nCells = A.algo == BGPU_KBAND ? (int)nCells : 0;   // SWAlign / AffineKBandAlign never set it
    G.qStart = G.tStart = 0; G.C0 = 0; G.nDB = 0; G.kmax = 0; G.hi0 = 0; G.score = 0;
    G.rowOff = 0; G.dblkOff = 0; G.arrowBytes = 0; G.runOff = runOff[job]; G.rowBufOff = rowBufOff[job];
    G.nRuns = G.nBlocks = G.nGaps = G.nGapLists = 0; G.qPos = G.tPos = 0; G.startR = G.startC = 0;
  }
}

__global__ void __launch_bounds__(128) dense_fill_kernel(BatchDev B, ScoreParams P, DenseArgs A, const uint32_t *order,
                                                         uint32_t nOrder, uint32_t *counter) {
  __shared__ int Mtab[25];
  __shared__ uint8_t lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = base_code((uint8_t)i);
  if (threadIdx.x < 25) Mtab[threadIdx.x] = P.M[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const bool sw = A.algo == BGPU_SW, qv = P.kind == BGPU_FN_QUALITY, ids = P.kind == BGPU_FN_IDS;
  const int at = P.alignType;
  for (;;) {
    uint32_t idx = 0;
    if (lane == 0) idx = atomicAdd(counter, 1u);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= nOrder) break;
    const uint32_t job = order[idx];
    JobGeom &G = B.geom[job];
    if (G.status != BGPU_JOB_OK) continue;
    const int R = G.Qn, T = G.Tn, k = G.band;          // rows 0..R, columns 0..T (already k-bounded for KBandAlign)
    const uint8_t *qb = B.q + B.qOff[job], *tb = B.tc + B.tOff[job];
    const uint8_t *qual = B.qual ? B.qual + B.qOff[job] : nullptr;
    const size_t qo = (size_t)B.qOff[job];
    uint8_t *arrows = B.arrows + A.arrowOff[job];
    int *row0 = B.rowBuf + G.rowBufOff, *row1 = row0 + (T + 2);
    const int nCols = sw ? T + 1 : 2 * k + 1;           // arrow row pitch
    const int ins = P.ins, del = P.del;
    const bool localFam = sw && (at == BGPU_LOCAL || at == BGPU_ENDANCHORED);

    // ---- row 0 (boundary): values into row0[], arrows into the first arrow row
    for (int t = lane; t <= T; t += 32) {
      int v = 0; uint8_t a = DN_NONE;
      if (sw) {                                          // SWAlign.h:49-138
        const bool delRow = at == BGPU_GLOBAL || at == BGPU_FRONTANCHORED || at == BGPU_TARGETFIT || at == BGPU_TPREFIXQSUFFIX;
        v = delRow ? del * t : 0;
        a = localFam ? DN_NONE : DN_LEFT;
        if (t == 0) a = DN_DIAG;                         // :140
        arrows[t] = a;
      } else {                                           // KBandAlign.h:119-130 (only t < tLen is initialised)
        const bool init = t >= 1 && t <= k && t < T && (at == BGPU_GLOBAL || at == BGPU_QUERYFIT || at == BGPU_FIT);
        if (init) v = at == BGPU_GLOBAL ? t * A.bndDel : 0;
      }
      row0[t] = v;
    }
    if (!sw) {
      for (int c = lane; c < nCols; c += 32) {
        const int t = c - k;
        uint8_t a = DN_NONE;
        if (t == 0) a = DN_DIAG;                         // :141-142
        else if (t >= 1 && t < T && (at == BGPU_GLOBAL || at == BGPU_QUERYFIT || at == BGPU_FIT)) a = DN_LEFT;
        arrows[c] = a;
      }
    }
    __syncwarp();

    // SW bookkeeping
    int lmVal = 0, lmR = 0, lmC = 0, lmDiag = 0;         // local minimum (0-based loop indices) :153-173
    int colBest = 0, colBestRow = 0, colDiagVal = 0; bool colSet = false; // last-column minimum (TargetFit families)
    int *prev = row0, *cur = row1;

    for (int r = 1; r <= R; r++) {
      const int tlo = sw ? 1 : max(1, r - k), thi = sw ? T : min(T, r + k);
      // boundary column 0 of this row
      int c0v = 0; uint8_t c0a = DN_NONE; bool c0in = sw || r <= k;
      if (sw) {
        const bool insCol = at == BGPU_GLOBAL || at == BGPU_FRONTANCHORED || at == BGPU_QUERYFIT || at == BGPU_OVERLAP || at == BGPU_TSUFFIXQPREFIX;
        c0v = insCol ? ins * r : 0; c0a = localFam ? DN_NONE : DN_UP;
      } else if (r <= k) {
        c0v = r * A.bndIns; c0a = DN_UP;                                       // KBandAlign.h:115-118
        if ((at == BGPU_TARGETFIT || at == BGPU_FIT) && r == k && r < R) c0v = 0; // :131-136 (band column 0 of row k)
      }
      uint8_t *arow = arrows + (size_t)r * nCols;
      if (lane == 0 && c0in) { cur[0] = c0v; if (sw) arow[0] = c0a; else arow[k - r] = c0a; }
      // k-band: cells of the band that are never computed keep NoArrow
      if (!sw) {
        for (int c = lane; c < nCols; c += 32) { const int t = r - k + c; if (t < 0 || t > T || (t == 0 && !c0in) ) arow[c] = DN_NONE; }
      }
      const uint8_t qraw = qb[r - 1], qch = lut[qraw];
      const int qvv = qv ? (int)qual[r - 1] : 0;
      // IDSScoreFunction.h:80-139: this row's tracks (KBandAlign passes the cell's own positions, KBandAlign.h:163-189)
      int rowIns = ins, subTag = 0, subQ = 0, delTag = 0, delQ = 0; bool hasDel = false;
      if (ids) {
        rowIns = (int)B.insQV[qo + r - 1]; subTag = (int)B.subTag[qo + r - 1]; subQ = (int)B.subQV[qo + r - 1];
        if (B.delQV) { hasDel = true; delTag = (int)B.delTag[qo + r - 1]; delQ = (int)B.delQV[qo + r - 1]; }
      }
      int carry = (sw || tlo > r - k) ? c0v : DBIG;      // value left of column tlo (k-band: none at the band edge)
      if (!sw && tlo > 1) carry = DBIG;                   // tlo == r-k >= 2: left band edge, deletion not allowed (:155-157)
      for (int base = tlo; base <= thi; base += 32) {
        const int t = base + lane;
        const bool act = t <= thi;
        int ms = DBIG, is = DBIG;
        int diag = 0, dcost = act ? del : 0;
        if (act) {
          diag = prev[t - 1];
          const uint8_t traw = tb[t - 1], tch = lut[traw];
          int m;
          if (ids) {
            m = (qraw == traw) ? 0 : (subTag == (int)traw ? subQ : P.subPrior);
            if (hasDel) dcost = (delTag != 'N' && delTag == (int)traw) ? delQ : P.delPrior;
          }
          else if (qv) m = ((qch == tch && qch < 4) ? -1 : 1) * qvv; else m = Mtab[qch * 5 + tch];
          ms = diag + m;
          if (sw || t != r + k) is = prev[t] + rowIns;    // right band edge: no insertion (:182-184)
        }
        int a0 = min(ms, is);
        if (localFam) a0 = min(a0, 0);                    // reset-to-zero == an extra zero candidate
        // D = inclusive prefix sum of the deletion costs of this chunk's columns
        int dpre = dcost;
        if (ids) {
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, dpre, o); if (lane >= o) dpre += u; }
        } else dpre = (lane + 1) * del;
        // S[t] = min_j (A[j] - D[j]) + D[t]  and the chunk's left neighbour
        int bv = act ? a0 - dpre : DBIG;
        bv = warp_prefix_min(bv, lane);
        int s = min(bv + dpre, carry >= DBIG ? DBIG : carry + dpre);
        int left = __shfl_up_sync(0xffffffffu, s, 1);
        if (lane == 0) left = carry;
        const int ds = left >= DBIG ? DBIG : left + dcost;
        const int best = min(min(ms, is), ds);            // the reference's minScore before any reset
        uint8_t arrow;
        if (sw) arrow = best == ms ? DN_DIAG : (best == is ? DN_UP : DN_LEFT);       // SWAlign.h:196-207
        else arrow = best == ms ? DN_DIAG : (best == ds ? DN_LEFT : DN_UP);          // KBandAlign.h:201-209
        if (localFam && best > 0) arrow = DN_NONE;        // :175-181 (value already 0 through the zero candidate)
        if (act) {
          cur[t] = s;
          if (sw) arow[t] = arrow; else arow[k + t - r] = arrow;
          if (sw && best < lmVal) { lmVal = best; lmR = r - 1; lmC = t - 1; lmDiag = diag; }
        }
        carry = __shfl_sync(0xffffffffu, s, 31);
      }
      __syncwarp();
      // last column bookkeeping (uniform): value of cell (r, T) if it was computed in this row
      if (thi == T && T >= 1) {
        const int v = cur[T];
        if (sw) { if (!colSet || v < colBest) { colBest = v; colBestRow = r; colSet = true; } }                 // SWAlign.h:275-298
        else if (r >= max(T - k, 1) && (!colSet || v <= colBest)) {    // KBandAlign.h:286-293 scans q2 downwards: the largest row wins ties
          colBest = v; colBestRow = r; colSet = true;
          // Fit/TargetFit start the traceback in the *corner's* band column of this row (:296,:301), i.e. at
          // column r + (T - R); remember that cell's value, it is the score the reference returns.
          const int ts = r + T - R;
          colDiagVal = (ts >= tlo && ts <= thi) ? cur[ts] : ((ts == 0 && c0in) ? c0v : 0);
        }
      }
      int *tmp = prev; prev = cur; cur = tmp;
    }
    // ---- reduce the per-lane local minimum to the first one in row-major order
    if (sw) {
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const int v = __shfl_xor_sync(0xffffffffu, lmVal, o), rr = __shfl_xor_sync(0xffffffffu, lmR, o);
        const int cc = __shfl_xor_sync(0xffffffffu, lmC, o), dd = __shfl_xor_sync(0xffffffffu, lmDiag, o);
        const bool take = v < lmVal || (v == lmVal && v < 0 && (rr < lmR || (rr == lmR && cc < lmC)));
        if (take) { lmVal = v; lmR = rr; lmC = cc; lmDiag = dd; }
      }
    }
    // ---- choose the traceback start and the returned score
    const int *lastRow = prev;                            // row R (row 0 when R == 0)
    int startR = R, startC = T, score = 0;
    // first minimum over the last row, columns [lo,hi]
    auto lastRowMin = [&](int lo, int hi, int &bestCol) {
      int bv = INT_MAX, bc = INT_MAX;
      for (int t = lo + lane; t <= hi; t += 32) { const int v = lastRow[t]; if (v < bv) { bv = v; bc = t; } }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const int v = __shfl_xor_sync(0xffffffffu, bv, o), c = __shfl_xor_sync(0xffffffffu, bc, o);
        if (v < bv || (v == bv && c < bc)) { bv = v; bc = c; }
      }
      bestCol = bc; return bv;
    };
    int status = BGPU_JOB_OK;
    if (sw) {
      if (at == BGPU_GLOBAL || at == BGPU_ENDANCHORED) { startR = R; startC = T; score = lastRow[T]; }
      else if (at == BGPU_LOCAL || at == BGPU_FRONTANCHORED) { startR = lmR; startC = lmC; score = lmDiag; }   // S at (localMinRow, localMinCol), SWAlign.h:243-246,388
      else if (at == BGPU_QUERYFIT || at == BGPU_OVERLAP || at == BGPU_TPREFIXQSUFFIX) {                        // :248-264,:302-314
        if (T < 1) status = BGPU_JOB_BAD_INPUT; else { int bc; score = lastRowMin(1, T, bc); startR = R; startC = bc; }
      } else {                                           // TargetFit :265-284, TSuffixQPrefix :285-301: last column, rows 1..R
        if (R < 1) status = BGPU_JOB_BAD_INPUT;
        else if (T >= 1) { startR = colBestRow; startC = T; score = colBest; }
        else { startR = 1; startC = 0; score = at == BGPU_TSUFFIXQPREFIX ? ins : 0; }
      }
    } else {
      int q = R, tbc = k - (R - T);                       // corner of the band, KBandAlign.h:247-248
      score = lastRow[T];
      int minRowV = score;
      if (at == BGPU_QUERYFIT || at == BGPU_FIT) {        // :255-279
        const int lo = max(1, R - k), hi = min(T, R + k);
        if (lo <= hi) { int bc; minRowV = lastRowMin(lo, hi, bc); tbc = k - (R - bc); score = minRowV; }
      }
      if (at == BGPU_TARGETFIT || at == BGPU_FIT) {       // :280-304
        if (!colSet) { if (at == BGPU_TARGETFIT) status = BGPU_JOB_REF_UNDEFINED; }
        else if (at == BGPU_TARGETFIT || colBest < minRowV) { tbc = k - (R - T); q = colBestRow; score = colDiagVal; }
      }
      startR = q; startC = tbc;                           // band coordinates
    }
    if (lane == 0) { G.startR = startR; G.startC = startC; G.score = score; if (status != BGPU_JOB_OK) G.status = status; }
  }
}

// one thread per job
__global__ void __launch_bounds__(64) dense_trace_kernel(BatchDev B, ScoreParams P, DenseArgs A, const uint32_t *order, uint32_t nOrder) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nOrder) return;
  const uint32_t job = order[idx];
  JobGeom &G = B.geom[job];
  if (G.status != BGPU_JOB_OK) return;
  const bool sw = A.algo == BGPU_SW;
  const int at = P.alignType, k = G.band, T = G.Tn;
  const int nCols = sw ? T + 1 : 2 * k + 1;
  const uint8_t *arrows = B.arrows + A.arrowOff[job];
  uint32_t *runs = B.runs + G.runOff;
  int r = G.startR, c = G.startC;
  int runType = -1; uint32_t runLen = 0, nRuns = 0, nBlocks = 0, nGaps = 0, pendGaps = 0;
  bool seenD = false, awry = false;
  auto push = [&](int type) {
    if (type == runType) { runLen++; return; }
    if (runType >= 0) runs[nRuns++] = ((uint32_t)runType << 30) | runLen;
    runType = type; runLen = 1;
    if (type == RUN_D) { if (seenD) nGaps += pendGaps; pendGaps = 0; seenD = true; nBlocks++; }
    else pendGaps++;
  };
  const bool traced = sw ? true : (at == BGPU_GLOBAL || at == BGPU_QUERYFIT || at == BGPU_FIT || at == BGPU_TARGETFIT);
  long guard = (long)(G.Qn + 2) * 2 + (long)nCols * 2 + T + 8;
  while (traced) {
    bool go;
    if (sw) {                                              // SWAlign.h:324-333
      if (at == BGPU_GLOBAL || at == BGPU_FRONTANCHORED) go = r > 0 || c > 0;
      else if (at == BGPU_QUERYFIT || at == BGPU_OVERLAP || at == BGPU_TSUFFIXQPREFIX) go = r > 0;
      else if (at == BGPU_TPREFIXQSUFFIX || at == BGPU_TARGETFIT) go = c > 0;
      else go = r > 0 && c > 0 && arrows[(size_t)r * nCols + c] != DN_NONE;
    } else {                                               // KBandAlign.h:327-383 (c is the band column)
      go = r > 0;
      if (at == BGPU_TARGETFIT) { if (r < k && k - r == c) go = false; }
      else { if (c < k && k - c == r) go = false; if (at == BGPU_FIT && r <= k && k - r == c) go = false; }
    }
    if (!go) break;
    if (r < 0 || c < 0 || c >= nCols || --guard < 0) { awry = true; break; }
    const uint8_t a = arrows[(size_t)r * nCols + c];
    if (a == DN_NONE) { if (sw) awry = true; break; }      // k-band: the loop simply stops (:330-332)
    if (a == DN_DIAG) { push(RUN_D); r--; if (sw) c--; }
    else if (a == DN_UP) { push(RUN_U); r--; if (!sw) c++; }
    else if (a == DN_LEFT) { push(RUN_L); c--; }
    else { awry = true; break; }
  }
  if (runType >= 0) runs[nRuns++] = ((uint32_t)runType << 30) | runLen;
  if (awry) { G.status = BGPU_JOB_PATH_AWRY; G.nRuns = G.nBlocks = G.nGaps = G.nGapLists = 0; return; }
  // no RemoveAlignmentPrefixGaps here: the leading gap list is kept, trailing gaps are dropped
  G.nRuns = nRuns; G.nBlocks = nBlocks; G.nGaps = seenD ? nGaps + pendGaps : 0; G.nGapLists = nRuns ? nBlocks + 1 : 0;
  uint32_t qPos = 0, tPos = 0;
  if (sw) {                                                // SWAlign.h:367-380
    if (at != BGPU_GLOBAL && at != BGPU_FRONTANCHORED && at != BGPU_OVERLAP) { qPos = (uint32_t)r; tPos = (uint32_t)c; }
  } else {                                                 // KBandAlign.h:389-399
    qPos = (uint32_t)r;
    tPos = c < k ? (uint32_t)((k - c) - r) : (uint32_t)((c - k) - r);
  }
  G.qPos = qPos; G.tPos = tPos; G.qStart = (int)qPos; G.tStart = (int)tPos;
}

// ---------------------------------------------------------------------------------------------------------------
// AffineKBandAlign (AffineKBandAlign.h:12-401): three band matrices -- main S, homopolymer-insertion H, insertion I.
// H and I of a cell depend on the row above only, S adds the in-row linear deletion, so the row sweep is the same
// min-plus scan as KBandAlign's.  One byte per cell: bits 0-1 main arrow, bit 2 / 3 = the I / H matrix arrow is Open
// (else Up), 0x80 = NoArrow.  INF_SCORE (:30) is DBIG here: it only ever loses comparisons or ties with itself.
enum { AK_DIAG = 0, AK_LEFT = 1, AK_ICLOSE = 2, AK_HCLOSE = 3, AK_IOPEN = 4, AK_HOPEN = 8, AK_NOARROW = 0x80 };

__global__ void __launch_bounds__(128) affine_kband_fill_kernel(BatchDev B, ScoreParams P, DenseArgs A, const uint32_t *order,
                                                                uint32_t nOrder, uint32_t *counter) {
  __shared__ int Mtab[25];
  __shared__ uint8_t lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = base_code((uint8_t)i);
  if (threadIdx.x < 25) Mtab[threadIdx.x] = P.M[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int at = P.alignType;
  const int hpO = A.hpInsOpen, hpE = A.hpInsExtend, inO = A.insOpen, inE = A.insExtend, del = A.bndDel;
  for (;;) {
    uint32_t idx = 0;
    if (lane == 0) idx = atomicAdd(counter, 1u);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= nOrder) break;
    const uint32_t job = order[idx];
    JobGeom &G = B.geom[job];
    if (G.status != BGPU_JOB_OK) continue;
    const int R = G.Qn, T = G.Tn, k = G.band, nCols = 2 * k + 1;
    const uint8_t *qb = B.q + B.qOff[job], *tb = B.tc + B.tOff[job];
    uint8_t *arrows = B.arrows + A.arrowOff[job];
    const int W = T + 2;
    int *pS = B.rowBuf + G.rowBufOff, *pH = pS + W, *pI = pH + W, *cS = pI + W, *cH = cS + W, *cI = cH + W;
    // ---- row 0 (:93-141): S = t * del with Left arrows for t <= k, H = I = INF right of the origin
    for (int t = lane; t <= T; t += 32) { pS[t] = t <= k ? t * del : 0; pH[t] = t ? DBIG : 0; pI[t] = t ? DBIG : 0; }
    for (int c = lane; c < nCols; c += 32) arrows[c] = c > k ? AK_LEFT : (c == k ? (AK_NOARROW | AK_IOPEN | AK_HOPEN) : AK_NOARROW);
    __syncwarp();
    for (int r = 1; r <= R; r++) {
      const int tlo = max(1, r - k), thi = min(T, r + k);
      uint8_t *arow = arrows + (size_t)r * nCols;
      for (int c = lane; c < nCols; c += 32) { const int t = r - k + c; if (t < 0 || t > T) arow[c] = AK_NOARROW; }   // t == 0: the boundary cell below
      // boundary column t = 0 (band column k - r), rows r <= k (:96-99,122-125,134-137)
      int carry = DBIG;                                    // S left of column tlo: none at the left band edge (:226-228)
      if (r <= k) {
        const int bI = r * inE + inO;
        if (lane == 0) { cS[0] = bI; cI[0] = bI; cH[0] = r * hpE + hpO; arow[k - r] = AK_ICLOSE; }
        carry = bI;
      }
      const uint8_t qraw = qb[r - 1];
      const int qch = lut[qraw];
      const bool hp = r > 1 && qraw == qb[r - 2];          // :180 (raw bytes)
      for (int base = tlo; base <= thi; base += 32) {
        const int t = base + lane;
        const bool act = t <= thi;
        int ms = DBIG, minI = DBIG, minH = DBIG; uint8_t flags = 0;
        if (act) {
          ms = pS[t - 1] + Mtab[qch * 5 + lut[tb[t - 1]]];                            // :248 (row = query)
          const bool inBand = t < r + k;                   // the cell above exists (:169-172,204-211)
          const int up = pS[t];
          const int hOpen = inBand ? up + hpO : DBIG, hExt = (hp && inBand) ? pH[t] + hpE : DBIG;
          if (hOpen < hExt) { flags |= AK_HOPEN; minH = hOpen; } else minH = hExt;   // strict '<': ties extend (:195-203)
          const int iOpen = inBand ? up + inO : DBIG, iExt = inBand ? pI[t] + inE : DBIG;
          if (iOpen < iExt) { flags |= AK_IOPEN; minI = iOpen; } else minI = iExt;   // :213-221
        }
        const int a0 = min(ms, min(minI, minH));
        int bv = act ? a0 - lane * del : DBIG;
        bv = warp_prefix_min(bv, lane);
        const int s = min(bv + lane * del, carry >= DBIG ? DBIG : carry + (lane + 1) * del);
        int left = __shfl_up_sync(0xffffffffu, s, 1);
        if (lane == 0) left = carry;
        const int ds = left >= DBIG ? DBIG : left + del;
        const int best = min(a0, ds);
        // tie order Diagonal > Left > AffineInsClose > AffineHPInsClose (:254-269)
        const uint8_t arrow = best == ms ? AK_DIAG : (best == ds ? AK_LEFT : (best == minI ? AK_ICLOSE : AK_HCLOSE));
        if (act) { cS[t] = s; cH[t] = minH; cI[t] = minI; arow[k + t - r] = arrow | flags; }
        carry = __shfl_sync(0xffffffffu, s, 31);
      }
      __syncwarp();
      int *x;
      x = pS; pS = cS; cS = x; x = pH; pH = cH; cH = x; x = pI; pI = cI; cI = x;
    }
    // ---- end cell: Global corner (:292-295) or the first minimum of the last row (QueryFit :296-311)
    int startC = k - (R - T), score = pS[T];
    if (at == BGPU_QUERYFIT) {
      const int lo = max(1, R - k), hi = min(T, R + k);
      int bv = INT_MAX, bc = INT_MAX;
      for (int t = lo + lane; t <= hi; t += 32) { const int v = pS[t]; if (v < bv) { bv = v; bc = t; } }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const int v = __shfl_xor_sync(0xffffffffu, bv, o), c = __shfl_xor_sync(0xffffffffu, bc, o);
        if (v < bv || (v == bv && c < bc)) { bv = v; bc = c; }
      }
      startC = k - (R - bc); score = bv;
    }
    if (lane == 0) { G.startR = R; G.startC = startC; G.score = score; }
  }
}

// one thread per job: the three-matrix walk of AffineKBandAlign.h:338-395
__global__ void __launch_bounds__(64) affine_kband_trace_kernel(BatchDev B, DenseArgs A, const uint32_t *order, uint32_t nOrder) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nOrder) return;
  const uint32_t job = order[idx];
  JobGeom &G = B.geom[job];
  if (G.status != BGPU_JOB_OK) return;
  const int k = G.band, nCols = 2 * k + 1;
  const uint8_t *arrows = B.arrows + A.arrowOff[job];
  uint32_t *runs = B.runs + G.runOff;
  int q = G.startR, t = G.startC, mat = 0;                  // 0 Match, 1 AffineHPIns, 2 AffineIns
  int runType = -1; uint32_t runLen = 0, nRuns = 0, nBlocks = 0, nGaps = 0, pendGaps = 0;
  bool seenD = false, awry = false;
  auto push = [&](int type) {
    if (type == runType) { runLen++; return; }
    if (runType >= 0) runs[nRuns++] = ((uint32_t)runType << 30) | runLen;
    runType = type; runLen = 1;
    if (type == RUN_D) { if (seenD) nGaps += pendGaps; pendGaps = 0; seenD = true; nBlocks++; }
    else pendGaps++;
  };
  long guard = 4l * (G.Qn + 1) * nCols + 16;
  while (q > 0 || (q == 0 && t > k)) {
    if (t < 0 || t >= nCols || --guard < 0) { awry = true; break; }
    const uint8_t a = arrows[(size_t)q * nCols + t];
    if (mat == 0) {
      if (a & AK_NOARROW) { awry = true; break; }           // the reference would never leave its loop here
      const int m = a & 3;
      if (m == AK_DIAG) { push(RUN_D); q--; }
      else if (m == AK_LEFT) { push(RUN_L); t--; }
      else mat = m == AK_ICLOSE ? 2 : 1;                    // closes change the matrix without moving
    } else {
 if (a & AK_NOARROW) { awry = true; break; }           // reference: assert(0)
      if (a & (mat == 1 ? AK_HOPEN : AK_IOPEN)) mat = 0;
      push(RUN_U); q--; t++;                                // every step inside an affine matrix emits Up (:363-390)
    }
  }
  if (runType >= 0) runs[nRuns++] = ((uint32_t)runType << 30) | runLen;
  if (awry) { G.status = BGPU_JOB_PATH_AWRY; G.nRuns = G.nBlocks = G.nGaps = G.nGapLists = 0; return; }
  G.nRuns = nRuns; G.nBlocks = nBlocks; G.nGaps = seenD ? nGaps + pendGaps : 0; G.nGapLists = nRuns ? nBlocks + 1 : 0;
  G.qPos = G.tPos = 0; G.qStart = G.tStart = 0;             // never set by the reference: the alignment starts at (0,0)
}

void kbounded_host(uint32_t tLength, uint32_t qLength, uint32_t k, uint32_t &tLen, uint32_t &qLen) { kbounded(tLength, qLength, k, tLen, qLen); }

void launch_dense_prep(const BatchDev &B, const ScoreParams &P, const DenseArgs &A, const uint64_t *rowBufOff,
                       const uint64_t *runOff, cudaStream_t s) {
  const unsigned grid = (B.nJobs + 3) / 4;
  if (grid) dense_prep_kernel<<<grid, 128, 0, s>>>(B, P, A, rowBufOff, runOff);
}
void launch_dense_fill(const BatchDev &B, const ScoreParams &P, const DenseArgs &A, const uint32_t *order, uint32_t nOrder,
                       uint32_t *counter, int nSM, cudaStream_t s) {
  unsigned grid = (unsigned)nSM * 8u;
  const unsigned need = (nOrder + 3) / 4;
  if (grid > need) grid = need;
  if (!grid) return;
  if (A.algo == BGPU_AFFINE_KBAND) affine_kband_fill_kernel<<<grid, 128, 0, s>>>(B, P, A, order, nOrder, counter);
  else dense_fill_kernel<<<grid, 128, 0, s>>>(B, P, A, order, nOrder, counter);
}
void launch_dense_trace(const BatchDev &B, const ScoreParams &P, const DenseArgs &A, const uint32_t *order, uint32_t nOrder,
                        cudaStream_t s) {
  const unsigned grid = (nOrder + 31) / 32;
  if (!grid) return;
  if (A.algo == BGPU_AFFINE_KBAND) affine_kband_trace_kernel<<<grid, 32, 0, s>>>(B, A, order, nOrder);
  else dense_trace_kernel<<<grid, 32, 0, s>>>(B, P, A, order, nOrder);
}

}  // namespace bgpu

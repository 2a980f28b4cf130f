// bgpu_trace.cu -- device-side traceback and alignment emission.
//
//   trace_guided_kernel : (one thread per job) walks the traceback words of one job from (qEnd-1,tEnd-1) to the origin with the
//                         reference's 3-matrix state machine (GuidedAlign.h:626-663,
//                         AffineGuidedAlign.h:377-468) and records the path as run-length runs (reversed).
//   scan_counts_kernel  : exclusive scan of per-job block / gap-list / gap counts -> arena offsets.
//   emit_kernel         : turns the runs into Block[] / GapList[] exactly as ArrowPathToAlignment
//                         (datastructures/alignment/Alignment.h:190-254) + RemoveAlignmentPrefixGaps
//                         (AlignmentUtils.h:620-644) do, and evaluates ComputeAlignmentStats
//                         (AlignmentUtils.h:535-584, :60-124) without building the three strings.
#include "bgpu_common.cuh"

namespace bgpu {

enum { RUN_D = 0, RUN_U = 1, RUN_L = 2 };   // diagonal / up (insertion, Gap::Target) / left (deletion, Gap::Query)

// One THREAD per job: a traceback is a serial pointer chase, so a warp walks 32 independent paths at once (jobs are
// ordered longest-first, neighbouring threads have similar path lengths).  The fill kernels store the arrows
// [d-block][row of 16 (linear) / 4 (affine) anti-diagonals][slot pair].  A walk only ever moves to lower
// anti-diagonals and at most one diagonal sideways per step, so the words it is going to read are known well ahead:
// each thread keeps a private window in shared memory -- chunk = 4 rows x 16 words (linear) or 8 rows x 8 words (affine) centred on the
// walk's current diagonal; see TrGeom) -- and copies the chunk BEFORE the current one with cp.async while it walks the
// current one, so a step costs a shared-memory load instead of an L2 round trip.  A walk that drifts out of its window
// (more than ~12 (linear) / ~6 (affine) diagonals sideways within two chunks) falls back to a global load for those steps.
// Linear words hold 16 two-bit arrows of one slot pair: a run of Diagonal arrows (every other field, the slot does
// not change) is consumed with one CLZ instead of one step each.
constexpr int TR_THREADS = 64;                           // threads per CTA
// rows per chunk / 16-byte pieces per row window (512 B of shared memory per walker, two chunks).  Measured on 100k pairs:
// linear (CR, WP) = (4,4) 7.95 ms, (4,2) 8.7, (2,4) 10.8; affine (4,4) 20.5 ms, (4,2) 17.6, (8,2) 15.7 -- chunk changes
// are the expensive part of a walk, and an affine chunk of 4 rows is only 16 anti-diagonals.
template <bool AFFINE> struct TrGeom { static constexpr int CR = AFFINE ? 8 : 4, WP = AFFINE ? 2 : 4; };
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool AFFINE>
__global__ void __launch_bounds__(TR_THREADS) trace_guided_kernel(BatchDev B, const uint32_t *orderBase, const PlanHead *plan, uint32_t spread) {
  constexpr int BITS = AFFINE ? 8 : 2, SPW = 32 / BITS, ROWS = 64 / SPW, CR = TrGeom<AFFINE>::CR, CPB = ROWS / CR, WP = TrGeom<AFFINE>::WP;
  // piece (buffer bi, row r, piece pc) of thread tid: interleaved over the CTA's threads, 16 B each
  __shared__ __align__(16) uint32_t win[2 * CR * WP * TR_THREADS * 4];
  // spread = 1 << k: only every spread-th thread walks (a small ticket has fewer walks than the chip has warps, and walks
  // that share a warp serialise on each other's branches: 64 walks of 10 kb take 6.1 ms packed, 2.8 ms one per warp)
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gtid & (spread - 1)) return;
  const uint32_t idx = gtid / spread;
  if ((plan->overflow & PLAN_OVF_ARROWS) || idx >= plan->traceCount) return;
  const uint32_t *order = orderBase + plan->traceBegin;
  const uint32_t job = order[idx];
  if (job == 0xffffffffu) return;
  JobGeom &G = B.geom[job];
  if (G.status != BGPU_JOB_OK) return;
  const int Qn = G.Qn, Tn = G.Tn, C0 = G.C0;
  const int lpj = cls_lpj(G.cls);
  const int unitW = ROWS * lpj;                          // words per arrow unit (one group of one d-block)
  const DBlock *dblk = B.dblk + G.dblkOff;
  const uint32_t *arrows = reinterpret_cast<const uint32_t *>(B.arrows + B.arrowOff[job]);
  uint32_t *runs = B.runs + G.runOff;
  uint32_t *mine = win + threadIdx.x * 4;
  auto piece = [&](int bi, int r, int pc) { return mine + ((bi * CR + r) * WP + pc) * (TR_THREADS * 4); };

  int q = Qn, t = Tn, mat = 0;
  int runType = -1; uint32_t runLen = 0, nRuns = 0;
  uint32_t nBlocks = 0, nGaps = 0, pendGaps = 0, pendQ = 0, pendT = 0;
  bool seenD = false, awry = false;
  auto push = [&](int type, uint32_t n) {              // (linear walk)
    if (type == runType) { runLen += n; return; }
    if (runType >= 0) runs[nRuns++] = ((uint32_t)runType << 30) | runLen;
    runType = type; runLen = n;
    if (type == RUN_D) { if (seenD) nGaps += pendGaps; pendGaps = 0; pendQ = pendT = 0; seenD = true; nBlocks++; }
    else pendGaps++;
  };
  // window of a chunk: np pieces from piece p0 of every row, centred on the diagonal the walk is on right now
  auto window = [&](const DBlock &db, int &p0, int &np) {
    const int rowPieces = (db.k * lpj) >> 2;
    np = min(WP, rowPieces);
    const int sg = (t - q + C0 - db.wbase) >> 1;
    p0 = max(0, min((sg - (2 * WP - 2)) >> 2, rowPieces - np));   // the walk's word sits in the middle of the window
  };
  auto stage = [&](int bi, int ci, const DBlock &db, int p0, int np) {
    const int rowWords = db.k * lpj;
    const uint32_t *src = arrows + (size_t)db.arrowUnit * unitW + (size_t)((ci % CPB) * CR) * rowWords + p0 * 4;
#pragma unroll
    for (int r = 0; r < CR; r++)
#pragma unroll
      for (int pc = 0; pc < WP; pc++)
        if (pc < np) cp_async16(piece(bi, r, pc), src + (size_t)r * rowWords + pc * 4);
  };

  const int d0 = Qn + Tn;
  int ci = (d0 >> 6) * CPB + ((d0 & 63) / SPW) / CR;     // chunk the walk starts in; it then visits ci-1, ci-2, ...
  DBlock dbW = dblk[ci / CPB];
  int p0W, npW;
  window(dbW, p0W, npW);
  stage(0, ci, dbW, p0W, npW);
  cp_commit();
  DBlock dbS = dblk[max(ci - 1, 0) / CPB];
  int bi = 0;
  for (;;) {
    int p0S = 0, npS = 0;
    if (ci >= 1) { window(dbS, p0S, npS); stage(bi ^ 1, ci - 1, dbS, p0S, npS); }
    cp_commit();
    const DBlock dbS2 = dblk[max(ci - 2, 0) / CPB];      // in flight while this chunk is walked
    cp_wait<1>();
    bool more = false;
    {
      const int bCur = ci / CPB, cCur = ci % CPB;
      const int wbase = dbW.wbase, rowWords = dbW.k * lpj;
      const uint32_t *gsrc = arrows + (size_t)dbW.arrowUnit * unitW;
      while (q >= 1 || t >= 1) {
        if (q < 0 || t < 0) { awry = true; break; }
        const int d = q + t, e = d & 63, row = e / SPW;
        if ((d >> 6) != bCur || row / CR != cCur) { more = true; break; }     // the walk has left this chunk
        const int s = t - q + C0 - wbase;
        if (s < 0 || (s >> 1) >= rowWords) { awry = true; break; }
        const int wo = (s >> 1) - p0W * 4;
        uint32_t word;
        if ((unsigned)wo < (unsigned)(npW * 4)) word = piece(bi, row - cCur * CR, wo >> 2)[wo & 3];
        else word = __ldg(gsrc + (size_t)row * rowWords + (s >> 1));
        const int pos = e % SPW;                                   // the first step of a word sits in its lowest field
        const uint32_t f = (word >> (BITS * pos)) & ((1u << BITS) - 1u);
        if (!AFFINE) {
          // (measured: the select form below costs the linear walk 3 %: its runs of Diagonal arrows are consumed a word at a
          // time, so the walkers of a warp branch far less often than the affine ones)
          if (f == TL_DIAG) {
            // earlier anti-diagonals of this slot sit in fields pos-2, pos-4, ...: take the whole run of Diagonal
            // arrows inside the word at once
            uint32_t x = word & (0x33333333u << (2 * (pos & 1)));
            x &= 0xffffffffu >> (30 - 2 * pos);
            int n = x == 0 ? (pos >> 1) + 1 : (pos - ((31 - __clz((int)x)) >> 1)) >> 1;
            n = min(n, min(q, t));
            if (n <= 0) { awry = true; break; }
            push(RUN_D, (uint32_t)n); q -= n; t -= n;
          }
          else if (f == TL_UP) { push(RUN_U, 1); pendQ++; q--; }
          else if (f == TL_LEFT) { push(RUN_L, 1); pendT++; t--; }
          else { awry = true; break; }
        } else {
          // The arrow is decoded into (run type, next matrix) with selects and two packed look-up words, and ONE copy of the
          // update code follows: the walkers of a warp sit on different arrows at every step, and a branch per arrow kind (the
          // reference's switch, AffineGuidedAlign.h:377-468) would run every kind's code one after the other for the whole
          // warp (40k walks of 10 kb: 7.2 -> 5.05 ms).
          const uint32_t tag = f & 7u;
          const bool m0 = mat == 0;
          if (tag == TB_NONE || (m0 && tag > TB_DCLOSE)) { awry = true; break; }
          const uint32_t open = (f >> (2 + mat)) & 1u;                           // mat 1: TB_IOPEN (bit 3), mat 2: TB_DOPEN (bit 4)
          // Match matrix: Diagonal -> D, Left / AffineDelClose -> L, Up / AffineInsClose -> U; the closes enter matrix 2 / 1.
          // Affine matrices: an open flag returns to the match matrix without a move, else one more U (matrix 1) / L (matrix 2)
          const uint32_t type = m0 ? (0x258u >> (2 * tag)) & 3u : (open ? 3u : (uint32_t)mat);
          const int newmat = m0 ? (int)((0x240u >> (2 * tag)) & 3u) : (open ? 0 : mat);
          if (type != 3u) {
            if ((int)type != runType) {
              if (runType >= 0) runs[nRuns++] = ((uint32_t)runType << 30) | runLen;
              runType = (int)type; runLen = 0;
              const bool isD = type == RUN_D;
              nGaps += (isD && seenD) ? pendGaps : 0u;
              pendGaps = isD ? 0u : pendGaps + 1u;
              pendQ = isD ? 0u : pendQ; pendT = isD ? 0u : pendT;
              nBlocks += isD ? 1u : 0u; seenD = seenD || isD;
            }
            runLen += 1u;
            const bool isU = type == RUN_U, isL = type == RUN_L;
            q -= isL ? 0 : 1; t -= isU ? 0 : 1;
            pendQ += isU ? 1u : 0u; pendT += isL ? 1u : 0u;
          }
          mat = newmat;
        }
      }
    }
    if (!more) break;
    if (ci == 0) { awry = true; break; }
    ci--; bi ^= 1; dbW = dbS; p0W = p0S; npW = npS; dbS = dbS2;
  }
  cp_wait<0>();
  if (runType >= 0) runs[nRuns++] = ((uint32_t)runType << 30) | runLen;
  if (awry) { G.status = BGPU_JOB_PATH_AWRY; G.nRuns = 0; G.nBlocks = G.nGaps = G.nGapLists = 0; }
  else {
    G.nRuns = nRuns; G.nBlocks = nBlocks; G.nGaps = nGaps; G.nGapLists = nRuns ? nBlocks + 1 : 0;
    // leading gaps fold into qPos/tPos only when a block follows (the all-gap path is cleared first)
    G.qPos = (uint32_t)G.qStart + (seenD ? pendQ : 0);
    G.tPos = (uint32_t)G.tStart + (seenD ? pendT : 0);
  }
}

// one CTA; counts -> exclusive offsets, totals in totals[0..2]
__global__ void __launch_bounds__(1024) scan_counts_kernel(BatchDev B, uint64_t *blockOff, uint64_t *listOff,
                                                           uint64_t *gapOff, uint64_t *totals) {
  __shared__ unsigned long long sh[3][32];
  __shared__ unsigned long long carry[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 3) carry[tid] = 0;
  __syncthreads();
  for (uint32_t base = 0; base < B.nJobs; base += 1024) {
    const uint32_t j = base + tid;
    unsigned long long v[3] = {0, 0, 0};
    if (j < B.nJobs) { const JobGeom &G = B.geom[j]; v[0] = G.nBlocks; v[1] = G.nGapLists; v[2] = G.nGaps; }
    unsigned long long inc[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      unsigned long long x = v[c];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { unsigned long long u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
      inc[c] = x;
      if (lane == 31) sh[c][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int c = 0; c < 3; c++) {
        unsigned long long x = sh[c][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
        sh[c][lane] = x;
      }
    }
    __syncthreads();
    unsigned long long off[3];
#pragma unroll
    for (int c = 0; c < 3; c++) off[c] = carry[c] + (warp ? sh[c][warp - 1] : 0) + inc[c] - v[c];
    if (j < B.nJobs) { blockOff[j] = off[0]; listOff[j] = off[1]; gapOff[j] = off[2]; }
    __syncthreads();
    if (tid < 3) carry[tid] += sh[tid][31];
    __syncthreads();
  }
  if (tid < 3) totals[tid] = carry[tid];
}

struct EmitOut {
  bgpu_result *results;
  bgpu_block *blocks; uint32_t *gapCounts; bgpu_gap *gaps;
  const uint64_t *blockOff, *listOff, *gapOff;
  uint32_t *runsOut;     // compact results: the kept runs in path order, job i from runsOut[blockOff[i] + gapOff[i]]; else NULL
  int32_t *rescoreOut;   // bgpu_rescore: only ComputeAlignmentScore(alignment, query, text, fn, affine) per job goes out; else NULL
};

__device__ __forceinline__ uint32_t warp_excl_u32(uint32_t v, int lane, uint32_t &total) {
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
  total = __shfl_sync(0xffffffffu, x, 31);
  return x - v;
}

// warp per job.  Every gapCounts entry of the job is written here (no zero-fill, no atomics): the lane holding block d
// writes the number of kept gap runs between blocks d-1 and d, lane 0 the (always empty) list after the last block.
// The per-base part of ComputeAlignmentStats is spread over the lanes by query position, not by run: the chunk's 32
// run starts go to shared memory and every lane finds the run its position falls into with a 5-step search, so a
// warp compares 32 consecutive query bases per step whatever the run lengths are.
constexpr int EMIT_WARPS = 4;
__global__ void __launch_bounds__(EMIT_WARPS * 32, 8) emit_kernel(BatchDev B, ScoreParams P, EmitOut O, int doStats, int statsAffine,
                                                   int keepLeading, PlanHead *plan) {
  __shared__ uint8_t lut[256];
  __shared__ int sM[64];
  __shared__ uint32_t sQ[EMIT_WARPS][32];
  __shared__ int sDelta[EMIT_WARPS][32];
  if (plan) {   // asynchronous path: the arena was sized before the counts were known; the host re-emits when it is too small
    if (plan->overflow & PLAN_OVF_ARROWS) return;
    if (plan->totals[0] > plan->caps[0] || plan->totals[1] > plan->caps[1] || plan->totals[2] > plan->caps[2]) {
      if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&plan->overflow, (uint32_t)PLAN_OVF_ARENA);
      return;
    }
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = base_code((uint8_t)i);
  if (threadIdx.x < 64) { const int r = threadIdx.x >> 3, c = threadIdx.x & 7; sM[threadIdx.x] = (r < 5 && c < 5) ? P.M[r * 5 + c] : 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t job = blockIdx.x * (blockDim.x >> 5) + warp;
  if (job >= B.nJobs) return;
  const JobGeom &G = B.geom[job];
  bgpu_result R;
  R.status = G.status; R.score = 0; R.qPos = 0; R.tPos = 0; R.nCells = 0;
  R.nMatch = R.nMismatch = R.nIns = R.nDel = 0; R.pctSimilarity = 0.f; R.statsScore = 0;
  R.nBlocks = 0; R.nGapLists = 0; R.nGaps = 0;
  R.blockOff = O.blockOff[job]; R.gapListOff = O.listOff[job]; R.gapOff = O.gapOff[job];
  const bool rescore = O.rescoreOut != nullptr;   // the Alignment overload of ComputeAlignmentScore (AlignmentUtils.h:127-169): blocks + one cost per Gap
  if (G.status != BGPU_JOB_OK) { if (lane == 0) { if (rescore) O.rescoreOut[job] = 0; else O.results[job] = R; } return; }
  R.score = G.score; R.qPos = G.qPos; R.tPos = G.tPos; R.nCells = G.nCells;
  R.nBlocks = G.nBlocks; R.nGapLists = G.nGapLists; R.nGaps = G.nGaps;
  const uint32_t nRuns = G.nRuns, nBlocks = G.nBlocks;
  const uint32_t *runs = B.runs + G.runOff;
  const long long qLenJ = (long long)(B.qOff[job + 1] - B.qOff[job]), tLenJ = (long long)(B.tOff[job + 1] - B.tOff[job]);
  const uint8_t *qb = B.q + B.qOff[job] + G.qStart;  // qb[x]: query base at path offset x
  const uint8_t *tb = B.t + B.tOff[job] + G.tStart;  // raw bytes, mapped through the same table as the query's
  int oob = 0;
  const uint32_t qPrefix = G.qPos - (uint32_t)G.qStart, tPrefix = G.tPos - (uint32_t)G.tStart;
  const bool compact = O.runsOut != nullptr;
  bgpu_block *blocks = (compact || rescore) ? nullptr : O.blocks + R.blockOff;
  uint32_t *gapCounts = (compact || rescore) ? nullptr : O.gapCounts + R.gapListOff;
  bgpu_gap *gaps = (compact || rescore) ? nullptr : O.gaps + R.gapOff;
  uint32_t *runsOut = (compact && !rescore) ? O.runsOut + R.blockOff + R.gapOff : nullptr;

  uint32_t cq = 0, ct = 0, cD = 0, cG = 0;            // carries: q/t consumed, D runs seen, kept gap runs seen
  uint32_t gAtPrevD = 0;                              // kept gap runs before the latest block seen so far
  int nMatch = 0, nMismatch = 0, nIns = 0, nDel = 0, score = 0; long long cols = 0;
  uint32_t runNext = (uint32_t)lane < nRuns ? runs[nRuns - 1 - lane] : 0;
  for (uint32_t base = 0; base < nRuns; base += 32) {
    const uint32_t f = base + lane;                   // forward run index
    const bool act = f < nRuns;
    uint32_t type = 3, len = 0;
    if (act) { const uint32_t r = runNext; type = r >> 30; len = r & 0x3fffffffu; }
    if (f + 32 < nRuns) runNext = runs[nRuns - 33 - f];   // the next chunk's run, in flight during this chunk
    const uint32_t dq = (type == RUN_D || type == RUN_U) ? len : 0, dt = (type == RUN_D || type == RUN_L) ? len : 0;
    uint32_t totQ, totT, totD, totG;
    const uint32_t pq = cq + warp_excl_u32(dq, lane, totQ), pt = ct + warp_excl_u32(dt, lane, totT);
    const bool isD = act && type == RUN_D;
    const uint32_t dBefore = cD + warp_excl_u32(isD ? 1u : 0u, lane, totD);
    // gap runs after the last block are dropped; the ones before the first block are folded into qPos/tPos by the
    // guided aligners (RemoveAlignmentPrefixGaps) and kept as gaps[0] by KBandAlign / SWAlign
    const bool kept = act && !isD && (keepLeading || dBefore >= 1) && dBefore < nBlocks;
    const uint32_t gBefore = cG + warp_excl_u32(kept ? 1u : 0u, lane, totG);
    // kept gap runs before the previous block: from the nearest lower lane holding a block, else the carry
    const uint32_t dMask = __ballot_sync(0xffffffffu, isD);
    const uint32_t below = dMask & ((1u << lane) - 1u);
    const uint32_t gPrevLane = __shfl_sync(0xffffffffu, gBefore, below ? 31 - __clz((int)below) : lane);
    bool cmp = false;                                 // this lane's block takes part in the per-base pass
    if (isD) {
      if (rescore) {}
      else if (compact) runsOut[dBefore + gBefore] = len;                 // RUN_D << 30 | len
      else {
        bgpu_block bl; bl.qPos = pq - qPrefix; bl.tPos = pt - tPrefix; bl.length = len;
        blocks[dBefore] = bl;
        gapCounts[dBefore] = gBefore - (below ? gPrevLane : gAtPrevD);
      }
      if (doStats) {
        const long long q0 = (long long)G.qStart + pq, t0 = (long long)G.tStart + pt;
        // KBandAlign can leave qPos/tPos pointing outside the sequences (KBandAlign.h:394-399); the reference then reads
        // out of bounds, here the job is flagged instead
        if (q0 < 0 || t0 < 0 || q0 + len > qLenJ || t0 + len > tLenJ) oob = 1; else cmp = true;
        cols += len;
      }
    } else if (kept && rescore) {
      // :141-160: every Gap of the lists between blocks on its own; the leading list (gaps[0]) is not scored
      if (dBefore >= 1) {
        const int lin = (int)len * (type == RUN_L ? P.del : P.ins), aff = P.open + (int)len * P.ext;
        score += (statsAffine && aff < lin) ? aff : lin;
      }
    } else if (kept) {
      if (compact) runsOut[dBefore + gBefore] = (type << 30) | len;       // 1 = Gap::Target (insertion), 2 = Gap::Query (deletion)
      else { bgpu_gap g; g.seq = (type == RUN_L) ? 0 : 1; g.length = (int32_t)len; gaps[gBefore] = g; }
      if (doStats) {
        if (type == RUN_L) nDel += (int)len; else nIns += (int)len;
        cols += len;
        if (!statsAffine) score += (int)len * (type == RUN_L ? P.del : P.ins);   // :111-116
        else {
          // affine: one cost per maximal run of gap columns, typed by its last column (:81-100);
          // the lane holding the last run of the group sums the group.
          const uint32_t nxt = runs[nRuns - 2 - f] >> 30;   // kept => a block follows eventually, f+1 < nRuns
          if (nxt == RUN_D) {
            long long L = len;
            for (uint32_t b = f; b-- > 0;) {
              const uint32_t r = runs[nRuns - 1 - b];
              if ((r >> 30) == RUN_D) break;
              L += r & 0x3fffffffu;
            }
            const int aff = (int)L * P.ext + P.open;
            const int lin = (int)L * (type == RUN_L ? P.del : P.ins);
            score += lin < aff ? lin : aff;
          }
        }
      }
    }
    if (dMask) gAtPrevD = __shfl_sync(0xffffffffu, gBefore, 31 - __clz((int)dMask));
    if (doStats) {
      // ---- per-base pass over the query positions [cq, cq + totQ) of this chunk (ComputeAlignmentStats :546-560,
      //      ComputeAlignmentScore :70-76): position x lies in the last run starting at or before x (runs that take no
      //      query base share their start with the run after them, which is the one found)
      const uint32_t cmpMask = __ballot_sync(0xffffffffu, cmp);
      if (cmpMask) {
        sQ[warp][lane] = pq; sDelta[warp][lane] = (int)(pt - pq);   // lanes past the last run hold pq = cq + totQ
        __syncwarp();
        const uint32_t xEnd = cq + totQ;
        const uint32_t *sq = sQ[warp]; const int *sd = sDelta[warp];
        auto find = [&](uint32_t x) { int r = 0;
#pragma unroll
          for (int st = 16; st; st >>= 1) if (sq[r + st] <= x) r += st;
          return r; };
        auto tally = [&](int qc, int tc) { if (qc == tc) nMatch++; else nMismatch++; score += sM[qc * 8 + tc]; };   // ComputeAlignmentScore :74 (row = query)
        // two positions per step: both searches, then all four byte loads, are in flight together
        for (uint32_t x = cq + lane; x < xEnd; x += 64) {
          const uint32_t y = x + 32; const bool hasY = y < xEnd;
          const int r0 = find(x), r1 = find(hasY ? y : x);
          const bool c0 = (cmpMask >> r0) & 1u, c1 = hasY && ((cmpMask >> r1) & 1u);
          uint8_t q0b = 0, t0b = 0, q1b = 0, t1b = 0;
          if (c0) { q0b = qb[x]; t0b = tb[(long long)x + sd[r0]]; }
          if (c1) { q1b = qb[y]; t1b = tb[(long long)y + sd[r1]]; }
          if (c0) tally(lut[q0b] & 7, lut[t0b] & 7);
          if (c1) tally(lut[q1b] & 7, lut[t1b] & 7);
        }
        __syncwarp();
      }
    }
    cq += totQ; ct += totT; cD += totD; cG += totG;
  }
  if (lane == 0 && R.nGapLists && !compact && !rescore) gapCounts[nBlocks] = 0;   // the list after the last block: its gap runs are dropped
  if (doStats) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      nMatch += __shfl_xor_sync(0xffffffffu, nMatch, o); nMismatch += __shfl_xor_sync(0xffffffffu, nMismatch, o);
      nIns += __shfl_xor_sync(0xffffffffu, nIns, o); nDel += __shfl_xor_sync(0xffffffffu, nDel, o);
      score += __shfl_xor_sync(0xffffffffu, score, o); cols += __shfl_xor_sync(0xffffffffu, cols, o);
    }
    oob = __reduce_or_sync(0xffffffffu, (unsigned)oob);
    if (rescore) { if (lane == 0) O.rescoreOut[job] = oob ? 0 : score; return; }
    if (oob) { nMatch = nMismatch = nIns = nDel = score = 0; cols = 0; R.status = BGPU_JOB_REF_UNDEFINED; }
    R.nMatch = nMatch; R.nMismatch = nMismatch; R.nIns = nIns; R.nDel = nDel; R.statsScore = score;
    R.pctSimilarity = cols > 0 ? (float)((nMatch * 2.0) / (double)(2 * cols) * 100) : 0.f;   // :566-576
  }
  if (lane == 0) O.results[job] = R;
}

// the traceback list is plan->traceCount jobs from order[plan->traceBegin]; nOrder (a host-side upper bound) sizes the grid
void launch_trace_guided(const BatchDev &B, bool affine, const uint32_t *order, const PlanHead *plan, uint32_t nOrder, cudaStream_t s) {
  const unsigned block = TR_THREADS;   // small CTAs spread the (few, long) walks over all SMs
  if (!nOrder) return;
  // one walk per warp while the walks fit the chip that way (about 7 resident CTAs of 64 threads per SM), else 2, 4, ... 32 per warp
  static int nSM = 0;
  if (!nSM) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, dev); }
  unsigned spread = 32;
  while (spread > 1 && (unsigned long long)nOrder * spread > (unsigned long long)nSM * 7 * TR_THREADS) spread >>= 1;
  const unsigned grid = (unsigned)(((unsigned long long)nOrder * spread + block - 1) / block);
  if (affine) trace_guided_kernel<true><<<grid, block, 0, s>>>(B, order, plan, spread);
  else trace_guided_kernel<false><<<grid, block, 0, s>>>(B, order, plan, spread);
}

void launch_scan_counts(const BatchDev &B, uint64_t *blockOff, uint64_t *listOff, uint64_t *gapOff, uint64_t *totals,
                        cudaStream_t s) {
  scan_counts_kernel<<<1, 1024, 0, s>>>(B, blockOff, listOff, gapOff, totals);
}

void launch_emit(const BatchDev &B, const ScoreParams &P, bgpu_result *results, bgpu_block *blocks,
                 uint32_t *gapCounts, bgpu_gap *gaps, const uint64_t *blockOff, const uint64_t *listOff,
                 const uint64_t *gapOff, int doStats, int statsAffine, int keepLeading, PlanHead *plan, uint32_t *runsOut,
                 cudaStream_t s) {
  EmitOut O{results, blocks, gapCounts, gaps, blockOff, listOff, gapOff, runsOut, nullptr};
  const unsigned grid = (B.nJobs + EMIT_WARPS - 1) / EMIT_WARPS;
  if (grid) emit_kernel<<<grid, EMIT_WARPS * 32, 0, s>>>(B, P, O, doStats, statsAffine, keepLeading, plan);
}

// bgpu_rescore: the emit pass again, under another score function, with only the score going out
void launch_rescore(const BatchDev &B, const ScoreParams &P, const uint64_t *blockOff, const uint64_t *listOff, const uint64_t *gapOff,
                    int useAffine, int32_t *out, cudaStream_t s) {
  EmitOut O{nullptr, nullptr, nullptr, nullptr, blockOff, listOff, gapOff, nullptr, out};
  const unsigned grid = (B.nJobs + EMIT_WARPS - 1) / EMIT_WARPS;
  if (grid) emit_kernel<<<grid, EMIT_WARPS * 32, 0, s>>>(B, P, O, 1, useAffine, 0, nullptr);
}

}  // namespace bgpu

// bgpu_plan.cu -- the schedule of a guided ticket, built on the device so that bgpu_submit never waits for the GPU.
//
// After prep every job has its geometry (class = lanes per job, d-blocks, typical window width).  The fill kernels want
//   * per class, jobs of similar width and length side by side (a warp sweeps 32 / LPJ jobs in lockstep),
//   * warp groups dispatched most expensive first (the work queue's tail is then made of short jobs),
//   * a traceback offset per job, and a longest-first traceback list.
// Round 1 read the geometry back and did this on the host (a blocking copy + O(n log n) sorts inside bgpu_submit).  Here the
// same ordering comes out of stable LSD radix sorts on packed keys -- ties keep job-index order, which is what the host's
// (key, index) pair sorts produced, so both planners emit the SAME order -- plus three small layout kernels and a scan.
#include "bgpu_common.cuh"

namespace bgpu {

constexpr uint32_t NOJOB_P = 0xffffffffu;
constexpr int SORT_TILE = 2048;

// ---- stable radix sort of (key, value) pairs, 8 bits per pass; the element count may live on the device ----
__global__ void __launch_bounds__(256) sort_hist_kernel(const uint32_t *keys, const uint32_t *nPtr, uint32_t nHost, int shift,
                                                        uint32_t *hist, uint32_t nTiles) {
  __shared__ uint32_t h[256];
  const uint32_t n = nPtr ? *nPtr : nHost, tile = blockIdx.x;
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lo = tile * SORT_TILE, hi = min(n, lo + SORT_TILE);
  for (uint32_t i = lo + threadIdx.x; i < hi; i += 256) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  __syncthreads();
  hist[threadIdx.x * nTiles + tile] = h[threadIdx.x];     // digit-major: one exclusive scan gives every (digit, tile) base
}

__global__ void __launch_bounds__(1024) sort_scan_kernel(uint32_t *hist, uint32_t total) {
  __shared__ uint32_t sh[32];
  __shared__ uint32_t carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < total; base += 1024) {
    const uint32_t i = base + tid;
    const uint32_t v = i < total ? hist[i] : 0;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
    if (lane == 31) sh[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t y = sh[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, y, o); if (lane >= o) y += u; }
      sh[lane] = y;
    }
    __syncthreads();
    if (i < total) hist[i] = carry + (warp ? sh[warp - 1] : 0) + x - v;
    __syncthreads();
    if (tid == 0) carry += sh[31];
    __syncthreads();
  }
}

// one warp per tile: chunks of 32 elements in order, rank inside a chunk by lane -> stable
__device__ __forceinline__ void scatter_chunk(const uint32_t key, const uint32_t val, const bool act, const int shift, const int lane,
                                              uint32_t *base, uint32_t *keysOut, uint32_t *valsOut) {
  const uint32_t d = (key >> shift) & 255u;
  const unsigned am = __ballot_sync(0xffffffffu, act);
  unsigned peers = __match_any_sync(0xffffffffu, act ? d : 256u + (uint32_t)lane) & am;
  const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
  uint32_t pos = 0;
  if (act) { pos = base[d] + rank; keysOut[pos] = key; valsOut[pos] = val; }
  __syncwarp();
  if (act && rank == (uint32_t)__popc(peers) - 1u) base[d] += (uint32_t)__popc(peers);
  __syncwarp();
}

__global__ void __launch_bounds__(32) sort_scatter_kernel(const uint32_t *keysIn, const uint32_t *valsIn, uint32_t *keysOut,
                                                          uint32_t *valsOut, const uint32_t *nPtr, uint32_t nHost, int shift,
                                                          const uint32_t *hist, uint32_t nTiles) {
  __shared__ uint32_t base[256];
  const int lane = threadIdx.x;
  const uint32_t n = nPtr ? *nPtr : nHost, tile = blockIdx.x;
  for (int b = lane; b < 256; b += 32) base[b] = hist[b * nTiles + tile];
  __syncwarp();
  const uint32_t lo = tile * SORT_TILE, hi = min(n, lo + SORT_TILE);
  for (uint32_t i0 = lo; i0 < hi; i0 += 32) {
    const uint32_t i = i0 + lane;
    const bool act = i < hi;
    scatter_chunk(act ? keysIn[i] : 0u, act ? valsIn[i] : 0u, act, shift, lane, base, keysOut, valsOut);
  }
}

// tickets of up to SORT_TILE elements (blasr-sized batches): every pass inside one warp, ping-pong in shared memory
__global__ void __launch_bounds__(32) sort_small_kernel(uint32_t *keys, uint32_t *vals, const uint32_t *nPtr, uint32_t nHost, int passes) {
  __shared__ uint32_t k[2][SORT_TILE], v[2][SORT_TILE];
  __shared__ uint32_t base[256];
  const int lane = threadIdx.x;
  const uint32_t n = min(nPtr ? *nPtr : nHost, (uint32_t)SORT_TILE);
  for (uint32_t i = lane; i < n; i += 32) { k[0][i] = keys[i]; v[0][i] = vals[i]; }
  __syncwarp();
  int cur = 0;
  for (int p = 0; p < passes; p++) {
    const int shift = 8 * p;
    for (int b = lane; b < 256; b += 32) base[b] = 0;
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) atomicAdd(&base[(k[cur][i] >> shift) & 255u], 1u);
    __syncwarp();
    {   // exclusive scan of the 256 counters: 8 per lane
      uint32_t c[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) { c[j] = base[lane * 8 + j]; sum += c[j]; }
      uint32_t x = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
      uint32_t run = x - sum;
#pragma unroll
      for (int j = 0; j < 8; j++) { base[lane * 8 + j] = run; run += c[j]; }
    }
    __syncwarp();
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
      const uint32_t i = i0 + lane;
      const bool act = i < n;
      scatter_chunk(act ? k[cur][i] : 0u, act ? v[cur][i] : 0u, act, shift, lane, base, k[cur ^ 1], v[cur ^ 1]);
    }
    cur ^= 1;
    __syncwarp();
  }
  for (uint32_t i = lane; i < n; i += 32) { keys[i] = k[cur][i]; vals[i] = v[cur][i]; }
}

struct SortBufs { uint32_t *keyA, *valA, *keyB, *valB, *hist; };

// sorts (keyA, valA) by the low 8 * passes bits of the key; the result is left in (keyA, valA).  nMax bounds the count.
static void radix_sort(const SortBufs &sb, const uint32_t *nPtr, uint32_t nMax, int passes, cudaStream_t s) {
  if (nMax == 0) return;
  if (nMax <= (uint32_t)SORT_TILE) { sort_small_kernel<<<1, 32, 0, s>>>(sb.keyA, sb.valA, nPtr, nMax, passes); return; }
  const uint32_t nTiles = (nMax + SORT_TILE - 1) / SORT_TILE;
  uint32_t *ki = sb.keyA, *vi = sb.valA, *ko = sb.keyB, *vo = sb.valB;
  for (int p = 0; p < passes; p++) {
    sort_hist_kernel<<<nTiles, 256, 0, s>>>(ki, nPtr, nMax, 8 * p, sb.hist, nTiles);
    sort_scan_kernel<<<1, 1024, 0, s>>>(sb.hist, 256u * nTiles);
    sort_scatter_kernel<<<nTiles, 32, 0, s>>>(ki, vi, ko, vo, nPtr, nMax, 8 * p, sb.hist, nTiles);
    std::swap(ki, ko); std::swap(vi, vo);
  }
  if (passes & 1) {   // odd pass count: the result sits in the B buffers
    cudaMemcpyAsync(sb.keyA, sb.keyB, sizeof(uint32_t) * nMax, cudaMemcpyDeviceToDevice, s);
    cudaMemcpyAsync(sb.valA, sb.valB, sizeof(uint32_t) * nMax, cudaMemcpyDeviceToDevice, s);
  }
}

// ---- planning kernels ----
struct PlanBufs {
  SortBufs jobs, groups, trace;   // (key, value) ping-pong buffers: jobs / trace hold nJobs pairs, groups nJobs + N_CLS
  uint32_t *grpFirst, *grpCount;  // per group id: first position in the sorted job list, members
  uint64_t *bound;                // per job: traceback bytes reserved
  uint64_t *slotBytes;            // per order slot
  uint32_t *jobStart, *grpStart;  // per class (+1): first sorted position / first group id
};

__global__ void __launch_bounds__(256) plan_keys_kernel(BatchDev B, PlanHead *plan, PlanBufs pb) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  unsigned long long cells = 0;
  if (i < B.nJobs) {
    const JobGeom &G = B.geom[i];
    uint32_t key = 0xffffffffu, tkey = 0xffffu;
    if (G.status == BGPU_JOB_OK) {
      const uint32_t nDB = (uint32_t)min(G.nDB, 32767);
      const uint32_t kt = (uint32_t)min((G.ksum + G.nDB / 2) / max(G.nDB, 1), 255);     // typical window width, in groups
      key = ((uint32_t)G.cls << 23) | ((255u - kt) << 15) | (32767u - nDB);              // class, widest first, longest first
      tkey = 32767u - nDB;
      atomicAdd(&plan->clsCount[G.cls], 1u);
      cells = (unsigned long long)G.nCells;
    }
    pb.jobs.keyA[i] = key; pb.jobs.valA[i] = i;
    pb.trace.keyA[i] = tkey; pb.trace.valA[i] = i;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
  if (lane == 0 && cells) atomicAdd(&plan->cells, cells);
}

__global__ void plan_layout_kernel(PlanHead *plan, PlanBufs pb) {
  if (threadIdx.x != 0) return;
  uint32_t js = 0, gs = 0, os = 0;
  for (int c = 0; c < N_CLS; c++) {
    const uint32_t nj = 32u / (uint32_t)cls_lpj(c), cnt = plan->clsCount[c], g = (cnt + nj - 1) / nj;
    pb.jobStart[c] = js; pb.grpStart[c] = gs;
    plan->nGroups[c] = g; plan->orderBegin[c] = os;
    js += cnt; gs += g; os += g * nj;
  }
  pb.jobStart[N_CLS] = js; pb.grpStart[N_CLS] = gs;
  plan->nOk = js; plan->nGroupsTotal = gs; plan->nSlots = os;
  plan->traceBegin = os; plan->traceCount = js;
}

__global__ void __launch_bounds__(128) plan_groups_kernel(BatchDev B, PlanHead *plan, PlanBufs pb, int affine) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= plan->nGroupsTotal) return;
  int c = 0;
  while (c + 1 < N_CLS && g >= pb.grpStart[c + 1]) c++;
  const uint32_t lpj = (uint32_t)cls_lpj(c), nj = 32u / lpj;
  const uint32_t first = pb.jobStart[c] + (g - pb.grpStart[c]) * nj;
  const uint32_t count = min(nj, pb.jobStart[c + 1] - first);
  int kmaxG = 1, nDBmax = 0; unsigned long long ksumMax = 0;
  for (uint32_t j = 0; j < count; j++) {
    const JobGeom &G = B.geom[pb.jobs.valA[first + j]];
    kmaxG = max(kmaxG, G.kmax); nDBmax = max(nDBmax, G.nDB); ksumMax = max(ksumMax, (unsigned long long)G.ksum);
  }
  const unsigned long long rowsPerBlock = affine ? 16 : 4;        // 64 anti-diagonals / steps per traceback word
  for (uint32_t j = 0; j < count; j++) {
    const uint32_t job = pb.jobs.valA[first + j];
    const unsigned long long bb = (unsigned long long)B.geom[job].nDB * (unsigned long long)kmaxG * rowsPerBlock * lpj * 4ull;
    pb.bound[job] = (bb + 255ull) & ~255ull;
  }
  const unsigned long long cost = max(ksumMax, (unsigned long long)nDBmax);
  pb.grpFirst[g] = first; pb.grpCount[g] = count;
  pb.groups.keyA[g] = ((uint32_t)c << 23) | (0x7fffffu - (uint32_t)min(cost, 0x7fffffull));   // class, most expensive first
  pb.groups.valA[g] = g;
  atomicAdd(&plan->laneSteps, ksumMax * 2048ull);                  // 64 steps x 32 lanes per group and d-block
}

__global__ void __launch_bounds__(128) plan_order_kernel(PlanHead *plan, PlanBufs pb, uint32_t *order) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;      // position in the dispatch order
  if (p >= plan->nGroupsTotal) return;
  const uint32_t g = pb.groups.valA[p];
  const int c = (int)(pb.groups.keyA[p] >> 23);
  const uint32_t nj = 32u / (uint32_t)cls_lpj(c);
  const uint32_t slot0 = plan->orderBegin[c] + (p - pb.grpStart[c]) * nj;
  const uint32_t first = pb.grpFirst[g], count = pb.grpCount[g];
  for (uint32_t j = 0; j < nj; j++) {
    const uint32_t job = j < count ? pb.jobs.valA[first + j] : NOJOB_P;
    order[slot0 + j] = job;
    pb.slotBytes[slot0 + j] = job != NOJOB_P ? pb.bound[job] : 0ull;
  }
}

// exclusive scan of the slot sizes -> per-job traceback offsets; one CTA
__global__ void __launch_bounds__(1024) plan_offsets_kernel(PlanHead *plan, PlanBufs pb, const uint32_t *order, uint64_t *arrowOff,
                                                            unsigned long long poolBytes) {
  __shared__ unsigned long long sh[32];
  __shared__ unsigned long long carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  const uint32_t n = plan->nSlots;
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + tid;
    const unsigned long long v = i < n ? pb.slotBytes[i] : 0ull;
    unsigned long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
    if (lane == 31) sh[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned long long y = sh[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, y, o); if (lane >= o) y += u; }
      sh[lane] = y;
    }
    __syncthreads();
    if (i < n && order[i] != NOJOB_P) arrowOff[order[i]] = carry + (warp ? sh[warp - 1] : 0ull) + x - v;
    __syncthreads();
    if (tid == 0) carry += sh[31];
    __syncthreads();
  }
  if (tid == 0) { plan->arrowBytes = carry; if (carry > poolBytes) plan->overflow |= PLAN_OVF_ARROWS; }
}

__global__ void __launch_bounds__(256) plan_trace_kernel(const PlanHead *plan, PlanBufs pb, uint32_t *order) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < plan->traceCount) order[plan->traceBegin + i] = pb.trace.valA[i];
}

// Scratch the planner needs, in uint32 units, for a ticket of n jobs (carved out of one allocation by the caller).
size_t plan_scratch_words(uint32_t n) {
  const size_t m = (size_t)n + N_CLS + 8, tiles = (m + SORT_TILE - 1) / SORT_TILE;
  return 4 * m * 3 + 256 * tiles * 3 + 2 * m + 2 * m /* bound */ + 2 * (2 * m + 64) /* slotBytes: <= 2 slots per job */ + 64;
}

// order must hold nSlots + nOk entries: at most 2 * n + 32 * N_CLS.
void launch_plan_guided(const BatchDev &B, bool affine, PlanHead *plan, uint32_t *scratch, uint32_t *order, uint64_t *arrowOff,
                        unsigned long long poolBytes, cudaStream_t s) {
  const uint32_t n = B.nJobs;
  if (!n) return;
  const size_t m = (size_t)n + N_CLS + 8, tiles = (m + SORT_TILE - 1) / SORT_TILE;
  PlanBufs pb;
  uint32_t *w = scratch;
  auto take = [&](size_t words) { uint32_t *p = w; w += words; return p; };
  for (SortBufs *sb : {&pb.jobs, &pb.groups, &pb.trace}) {
    sb->keyA = take(m); sb->valA = take(m); sb->keyB = take(m); sb->valB = take(m); sb->hist = take(256 * tiles);
  }
  pb.grpFirst = take(m); pb.grpCount = take(m);
  pb.jobStart = take(16); pb.grpStart = take(16);
  if ((uintptr_t)w & 7) w++;
  pb.bound = reinterpret_cast<uint64_t *>(take(2 * m));
  pb.slotBytes = reinterpret_cast<uint64_t *>(take(2 * (2 * m + 64)));
  plan_keys_kernel<<<(n + 255) / 256, 256, 0, s>>>(B, plan, pb);
  plan_layout_kernel<<<1, 32, 0, s>>>(plan, pb);
  radix_sort(pb.jobs, nullptr, n, 4, s);
  const uint32_t gMax = n + N_CLS;                                  // groups: at most one per job plus one partial group per class
  plan_groups_kernel<<<(gMax + 127) / 128, 128, 0, s>>>(B, plan, pb, affine ? 1 : 0);
  radix_sort(pb.groups, &plan->nGroupsTotal, gMax, 4, s);
  plan_order_kernel<<<(gMax + 127) / 128, 128, 0, s>>>(plan, pb, order);
  plan_offsets_kernel<<<1, 1024, 0, s>>>(plan, pb, order, arrowOff, poolBytes);
  radix_sort(pb.trace, nullptr, n, 2, s);
  plan_trace_kernel<<<(n + 255) / 256, 256, 0, s>>>(plan, pb, order);
}

}  // namespace bgpu

// bgpu_api.cu -- host side of the C ABI declared in include/blasr_gpu.h.
//
// Owns device memory, pinned staging, the streams and the kernel schedule of one batch:
//   H2D -> prep (guide rows, d-block windows) -> planner kernels (bgpu_plan.cu: classify by window width, order
//   longest-first, lay out the traceback pool) -> fill kernels (one per job class, concurrently) -> traceback ->
//   count scan -> emit (blocks, gaps, stats) -> D2H.
// bgpu_submit enqueues all of it and returns (enqueue_guided_fast); a ticket whose traceback does not fit the pool in one
// wave -- or BGPU_HOST_PLAN=1 -- takes the round-1 path instead (enqueue_guided: geometry read back, waves cut on the host).
// No CPU implementation of any aligner exists here: without a CUDA device every entry point fails.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "bgpu_common.cuh"

namespace bgpu {
int run_sdp(const bgpu_scorefn *, const int *, float, int, int, uint32_t, const uint8_t *, const uint64_t *, const uint8_t *, const uint64_t *,
            uint8_t *, size_t, unsigned, uint32_t *, bgpu_result *, bgpu_block *, const uint64_t *, cudaStream_t);
void launch_gather_reference(uint32_t, const uint8_t *, uint64_t, const uint64_t *, const uint8_t *, const uint64_t *, uint8_t *, cudaStream_t);
void launch_prep_guided(const BatchDev &, const ScoreParams &, int, const uint64_t *, const uint64_t *,
                        const uint64_t *, cudaStream_t);
void launch_fill_guided(const BatchDev &, const ScoreParams &, int, const uint32_t *, const PlanHead *, uint32_t, uint32_t *, int,
                        cudaStream_t);
void launch_trace_guided(const BatchDev &, bool, const uint32_t *, const PlanHead *, uint32_t, cudaStream_t);
void launch_unpack_guide(uint32_t, const uint64_t *, const uint8_t *, const uint32_t *, uint64_t, bgpu_block *, cudaStream_t);
void launch_plan_guided(const BatchDev &, bool, PlanHead *, uint32_t *, uint32_t *, uint64_t *, unsigned long long, cudaStream_t);
size_t plan_scratch_words(uint32_t);
void launch_scan_counts(const BatchDev &, uint64_t *, uint64_t *, uint64_t *, uint64_t *, cudaStream_t);
void launch_emit(const BatchDev &, const ScoreParams &, bgpu_result *, bgpu_block *, uint32_t *, bgpu_gap *,
                 const uint64_t *, const uint64_t *, const uint64_t *, int, int, int, PlanHead *, uint32_t *, cudaStream_t);
void launch_rescore(const BatchDev &, const ScoreParams &, const uint64_t *, const uint64_t *, const uint64_t *, int, int32_t *, cudaStream_t);
void launch_dense_prep(const BatchDev &, const ScoreParams &, const DenseArgs &, const uint64_t *, const uint64_t *, cudaStream_t);
void launch_dense_fill(const BatchDev &, const ScoreParams &, const DenseArgs &, const uint32_t *, uint32_t, uint32_t *, int,
                       cudaStream_t);
void launch_dense_trace(const BatchDev &, const ScoreParams &, const DenseArgs &, const uint32_t *, uint32_t, cudaStream_t);
void launch_fmt_count(const BatchDev &, uint32_t *, uint32_t *, cudaStream_t);
void launch_fmt_scan(uint32_t, const uint32_t *, const uint32_t *, uint64_t *, uint64_t *, uint64_t *, cudaStream_t);
void launch_fmt_cigar(const BatchDev &, const uint64_t *, const uint64_t *, uint32_t *, uint32_t *, uint32_t *, const uint32_t *,
                      const uint8_t *, cudaStream_t);
void launch_fmt_strings(const BatchDev &, const uint64_t *, char *, char *, char *, cudaStream_t);
double measure_int_peak(int nSM, cudaStream_t s, double *clockMHz);
struct AnchorState;
int anchor_set_index(int, const uint8_t *, uint64_t, uint64_t, const uint32_t *, uint64_t, const uint32_t *, const uint32_t *, uint32_t, std::string &);
int anchor_map(int, const uint8_t *, uint64_t, uint64_t, cudaStream_t, AnchorState **, const bgpu_anchor_params *, const uint8_t *, const uint64_t *, uint32_t,
               const uint32_t *, const uint32_t *, uint64_t *, const bgpu_match **, std::string &);
int anchor_rerun(AnchorState *, cudaStream_t, std::string &);
int anchor_timing(const AnchorState *, double *, uint64_t *, uint64_t *, uint64_t *);
void anchor_free_state(AnchorState *);
int build_lookup_table(const uint8_t *, uint64_t, const uint32_t *, uint32_t, uint32_t *, uint32_t *);
extern double g_peakByMode[4];
}  // namespace bgpu

using namespace bgpu;

// ---- phase gates.  Several host threads, each with its own context, drive one GPU (blasr's MapReads pthreads).  Left
// alone they fall into lockstep -- all copy in, then all compute, then all copy out -- and nothing overlaps.  One gate
// per device and phase (H2D, kernels, D2H) admits a bounded number of tickets at a time, so while one ticket computes
// the next one copies in and the previous one copies out.  BGPU_GATES="h,c,d" sets the three widths (0 = no gate).
struct Gate {
  std::mutex mu; std::condition_variable cv; int width = 1, held = 0;
  void acquire() { if (width <= 0) return; std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return held < width; }); held++; }
  void release() { if (width <= 0) return; { std::lock_guard<std::mutex> lk(mu); held--; } cv.notify_one(); }
};
enum { GATE_H2D = 0, GATE_COMPUTE = 1, GATE_D2H = 2 };
static Gate g_gates[64][3];
static std::once_flag g_gateOnce;
static Gate &gate(int device, int phase) {
  std::call_once(g_gateOnce, [] {
    int w[3] = {1, 3, 1};   // measured best on B200 / PCIe gen5: a few tickets' kernels overlap each other's tails
    if (const char *e = getenv("BGPU_GATES")) sscanf(e, "%d,%d,%d", &w[0], &w[1], &w[2]);
    for (auto &d : g_gates) for (int p = 0; p < 3; p++) d[p].width = w[p];
  });
  return g_gates[device & 63][phase];
}
static void CUDART_CB gate_release_cb(void *g) { static_cast<Gate *>(g)->release(); }

// BGPU_TRACE=1: every collected ticket prints the device-side times of its stages relative to one process-wide reference
// event (diagnostic for the pipelining of concurrent contexts; stderr, one line per ticket).
static cudaEvent_t g_refEvent = nullptr;
static std::once_flag g_refOnce;
static bool g_trace = false;
static void trace_init(cudaStream_t s) {
  std::call_once(g_refOnce, [&] {
    const char *e = getenv("BGPU_TRACE");
    g_trace = e && *e && *e != '0';
    if (g_trace) { cudaEventCreate(&g_refEvent); cudaEventRecord(g_refEvent, s); cudaEventSynchronize(g_refEvent); }
  });
}

// ---- memory.  A ticket makes ~40 allocations whose sizes follow the batch; a stream of blasr-sized tickets (the candidates
// of a few reads) never repeats a size, so a per-size cache keeps missing and every miss is a cudaMalloc / cudaHostAlloc
// (measured: 5-45 ms per ticket).  Instead each context keeps SLABS: a ticket takes whole slabs from the free list
// (smallest one that holds the request and the ticket's size hint), bump-allocates inside them and gives them back on
// release.  After a few tickets the free list covers the working set and allocation is pointer arithmetic.
struct Slab { char *base; size_t cap; };
struct SlabPool {
  bool pinned;
  std::vector<Slab> free_;
  size_t lastCap = 0;
  uint32_t nAlloc = 0;                           // cudaMalloc / cudaHostAlloc calls so far
  explicit SlabPool(bool p) : pinned(p) {}
};
struct TicketMem { std::vector<Slab> owned; size_t used = 0, hint = 0; };

struct bgpu_ctx {
  int device = 0, nSM = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t aux[N_CLS] = {};                  // one per job class: the class kernels of a wave run concurrently
  cudaEvent_t evFork = nullptr, evJoin[N_CLS] = {};
  cudaEvent_t evSync = nullptr;                  // blocking-sync event: waiting host threads sleep instead of spinning
  cudaStream_t copyStream = nullptr;             // bgpu_collect's copies (and late emit): not behind the next ticket's kernels
  std::string err;
  std::mutex mu;
  size_t arrowPoolCap = 0;                       // max bytes of traceback pool per wave
  SlabPool devPool{false}, pinPool{true};        // cached device / pinned slabs, handed to tickets whole
  bgpu_ticket lastSync = nullptr;                // ticket owned by bgpu_align
  void *sdpPinned = nullptr; size_t sdpPinnedBytes = 0;   // result arena of the last bgpu_sdp_align
  bool sdpStackSet = false;
  bgpu::AnchorState *anchor = nullptr;                    // buffers of the last bgpu_map_reads (bgpu_anchor.cu)
};

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      char buf_[256];                                                                             \
      snprintf(buf_, sizeof buf_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      ctx->err = buf_;                                                                            \
      return (e_ == cudaErrorMemoryAllocation) ? BGPU_E_OOM : BGPU_E_CUDA;                         \
    }                                                                                             \
  } while (0)

// Waits for everything queued on the context's stream.  By default the waiting thread polls an event with short naps (5 us
// growing to 200 us, wait_event below): close to the latency of a spinning wait without its CPU cost.  BGPU_SPIN_SYNC=1 spins
// in cudaStreamSynchronize, BGPU_BLOCKING_SYNC=1 sleeps on an event created with cudaEventBlockingSync: for hosts that run
// more waiting threads than they have cores (blasr's one pthread per core, several GPUs per box).
static bool blocking_sync() { static const bool on = [] { const char *e = getenv("BGPU_BLOCKING_SYNC"); return e && *e && *e != '0'; }(); return on; }
static bool spin_sync() { static const bool on = [] { const char *e = getenv("BGPU_SPIN_SYNC"); return e && *e && *e != '0'; }(); return on; }
static cudaError_t wait_event(cudaEvent_t ev);
static cudaError_t wait_stream(bgpu_ctx *ctx, cudaStream_t st = nullptr) {
  if (!st) st = ctx->stream;
  if (spin_sync()) return cudaStreamSynchronize(st);
  cudaError_t e = cudaEventRecord(ctx->evSync, st);
  if (e != cudaSuccess) return e;
  if (blocking_sync()) return cudaEventSynchronize(ctx->evSync);
  return wait_event(ctx->evSync);
}
static cudaError_t wait_event(cudaEvent_t ev) {
  if (spin_sync() || blocking_sync()) return cudaEventSynchronize(ev);
  cudaError_t e;
  // default: poll with short naps that grow to 200 us -- a waiting host thread costs (almost) no CPU, which matters as soon as
  // a box drives several GPUs with several threads each, and wakes within the time a PCIe copy of the results takes anyway
  unsigned nap = 5;
  for (;;) {
    e = cudaEventQuery(ev);
    if (e != cudaErrorNotReady) return e;
    cudaGetLastError();
    std::this_thread::sleep_for(std::chrono::microseconds(nap));
    if (nap < 200) nap += nap;
  }
}

static cudaError_t raw_alloc(bool pinned, void **p, size_t bytes) {
  return pinned ? cudaHostAlloc(p, bytes, cudaHostAllocDefault) : cudaMalloc(p, bytes);
}
static void raw_free(bool pinned, void *p) { if (pinned) cudaFreeHost(p); else cudaFree(p); }

// every live context of the process: when an allocation fails, the idle slabs cached by the OTHER contexts of the same device
// are given back to the driver before the retry (several host threads, one context each, share a GPU)
static std::mutex g_ctxMu;
static std::vector<bgpu_ctx *> g_ctxs;
static void trim_pool(SlabPool &pool) {
  for (auto &f : pool.free_) raw_free(pool.pinned, f.base);
  pool.free_.clear();
}
static void trim_others(bgpu_ctx *self, bool pinned);

static int slab_alloc(bgpu_ctx *ctx, SlabPool &pool, TicketMem &tm, void **p, size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (!tm.owned.empty() && tm.used + bytes <= tm.owned.back().cap) {
    *p = tm.owned.back().base + tm.used; tm.used += bytes; return BGPU_OK;
  }
  // a new slab: the smallest free one that holds this request plus what the ticket still expects to ask for
  const size_t minSlab = pool.pinned ? (1u << 20) : (4u << 20);
  const size_t want = std::max(bytes + std::min(tm.hint, (size_t)1 << 30), minSlab);
  int best = -1, fits = -1;
  for (int i = 0; i < (int)pool.free_.size(); i++) {
    const size_t c = pool.free_[i].cap;
    if (c >= want && (best < 0 || c < pool.free_[best].cap)) best = i;
    if (c >= bytes && (fits < 0 || c > pool.free_[fits].cap)) fits = i;       // else the largest one that holds the request
  }
  if (best < 0) best = fits;
  Slab sl{};
  if (best >= 0) { sl = pool.free_[best]; pool.free_.erase(pool.free_.begin() + best); }
  else {
    size_t cap = std::max(want, std::min(2 * pool.lastCap, (size_t)1 << 30));
    cap = (cap + 0xfffff) & ~(size_t)0xfffff;
    void *v = nullptr;
    pool.nAlloc++;
    cudaError_t e = raw_alloc(pool.pinned, &v, cap);
    if (e != cudaSuccess && cap > bytes) { cudaGetLastError(); cap = (bytes + 0xfffff) & ~(size_t)0xfffff; e = raw_alloc(pool.pinned, &v, cap); }
    if (e != cudaSuccess) {   // drop this context's cache, then the other contexts' idle slabs, and retry
      cudaGetLastError();
      trim_pool(pool);
      e = raw_alloc(pool.pinned, &v, cap);
      if (e != cudaSuccess) { cudaGetLastError(); trim_others(ctx, pool.pinned); e = raw_alloc(pool.pinned, &v, cap); }
    }
    if (e != cudaSuccess) { ctx->err = std::string(pool.pinned ? "cudaHostAlloc: " : "cudaMalloc: ") + cudaGetErrorString(e); cudaGetLastError(); return BGPU_E_OOM; }
    sl.base = (char *)v; sl.cap = cap; pool.lastCap = cap;
  }
  tm.hint = tm.hint > sl.cap - bytes ? tm.hint - (sl.cap - bytes) : 0;
  tm.owned.push_back(sl); tm.used = bytes;
  *p = sl.base;
  return BGPU_OK;
}
static void slab_release(SlabPool &pool, TicketMem &tm) {
  for (auto &sl : tm.owned) pool.free_.push_back(sl);
  tm.owned.clear(); tm.used = 0; tm.hint = 0;
}

static void trim_others(bgpu_ctx *self, bool pinned) {
  std::lock_guard<std::mutex> lk(g_ctxMu);
  for (bgpu_ctx *c : g_ctxs) {
    if (c == self || c->device != self->device) continue;
    std::unique_lock<std::mutex> cl(c->mu, std::try_to_lock);   // a context busy in a call keeps its slabs
    if (cl.owns_lock()) trim_pool(pinned ? c->pinPool : c->devPool);
  }
}

struct Wave { uint32_t begin[N_CLS], count[N_CLS]; uint32_t traceBegin, traceCount; };  // index by job class

struct bgpu_ticket_s {
  uint32_t nJobs = 0;
  bgpu_params params{};
  ScoreParams sp{};
  BatchDev B{};
  TicketMem dev, pin;         // the slabs this ticket allocates from (returned to the context's pools on release)
  uint64_t *d_rowOff = nullptr, *d_dblkOff = nullptr, *d_runOff = nullptr, *d_arrowOff = nullptr;
  uint32_t *d_order = nullptr, *d_counters = nullptr;
  uint64_t *d_blockOff = nullptr, *d_listOff = nullptr, *d_gapOff = nullptr, *d_totals = nullptr;
  bgpu_result *d_results = nullptr;
  bgpu_block *d_blocks = nullptr; uint32_t *d_gapCounts = nullptr; bgpu_gap *d_gaps = nullptr;
  uint32_t *d_runsOut = nullptr, *h_runsOut = nullptr;          // params.compactResults: the run-length paths instead
  // pinned host
  JobGeom *h_geom = nullptr; uint64_t *h_totals = nullptr;
  bgpu_result *h_results = nullptr; bgpu_block *h_blocks = nullptr; uint32_t *h_gapCounts = nullptr; bgpu_gap *h_gaps = nullptr;
  uint64_t totals[3] = {0, 0, 0};
  std::vector<Wave> waves;
  uint32_t nCounters = 0;
  bool arenaReady = false, collected = false, dense = false;
  bool holdsH2D = false;      // the ticket is inside the H2D gate and its release has not been queued on the stream yet
  bool holdsCompute = false;  // ... inside the kernel gate and its release has not been queued on the stream yet
  double msUpload = 0;            // host time spent staging + enqueueing the input copies
  cudaEvent_t evDone = nullptr;   // recorded behind everything bgpu_submit enqueued (bgpu_query)
  cudaEvent_t ev[6] = {};     // start, prepEnd, fillTraceEnd(unused), scanEnd, emitEnd
  std::vector<cudaEvent_t> waveEv;   // per wave: fillStart, fillEnd, traceEnd
  bgpu_timing timing{};
  DenseArgs dargs{};
  std::vector<uint64_t> h_arrowBytes;   // dense: host-computed traceback bytes per job
  std::vector<uint64_t> h_cellsMetric;  // dense: SURVEY 8(d) cell count per job
  bool denseSmall = false;              // dense: every matrix is small and all of them fit one wave: no read-back, no host planning
  uint64_t *h_aoffFast = nullptr; uint64_t denseWaveBytes = 0;   // ... traceback offsets laid out by bgpu_submit itself
  uint32_t *h_cigar = nullptr; uint64_t *h_cigarOff = nullptr;   // bgpu_cigar results (pinned), once built
  uint32_t *d_fmtOps = nullptr, *d_fmtCols = nullptr;            // per-job CIGAR op / alignment column counts, once counted
  char *h_str = nullptr; uint64_t *h_strOff = nullptr; size_t strTotal = 0;   // bgpu_strings results (pinned), once built
  // asynchronous guided path: the schedule is built on the device (bgpu_plan.cu), one wave, speculative sizes
  bool fast = false;            // submit enqueued everything without waiting for the device
  bool gated = false;           // large ticket: goes through the H2D / kernel / D2H phase gates
  bool arenaInline = false;     // the (small) result arena was copied back by submit already
  PlanHead *d_plan = nullptr;   // device: one head per wave (the asynchronous path has exactly one)
  PlanHead *h_planInit = nullptr, *h_plan = nullptr;   // pinned: the head as uploaded / as read back after the kernels
  uint32_t *d_planScratch = nullptr;
  uint64_t poolBytes = 0, arenaCap[3] = {0, 0, 0};
};

template <typename T>
static int talloc_dev(bgpu_ctx *ctx, bgpu_ticket t, T **p, size_t n) {
  void *v = nullptr; int rc = slab_alloc(ctx, ctx->devPool, t->dev, &v, n * sizeof(T)); if (rc) return rc;
  *p = (T *)v; return BGPU_OK;
}
template <typename T>
static int talloc_pin(bgpu_ctx *ctx, bgpu_ticket t, T **p, size_t n) {
  void *v = nullptr; int rc = slab_alloc(ctx, ctx->pinPool, t->pin, &v, n * sizeof(T)); if (rc) return rc;
  *p = (T *)v; return BGPU_OK;
}
#define RC(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

// H2D of caller memory: through pinned staging unless the caller's buffer is already pinned.
static int upload(bgpu_ctx *ctx, bgpu_ticket t, void *dst, const void *src, size_t bytes) {
  if (!bytes) return BGPU_OK;
  struct Tm { bgpu_ticket t; std::chrono::steady_clock::time_point t0; ~Tm() { t->msUpload += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); } } tm{t, std::chrono::steady_clock::now()};
  cudaPointerAttributes at{};
  const bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
  cudaGetLastError();
  const void *from = src;
  if (!pinned) {
    void *stage = nullptr; RC(talloc_pin(ctx, t, (uint8_t **)&stage, bytes));
    memcpy(stage, src, bytes);
    from = stage;
  }
  CK(cudaMemcpyAsync(dst, from, bytes, cudaMemcpyHostToDevice, ctx->stream));
  t->timing.h2dBytes += bytes;
  return BGPU_OK;
}

// the reference (genome) resident per device: shared by every context on it
struct DeviceRef { uint8_t *d = nullptr; uint64_t n = 0, gen = 0; };   // gen: bumped by every bgpu_set_reference
static std::mutex g_refMu;
static DeviceRef g_ref[64];

extern "C" int bgpu_set_reference(bgpu_ctx *ctx, const uint8_t *bases, uint64_t n) {
  if (!ctx || (n && !bases) || ctx->device >= 64) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  std::lock_guard<std::mutex> lr(g_refMu);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return BGPU_E_CUDA;
  DeviceRef &r = g_ref[ctx->device];
  CK(cudaDeviceSynchronize());                       // tickets in flight may still be gathering from the old one
  if (r.d) { cudaFree(r.d); r.d = nullptr; r.n = 0; }
  r.gen++;
  if (!n) return BGPU_OK;
  CK(cudaMalloc(&r.d, n + 16));
  CK(cudaMemcpy(r.d, bases, n, cudaMemcpyHostToDevice));
  r.n = n;
  return BGPU_OK;
}

extern "C" int bgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
extern "C" int bgpu_version(void) { return BGPU_VERSION; }
extern "C" int bgpu_base_code(int c) { return bgpu::base_code((uint8_t)c); }

extern "C" int bgpu_create(bgpu_ctx **out, int device) {
  if (!out) return BGPU_E_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return BGPU_E_NO_DEVICE; }
  if (device < 0 || device >= n) return BGPU_E_INVALID;
  bgpu_ctx *ctx = new bgpu_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return BGPU_E_CUDA; }
  cudaDeviceProp pr{};
  cudaGetDeviceProperties(&pr, device);
  ctx->nSM = pr.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return BGPU_E_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return BGPU_E_CUDA; }
  for (int c = 0; c < N_CLS; c++) {
    if (cudaStreamCreateWithFlags(&ctx->aux[c], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evJoin[c], cudaEventDisableTiming) != cudaSuccess) { delete ctx; return BGPU_E_CUDA; }
  }
  if (cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->evSync, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) { delete ctx; return BGPU_E_CUDA; }
  size_t freeB = 0, totalB = 0;
  cudaMemGetInfo(&freeB, &totalB);
  ctx->arrowPoolCap = freeB / 5 * 3;   // one wave holds ~100 GB of affine arrows (100k pairs of 1-20 kb): a B200 has the HBM for it
  const char *env = getenv("BGPU_ARROW_POOL_MB");
  if (env) ctx->arrowPoolCap = (size_t)atoll(env) << 20;
  trace_init(ctx->stream);
  { std::lock_guard<std::mutex> lk(g_ctxMu); g_ctxs.push_back(ctx); }
  *out = ctx;
  return BGPU_OK;
}

extern "C" int bgpu_trim(bgpu_ctx *ctx) {
  if (!ctx) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return BGPU_E_CUDA;
  cudaStreamSynchronize(ctx->stream);
  trim_pool(ctx->devPool); trim_pool(ctx->pinPool);
  return BGPU_OK;
}

extern "C" void bgpu_destroy(bgpu_ctx *ctx) {
  if (!ctx) return;
  { std::lock_guard<std::mutex> lk(g_ctxMu); g_ctxs.erase(std::remove(g_ctxs.begin(), g_ctxs.end(), ctx), g_ctxs.end()); }
  cudaSetDevice(ctx->device);
  if (ctx->lastSync) bgpu_release(ctx, ctx->lastSync);
  cudaStreamSynchronize(ctx->stream);
  for (auto &sl : ctx->devPool.free_) cudaFree(sl.base);
  for (auto &sl : ctx->pinPool.free_) cudaFreeHost(sl.base);
  for (int c = 0; c < N_CLS; c++) { cudaStreamDestroy(ctx->aux[c]); cudaEventDestroy(ctx->evJoin[c]); }
  cudaEventDestroy(ctx->evFork); cudaEventDestroy(ctx->evSync);
  if (ctx->sdpPinned) cudaFreeHost(ctx->sdpPinned);
  anchor_free_state(ctx->anchor);
  cudaStreamDestroy(ctx->copyStream);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char *bgpu_last_error(const bgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

static void fill_score_params(ScoreParams &sp, const bgpu_scorefn *fn, const bgpu_params *p) {
  memcpy(sp.M, fn->M, sizeof sp.M);
  sp.ins = fn->ins; sp.del = fn->del; sp.open = fn->affineOpen; sp.ext = fn->affineExtend;
  sp.kind = fn->kind; sp.alignType = p->alignType; sp.affine = (p->algo == BGPU_AFFINE_GUIDED); sp.pad = 0;
  sp.subPrior = fn->substitutionPrior; sp.delPrior = fn->globalDeletionPrior;
}

// IDSScoreFunction (IDSScoreFunction.h:80-139) reads insertionQV / substitutionQV / substitutionTag unconditionally and
// deletionQV / deletionTag when both are present; SWAlign hands it transposed positions (SWAlign.h:166-167) and the
// reference then reads past the ends of the tracks, so that combination is refused.
static int check_ids_tracks(bgpu_ctx *ctx, const bgpu_scorefn *fn, const bgpu_params *p, const bgpu_batch *b) {
  if (fn->kind != BGPU_FN_IDS) return BGPU_OK;
  if (p->algo == BGPU_AFFINE_KBAND) { ctx->err = "AffineKBandAlign takes a match matrix, not a score function (AffineKBandAlign.h:13)"; return BGPU_E_INVALID; }
  if (p->algo == BGPU_SW) { ctx->err = "SWAlign x BGPU_FN_IDS is undefined in the reference (out-of-bounds track reads)"; return BGPU_E_INVALID; }
  if (b->nJobs && (!b->insQV || !b->subQV || !b->subTag)) { ctx->err = "BGPU_FN_IDS needs batch.insQV, subQV and subTag"; return BGPU_E_INVALID; }
  if ((b->delQV == nullptr) != (b->delTag == nullptr)) { ctx->err = "BGPU_FN_IDS: delQV and delTag come as a pair"; return BGPU_E_INVALID; }
  return BGPU_OK;
}
static int upload_ids_tracks(bgpu_ctx *ctx, bgpu_ticket t, const bgpu_scorefn *fn, const bgpu_batch *b, uint64_t totQ) {
  BatchDev &B = t->B;
  B.insQV = B.delQV = B.subQV = B.delTag = B.subTag = nullptr;
  if (fn->kind != BGPU_FN_IDS) return BGPU_OK;
  const uint8_t *src[5] = {b->insQV, b->delQV, b->subQV, b->delTag, b->subTag};
  const uint8_t **dst[5] = {&B.insQV, &B.delQV, &B.subQV, &B.delTag, &B.subTag};
  for (int i = 0; i < 5; i++) {
    if (!src[i]) continue;
    uint8_t *d = nullptr;
    RC(talloc_dev(ctx, t, &d, totQ + 16));
    RC(upload(ctx, t, d, src[i], totQ));
    *dst[i] = d;
  }
  return BGPU_OK;
}

// ---- kernel schedule of a guided ticket (used by submit and rerun) ----
static int enqueue_guided(bgpu_ctx *ctx, bgpu_ticket t, bool firstRun) {
  cudaStream_t s = ctx->stream;
  CK(cudaEventRecord(t->ev[0], s));
  launch_prep_guided(t->B, t->sp, t->params.band, t->d_rowOff, t->d_dblkOff, t->d_runOff, s);
  t->timing.kernelLaunches = 1;
  CK(cudaEventRecord(t->ev[1], s));
  if (firstRun) {
    // geometry back to the host: class lists, warp groups and wave cutting need it
    CK(cudaMemcpyAsync(t->h_geom, t->B.geom, sizeof(JobGeom) * t->nJobs, cudaMemcpyDeviceToHost, s));
    CK(wait_stream(ctx));
    const uint32_t n = t->nJobs;
    const bool affine = t->sp.affine != 0;
    const uint64_t rowsPerBlock = affine ? 16 : 4;             // 64 anti-diagonals / steps per traceback word
    std::vector<uint32_t> byCls[N_CLS];
    for (uint32_t i = 0; i < n; i++) if (t->h_geom[i].status == BGPU_JOB_OK) byCls[t->h_geom[i].cls].push_back(i);
    // a warp sweeps 32/LPJ jobs in lockstep with k = the widest member's need: put jobs of similar typical width and
    // length side by side (typical width first, longest first inside it)
    struct Group { uint32_t first, count; uint64_t bytes, cost; int cls; };
    std::vector<Group> groups;
    std::vector<uint32_t> sorted;                                // job indices, class-major
    std::vector<uint64_t> bound(n, 0);                           // traceback bytes reserved per job
    uint64_t laneSteps = 0;
    for (int c = 0; c < N_CLS; c++) {
      auto &v = byCls[c];
      auto kt = [&](uint32_t i) { const JobGeom &g = t->h_geom[i]; return (g.ksum + g.nDB / 2) / std::max(g.nDB, 1); };
      {   // (typical width, d-blocks) descending, index ascending -- sorted on packed keys, not through the geometry array
        std::vector<std::pair<uint64_t, uint32_t>> keyed; keyed.reserve(v.size());
        for (uint32_t i : v) keyed.emplace_back(~(((uint64_t)(uint32_t)kt(i) << 32) | (uint32_t)t->h_geom[i].nDB), i);
        std::sort(keyed.begin(), keyed.end());
        for (size_t i = 0; i < keyed.size(); i++) v[i] = keyed[i].second;
      }
      const uint32_t lpj = (uint32_t)cls_lpj(c), nj = 32 / lpj;
      for (size_t i0 = 0; i0 < v.size(); i0 += nj) {
        const size_t i1 = std::min(v.size(), i0 + nj);
        int kmaxG = 1, nDBmax = 0; uint64_t ksumMax = 0;
        for (size_t i = i0; i < i1; i++) {
          kmaxG = std::max(kmaxG, t->h_geom[v[i]].kmax); nDBmax = std::max(nDBmax, t->h_geom[v[i]].nDB);
          ksumMax = std::max<uint64_t>(ksumMax, (uint64_t)t->h_geom[v[i]].ksum);
        }
        Group g{(uint32_t)sorted.size(), (uint32_t)(i1 - i0), 0, std::max<uint64_t>(ksumMax, (uint64_t)nDBmax), c};
        for (size_t i = i0; i < i1; i++) {
          const uint64_t bb = (uint64_t)t->h_geom[v[i]].nDB * (uint64_t)kmaxG * rowsPerBlock * lpj * 4ull;
          bound[v[i]] = (bb + 255) & ~255ull; g.bytes += bound[v[i]];
          sorted.push_back(v[i]);
        }
        laneSteps += ksumMax * 64ull * 32ull;                    // lower bound of the cell slots the warp executes
        groups.push_back(g);
      }
    }
    // dispatch order inside a class: most expensive warp group first (the queue's tail is then made of short jobs)
    std::stable_sort(groups.begin(), groups.end(), [](const Group &a, const Group &b) { return a.cls != b.cls ? a.cls < b.cls : a.cost > b.cost; });
    // cut into waves by reserved traceback bytes (groups stay whole)
    std::vector<uint64_t> arrowOff(n, 0);
    std::vector<uint32_t> order;
    size_t maxWaveBytes = 0, g0 = 0;
    while (g0 < groups.size()) {
      size_t bytes = 0, g1 = g0;
      while (g1 < groups.size()) {
        if (g1 > g0 && bytes + groups[g1].bytes > ctx->arrowPoolCap) break;
        bytes += groups[g1].bytes; g1++;
      }
      Wave w{};
      size_t off = 0;
      for (int c = 0; c < N_CLS; c++) {
        const uint32_t nj = 32 / (uint32_t)cls_lpj(c);
        w.begin[c] = (uint32_t)order.size(); w.count[c] = 0;
        for (size_t gi = g0; gi < g1; gi++) {
          const Group &g = groups[gi];
          if (g.cls != c) continue;
          for (uint32_t j = 0; j < nj; j++) {
            if (j < g.count) { const uint32_t job = sorted[g.first + j]; arrowOff[job] = off; off += bound[job]; order.push_back(job); }
            else order.push_back(0xffffffffu);
          }
          w.count[c]++;
        }
      }
      // traceback list: longest jobs first
      w.traceBegin = (uint32_t)order.size();
      std::vector<uint32_t> tl;
      for (size_t gi = g0; gi < g1; gi++) for (uint32_t j = 0; j < groups[gi].count; j++) tl.push_back(sorted[groups[gi].first + j]);
      {
        std::vector<std::pair<uint64_t, uint32_t>> keyed; keyed.reserve(tl.size());
        for (uint32_t i : tl) keyed.emplace_back(~(uint64_t)(uint32_t)t->h_geom[i].nDB, i);
        std::sort(keyed.begin(), keyed.end());
        for (size_t i = 0; i < keyed.size(); i++) tl[i] = keyed[i].second;
      }
      order.insert(order.end(), tl.begin(), tl.end());
      w.traceCount = (uint32_t)tl.size();
      maxWaveBytes = std::max(maxWaveBytes, off);
      t->waves.push_back(w);
      g0 = g1;
    }
    RC(talloc_dev(ctx, t, &t->d_order, std::max<size_t>(order.size(), 1)));
    RC(talloc_dev(ctx, t, &t->d_arrowOff, std::max<uint32_t>(n, 1)));
    {   // the kernels read their part of the schedule from a device-resident head per wave
      const size_t nw = std::max<size_t>(t->waves.size(), 1);
      PlanHead *h_heads = nullptr;
      RC(talloc_dev(ctx, t, &t->d_plan, nw)); RC(talloc_pin(ctx, t, &h_heads, nw));
      memset(h_heads, 0, sizeof(PlanHead) * nw);
      for (size_t w = 0; w < t->waves.size(); w++) {
        for (int c = 0; c < N_CLS; c++) { h_heads[w].nGroups[c] = t->waves[w].count[c]; h_heads[w].orderBegin[c] = t->waves[w].begin[c]; }
        h_heads[w].traceBegin = t->waves[w].traceBegin; h_heads[w].traceCount = t->waves[w].traceCount;
      }
      CK(cudaMemcpyAsync(t->d_plan, h_heads, sizeof(PlanHead) * nw, cudaMemcpyHostToDevice, s));
    }
    t->nCounters = (uint32_t)t->waves.size() * 16 + 16;          // per wave: [c] = work queue of class c's fill kernel
    RC(talloc_dev(ctx, t, &t->d_counters, t->nCounters));
    uint8_t *arrows = nullptr;
    RC(talloc_dev(ctx, t, &arrows, std::max<size_t>(maxWaveBytes, 16)));
    t->B.arrows = arrows; t->B.arrowOff = t->d_arrowOff; t->B.order = t->d_order; t->B.counters = t->d_counters;
    uint32_t *h_order = nullptr; uint64_t *h_aoff = nullptr;
    RC(talloc_pin(ctx, t, &h_order, std::max<size_t>(order.size(), 1)));
    RC(talloc_pin(ctx, t, &h_aoff, std::max<uint32_t>(n, 1)));
    memcpy(h_order, order.data(), order.size() * sizeof(uint32_t));
    memcpy(h_aoff, arrowOff.data(), n * sizeof(uint64_t));
    CK(cudaMemcpyAsync(t->d_order, h_order, order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(t->d_arrowOff, h_aoff, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    t->waveEv.resize(t->waves.size() * 3);
    for (auto &e : t->waveEv) CK(cudaEventCreate(&e));
    uint64_t cells = 0;
    for (uint32_t i = 0; i < n; i++) if (t->h_geom[i].status == BGPU_JOB_OK) cells += (uint64_t)t->h_geom[i].nCells;
    t->timing.cells = cells;
    t->timing.fillCells = laneSteps;
  }
  if (firstRun && t->gated) { gate(ctx->device, GATE_COMPUTE).acquire(); t->holdsCompute = true; }   // released by the stream itself once the last kernel is done
  CK(cudaMemsetAsync(t->d_counters, 0, sizeof(uint32_t) * t->nCounters, s));
  CK(cudaMemsetAsync(t->d_totals + 3, 0, sizeof(uint64_t), s));
  t->B.cellSlots = reinterpret_cast<unsigned long long *>(t->d_totals + 3);
  for (size_t w = 0; w < t->waves.size(); w++) {
    const Wave &W = t->waves[w];
    CK(cudaEventRecord(t->waveEv[3 * w], s));
    // the class kernels are independent: fork them onto their own streams (heaviest classes first), join on s
    CK(cudaEventRecord(ctx->evFork, s));
    for (int c = N_CLS - 1; c >= 0; c--)
      if (W.count[c]) {
        CK(cudaStreamWaitEvent(ctx->aux[c], ctx->evFork, 0));
        launch_fill_guided(t->B, t->sp, c, t->d_order, t->d_plan + w, W.count[c], t->d_counters + 16 * w + c, ctx->nSM, ctx->aux[c]);
        CK(cudaEventRecord(ctx->evJoin[c], ctx->aux[c]));
        CK(cudaStreamWaitEvent(s, ctx->evJoin[c], 0));
        t->timing.kernelLaunches++;
      }
    CK(cudaEventRecord(t->waveEv[3 * w + 1], s));
    if (W.traceCount) {
      launch_trace_guided(t->B, t->sp.affine != 0, t->d_order, t->d_plan + w, W.traceCount, s);
      t->timing.kernelLaunches++;
    }
    CK(cudaEventRecord(t->waveEv[3 * w + 2], s));
  }
  launch_scan_counts(t->B, t->d_blockOff, t->d_listOff, t->d_gapOff, t->d_totals, s);
  t->timing.kernelLaunches++;
  CK(cudaEventRecord(t->ev[3], s));
  if (firstRun && t->holdsCompute) {
    if (gate(ctx->device, GATE_COMPUTE).width > 0) CK(cudaLaunchHostFunc(s, gate_release_cb, &gate(ctx->device, GATE_COMPUTE)));
    t->holdsCompute = false;
  }
  CK(cudaGetLastError());
  return BGPU_OK;
}

static int enqueue_emit(bgpu_ctx *ctx, bgpu_ticket t, cudaStream_t st = nullptr);

// result arena for nB blocks, nL gap lists, nG gaps (device + pinned host): Block / Gap arrays, or the run-length paths
static int alloc_arena(bgpu_ctx *ctx, bgpu_ticket t, uint64_t nB, uint64_t nL, uint64_t nG) {
  if (t->params.compactResults && !t->dense) {
    RC(talloc_dev(ctx, t, &t->d_runsOut, nB + nG + 1)); RC(talloc_pin(ctx, t, &t->h_runsOut, nB + nG + 1));
  } else {
    RC(talloc_dev(ctx, t, &t->d_blocks, nB + 1)); RC(talloc_dev(ctx, t, &t->d_gapCounts, nL + 1)); RC(talloc_dev(ctx, t, &t->d_gaps, nG + 1));
    RC(talloc_pin(ctx, t, &t->h_blocks, nB + 1)); RC(talloc_pin(ctx, t, &t->h_gapCounts, nL + 1)); RC(talloc_pin(ctx, t, &t->h_gaps, nG + 1));
  }
  t->arenaReady = true;
  return BGPU_OK;
}
static uint64_t arena_bytes(bgpu_ticket t, uint64_t nB, uint64_t nL, uint64_t nG) {
  if (t->params.compactResults && !t->dense) return sizeof(uint32_t) * (nB + nG);
  return sizeof(bgpu_block) * nB + sizeof(uint32_t) * nL + sizeof(bgpu_gap) * nG;
}
static int copy_arena(bgpu_ctx *ctx, bgpu_ticket t, uint64_t nB, uint64_t nL, uint64_t nG, cudaStream_t s = nullptr) {
  if (!s) s = ctx->stream;
  if (t->d_runsOut) { CK(cudaMemcpyAsync(t->h_runsOut, t->d_runsOut, sizeof(uint32_t) * (nB + nG), cudaMemcpyDeviceToHost, s)); return BGPU_OK; }
  CK(cudaMemcpyAsync(t->h_blocks, t->d_blocks, sizeof(bgpu_block) * nB, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(t->h_gapCounts, t->d_gapCounts, sizeof(uint32_t) * nL, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(t->h_gaps, t->d_gaps, sizeof(bgpu_gap) * nG, cudaMemcpyDeviceToHost, s));
  return BGPU_OK;
}

// ---- the asynchronous schedule of a guided ticket: nothing here waits for the device.  prep -> planner kernels (classes,
// warp groups, dispatch order, traceback offsets: bgpu_plan.cu) -> the fill kernel of every class (each reads its share of
// the schedule from the device-resident PlanHead; empty classes exit at once) -> traceback -> count scan -> emit into an
// arena sized from the batch -> the copies back.  The traceback pool and the arena are sized BEFORE the device knows the
// exact needs; the kernels check, flag PLAN_OVF_* and skip, and bgpu_collect then re-plans that ticket on the host path.
static int enqueue_guided_fast(bgpu_ctx *ctx, bgpu_ticket t, bool firstRun) {
  cudaStream_t s = ctx->stream;
  const uint32_t n = t->nJobs;
  CK(cudaEventRecord(t->ev[0], s));
  launch_prep_guided(t->B, t->sp, t->params.band, t->d_rowOff, t->d_dblkOff, t->d_runOff, s);
  CK(cudaEventRecord(t->ev[1], s));
  CK(cudaMemcpyAsync(t->d_plan, t->h_planInit, sizeof(PlanHead), cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(t->d_counters, 0, sizeof(uint32_t) * t->nCounters, s));
  CK(cudaMemsetAsync(t->d_totals + 3, 0, sizeof(uint64_t), s));
  t->B.cellSlots = reinterpret_cast<unsigned long long *>(t->d_totals + 3);
  launch_plan_guided(t->B, t->sp.affine != 0, t->d_plan, t->d_planScratch, t->d_order, t->d_arrowOff, t->poolBytes, s);
  if (firstRun && t->gated) { gate(ctx->device, GATE_COMPUTE).acquire(); t->holdsCompute = true; }
  CK(cudaEventRecord(t->waveEv[0], s));
  CK(cudaEventRecord(ctx->evFork, s));
  t->timing.kernelLaunches = 2;
  // BGPU_SERIAL_CLASSES=1 (diagnostic): the class kernels one after the other on the ticket's stream instead of concurrently
  static const bool serialClasses = [] { const char *e = getenv("BGPU_SERIAL_CLASSES"); return e && *e && *e != '0'; }();
  for (int c = N_CLS - 1; c >= 0; c--) {
    if (serialClasses) { launch_fill_guided(t->B, t->sp, c, t->d_order, t->d_plan, n + N_CLS, t->d_counters + c, ctx->nSM, s); t->timing.kernelLaunches++; continue; }
    CK(cudaStreamWaitEvent(ctx->aux[c], ctx->evFork, 0));
    launch_fill_guided(t->B, t->sp, c, t->d_order, t->d_plan, n + N_CLS, t->d_counters + c, ctx->nSM, ctx->aux[c]);
    CK(cudaEventRecord(ctx->evJoin[c], ctx->aux[c]));
    CK(cudaStreamWaitEvent(s, ctx->evJoin[c], 0));
    t->timing.kernelLaunches++;
  }
  CK(cudaEventRecord(t->waveEv[1], s));
  launch_trace_guided(t->B, t->sp.affine != 0, t->d_order, t->d_plan, n, s);
  CK(cudaEventRecord(t->waveEv[2], s));
  launch_scan_counts(t->B, t->d_blockOff, t->d_listOff, t->d_gapOff, reinterpret_cast<uint64_t *>(&t->d_plan->totals[0]), s);
  t->timing.kernelLaunches += 2;
  CK(cudaEventRecord(t->ev[3], s));
  if (t->arenaReady) RC(enqueue_emit(ctx, t));
  if (firstRun && t->holdsCompute) {
    if (gate(ctx->device, GATE_COMPUTE).width > 0) CK(cudaLaunchHostFunc(s, gate_release_cb, &gate(ctx->device, GATE_COMPUTE)));
    t->holdsCompute = false;
  }
  if (firstRun) {
    // the head (flags, totals, cell counts), the per-job results and -- when it is small -- the arena itself go back now
    CK(cudaMemcpyAsync(t->h_plan, t->d_plan, sizeof(PlanHead), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(t->h_totals, t->d_totals, sizeof(uint64_t) * 4, cudaMemcpyDeviceToHost, s));
    if (t->arenaReady) CK(cudaMemcpyAsync(t->h_results, t->d_results, sizeof(bgpu_result) * n, cudaMemcpyDeviceToHost, s));
    if (t->arenaInline) RC(copy_arena(ctx, t, t->arenaCap[0], t->arenaCap[1], t->arenaCap[2]));
  }
  CK(cudaGetLastError());
  return BGPU_OK;
}

static int enqueue_emit(bgpu_ctx *ctx, bgpu_ticket t, cudaStream_t st) {
  cudaStream_t s = st ? st : ctx->stream;
  launch_emit(t->B, t->sp, t->d_results, t->d_blocks, t->d_gapCounts, t->d_gaps, t->d_blockOff, t->d_listOff,
              t->d_gapOff, t->params.doStats, t->params.statsAffine, t->dense ? 1 : 0, t->fast ? t->d_plan : nullptr, t->d_runsOut, s);
  t->timing.kernelLaunches++;
  CK(cudaEventRecord(t->ev[4], s));
  CK(cudaGetLastError());
  return BGPU_OK;
}

static void gather_timing(bgpu_ticket t) {
  float ms = 0;
  if (cudaEventElapsedTime(&ms, t->ev[0], t->ev[1]) == cudaSuccess) t->timing.msPrep = ms;
  double fill = 0, trace = 0;
  for (size_t w = 0; w < t->waves.size(); w++) {
    if (cudaEventElapsedTime(&ms, t->waveEv[3 * w], t->waveEv[3 * w + 1]) == cudaSuccess) fill += ms;
    if (cudaEventElapsedTime(&ms, t->waveEv[3 * w + 1], t->waveEv[3 * w + 2]) == cudaSuccess) trace += ms;
  }
  t->timing.msFill = fill; t->timing.msTrace = trace;
  if (cudaEventElapsedTime(&ms, t->ev[3], t->ev[4]) == cudaSuccess) t->timing.msEmit = ms;
  if (cudaEventElapsedTime(&ms, t->ev[0], t->ev[4]) == cudaSuccess) t->timing.msTotal = ms;
  cudaGetLastError();
}

static int submit_guided(bgpu_ctx *ctx, const bgpu_scorefn *fn, const bgpu_params *p, const bgpu_batch *b, bgpu_ticket t) {
  const uint32_t n = b->nJobs;
  if (!b->qOff || !b->tOff || !b->guideOff || (n && (!b->qBases || (!b->tBases && !b->tRefOff)))) { ctx->err = "null batch arrays"; return BGPU_E_INVALID; }
  if (fn->kind == BGPU_FN_QUALITY && !b->qual) { ctx->err = "BGPU_FN_QUALITY needs batch.qual"; return BGPU_E_INVALID; }
  RC(check_ids_tracks(ctx, fn, p, b));
  const uint64_t totQ = b->qOff[n], totT = b->tOff[n], totG = b->guideOff[n];
  fill_score_params(t->sp, fn, p);
  // what this ticket is going to ask for besides the traceback pool (sequences, band table, runs, job tables, results)
  t->dev.hint = 18 * totQ + 7 * totT + 12 * totG + (512ull + 16 * ROW_PAD) * n + (1u << 16);
  t->pin.hint = 4 * totQ + totT + 12 * totG + 320ull * n + (1u << 16);
  BatchDev &B = t->B;
  B.nJobs = n;
  uint64_t *d_qOff, *d_tOff, *d_gOff; uint8_t *d_q, *d_t, *d_tc, *d_qc, *d_qual = nullptr; bgpu_block *d_guide; int32_t *d_band = nullptr;
  // tc / qc / qual are read by unchecked 16-byte-aligned bulk copies that start up to a window before and end up to a
  // window behind a job's bytes: BYTE_PAD bytes on either side
  RC(talloc_dev(ctx, t, &d_q, totQ + 16)); RC(talloc_dev(ctx, t, &d_t, totT + 16)); RC(talloc_dev(ctx, t, &d_tc, totT + 2 * BYTE_PAD));
  RC(talloc_dev(ctx, t, &d_qc, totQ + 2 * BYTE_PAD));
  RC(talloc_dev(ctx, t, &d_qOff, n + 1)); RC(talloc_dev(ctx, t, &d_tOff, n + 1)); RC(talloc_dev(ctx, t, &d_gOff, n + 1));
  RC(talloc_dev(ctx, t, &d_guide, totG + 1));
  if (b->qual) { RC(talloc_dev(ctx, t, &d_qual, totQ + 2 * BYTE_PAD)); d_qual += BYTE_PAD; }
  if (b->band) RC(talloc_dev(ctx, t, &d_band, n));
  // prep writes tc only inside [tStart, tEnd) of each job, the fill kernels also stage the boundary column t' = 0 and the
  // columns past the guide's end: those bytes must be valid codes (0), not whatever the cached allocation last held
  CK(cudaMemsetAsync(d_tc, 0, totT + 2 * BYTE_PAD, ctx->stream));
  CK(cudaMemsetAsync(d_qc, 0, totQ + 2 * BYTE_PAD, ctx->stream));      // likewise the coded query outside [qStart, qEnd)
  d_tc += BYTE_PAD; d_qc += BYTE_PAD;
  // the phase gates keep LARGE tickets of concurrent contexts pipelined (copy in / compute / copy out); small tickets
  // (the candidates of a few reads) would only pay their host round trips
  t->gated = totQ + (b->tRefOff ? 0 : totT) + (b->guidePacked ? 3 : sizeof(bgpu_block)) * totG > (32u << 20);
  if (t->gated) { gate(ctx->device, GATE_H2D).acquire(); t->holdsH2D = true; }   // until the uploads below are done
  RC(upload(ctx, t, d_q, b->qBases, totQ));
  RC(upload(ctx, t, d_qOff, b->qOff, sizeof(uint64_t) * (n + 1)));
  RC(upload(ctx, t, d_tOff, b->tOff, sizeof(uint64_t) * (n + 1)));
  if (b->tRefOff) {           // targets are windows of the reference resident on this device: 8 bytes per job over PCIe
    const uint8_t *refD; uint64_t refN;
    { std::lock_guard<std::mutex> lr(g_refMu); refD = ctx->device < 64 ? g_ref[ctx->device].d : nullptr; refN = ctx->device < 64 ? g_ref[ctx->device].n : 0; }
    if (!refD) { ctx->err = "batch.tRefOff without bgpu_set_reference on this device"; return BGPU_E_INVALID; }
    uint64_t *d_refOff = nullptr; uint8_t *d_rc = nullptr;
    RC(talloc_dev(ctx, t, &d_refOff, std::max<uint32_t>(n, 1))); RC(upload(ctx, t, d_refOff, b->tRefOff, sizeof(uint64_t) * n));
    if (b->tRefRc) { RC(talloc_dev(ctx, t, &d_rc, (size_t)n + 16)); RC(upload(ctx, t, d_rc, b->tRefRc, n)); }
    launch_gather_reference(n, refD, refN, d_refOff, d_rc, d_tOff, d_t, ctx->stream);
  } else {
    RC(upload(ctx, t, d_t, b->tBases, totT));
  }
  RC(upload(ctx, t, d_gOff, b->guideOff, sizeof(uint64_t) * (n + 1)));
  if (b->guidePacked) {       // three bytes per block over PCIe, expanded into Block form on the device
    uint8_t *d_packed = nullptr; uint32_t *d_wide = nullptr;
    RC(talloc_dev(ctx, t, &d_packed, 3 * totG + 16)); RC(talloc_dev(ctx, t, &d_wide, 4 * b->nGuideWide + 4));
    RC(upload(ctx, t, d_packed, b->guidePacked, 3 * totG));
    if (b->nGuideWide) RC(upload(ctx, t, d_wide, b->guideWide, sizeof(uint32_t) * 4 * b->nGuideWide));
    launch_unpack_guide(n, d_gOff, d_packed, d_wide, b->nGuideWide, d_guide, ctx->stream);
  } else {
    RC(upload(ctx, t, d_guide, b->guide, sizeof(bgpu_block) * totG));
  }
  if (b->qual) RC(upload(ctx, t, d_qual, b->qual, totQ));
  if (b->band) RC(upload(ctx, t, d_band, b->band, sizeof(int32_t) * n));
  B.q = d_q; B.qOff = d_qOff; B.t = d_t; B.tc = d_tc; B.qc = d_qc; B.tOff = d_tOff; B.qual = d_qual; B.guide = d_guide; B.guideOff = d_gOff; B.band = d_band;
  RC(upload_ids_tracks(ctx, t, fn, b, totQ));
  // capacities from sequence lengths (upper bounds of the guide extents)
  uint64_t *h_off = nullptr;
  RC(talloc_pin(ctx, t, &h_off, 3 * (size_t)n + 3));
  uint64_t rowTot = 0, dbTot = 0, runTot = 0, poolEst = 0;
  const uint64_t rowsPerBlock = t->sp.affine ? 16 : 4;
  for (uint32_t i = 0; i < n; i++) {
    const uint64_t ql = b->qOff[i + 1] - b->qOff[i], tl = b->tOff[i + 1] - b->tOff[i];
    h_off[i] = rowTot; h_off[n + i] = dbTot; h_off[2 * (size_t)n + i] = runTot;
    rowTot += ql + 1 + 2 * ROW_PAD; dbTot += (ql + tl + 1) / 64 + 2; runTot += ql + tl + 2;
    // traceback bytes this job is expected to reserve: d-blocks x rows per block x words per row (window of the band plus
    // drift and class quantisation) -- an estimate, the planner kernels check the real sum against the pool
    const int64_t bd = std::max<int64_t>(b->band ? b->band[i] : p->band, 0);
    poolEst += ((((ql + tl + 1) / 64 + 2) * rowsPerBlock * (uint64_t)(bd + 24 + (bd > 40 ? 16 : 0)) * 4ull) + 255ull) & ~255ull;
  }
  RC(talloc_dev(ctx, t, &t->d_rowOff, 3 * (size_t)n + 3));
  t->d_dblkOff = t->d_rowOff + n; t->d_runOff = t->d_rowOff + 2 * (size_t)n;
  CK(cudaMemcpyAsync(t->d_rowOff, h_off, sizeof(uint64_t) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
  // the stream leaves the H2D gate as soon as the last input byte has landed
  if (t->holdsH2D) {
    if (gate(ctx->device, GATE_H2D).width > 0) CK(cudaLaunchHostFunc(ctx->stream, gate_release_cb, &gate(ctx->device, GATE_H2D)));
    t->holdsH2D = false;
  }
  RC(talloc_dev(ctx, t, &B.geom, n)); RC(talloc_dev(ctx, t, &B.rows, rowTot + 1)); RC(talloc_dev(ctx, t, &B.dblk, dbTot + 1));
  RC(talloc_dev(ctx, t, &B.dmin, dbTot + 1)); RC(talloc_dev(ctx, t, &B.dmax, dbTot + 1)); RC(talloc_dev(ctx, t, &B.runs, runTot + 1));
  RC(talloc_dev(ctx, t, &t->d_blockOff, 3 * (size_t)n + 3));
  t->d_listOff = t->d_blockOff + n; t->d_gapOff = t->d_blockOff + 2 * (size_t)n;
  RC(talloc_dev(ctx, t, &t->d_totals, 4)); RC(talloc_dev(ctx, t, &t->d_results, n));
  RC(talloc_pin(ctx, t, &t->h_geom, n)); RC(talloc_pin(ctx, t, &t->h_totals, 4)); RC(talloc_pin(ctx, t, &t->h_results, n));
  static const bool hostPlan = [] { const char *e = getenv("BGPU_HOST_PLAN"); return e && *e && *e != '0'; }();
  if (hostPlan || n == 0 || poolEst > ctx->arrowPoolCap) return enqueue_guided(ctx, t, true);   // multi-wave: planned on the host
  // ---- asynchronous path
  t->fast = true;
  t->poolBytes = poolEst;
  uint8_t *arrows = nullptr;
  RC(talloc_dev(ctx, t, &t->d_order, 2 * (size_t)n + 64)); RC(talloc_dev(ctx, t, &t->d_arrowOff, n));
  t->nCounters = 16;
  RC(talloc_dev(ctx, t, &t->d_counters, t->nCounters));
  RC(talloc_dev(ctx, t, &t->d_plan, 1)); RC(talloc_dev(ctx, t, &t->d_planScratch, plan_scratch_words(n)));
  RC(talloc_pin(ctx, t, &t->h_planInit, 1)); RC(talloc_pin(ctx, t, &t->h_plan, 1));
  RC(talloc_dev(ctx, t, &arrows, std::max<uint64_t>(poolEst, 16)));
  B.arrows = arrows; B.arrowOff = t->d_arrowOff; B.order = t->d_order; B.counters = t->d_counters;
  memset(t->h_planInit, 0, sizeof(PlanHead));
  // result arena sized from the batch (a block needs a matching base, gap runs sit between blocks): speculative when that
  // is small enough to keep around, else sized exactly by bgpu_collect once the counts are known
  t->arenaCap[0] = totQ / 4 + 2ull * n + 16; t->arenaCap[1] = t->arenaCap[0] + n; t->arenaCap[2] = 2 * t->arenaCap[0];
  const uint64_t arenaBytes = arena_bytes(t, t->arenaCap[0], t->arenaCap[1], t->arenaCap[2]);
  if (arenaBytes <= (64u << 20)) {
    RC(alloc_arena(ctx, t, t->arenaCap[0], t->arenaCap[1], t->arenaCap[2]));
    t->arenaInline = arenaBytes <= (1u << 20);
    for (int k = 0; k < 3; k++) t->h_planInit->caps[k] = t->arenaCap[k];
  } else {
    for (int k = 0; k < 3; k++) t->h_planInit->caps[k] = ~0ull;
  }
  t->waves.assign(1, Wave{});
  t->waveEv.resize(3);
  for (auto &e : t->waveEv) CK(cudaEventCreate(&e));
  return enqueue_guided_fast(ctx, t, true);
}


// ---- kernel schedule of a KBandAlign / SWAlign ticket ----
static int enqueue_dense(bgpu_ctx *ctx, bgpu_ticket t, bool firstRun) {
  cudaStream_t s = ctx->stream;
  CK(cudaEventRecord(t->ev[0], s));
  launch_dense_prep(t->B, t->sp, t->dargs, t->d_dblkOff /* row-buffer offsets */, t->d_runOff, s);
  t->timing.kernelLaunches = 1;
  CK(cudaEventRecord(t->ev[1], s));
  if (firstRun && t->denseSmall) {
    // batches of small matrices (the ~80-cell AffineKBandAlign gap fills of -alignContigs, the SWAlign fills of SDPAlign): every
    // job is dispatched in submission order -- the kernels skip the ones prep flagged -- with the traceback offsets bgpu_submit
    // laid out from the lengths alone: nothing is read back, the host does not wait for the device
    const uint32_t n = t->nJobs;
    Wave w{};
    w.begin[0] = 0; w.count[0] = n; w.traceBegin = 0; w.traceCount = n;
    t->waves.push_back(w);
    RC(talloc_dev(ctx, t, &t->d_order, std::max<uint32_t>(n, 1)));
    RC(talloc_dev(ctx, t, &t->d_arrowOff, std::max<uint32_t>(n, 1)));
    t->nCounters = 16;
    RC(talloc_dev(ctx, t, &t->d_counters, t->nCounters));
    uint8_t *arrows = nullptr;
    RC(talloc_dev(ctx, t, &arrows, std::max<size_t>(t->denseWaveBytes, 16)));
    t->B.arrows = arrows; t->B.arrowOff = t->d_arrowOff; t->dargs.arrowOff = t->d_arrowOff;
    uint32_t *h_order = nullptr;
    RC(talloc_pin(ctx, t, &h_order, std::max<uint32_t>(n, 1)));
    for (uint32_t i = 0; i < n; i++) h_order[i] = i;
    CK(cudaMemcpyAsync(t->d_order, h_order, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(t->d_arrowOff, t->h_aoffFast, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    t->waveEv.resize(3);
    for (auto &e : t->waveEv) CK(cudaEventCreate(&e));
    uint64_t cells = 0;
    for (uint32_t i = 0; i < n; i++) cells += t->h_cellsMetric[i];     // refused jobs are taken out again by bgpu_collect
    t->timing.cells = cells; t->timing.fillCells = cells;
  } else if (firstRun) {
    CK(cudaMemcpyAsync(t->h_geom, t->B.geom, sizeof(JobGeom) * t->nJobs, cudaMemcpyDeviceToHost, s));
    CK(wait_stream(ctx));
    const uint32_t n = t->nJobs;
    std::vector<uint32_t> idx; idx.reserve(n);
    uint64_t maxBytes = 0;
    for (uint32_t i = 0; i < n; i++) if (t->h_geom[i].status == BGPU_JOB_OK) { idx.push_back(i); maxBytes = std::max(maxBytes, t->h_arrowBytes[i]); }
    // largest matrices first, so the queue's tail is made of small jobs; batches of small gap fills (every matrix under
    // 64 KiB: the ~80-cell AffineKBandAlign jobs of -alignContigs) keep their order, a sort would cost more than it saves
    if (maxBytes > (64u << 10)) {
      std::vector<std::pair<uint64_t, uint32_t>> keyed; keyed.reserve(idx.size());
      for (uint32_t i : idx) keyed.emplace_back(~t->h_arrowBytes[i], i);      // ascending (~bytes, index) = bytes descending, index ascending
      std::sort(keyed.begin(), keyed.end());
      for (size_t i = 0; i < keyed.size(); i++) idx[i] = keyed[i].second;
    }
    std::vector<uint64_t> arrowOff(n, 0);
    std::vector<uint32_t> order; order.reserve(idx.size());
    size_t maxWaveBytes = 0, i0 = 0;
    while (i0 < idx.size()) {
      size_t bytes = 0, i1 = i0;
      while (i1 < idx.size()) {
        const uint64_t ab = (t->h_arrowBytes[idx[i1]] + 15) & ~15ull;
        if (i1 > i0 && bytes + ab > ctx->arrowPoolCap) break;
        arrowOff[idx[i1]] = bytes; bytes += ab; i1++;
      }
      maxWaveBytes = std::max(maxWaveBytes, bytes);
      Wave w{};
      w.begin[0] = (uint32_t)order.size();
      for (size_t i = i0; i < i1; i++) order.push_back(idx[i]);
      w.count[0] = (uint32_t)(i1 - i0); w.traceBegin = w.begin[0]; w.traceCount = w.count[0];
      t->waves.push_back(w);
      i0 = i1;
    }
    RC(talloc_dev(ctx, t, &t->d_order, std::max<size_t>(order.size(), 1)));
    RC(talloc_dev(ctx, t, &t->d_arrowOff, std::max<uint32_t>(n, 1)));
    t->nCounters = (uint32_t)t->waves.size() * 8 + 8;
    RC(talloc_dev(ctx, t, &t->d_counters, t->nCounters));
    uint8_t *arrows = nullptr;
    RC(talloc_dev(ctx, t, &arrows, std::max<size_t>(maxWaveBytes, 16)));
    t->B.arrows = arrows; t->B.arrowOff = t->d_arrowOff; t->dargs.arrowOff = t->d_arrowOff;
    uint32_t *h_order = nullptr; uint64_t *h_aoff = nullptr;
    RC(talloc_pin(ctx, t, &h_order, std::max<size_t>(order.size(), 1)));
    RC(talloc_pin(ctx, t, &h_aoff, std::max<uint32_t>(n, 1)));
    memcpy(h_order, order.data(), order.size() * sizeof(uint32_t));
    memcpy(h_aoff, arrowOff.data(), n * sizeof(uint64_t));
    CK(cudaMemcpyAsync(t->d_order, h_order, order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(t->d_arrowOff, h_aoff, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    t->waveEv.resize(t->waves.size() * 3);
    for (auto &e : t->waveEv) CK(cudaEventCreate(&e));
    uint64_t cells = 0;
    for (uint32_t i : idx) cells += t->h_cellsMetric[i];
    t->timing.cells = cells; t->timing.fillCells = cells;
  }
  CK(cudaMemsetAsync(t->d_counters, 0, sizeof(uint32_t) * t->nCounters, s));
  for (size_t w = 0; w < t->waves.size(); w++) {
    const Wave &W = t->waves[w];
    CK(cudaEventRecord(t->waveEv[3 * w], s));
    if (W.count[0]) { launch_dense_fill(t->B, t->sp, t->dargs, t->d_order + W.begin[0], W.count[0], t->d_counters + 8 * w, ctx->nSM, s); t->timing.kernelLaunches++; }
    CK(cudaEventRecord(t->waveEv[3 * w + 1], s));
    if (W.traceCount) { launch_dense_trace(t->B, t->sp, t->dargs, t->d_order + W.traceBegin, W.traceCount, s); t->timing.kernelLaunches++; }
    CK(cudaEventRecord(t->waveEv[3 * w + 2], s));
  }
  launch_scan_counts(t->B, t->d_blockOff, t->d_listOff, t->d_gapOff, t->d_totals, s);
  t->timing.kernelLaunches++;
  CK(cudaEventRecord(t->ev[3], s));
  CK(cudaGetLastError());
  return BGPU_OK;
}

namespace bgpu { void kbounded_host(uint32_t, uint32_t, uint32_t, uint32_t &, uint32_t &); }

static int submit_dense(bgpu_ctx *ctx, const bgpu_scorefn *fn, const bgpu_params *p, const bgpu_batch *b, bgpu_ticket t) {
  const uint32_t n = b->nJobs;
  if (!b->qOff || !b->tOff || (n && (!b->qBases || !b->tBases))) { ctx->err = "null batch arrays"; return BGPU_E_INVALID; }
  if (fn->kind == BGPU_FN_QUALITY && !b->qual) { ctx->err = "BGPU_FN_QUALITY needs batch.qual"; return BGPU_E_INVALID; }
  RC(check_ids_tracks(ctx, fn, p, b));
  const uint64_t totQ = b->qOff[n], totT = b->tOff[n];
  fill_score_params(t->sp, fn, p);
  t->dev.hint = 8 * totQ + 32 * totT + 512ull * n + (1u << 16);
  t->pin.hint = 4 * totQ + 4 * totT + 320ull * n + (1u << 16);
  t->dense = true;
  t->dargs.algo = p->algo; t->dargs.defaultBand = p->band; t->dargs.bndIns = p->bndIns; t->dargs.bndDel = p->bndDel;
  t->dargs.hpInsOpen = p->hpInsOpen; t->dargs.hpInsExtend = p->hpInsExtend; t->dargs.insOpen = p->insOpen; t->dargs.insExtend = p->insExtend;
  BatchDev &B = t->B;
  B.nJobs = n;
  uint64_t *d_qOff, *d_tOff; uint8_t *d_q, *d_t, *d_tc, *d_qual = nullptr; int32_t *d_band = nullptr;
  RC(talloc_dev(ctx, t, &d_q, totQ + 16)); RC(talloc_dev(ctx, t, &d_t, totT + 16)); RC(talloc_dev(ctx, t, &d_tc, totT + 16));
  RC(talloc_dev(ctx, t, &d_qOff, n + 1)); RC(talloc_dev(ctx, t, &d_tOff, n + 1));
  if (b->qual) RC(talloc_dev(ctx, t, &d_qual, totQ + 16));
  const bool banded = p->algo == BGPU_KBAND || p->algo == BGPU_AFFINE_KBAND;
  if (b->band && banded) RC(talloc_dev(ctx, t, &d_band, std::max<uint32_t>(n, 1)));
  RC(upload(ctx, t, d_q, b->qBases, totQ)); RC(upload(ctx, t, d_t, b->tBases, totT));
  RC(upload(ctx, t, d_qOff, b->qOff, sizeof(uint64_t) * (n + 1)));
  RC(upload(ctx, t, d_tOff, b->tOff, sizeof(uint64_t) * (n + 1)));
  if (b->qual) RC(upload(ctx, t, d_qual, b->qual, totQ));
  if (d_band) RC(upload(ctx, t, d_band, b->band, sizeof(int32_t) * n));
  B.q = d_q; B.qOff = d_qOff; B.t = d_t; B.tc = d_tc; B.tOff = d_tOff; B.qual = d_qual; B.guide = nullptr; B.guideOff = nullptr; B.band = d_band;
  RC(upload_ids_tracks(ctx, t, fn, b, totQ));
  uint64_t *h_off = nullptr;
  RC(talloc_pin(ctx, t, &h_off, 2 * (size_t)n + 2));
  t->h_arrowBytes.assign(n, 0); t->h_cellsMetric.assign(n, 0);
  RC(talloc_pin(ctx, t, &t->h_aoffFast, std::max<uint32_t>(n, 1)));
  uint64_t rbTot = 0, runTot = 0, maxBytes = 0, waveBytes = 0;
  for (uint32_t i = 0; i < n; i++) {
    const uint64_t ql = b->qOff[i + 1] - b->qOff[i], tl = b->tOff[i + 1] - b->tOff[i];
    uint32_t qb = (uint32_t)ql, tb = (uint32_t)tl;
    uint64_t bytes, cells;
    if (banded) {
      const int k = d_band ? b->band[i] : p->band;
      if (k >= 0) kbounded_host((uint32_t)tl, (uint32_t)ql, (uint32_t)k, tb, qb);
      bytes = k >= 0 ? ((uint64_t)qb + 1) * (2ull * (uint64_t)k + 1) : 16; cells = bytes;
    } else { bytes = (ql + 1) * (tl + 1); cells = bytes; }
    if (bytes > (1ull << 31)) bytes = 16;          // rejected by the prep kernel (matrix size is an int in the reference)
    t->h_arrowBytes[i] = bytes; t->h_cellsMetric[i] = cells;
    t->h_aoffFast[i] = waveBytes; waveBytes += (bytes + 15) & ~15ull; maxBytes = std::max(maxBytes, bytes);
    h_off[i] = rbTot; h_off[n + i] = runTot;
    rbTot += (p->algo == BGPU_AFFINE_KBAND ? 6 : 2) * ((uint64_t)tb + 2); runTot += ql + tl + 2;   // ping-pong rows (x3 matrices)
  }
  static const bool noFastDense = [] { const char *e = getenv("BGPU_DENSE_PLANNED"); return e && *e && *e != '0'; }();   // diagnostic: always plan on the host
  t->denseSmall = !noFastDense && maxBytes <= (64u << 10) && waveBytes <= ctx->arrowPoolCap;
  t->denseWaveBytes = waveBytes;
  RC(talloc_dev(ctx, t, &t->d_dblkOff, 2 * (size_t)n + 2));
  t->d_runOff = t->d_dblkOff + n;
  CK(cudaMemcpyAsync(t->d_dblkOff, h_off, sizeof(uint64_t) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
  RC(talloc_dev(ctx, t, &B.geom, std::max<uint32_t>(n, 1))); RC(talloc_dev(ctx, t, &B.rowBuf, rbTot + 4)); RC(talloc_dev(ctx, t, &B.runs, runTot + 1));
  RC(talloc_dev(ctx, t, &t->d_blockOff, 3 * (size_t)n + 3));
  t->d_listOff = t->d_blockOff + n; t->d_gapOff = t->d_blockOff + 2 * (size_t)n;
  RC(talloc_dev(ctx, t, &t->d_totals, 4)); RC(talloc_dev(ctx, t, &t->d_results, std::max<uint32_t>(n, 1)));
  RC(talloc_pin(ctx, t, &t->h_geom, std::max<uint32_t>(n, 1))); RC(talloc_pin(ctx, t, &t->h_totals, 4)); RC(talloc_pin(ctx, t, &t->h_results, std::max<uint32_t>(n, 1)));
  return enqueue_dense(ctx, t, true);
}

extern "C" int bgpu_submit(bgpu_ctx *ctx, const bgpu_scorefn *fn, const bgpu_params *p, const bgpu_batch *b, bgpu_ticket *out) {
  if (!ctx) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!fn || !p || !b || !out) { ctx->err = "null argument"; return BGPU_E_INVALID; }
  if (p->algo < BGPU_GUIDED || p->algo > BGPU_AFFINE_KBAND) { ctx->err = "unknown algo"; return BGPU_E_INVALID; }
  if (fn->kind < BGPU_FN_DISTANCE || fn->kind > BGPU_FN_IDS) { ctx->err = "unknown score function kind"; return BGPU_E_INVALID; }
  if (cudaSetDevice(ctx->device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return BGPU_E_CUDA; }
  const auto h0 = std::chrono::steady_clock::now();
  bgpu_ticket t = new bgpu_ticket_s();
  t->nJobs = b->nJobs; t->params = *p;
  for (auto &e : t->ev) cudaEventCreate(&e);
  int rc;
  if (p->algo == BGPU_GUIDED || p->algo == BGPU_AFFINE_GUIDED) rc = submit_guided(ctx, fn, p, b, t);
  else rc = submit_dense(ctx, fn, p, b, t);
  if (rc != BGPU_OK) {
    cudaStreamSynchronize(ctx->stream);
    if (t->holdsH2D) { gate(ctx->device, GATE_H2D).release(); t->holdsH2D = false; }
    if (t->holdsCompute) { gate(ctx->device, GATE_COMPUTE).release(); t->holdsCompute = false; }
    slab_release(ctx->devPool, t->dev); slab_release(ctx->pinPool, t->pin);
    for (auto &e : t->ev) cudaEventDestroy(e);
    for (auto &e : t->waveEv) cudaEventDestroy(e);
    delete t;
    return rc;
  }
  if (cudaEventCreateWithFlags(&t->evDone, cudaEventDisableTiming) == cudaSuccess) cudaEventRecord(t->evDone, ctx->stream);
  t->timing.msHostSubmit = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
  static const bool hostTrace = getenv("BGPU_HOST_TRACE") != nullptr;
  if (hostTrace && t->timing.msHostSubmit > 5.0)
    fprintf(stderr, "BGPU_HOST_TRACE submit %.2f ms: jobs %u, upload %.2f ms (%llu bytes), dev allocs %u pin allocs %u\n", t->timing.msHostSubmit,
            t->nJobs, t->msUpload, (unsigned long long)t->timing.h2dBytes, ctx->devPool.nAlloc, ctx->pinPool.nAlloc);
  *out = t;
  return BGPU_OK;
}

extern "C" int bgpu_query(bgpu_ctx *ctx, bgpu_ticket t) {
  if (!ctx || !t) return BGPU_E_INVALID;
  if (t->collected || !t->evDone) return 1;
  const cudaError_t e = cudaEventQuery(t->evDone);
  if (e == cudaSuccess) return 1;
  if (e == cudaErrorNotReady) { cudaGetLastError(); return 0; }
  cudaGetLastError();
  return BGPU_E_CUDA;
}

extern "C" int bgpu_submit_jobs(bgpu_ctx *ctx, const bgpu_scorefn *fn, const bgpu_params *p, const bgpu_job *jobs,
                                uint32_t nJobs, bgpu_ticket *out) {
  if (!ctx || !jobs) return BGPU_E_INVALID;
  std::vector<uint64_t> qOff(nJobs + 1, 0), tOff(nJobs + 1, 0), gOff(nJobs + 1, 0);
  bool anyQual = false, anyTrack[5] = {false, false, false, false, false};
  auto track = [](const bgpu_job &j, int k) { return k == 0 ? j.insQV : k == 1 ? j.delQV : k == 2 ? j.subQV : k == 3 ? j.delTag : j.subTag; };
  for (uint32_t i = 0; i < nJobs; i++) {
    qOff[i + 1] = qOff[i] + jobs[i].qLen; tOff[i + 1] = tOff[i] + jobs[i].tLen; gOff[i + 1] = gOff[i] + jobs[i].nGuide;
    anyQual |= jobs[i].qual != nullptr;
    for (int k = 0; k < 5; k++) anyTrack[k] |= track(jobs[i], k) != nullptr;
  }
  std::vector<uint8_t> tracks[5];
  for (int k = 0; k < 5; k++) if (anyTrack[k]) tracks[k].assign(qOff[nJobs] + 1, 0);
  std::vector<uint8_t> q(qOff[nJobs] + 1), tt(tOff[nJobs] + 1), qual(anyQual ? qOff[nJobs] + 1 : 0);
  std::vector<bgpu_block> g(gOff[nJobs] + 1);
  std::vector<int32_t> band(nJobs + 1);
  for (uint32_t i = 0; i < nJobs; i++) {
    if (jobs[i].qLen) memcpy(&q[qOff[i]], jobs[i].q, jobs[i].qLen);
    if (jobs[i].tLen) memcpy(&tt[tOff[i]], jobs[i].t, jobs[i].tLen);
    if (anyQual && jobs[i].qual && jobs[i].qLen) memcpy(&qual[qOff[i]], jobs[i].qual, jobs[i].qLen);
    if (jobs[i].nGuide) memcpy(&g[gOff[i]], jobs[i].guide, sizeof(bgpu_block) * jobs[i].nGuide);
    band[i] = jobs[i].band;
    for (int k = 0; k < 5; k++) if (anyTrack[k] && track(jobs[i], k) && jobs[i].qLen) memcpy(&tracks[k][qOff[i]], track(jobs[i], k), jobs[i].qLen);
  }
  bgpu_batch b{};
  b.nJobs = nJobs; b.qBases = q.data(); b.qOff = qOff.data(); b.tBases = tt.data(); b.tOff = tOff.data();
  b.qual = anyQual ? qual.data() : nullptr; b.guide = g.data(); b.guideOff = gOff.data(); b.band = band.data();
  b.insQV = anyTrack[0] ? tracks[0].data() : nullptr; b.delQV = anyTrack[1] ? tracks[1].data() : nullptr;
  b.subQV = anyTrack[2] ? tracks[2].data() : nullptr; b.delTag = anyTrack[3] ? tracks[3].data() : nullptr;
  b.subTag = anyTrack[4] ? tracks[4].data() : nullptr;
  return bgpu_submit(ctx, fn, p, &b, out);   // inputs are staged into pinned memory before this returns
}

static int ensure_arena(bgpu_ctx *ctx, bgpu_ticket t) {
  if (t->arenaReady) return BGPU_OK;
  CK(cudaMemcpyAsync(t->h_totals, t->d_totals, sizeof(uint64_t) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(wait_stream(ctx));
  for (int i = 0; i < 3; i++) t->totals[i] = t->h_totals[i];
  if (!t->dense) t->timing.fillCells = t->h_totals[3];
  RC(alloc_arena(ctx, t, t->totals[0], t->totals[1], t->totals[2]));
  return BGPU_OK;
}

extern "C" int bgpu_collect(bgpu_ctx *ctx, bgpu_ticket t, bgpu_result *results, bgpu_arena *arena) {
  if (!ctx || !t) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return BGPU_E_CUDA;
  cudaStream_t s = ctx->stream;
  const auto h0 = std::chrono::steady_clock::now();
  if (!t->collected) {
    bool copied = false;
    if (t->fast) {
      // everything bgpu_submit enqueued for THIS ticket (the context's stream may already hold the next ticket: a host
      // thread submits sub-batch i+1 before it collects sub-batch i); what follows runs on the copy stream for the same reason
      CK(t->evDone ? wait_event(t->evDone) : wait_stream(ctx));
      cudaStream_t cs = ctx->copyStream;
      if (t->h_plan->overflow & PLAN_OVF_ARROWS) {
        // the traceback pool was under-estimated (adversarial guides): plan this ticket on the host, in waves
        t->fast = false; t->arenaReady = false; t->arenaInline = false;
        for (auto &e : t->waveEv) cudaEventDestroy(e);
        t->waveEv.clear(); t->waves.clear();
        t->d_blocks = nullptr; t->d_runsOut = nullptr; t->h_runsOut = nullptr;
        RC(enqueue_guided(ctx, t, true));
      } else {
        for (int i = 0; i < 3; i++) t->totals[i] = t->h_plan->totals[i];
        t->timing.cells = t->h_plan->cells; t->timing.fillCells = t->h_totals[3];
        const bool emitted = t->arenaReady && !(t->h_plan->overflow & PLAN_OVF_ARENA);
        if (!emitted) {                                  // no speculative arena, or it was too small: exact sizes now
          t->fast = false;                               // (emit without the capacity check)
          RC(alloc_arena(ctx, t, t->totals[0], t->totals[1], t->totals[2]));
          t->arenaInline = false;
          for (int k = 0; k < 3; k++) t->h_planInit->caps[k] = ~0ull;   // a later bgpu_rerun emits into this exact-size arena
          RC(enqueue_emit(ctx, t, cs));
          t->fast = true;
        }
        if (!t->arenaInline) {
          struct Hold { Gate *g; Hold(Gate *x) : g(x) { if (g) g->acquire(); } ~Hold() { if (g) g->release(); } } hold(t->gated ? &gate(ctx->device, GATE_D2H) : nullptr);
          if (!emitted) CK(cudaMemcpyAsync(t->h_results, t->d_results, sizeof(bgpu_result) * t->nJobs, cudaMemcpyDeviceToHost, cs));
          RC(copy_arena(ctx, t, t->totals[0], t->totals[1], t->totals[2], cs));
          CK(cudaEventRecord(t->ev[5], cs));
          CK(wait_stream(ctx, cs));
        }
        copied = true;
      }
    }
    if (!copied) {
      RC(ensure_arena(ctx, t));
      RC(enqueue_emit(ctx, t));                          // a kernel: outside the D2H gate, which only covers the copies
      struct Hold { Gate *g; Hold(Gate *x) : g(x) { if (g) g->acquire(); } ~Hold() { if (g) g->release(); } } hold(t->gated || t->dense ? &gate(ctx->device, GATE_D2H) : nullptr);
      CK(cudaMemcpyAsync(t->h_results, t->d_results, sizeof(bgpu_result) * t->nJobs, cudaMemcpyDeviceToHost, s));
      RC(copy_arena(ctx, t, t->totals[0], t->totals[1], t->totals[2]));
      CK(cudaEventRecord(t->ev[5], s));
      CK(wait_stream(ctx));
      if (t->dense && t->denseSmall) {                   // the cell metric counted every submitted job: take the refused ones out
        uint64_t cells = t->timing.cells;
        for (uint32_t i = 0; i < t->nJobs; i++) if (t->h_results[i].status != BGPU_JOB_OK) cells -= std::min(cells, t->h_cellsMetric[i]);
        t->timing.cells = cells; t->timing.fillCells = cells;
      }
    }
    t->timing.d2hBytes = sizeof(bgpu_result) * (uint64_t)t->nJobs + arena_bytes(t, t->totals[0], t->totals[1], t->totals[2]);
    gather_timing(t);
    if (g_trace) {
      auto at = [&](cudaEvent_t e) { float ms = -1; if (e) cudaEventElapsedTime(&ms, g_refEvent, e); return ms; };
      fprintf(stderr, "BGPU_TRACE ctx %p jobs %u start %.2f prepEnd %.2f fill0 %.2f fillEnd %.2f traceEnd %.2f scanEnd %.2f emitEnd %.2f d2hEnd %.2f\n",
              (void *)ctx, t->nJobs, at(t->ev[0]), at(t->ev[1]), t->waveEv.empty() ? -1.f : at(t->waveEv[0]),
              t->waveEv.empty() ? -1.f : at(t->waveEv[t->waveEv.size() - 2]), t->waveEv.empty() ? -1.f : at(t->waveEv.back()),
              at(t->ev[3]), at(t->ev[4]), at(t->ev[5]));
      cudaGetLastError();
    }
    t->collected = true;
  }
  if (results && t->nJobs) memcpy(results, t->h_results, sizeof(bgpu_result) * t->nJobs);
  t->timing.msHostCollect = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
  if (arena) {
    arena->blocks = t->h_blocks; arena->nBlocks = t->totals[0];
    arena->gapCounts = t->h_gapCounts; arena->nGapLists = t->totals[1];
    arena->gaps = t->h_gaps; arena->nGaps = t->totals[2];
    arena->runs = t->h_runsOut; arena->nRuns = t->h_runsOut ? t->totals[0] + t->totals[2] : 0;
  }
  return BGPU_OK;
}

// per-job op / column counts of a collected guided ticket (one kernel, cached in the ticket)
static int fmt_counts(bgpu_ctx *ctx, bgpu_ticket t) {
  if (t->d_fmtOps) return BGPU_OK;
  RC(talloc_dev(ctx, t, &t->d_fmtOps, (size_t)t->nJobs + 1)); RC(talloc_dev(ctx, t, &t->d_fmtCols, (size_t)t->nJobs + 1));
  launch_fmt_count(t->B, t->d_fmtOps, t->d_fmtCols, ctx->stream);
  return BGPU_OK;
}

static int cigar_impl(bgpu_ctx *ctx, bgpu_ticket t, const uint32_t *clips, const uint8_t *tStrand, const uint32_t **ops,
                      const uint64_t **cigarOff) {
  if (!t->collected) { ctx->err = "bgpu_cigar needs a collected ticket"; return BGPU_E_BUSY; }
  if (t->dense) { ctx->err = "bgpu_cigar: GuidedAlign / AffineGuidedAlign tickets only"; return BGPU_E_INVALID; }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return BGPU_E_CUDA;
  cudaStream_t s = ctx->stream;
  const uint32_t n = t->nJobs;
  RC(fmt_counts(ctx, t));
  uint32_t *d_clips = nullptr; uint8_t *d_strand = nullptr;
  if (clips && n) { RC(talloc_dev(ctx, t, &d_clips, 4 * (size_t)n)); RC(upload(ctx, t, d_clips, clips, sizeof(uint32_t) * 4 * n)); }
  if (tStrand && n) { RC(talloc_dev(ctx, t, &d_strand, (size_t)n)); RC(upload(ctx, t, d_strand, tStrand, n)); }
  uint64_t *d_off = nullptr, *d_core = nullptr, *d_tot = nullptr, *h_off = nullptr, *h_tot = nullptr;
  RC(talloc_dev(ctx, t, &d_off, (size_t)n + 1)); RC(talloc_dev(ctx, t, &d_core, (size_t)n + 1)); RC(talloc_dev(ctx, t, &d_tot, 2));
  RC(talloc_pin(ctx, t, &h_off, (size_t)n + 1)); RC(talloc_pin(ctx, t, &h_tot, 2));
  launch_fmt_scan(n, t->d_fmtOps, d_clips, d_off, d_core, d_tot, s);      // offsets on the device: only the two totals come back
  CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(uint64_t) * 2, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h_off, d_off, sizeof(uint64_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost, s));
  CK(wait_stream(ctx));
  uint32_t *d_ops = nullptr, *h_ops = nullptr, *d_tc = nullptr, *d_tp = nullptr;
  RC(talloc_dev(ctx, t, &d_ops, h_tot[0] + 1)); RC(talloc_pin(ctx, t, &h_ops, h_tot[0] + 1));
  RC(talloc_dev(ctx, t, &d_tc, h_tot[1] + 1)); RC(talloc_dev(ctx, t, &d_tp, h_tot[1] + 1));
  launch_fmt_cigar(t->B, d_off, d_core, d_tc, d_tp, d_ops, d_clips, d_strand, s);
  CK(cudaMemcpyAsync(h_ops, d_ops, sizeof(uint32_t) * h_tot[0], cudaMemcpyDeviceToHost, s));
  CK(wait_stream(ctx));
  CK(cudaGetLastError());
  *ops = h_ops; *cigarOff = h_off;
  return BGPU_OK;
}

extern "C" int bgpu_cigar(bgpu_ctx *ctx, bgpu_ticket t, const uint32_t **ops, const uint64_t **cigarOff) {
  if (!ctx || !t || !ops || !cigarOff) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!t->h_cigarOff) {
    const uint32_t *o = nullptr; const uint64_t *f = nullptr;
    RC(cigar_impl(ctx, t, nullptr, nullptr, &o, &f));
    t->h_cigar = const_cast<uint32_t *>(o); t->h_cigarOff = const_cast<uint64_t *>(f);
  }
  *ops = t->h_cigar; *cigarOff = t->h_cigarOff;
  return BGPU_OK;
}

// ComputeAlignmentScore of every alignment of a collected guided ticket under another score function (StoreMapQVs)
extern "C" int bgpu_rescore(bgpu_ctx *ctx, bgpu_ticket t, const bgpu_scorefn *fn, int useAffinePenalty, int32_t *scores) {
  if (!ctx || !t || !fn || (!scores && t->nJobs)) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!t->collected) { ctx->err = "bgpu_rescore needs a collected ticket"; return BGPU_E_BUSY; }
  if (t->dense) { ctx->err = "bgpu_rescore: GuidedAlign / AffineGuidedAlign tickets only"; return BGPU_E_INVALID; }
  if (fn->kind != BGPU_FN_DISTANCE) { ctx->err = "bgpu_rescore takes a DistanceMatrixScoreFunction (what StoreMapQVs builds, Blasr.cpp:2768-2771)"; return BGPU_E_INVALID; }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return BGPU_E_CUDA;
  const uint32_t n = t->nJobs;
  if (!n) return BGPU_OK;
  ScoreParams sp; fill_score_params(sp, fn, &t->params);
  int32_t *d_out = nullptr, *h_out = nullptr;
  RC(talloc_dev(ctx, t, &d_out, n)); RC(talloc_pin(ctx, t, &h_out, n));
  cudaStream_t s = ctx->stream;
  launch_rescore(t->B, sp, t->d_blockOff, t->d_listOff, t->d_gapOff, useAffinePenalty ? 1 : 0, d_out, s);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h_out, d_out, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  CK(wait_stream(ctx));
  memcpy(scores, h_out, sizeof(int32_t) * n);
  return BGPU_OK;
}

extern "C" int bgpu_cigar_clipped(bgpu_ctx *ctx, bgpu_ticket t, const uint32_t *clips, const uint8_t *tStrand, const uint32_t **ops,
                                  const uint64_t **cigarOff) {
  if (!ctx || !t || !ops || !cigarOff) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return cigar_impl(ctx, t, clips, tStrand, ops, cigarOff);
}

extern "C" int bgpu_strings(bgpu_ctx *ctx, bgpu_ticket t, const char **text, const char **align, const char **query,
                            const uint64_t **strOff) {
  if (!ctx || !t || !text || !align || !query || !strOff) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!t->collected) { ctx->err = "bgpu_strings needs a collected ticket"; return BGPU_E_BUSY; }
  if (t->dense) { ctx->err = "bgpu_strings: GuidedAlign / AffineGuidedAlign tickets only"; return BGPU_E_INVALID; }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return BGPU_E_CUDA;
  if (!t->h_strOff) {
    cudaStream_t s = ctx->stream;
    const uint32_t n = t->nJobs;
    RC(fmt_counts(ctx, t));
    uint64_t *d_off = nullptr, *d_tot = nullptr, *h_off = nullptr, *h_tot = nullptr;
    RC(talloc_dev(ctx, t, &d_off, (size_t)n + 1)); RC(talloc_dev(ctx, t, &d_tot, 2));
    RC(talloc_pin(ctx, t, &h_off, (size_t)n + 1)); RC(talloc_pin(ctx, t, &h_tot, 2));
    launch_fmt_scan(n, t->d_fmtCols, nullptr, d_off, nullptr, d_tot, s);
    CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(uint64_t) * 2, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h_off, d_off, sizeof(uint64_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost, s));
    CK(wait_stream(ctx));
    const size_t tot = h_tot[0];
    char *d_str = nullptr, *h_str = nullptr;
    RC(talloc_dev(ctx, t, &d_str, 3 * tot + 16)); RC(talloc_pin(ctx, t, &h_str, 3 * tot + 16));
    launch_fmt_strings(t->B, d_off, d_str, d_str + tot, d_str + 2 * tot, s);
    CK(cudaMemcpyAsync(h_str, d_str, 3 * tot, cudaMemcpyDeviceToHost, s));
    CK(wait_stream(ctx));
    CK(cudaGetLastError());
    t->h_str = h_str; t->h_strOff = h_off; t->strTotal = tot;
  }
  *text = t->h_str; *align = t->h_str + t->strTotal; *query = t->h_str + 2 * t->strTotal; *strOff = t->h_strOff;
  return BGPU_OK;
}

extern "C" int bgpu_rerun(bgpu_ctx *ctx, bgpu_ticket t) {
  if (!ctx || !t) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!t->collected) { ctx->err = "bgpu_rerun needs a collected ticket"; return BGPU_E_BUSY; }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return BGPU_E_CUDA;
  if (t->dense) { RC(enqueue_dense(ctx, t, false)); RC(enqueue_emit(ctx, t)); }
  else if (t->fast) RC(enqueue_guided_fast(ctx, t, false));   // emit included (the arena exists by now)
  else { RC(enqueue_guided(ctx, t, false)); RC(enqueue_emit(ctx, t)); }
  CK(wait_stream(ctx));
  gather_timing(t);
  return BGPU_OK;
}

extern "C" int bgpu_timing_of(bgpu_ctx *ctx, bgpu_ticket t, bgpu_timing *out) {
  if (!ctx || !t || !out) return BGPU_E_INVALID;
  *out = t->timing;
  out->devAllocs = ctx->devPool.nAlloc; out->pinAllocs = ctx->pinPool.nAlloc;
  return BGPU_OK;
}

extern "C" int bgpu_release(bgpu_ctx *ctx, bgpu_ticket t) {
  if (!ctx || !t) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  // a collected ticket has nothing in flight (the context's stream may hold LATER tickets: do not wait for those)
  if (!t->collected || !t->fast || !t->evDone) cudaStreamSynchronize(ctx->stream);
  slab_release(ctx->devPool, t->dev); slab_release(ctx->pinPool, t->pin);
  for (auto &e : t->ev) cudaEventDestroy(e);
  for (auto &e : t->waveEv) cudaEventDestroy(e);
  if (t->evDone) cudaEventDestroy(t->evDone);
  if (ctx->lastSync == t) ctx->lastSync = nullptr;
  delete t;
  return BGPU_OK;
}

extern "C" int bgpu_align(bgpu_ctx *ctx, const bgpu_scorefn *fn, const bgpu_params *p, const bgpu_batch *b,
                          bgpu_result *results, bgpu_arena *arena) {
  if (!ctx) return BGPU_E_INVALID;
  if (ctx->lastSync) bgpu_release(ctx, ctx->lastSync);
  bgpu_ticket t = nullptr;
  int rc = bgpu_submit(ctx, fn, p, b, &t);
  if (rc) return rc;
  rc = bgpu_collect(ctx, t, results, arena);
  if (rc) { bgpu_release(ctx, t); return rc; }
  ctx->lastSync = t;
  return BGPU_OK;
}


// ---- SDPAlign: synchronous, one thread per job (bgpu_sdp.cu) ----
extern "C" int bgpu_sdp_align(bgpu_ctx *ctx, const bgpu_scorefn *fn, const bgpu_sdp_params *p, const bgpu_batch *b,
                              bgpu_result *results, bgpu_arena *arena) {
  if (!ctx || !fn || !p || !b || (!results && b->nJobs) || !arena) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return BGPU_E_CUDA;
  memset(arena, 0, sizeof *arena);
  const uint32_t n = b->nJobs;
  if (n == 0) return BGPU_OK;
  if (!b->qOff || !b->tOff || !b->qBases || !b->tBases) { ctx->err = "null batch arrays"; return BGPU_E_INVALID; }
  if (fn->kind != BGPU_FN_DISTANCE) { ctx->err = "bgpu_sdp_align takes a DistanceMatrixScoreFunction (what blasr passes, Blasr.cpp:1716)"; return BGPU_E_INVALID; }
  if (p->wordSize < 1 || p->wordSize > 15 || (p->alignType != BGPU_LOCAL && p->alignType != BGPU_GLOBAL) || p->recurse < 0 || p->recurse > 8) {
    ctx->err = "bgpu_sdp_align: wordSize 1..15, alignType Local / Global, recurse 0..8"; return BGPU_E_INVALID;
  }
  if (!ctx->sdpStackSet) {       // sdp_align recurses (recurse + 1 frames) and sorts with an explicit 1.7 KB stack
    size_t cur = 0; cudaDeviceGetLimit(&cur, cudaLimitStackSize);
    if (cur < 16384) CK(cudaDeviceSetLimit(cudaLimitStackSize, 16384));
    ctx->sdpStackSet = true;
  }
  const uint64_t totQ = b->qOff[n], totT = b->tOff[n];
  std::vector<uint64_t> blockOff(n + 1, 0);
  uint64_t qMax = 0, tMax = 0;
  for (uint32_t i = 0; i < n; i++) {
    const uint64_t ql = b->qOff[i + 1] - b->qOff[i], tl = b->tOff[i + 1] - b->tOff[i];
    qMax = std::max(qMax, ql); tMax = std::max(tMax, tl);
    blockOff[i + 1] = blockOff[i] + std::min(ql, tl) + 2;      // blocks are disjoint in both sequences
  }
  // scratch slice per thread: the k-mer table of the target (40 B / base) + 80 B per fragment, room for 2 (|q| + |t|) + 4096
  // fragments (a 10 kb pair at 15 % error has ~2,000); a job that needs more comes back BGPU_JOB_RANGE
  const size_t slice = (40 * tMax + 80 * (2 * (qMax + tMax) + 4096) + 65536 + 255) & ~(size_t)255;
  size_t freeB = 0, totalB = 0; cudaMemGetInfo(&freeB, &totalB);
  const size_t budget = std::min<size_t>(freeB / 4, (size_t)24 << 30);
  unsigned slices = (unsigned)std::min<size_t>(std::min<size_t>(budget / slice, (size_t)ctx->nSM * 48), ((size_t)n + 1) / 2 * 2);   // warps
  slices = slices / 2 * 2;
  if (slices < 2) { ctx->err = "bgpu_sdp_align: not enough device memory for the scratch arena"; return BGPU_E_OOM; }
  uint8_t *d_q = nullptr, *d_t = nullptr, *d_arena = nullptr; uint64_t *d_off = nullptr; uint32_t *d_counter = nullptr;
  bgpu_result *d_res = nullptr; bgpu_block *d_blocks = nullptr;
  auto freeAll = [&]() { cudaFree(d_q); cudaFree(d_t); cudaFree(d_arena); cudaFree(d_off); cudaFree(d_counter); cudaFree(d_res); cudaFree(d_blocks); };
#define SCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { char buf_[256]; snprintf(buf_, sizeof buf_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); ctx->err = buf_; freeAll(); return e_ == cudaErrorMemoryAllocation ? BGPU_E_OOM : BGPU_E_CUDA; } } while (0)
  SCK(cudaMalloc(&d_q, totQ + 16)); SCK(cudaMalloc(&d_t, totT + 16));
  SCK(cudaMalloc(&d_off, sizeof(uint64_t) * 3 * ((size_t)n + 1))); SCK(cudaMalloc(&d_counter, 64));
  SCK(cudaMalloc(&d_res, sizeof(bgpu_result) * n)); SCK(cudaMalloc(&d_blocks, sizeof(bgpu_block) * (blockOff[n] + 1)));
  SCK(cudaMalloc(&d_arena, slice * slices));
  cudaStream_t s = ctx->stream;
  SCK(cudaMemcpyAsync(d_q, b->qBases, totQ, cudaMemcpyHostToDevice, s)); SCK(cudaMemcpyAsync(d_t, b->tBases, totT, cudaMemcpyHostToDevice, s));
  SCK(cudaMemcpyAsync(d_off, b->qOff, sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
  SCK(cudaMemcpyAsync(d_off + (n + 1), b->tOff, sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
  SCK(cudaMemcpyAsync(d_off + 2 * ((size_t)n + 1), blockOff.data(), sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
  const int prm[8] = {p->wordSize, p->alignType, p->detailed, p->extendFront, p->sdpPrefix, p->recurse, p->noRecurseUnder, p->maxMatches};
  SCK((cudaError_t)run_sdp(fn, prm, p->indelRate, p->sdpIns, p->sdpDel, n, d_q, d_off, d_t, d_off + (n + 1), d_arena, slice, slices, d_counter,
                           d_res, d_blocks, d_off + 2 * ((size_t)n + 1), s));
  const size_t needPin = sizeof(bgpu_block) * (blockOff[n] + 1);
  if (ctx->sdpPinnedBytes < needPin) {
    if (ctx->sdpPinned) cudaFreeHost(ctx->sdpPinned);
    ctx->sdpPinned = nullptr; ctx->sdpPinnedBytes = 0;
    SCK(cudaHostAlloc(&ctx->sdpPinned, needPin, cudaHostAllocDefault));
    ctx->sdpPinnedBytes = needPin;
  }
  SCK(cudaMemcpyAsync(results, d_res, sizeof(bgpu_result) * n, cudaMemcpyDeviceToHost, s));
  SCK(cudaMemcpyAsync(ctx->sdpPinned, d_blocks, needPin, cudaMemcpyDeviceToHost, s));
  SCK(cudaStreamSynchronize(s));
#undef SCK
  freeAll();
  arena->blocks = (const bgpu_block *)ctx->sdpPinned; arena->nBlocks = blockOff[n];
  return BGPU_OK;
}

// ---- suffix-array anchoring: kernels and buffers in bgpu_anchor.cu ----
extern "C" int bgpu_set_suffix_array(bgpu_ctx *ctx, const uint32_t *index, uint64_t n, const uint32_t *startPosTable,
                                     const uint32_t *endPosTable, uint32_t lookupPrefixLength) {
  if (!ctx) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  std::lock_guard<std::mutex> lr(g_refMu);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return BGPU_E_CUDA;
  const DeviceRef &r = g_ref[ctx->device < 64 ? ctx->device : 0];
  return anchor_set_index(ctx->device, r.d, r.n, r.gen, index, n, startPosTable, endPosTable, lookupPrefixLength, ctx->err);
}

extern "C" int bgpu_map_reads(bgpu_ctx *ctx, const bgpu_anchor_params *p, const uint8_t *reads, const uint64_t *readOff, uint32_t nReads,
                              const uint32_t *subreadStart, const uint32_t *subreadEnd, uint64_t *matchOff, const bgpu_match **matches) {
  if (!ctx || !p || !matchOff || !matches || (nReads && (!reads || !readOff))) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return BGPU_E_CUDA;
  const uint8_t *refD; uint64_t refN, refGen;
  { std::lock_guard<std::mutex> lr(g_refMu); const DeviceRef &r = g_ref[ctx->device < 64 ? ctx->device : 0]; refD = r.d; refN = r.n; refGen = r.gen; }
  return anchor_map(ctx->device, refD, refN, refGen, ctx->stream, &ctx->anchor, p, reads, readOff, nReads, subreadStart, subreadEnd, matchOff, matches, ctx->err);
}

extern "C" int bgpu_build_lookup_table(const uint8_t *genome, uint64_t n, const uint32_t *index, uint32_t lookupPrefixLength,
                                       uint32_t *startPosTable, uint32_t *endPosTable) {
  return build_lookup_table(genome, n, index, lookupPrefixLength, startPosTable, endPosTable);
}

extern "C" int bgpu_map_rerun(bgpu_ctx *ctx) {
  if (!ctx) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return BGPU_E_CUDA;
  return anchor_rerun(ctx->anchor, ctx->stream, ctx->err);
}

extern "C" int bgpu_map_timing(bgpu_ctx *ctx, double ms[2], uint64_t *positions, uint64_t *h2dBytes, uint64_t *d2hBytes) {
  if (!ctx) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return anchor_timing(ctx->anchor, ms, positions, h2dBytes, d2hBytes);
}

extern "C" int bgpu_measure_int_peak(bgpu_ctx *ctx, double *opsPerSec, double *smClockMHz) {
  if (!ctx || !opsPerSec) return BGPU_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return BGPU_E_CUDA;
  double mhz = 0;
  *opsPerSec = measure_int_peak(ctx->nSM, ctx->stream, &mhz);
  if (smClockMHz) *smClockMHz = mhz;
  CK(cudaGetLastError());
  return BGPU_OK;
}

extern "C" int bgpu_int_peak_modes(double out[4]) {
  if (!out) return BGPU_E_INVALID;
  for (int i = 0; i < 4; i++) out[i] = bgpu::g_peakByMode[i];
  return BGPU_OK;
}

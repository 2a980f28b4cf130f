// baseline/gpu_refine.hpp -- the patch of INTEGRATION.md section 2, compiled for real into baseline/_ref/blasrmc_gpu.
//
// Included by the patched copy of alignment/Blasr.cpp right above RefineAlignments (Blasr.cpp:2163), whose first
// statement becomes `if (BgpuRefineAlignments(...)) return;`.  Everything else of the reference program -- anchoring,
// SDPAlign, filters, mapQV, printers, the MapReads loop -- is the reference's own code, unmodified.
//
// What it replaces, for the default path (useGuidedAlign, not -global):
//   RefineAlignment: slices (Blasr.cpp:850-859), AffineGuidedAlign / GuidedAlign (:862-873), ComputeAlignmentStats
//   (:875-878), copy-back (:888-914); RefineAlignments' sort (:2178-2180).
//
// Thread driver.  blasr runs MapReads once per -nproc on its own pthread (Blasr.cpp:4838) and each instance refines one
// read at a time, synchronously.  With the refinement on the GPU a thread would sit idle for the latency of its ticket, so
// the two pthread calls of that site are replaced by BgpuSpawn / BgpuJoin: the -nproc MapReads instances become user-level
// FIBERS (ucontext) spread over as many pthreads as the host has cores (BGPU_THREADS overrides).  A fiber that reaches
// RefineAlignments hands its candidates to blasr_gpu::RefineService and yields; its pthread goes on with another
// fiber's read (anchoring, SDPAlign, printing) until the ticket is back.  CPU stages and GPU refinement of different reads
// overlap without oversubscribing the cores.  `-nproc N` is therefore the number of reads in flight; 4 x cores is a good value.
#ifndef BGPU_GPU_REFINE_HPP_
#define BGPU_GPU_REFINE_HPP_
#include <cstdlib>
#include <ucontext.h>
#include <sys/mman.h>
#include <unistd.h>
#include <mutex>
#include "blasr_gpu_adapter.hpp"

struct BgpuFiber;
struct BgpuWorker {
  pthread_t thread;
  std::vector<BgpuFiber *> fibers;
  ucontext_t sched;
  BgpuFiber *current;
  BgpuWorker() : current(NULL) {}
};
struct BgpuFiber : blasr_gpu::RefineService::Waiter {
  ucontext_t ctx;
  void *(*fn)(void *); void *arg;
  bool finished;
  const std::atomic<bool> *waitingOn;         // NULL: runnable
  BgpuWorker *worker;
  BgpuFiber() : fn(NULL), arg(NULL), finished(false), waitingOn(NULL), worker(NULL) {}
  // called on the fiber: back to the worker's scheduler (which polls the service between fibers) until the request is done
  void Wait(blasr_gpu::RefineService &, const std::atomic<bool> &done) {
    while (!done.load(std::memory_order_acquire)) { waitingOn = &done; swapcontext(&ctx, &worker->sched); }
    waitingOn = NULL;
  }
};
static __thread BgpuFiber *bgpuCurrentFiber = NULL;

static std::vector<BgpuWorker *> &BgpuWorkers() { static std::vector<BgpuWorker *> w; return w; }
static std::vector<BgpuFiber *> &BgpuFibers() { static std::vector<BgpuFiber *> f; return f; }

static void BgpuFiberMain(unsigned lo, unsigned hi) {
  BgpuFiber *f = (BgpuFiber *)(((uintptr_t)hi << 32) | (uintptr_t)lo);
  f->fn(f->arg);
  f->finished = true;
  swapcontext(&f->ctx, &f->worker->sched);
}

static blasr_gpu::RefineService &BgpuService();

static void *BgpuWorkerMain(void *p) {
  BgpuWorker *w = (BgpuWorker *)p;
  for (;;) {
    bool alive = false, ran = false;
    for (size_t i = 0; i < w->fibers.size(); i++) {
      BgpuFiber *f = w->fibers[i];
      if (f->finished) continue;
      alive = true;
      if (f->waitingOn) {
        if (!f->waitingOn->load(std::memory_order_acquire)) continue;
      }
      w->current = f; bgpuCurrentFiber = f;
      swapcontext(&w->sched, &f->ctx);
      bgpuCurrentFiber = NULL;
      ran = true;
      if (f->waitingOn) BgpuService().Poll();     // a request has just been queued: start a ticket if a context is free
    }
    if (!alive) return NULL;
    // between rounds: start / finish tickets; with every fiber of this pthread waiting for the GPU, nap instead of spinning
    if (!BgpuService().Poll() && !ran) usleep(50);
  }
}

// replaces the pthread_exit(NULL) at the end of MapReads (Blasr.cpp:3915): only this fiber ends, not its pthread
static void BgpuFiberExit() {
  BgpuFiber *f = bgpuCurrentFiber;
  if (!f) pthread_exit(NULL);
  f->finished = true;
  swapcontext(&f->ctx, &f->worker->sched);
}

// replaces pthread_create(&threads[i], &attr[i], MapReads, &mapdb[i]) at Blasr.cpp:4838
static int BgpuSpawn(pthread_t *, const pthread_attr_t *, void *(*fn)(void *), void *arg, int index, int nProc) {
  std::vector<BgpuWorker *> &workers = BgpuWorkers();
  if (workers.empty()) {
    long cores = sysconf(_SC_NPROCESSORS_ONLN);
    if (getenv("BGPU_THREADS")) cores = atoi(getenv("BGPU_THREADS"));
    const int nw = (int)std::max(1L, std::min((long)nProc, cores));
    for (int i = 0; i < nw; i++) workers.push_back(new BgpuWorker());
  }
  BgpuFiber *f = new BgpuFiber();
  f->fn = fn; f->arg = arg; f->worker = workers[index % workers.size()];
  const size_t stackBytes = (size_t)64 << 20;      // reserved, touched on demand
  void *stack = mmap(NULL, stackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (stack == MAP_FAILED) { perror("mmap fiber stack"); exit(1); }
  getcontext(&f->ctx);
  f->ctx.uc_stack.ss_sp = stack; f->ctx.uc_stack.ss_size = stackBytes; f->ctx.uc_link = NULL;
  makecontext(&f->ctx, (void (*)())BgpuFiberMain, 2, (unsigned)((uintptr_t)f & 0xffffffffu), (unsigned)((uintptr_t)f >> 32));
  f->worker->fibers.push_back(f);
  BgpuFibers().push_back(f);
  if (index == nProc - 1)
    for (size_t i = 0; i < workers.size(); i++) pthread_create(&workers[i]->thread, NULL, BgpuWorkerMain, workers[i]);
  return 0;
}
// replaces pthread_join(threads[i], NULL) at Blasr.cpp:4841
static int BgpuJoin(pthread_t, int index) {
  if (index == 0) for (size_t i = 0; i < BgpuWorkers().size(); i++) pthread_join(BgpuWorkers()[i]->thread, NULL);
  return 0;
}

static blasr_gpu::RefineService &BgpuService() {
  static blasr_gpu::RefineService svc(getenv("BGPU_DEVICE") ? atoi(getenv("BGPU_DEVICE")) : 0,
                                      getenv("BGPU_SERVICE_CONTEXTS") ? atoi(getenv("BGPU_SERVICE_CONTEXTS")) : 3);
  return svc;
}

template<typename T_RefSequence, typename T_Sequence>
bool BgpuRefineAlignments(vector<T_Sequence*> &bothQueryStrands, T_RefSequence &genome,
                          vector<T_AlignmentCandidate*> &alignmentPtrs, MappingParameters &params,
                          MappingBuffers &mappingBuffers) {
  if (params.doGlobalAlignment || !params.useGuidedAlign) return false;     // the other branches stay the reference's
  static const bool fiberTest = getenv("BGPU_FIBER_TEST") != NULL;           // diagnostic: yield once, then the CPU path
  if (fiberTest) {
    if (bgpuCurrentFiber) { std::atomic<bool> never(false); BgpuFiber *f = bgpuCurrentFiber; f->waitingOn = NULL; swapcontext(&f->ctx, &f->worker->sched); }
    return false;
  }
  DistanceMatrixScoreFunction<DNASequence, FASTQSequence> distScoreFn;
  params.InitializeScoreFunction(distScoreFn);
  distScoreFn.InitializeScoreMatrix(SMRTDistanceMatrix);

  T_Sequence &query = *bothQueryStrands[0];
  blasr_gpu::RefineBatch batch;
  vector<DNASequence> tSeqs(alignmentPtrs.size());
  vector<FASTQSequence> qSeqs(alignmentPtrs.size());
  for (UInt i = 0; i < alignmentPtrs.size(); i++) {
    T_AlignmentCandidate &c = *alignmentPtrs[i];
    if (c.blocks.size() == 0) continue;
    int lastBlock = c.blocks.size() - 1;
    tSeqs[i].Copy(c.tAlignedSeq, c.tPos, c.blocks[lastBlock].tPos + c.blocks[lastBlock].length);
    qSeqs[i].ReferenceSubstring(query, c.qAlignedSeqPos + c.qPos, c.blocks[lastBlock].qPos + c.blocks[lastBlock].length);
    batch.Add(qSeqs[i].seq, qSeqs[i].length, tSeqs[i].seq, tSeqs[i].length, c.blocks);
  }
  if (batch.size() > 0)
    BgpuService().Run(batch, distScoreFn, params.affineAlign ? params.bandSize : params.guidedAlignBandSize,
                                  params.affineAlign, BGPU_GLOBAL, bgpuCurrentFiber);   // NULL (-nproc 1): blocks the thread
  UInt j = 0;
  for (UInt i = 0; i < alignmentPtrs.size(); i++) {
    T_AlignmentCandidate &c = *alignmentPtrs[i];
    if (c.blocks.size() == 0) continue;
    // a candidate the device refuses (a window wider than the widest kernel, scores beyond the kernels' number format: see
    // INTEGRATION.md section 6) stays with the reference's own RefineAlignment -- c is still untouched at this point
    const int st = batch.Result(j).status;
    if (st == BGPU_JOB_TOO_WIDE || st == BGPU_JOB_RANGE) {
      j++;
      tSeqs[i].Free();
      RefineAlignment(query, genome, c, params, mappingBuffers);
      continue;
    }
    T_AlignmentCandidate refinedAlignment;
    batch.Store(j++, refinedAlignment);
    c.blocks.clear();
    c.blocks = refinedAlignment.blocks;
    c.CopyStats(refinedAlignment);
    c.gaps = refinedAlignment.gaps;
    c.score = refinedAlignment.score;
    c.nCells = refinedAlignment.nCells;
    c.tAlignedSeq.TakeOwnership(tSeqs[i]);
    c.ReassignQSequence(qSeqs[i]);
    c.tAlignedSeqPos += c.tPos;
    c.qAlignedSeqPos += c.qPos;
    c.tPos = refinedAlignment.tPos;
    c.qPos = refinedAlignment.qPos;
    c.tAlignedSeqLength = tSeqs[i].length;
    c.qAlignedSeqLength = qSeqs[i].length;
  }
  if (params.sortRefinedAlignments)
    std::sort(alignmentPtrs.begin(), alignmentPtrs.end(), SortAlignmentPointersByScore());
  return true;
}
// ---- anchoring.  The two MapReadToGenome calls of MapRead (Blasr.cpp:2282-2296: the read, then its reverse
// complement) become ONE bgpu_map_reads call of two reads on the calling pthread's own context (stream + buffers; the index
// and the genome are loaded to the device once, from the program's own DNASuffixArray / genome objects).  The call is
// synchronous: a 10 kb read pair is ~20,000 independent searches, a fraction of a millisecond of device time against ~6 ms of
// this thread's CPU time in the reference.  Parameter sets the library refuses (it refuses exactly where the reference
// asserts or reads out of bounds) and the -lcpBounds debug output go to the reference's own function.
static __thread blasr_gpu::Context *bgpuAnchorCtx = NULL;
static __thread int bgpuRcCount = -1;                         // >= 0: the forward call has filled rcMatchPosList already

template<typename T_RefSequence, typename T_SuffixArray, typename T_Sequence, typename T_MatchPos>
int BgpuMapReadToGenome(T_Sequence *readRC, vector<T_MatchPos> *rcMatchPosList, bool forwardOnly,
                        T_RefSequence &genome, T_SuffixArray &sa, T_Sequence &read, unsigned int minPrefixMatchLength,
                        vector<T_MatchPos> &matchPosList, AnchorParameters &ap) {
  static const bool off = getenv("BGPU_NO_ANCHOR") != NULL;
  bgpuRcCount = -1;
  if (off || ap.lcpBoundsOutPtr != NULL || ap.removeEncompassedMatches || ap.expand > 14)
    return MapReadToGenome(genome, sa, read, minPrefixMatchLength, matchPosList, ap);
  static std::once_flag loaded;
  if (!bgpuAnchorCtx) bgpuAnchorCtx = new blasr_gpu::Context(getenv("BGPU_DEVICE") ? atoi(getenv("BGPU_DEVICE")) : 0);
  std::call_once(loaded, [&] { blasr_gpu::AnchorBatch::LoadIndex(*bgpuAnchorCtx, sa, genome); });
  blasr_gpu::AnchorBatch batch;
  batch.Add(read);
  if (!forwardOnly) batch.Add(*readRC);
  try {
    batch.Run(*bgpuAnchorCtx, minPrefixMatchLength, ap);
  } catch (const blasr_gpu::Error &e) {
    if (e.code != BGPU_E_INVALID) throw;
    return MapReadToGenome(genome, sa, read, minPrefixMatchLength, matchPosList, ap);
  }
  // MapReadToGenome clears the list only on its early return (MapBySuffixArray.h:219-222); MapRead clears both before the calls
  if (!forwardOnly) bgpuRcCount = batch.Store(1, *rcMatchPosList);
  return batch.Store(0, matchPosList);
}

template<typename T_RefSequence, typename T_SuffixArray, typename T_Sequence, typename T_MatchPos>
int BgpuMapReadToGenomeRC(T_RefSequence &genome, T_SuffixArray &sa, T_Sequence &readRC, unsigned int minPrefixMatchLength,
                          vector<T_MatchPos> &rcMatchPosList, AnchorParameters &ap) {
  if (bgpuRcCount >= 0) { const int n = bgpuRcCount; bgpuRcCount = -1; return n; }
  return MapReadToGenome(genome, sa, readRC, minPrefixMatchLength, rcMatchPosList, ap);
}
#endif

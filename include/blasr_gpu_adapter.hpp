// blasr_gpu_adapter.hpp -- header-only C++ glue between blasr's own containers and the C ABI (blasr_gpu.h).
//
// This is the binding a blasr maintainer adds next to alignment/Blasr.cpp.  It is written
// against *duck-typed* template parameters, so it compiles both against the reference's real types
//   T_AlignmentCandidate   common/datastructures/alignment/AlignmentCandidate.h:306
//   DNASequence            common/DNASequence.h            (seq, length)
//   Block / Gap / GapList  common/datastructures/alignment/AlignmentBlock.h:9-41, AlignmentGapList.h:9-24
//   DistanceMatrixScoreFunction  common/algorithms/alignment/DistanceMatrixScoreFunction.h:11
// (oracle/adapter_check.cpp and baseline/gpu_refine.hpp compile it that way) and against any stand-in with the same members;
// it contains no alignment arithmetic.
//
// Replaces, per batch of candidates instead of per candidate:
//   AffineGuidedAlign / GuidedAlign + ComputeAlignmentStats      alignment/Blasr.cpp:863-878   (RefineBatch)
//   KBandAlign / AffineKBandAlign / SWAlign                      Blasr.cpp:717-730,820-824,1067-1076; SDPAlign.h:440,503,563   (DenseBatch)
//   SDPAlign                                                     Blasr.cpp:1716-1722,1080-1090   (SdpBatch)
//   MapReadToGenome (read and reverse complement)                Blasr.cpp:2282-2296             (AnchorBatch)
#ifndef BLASR_GPU_ADAPTER_HPP_
#define BLASR_GPU_ADAPTER_HPP_
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include "blasr_gpu.h"

namespace blasr_gpu {

struct Error : std::runtime_error { int code; Error(int c, const std::string &m) : std::runtime_error(m), code(c) {} };

// One GPU context = one stream + one allocation cache.  Calls into a context are serialised inside the library, so the
// throughput configuration is ONE CONTEXT PER HOST THREAD of the blasr thread driver (MapReads, Blasr.cpp:3193) on the shared
// device -- the library's per-device phase gates keep the threads' tickets pipelined -- or a RefineService (below), which owns
// a few contexts and merges the requests of all threads into tickets.  A context shared by several pthreads works, serialised.
class Context {
 public:
  explicit Context(int device = 0) {
    int rc = bgpu_create(&ctx_, device);
    if (rc != BGPU_OK) throw Error(rc, rc == BGPU_E_NO_DEVICE ? "blasr_gpu: no CUDA device (there is no CPU fallback)" : "blasr_gpu: bgpu_create failed");
  }
  ~Context() { if (ctx_) bgpu_destroy(ctx_); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  bgpu_ctx *get() const { return ctx_; }
 private:
  bgpu_ctx *ctx_ = nullptr;
};

template <typename T_ScoreFn>
inline bgpu_scorefn MakeScoreFn(const T_ScoreFn &fn, int kind = BGPU_FN_DISTANCE) {
  bgpu_scorefn s;
  std::memset(&s, 0, sizeof s);
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) s.M[i * 5 + j] = fn.scoreMatrix[i][j];
  s.ins = fn.ins; s.del = fn.del; s.affineOpen = fn.affineOpen; s.affineExtend = fn.affineExtend; s.kind = kind;
  s.substitutionPrior = fn.substitutionPrior; s.globalDeletionPrior = fn.globalDeletionPrior;   // BaseScoreFunction.h:8-9
  return s;
}

// A ticket shared by the RefineBatch objects whose jobs RefineService merged into one submission: released when the
// last of them lets go.
struct SharedTicket {
  bgpu_ctx *ctx; bgpu_ticket ticket; int refs; std::mutex mu; uint32_t nJobs;
  SharedTicket(bgpu_ctx *c, bgpu_ticket t, int r, uint32_t n = 0) : ctx(c), ticket(t), refs(r), nJobs(n) {}
};

class RefineService;

// Collects the per-candidate slices RefineAlignment builds (Blasr.cpp:850-859) and runs them as one batch.
class RefineBatch {
  friend class RefineService;
 public:
  // q/t: the slices handed to (Affine)GuidedAlign; blocks: candidate.blocks (used raw as the guide).
  template <typename T_BlockVector>
  void Add(const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, const T_BlockVector &blocks,
           const uint8_t *qual = nullptr) {
    // the guide travels packed (bgpu_batch::guidePacked): per block the gap to the previous block's end in q and in t and the
    // length, one byte each; the rare block that does not fit goes to the side list
    const size_t g0 = nGuide_;
    guideP_.resize(3 * (g0 + blocks.size()));
    uint64_t qe = 0, te = 0;
    for (size_t i = 0; i < blocks.size(); i++) {
      const uint64_t dq = (uint64_t)blocks[i].qPos - qe, dt = (uint64_t)blocks[i].tPos - te, len = blocks[i].length;
      uint8_t *p = &guideP_[3 * (g0 + i)];
      if (dq < 255 && dt < 255 && len < 255 && blocks[i].qPos >= qe && blocks[i].tPos >= te) { p[0] = (uint8_t)dq; p[1] = (uint8_t)dt; p[2] = (uint8_t)len; }
      else {
        p[0] = p[1] = p[2] = 255;
        guideW_.push_back((uint32_t)(g0 + i)); guideW_.push_back((uint32_t)dq); guideW_.push_back((uint32_t)dt); guideW_.push_back((uint32_t)len);
      }
      qe = (uint64_t)blocks[i].qPos + len; te = (uint64_t)blocks[i].tPos + len;
    }
    nGuide_ = g0 + blocks.size();
    q_.insert(q_.end(), q, q + qLen); t_.insert(t_.end(), t, t + tLen);
    if (qual) { qual_.resize(q_.size() - qLen, 0); qual_.insert(qual_.end(), qual, qual + qLen); }
    qOff_.push_back(q_.size()); tOff_.push_back(t_.size()); gOff_.push_back(nGuide_);
  }
  uint32_t size() const { return (uint32_t)qOff_.size() - 1; }

  // Runs AffineGuidedAlign (affine=true: what blasr does, MappingParameters.h:539) or GuidedAlign on every job and
  // ComputeAlignmentStats afterwards.  Results stay valid until the next Run()/destruction.
  template <typename T_ScoreFn>
  void Run(Context &ctx, const T_ScoreFn &fn, int bandSize, bool affine, int alignType = BGPU_GLOBAL) {
    bgpu_scorefn s = MakeScoreFn(fn);
    bgpu_params p;
    std::memset(&p, 0, sizeof p);
    p.algo = affine ? BGPU_AFFINE_GUIDED : BGPU_GUIDED; p.alignType = alignType; p.band = bandSize;
    p.doStats = 1; p.statsAffine = affine ? 1 : 0;
    p.compactResults = 1;                                              // run-length paths back, expanded by Store()
    if (!qual_.empty()) qual_.resize(q_.size(), 0);
    bgpu_batch b;
    std::memset(&b, 0, sizeof b);   // no rich QV tracks on this path (DistanceMatrixScoreFunction)
    b.nJobs = size(); b.qBases = q_.data(); b.qOff = qOff_.data(); b.tBases = t_.data(); b.tOff = tOff_.data();
    b.qual = qual_.empty() ? nullptr : qual_.data(); b.guideOff = gOff_.data(); b.band = nullptr;
    b.guidePacked = guideP_.data(); b.guideWide = guideW_.data(); b.nGuideWide = guideW_.size() / 4;
    results_.resize(b.nJobs);
    Release();
    int rc = bgpu_submit(ctx.get(), &s, &p, &b, &ticket_);
    if (rc == BGPU_OK) { owner_ = ctx.get(); rc = bgpu_collect(owner_, ticket_, results_.data(), &arena_); }
    if (rc != BGPU_OK) throw Error(rc, std::string("blasr_gpu: ") + bgpu_last_error(ctx.get()));
    cigarOps_ = nullptr; cigarOff_ = nullptr; base_ = 0; rescored_.clear();
  }
  ~RefineBatch() { Release(); }

  // The SAM CIGAR core of job i as CreateNoClippingCigarOps + CigarOpsToString print it (SAMPrinter.h:203-293,318-327):
  // '=' / 'X' runs per block, 'I' / 'D' per stored gap.  Clipping ops and the tStrand reversal stay with the caller
  // (CreateCIGARString :366-397).  Built on the device for the whole batch at the first call.
  std::string Cigar(uint32_t i) {
    if (!cigarOps_) {
      int rc = bgpu_cigar(owner_, shared_ ? shared_->ticket : ticket_, &cigarOps_, &cigarOff_);
      if (rc != BGPU_OK) throw Error(rc, std::string("blasr_gpu: ") + bgpu_last_error(owner_));
    }
    std::string out;
    i += base_;                                   // position of this batch's first job inside a merged ticket
    for (uint64_t k = cigarOff_[i]; k < cigarOff_[i + 1]; k++) { out += std::to_string(cigarOps_[k] >> 4); out += "MIDNSHP=X"[cigarOps_[k] & 15]; }
    return out;
  }

  // The rescoring step of StoreMapQVs (alignment/Blasr.cpp:2768-2780): ComputeAlignmentScore(alignment, qAlignedSeq,
  // tAlignedSeq, fn, useAffinePenalty) -- the Alignment overload, AlignmentUtils.h:127-169 -- of job i under ANOTHER score
  // function (blasr: the run's gap costs with SMRTLogProbMatrix; probScore = -Rescore(i, fn, affine) / 10.0).  Evaluated on the
  // device for the whole ticket at the first call.
  template <typename T_ScoreFn>
  int Rescore(uint32_t i, const T_ScoreFn &fn, bool useAffinePenalty) {
    if (rescored_.empty()) {
      const bgpu_scorefn s = MakeScoreFn(fn);
      rescored_.resize(shared_ ? shared_->nJobs : size());
      const int rc = bgpu_rescore(owner_, shared_ ? shared_->ticket : ticket_, &s, useAffinePenalty ? 1 : 0, rescored_.data());
      if (rc != BGPU_OK) { rescored_.clear(); throw Error(rc, std::string("blasr_gpu: ") + bgpu_last_error(owner_)); }
    }
    return rescored_[base_ + i];
  }

  const bgpu_result &Result(uint32_t i) const { return results_[i]; }

  // Writes job i into a reference-style alignment exactly as Blasr.cpp:888-914 copies refinedAlignment back.
  // T_Alignment needs: blocks (vector of {qPos,tPos,length}), gaps (vector<vector<T_Gap>>), score, nCells, qPos, tPos,
  // nMatch, nMismatch, nIns, nDel, pctSimilarity.  T_Gap needs {seq, length} with seq convertible from int
  // (Gap::Query = 0, Gap::Target = 1, AlignmentGapList.h:11-14).
  template <typename T_Alignment>
  void Store(uint32_t i, T_Alignment &out) const {
    const bgpu_result &r = results_[i];
    if (r.status == BGPU_JOB_PATH_AWRY)   // GuidedAlign.h:637-651 prints and exit(1)s here
      throw Error(r.status, "ERROR, this path has gone awry");
    if (r.status != BGPU_JOB_OK && r.status != BGPU_JOB_EMPTY_GUIDE) throw Error(r.status, "blasr_gpu: job rejected");
    out.blocks.resize(r.nBlocks);
    out.gaps.clear(); out.gaps.resize(r.nGapLists);
    if (arena_.runs) {
      // compact results: the path as runs (type << 30 | length) from the first block to the last; block positions are relative
      // to qPos / tPos, gaps[b + 1] holds the gap runs after block b (gaps[0] and the last list stay empty)
      const uint32_t *run = arena_.runs + r.blockOff + r.gapOff;
      uint32_t q = 0, t = 0, b = 0;
      for (uint32_t k = 0; k < r.nBlocks + r.nGaps; k++) {
        const uint32_t type = run[k] >> 30, len = run[k] & 0x3fffffffu;
        if (type == 0) { out.blocks[b].qPos = q; out.blocks[b].tPos = t; out.blocks[b].length = len; b++; q += len; t += len; }
        else {
          out.gaps[b].resize(out.gaps[b].size() + 1);
          typedef decltype(out.gaps[b][0].seq) seq_t;
          out.gaps[b].back().seq = (seq_t)(type == 1 ? 1 : 0); out.gaps[b].back().length = (int)len;   // Gap::Target = 1, Gap::Query = 0
          if (type == 1) q += len; else t += len;
        }
      }
    } else {
      for (uint32_t k = 0; k < r.nBlocks; k++) {
        const bgpu_block &bk = arena_.blocks[r.blockOff + k];
        out.blocks[k].qPos = bk.qPos; out.blocks[k].tPos = bk.tPos; out.blocks[k].length = bk.length;
      }
      uint64_t g = r.gapOff;
      for (uint32_t k = 0; k < r.nGapLists; k++) {
        const uint32_t c = arena_.gapCounts[r.gapListOff + k];
        out.gaps[k].resize(c);
        for (uint32_t x = 0; x < c; x++, g++) {
          typedef decltype(out.gaps[k][x].seq) seq_t;
          out.gaps[k][x].seq = (seq_t)arena_.gaps[g].seq; out.gaps[k][x].length = arena_.gaps[g].length;
        }
      }
    }
    out.qPos = r.qPos; out.tPos = r.tPos; out.nCells = r.nCells;
    out.nMatch = r.nMatch; out.nMismatch = r.nMismatch; out.nIns = r.nIns; out.nDel = r.nDel;
    out.pctSimilarity = r.pctSimilarity; out.score = r.statsScore;   // ComputeAlignmentStats overwrites score, AlignmentUtils.h:578
  }

  void Release() {
    if (shared_) {
      bool last;
      { std::lock_guard<std::mutex> lk(shared_->mu); last = --shared_->refs == 0; }
      if (last) { bgpu_release(shared_->ctx, shared_->ticket); delete shared_; }
      shared_ = nullptr; ticket_ = nullptr;
    }
    if (ticket_) { bgpu_release(owner_, ticket_); ticket_ = nullptr; }
  }
  void Clear() { Release(); q_.clear(); t_.clear(); qual_.clear(); guideP_.clear(); guideW_.clear(); nGuide_ = 0; qOff_.assign(1, 0); tOff_.assign(1, 0); gOff_.assign(1, 0); results_.clear(); }

 private:
  SharedTicket *shared_ = nullptr;                                 // set when the jobs ran inside a merged submission
  uint32_t base_ = 0;                                              // ... and the index of this batch's first job in it
  std::vector<uint8_t> q_, t_, qual_;
  std::vector<uint8_t> guideP_;                                    // packed guide: 3 bytes per block
  std::vector<uint32_t> guideW_;                                   // ... and the side list {block, dq, dt, length}
  uint64_t nGuide_ = 0;
  std::vector<uint64_t> qOff_{0}, tOff_{0}, gOff_{0};
  std::vector<bgpu_result> results_;
  bgpu_arena arena_{};
  bgpu_ctx *owner_ = nullptr; bgpu_ticket ticket_ = nullptr;       // results and the arena live until Release()
  const uint32_t *cigarOps_ = nullptr; const uint64_t *cigarOff_ = nullptr;
  std::vector<int32_t> rescored_;                                  // Rescore(): the whole ticket's scores
};

// Cross-thread batching for blasr's thread driver.  MapReads (Blasr.cpp:3193) runs one instance per -nproc, and each
// calls RefineAlignments synchronously with the handful of candidates of ONE read -- far too few jobs to fill a GPU.
// RefineService keeps that structure untouched: every caller hands its RefineBatch to Run() and waits.  There are no
// service threads: whoever waits drives the service by calling Poll(), which (1) starts a ticket from ALL requests that are
// pending whenever one of the contexts is free (bgpu_submit only enqueues, it never waits for the device) and (2) collects
// every ticket whose kernels are done (bgpu_query) and hands each caller its slice of the results.  While the contexts are
// busy the next requests pile up, so tickets grow with the load (batching by back-pressure); several contexts keep several
// tickets in flight (a ticket's latency is the serial sweep of its longest read, whatever the number of reads in it).
//
// How a caller waits is pluggable (Waiter): the default polls and naps on the calling thread; a host that runs its
// per-read loop on user-level fibers (baseline/gpu_refine.hpp does, to overlap the CPU stages of one read with the GPU
// refinement of another without oversubscribing the cores) yields to its scheduler, which polls between fibers.
class RefineService {
 public:
  struct Waiter {                       // Wait() returns once done is true
    virtual void Wait(RefineService &svc, const std::atomic<bool> &done) = 0;
    virtual ~Waiter() {}
  };

  RefineService(int device, int nContexts = 3) {
    for (int i = 0; i < (nContexts < 1 ? 1 : nContexts); i++) ctxs_.push_back(std::unique_ptr<Context>(new Context(device)));
    flights_.resize(ctxs_.size());
  }
  ~RefineService() {
    if (getenv("BGPU_SERVICE_STATS"))
      fprintf(stderr, "RefineService: %llu tickets, %llu jobs (%.1f per ticket); per ticket: bgpu_submit %.2f ms, submit -> collected %.2f ms, bgpu_collect %.2f ms, kernels on the device %.2f ms; cudaMalloc + cudaHostAlloc calls of the busiest context %llu; per request: queued %.2f ms before its ticket started, resumed %.2f ms after it was done\n",
              (unsigned long long)tickets_, (unsigned long long)jobs_, tickets_ ? (double)jobs_ / tickets_ : 0.0,
              tickets_ ? submitMs_ / tickets_ : 0.0, tickets_ ? flightMs_ / tickets_ : 0.0, tickets_ ? collectMs_ / tickets_ : 0.0,
              tickets_ ? deviceMs_ / tickets_ : 0.0, (unsigned long long)allocs_, requests_ ? pendingMs_ / requests_ : 0.0,
              requests_ ? resumeMs_ / requests_ : 0.0);
  }

  template <typename T_ScoreFn>
  void Run(RefineBatch &batch, const T_ScoreFn &fn, int bandSize, bool affine, int alignType = BGPU_GLOBAL, Waiter *waiter = nullptr) {
    Request r;
    r.batch = &batch; r.fn = MakeScoreFn(fn);
    std::memset(&r.p, 0, sizeof r.p);
    r.p.algo = affine ? BGPU_AFFINE_GUIDED : BGPU_GUIDED; r.p.alignType = alignType; r.p.band = bandSize;
    r.p.doStats = 1; r.p.statsAffine = affine ? 1 : 0; r.p.compactResults = 1;
    batch.Release();
    r.tEnqueue = std::chrono::steady_clock::now();
    { std::lock_guard<std::mutex> lk(mu_); pending_.push_back(&r); }
    NappingWaiter nw;
    (waiter ? waiter : &nw)->Wait(*this, r.done);
    {
      std::lock_guard<std::mutex> lk(statMu_);
      requests_++;
      pendingMs_ += std::chrono::duration<double, std::milli>(r.tStart - r.tEnqueue).count();
      resumeMs_ += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - r.tDone).count();
    }
    if (r.rc != BGPU_OK) throw Error(r.rc, "blasr_gpu: " + r.err);
  }

  // Drives the service from the calling thread; returns true when it started or finished a ticket.
  bool Poll() {
    bool did = false;
    std::unique_lock<std::mutex> lk(mu_, std::try_to_lock);
    if (!lk.owns_lock()) return false;              // someone else is driving right now
    // completion queries go through the CUDA driver's lock: at most one every 20 us, whoever calls
    const auto now = std::chrono::steady_clock::now();
    const bool ask = now - lastQuery_ >= std::chrono::microseconds(20);
    if (ask) lastQuery_ = now;
    for (size_t c = 0; ask && c < flights_.size(); c++) {
      Flight &f = flights_[c];
      if (f.state == Flight::FLYING && bgpu_query(ctxs_[c]->get(), f.ticket) != 0) {
        f.state = Flight::LANDING;
        lk.unlock();
        Finish(*ctxs_[c], f);
        lk.lock();
        f.state = Flight::IDLE;
        did = true;
      }
    }
    if (!pending_.empty())
      for (size_t c = 0; c < flights_.size(); c++) {
        Flight &f = flights_[c];
        if (f.state != Flight::IDLE) continue;
        Request *first = pending_.front();
        std::vector<Request *> rest;                // one ticket = one (score function, parameters) pair
        f.reqs.clear();
        for (Request *x : pending_)
          (std::memcmp(&x->fn, &first->fn, sizeof first->fn) == 0 && std::memcmp(&x->p, &first->p, sizeof first->p) == 0 ? f.reqs : rest).push_back(x);
        pending_.swap(rest);
        f.state = Flight::STARTING;
        lk.unlock();
        const bool ok = Start(*ctxs_[c], f);
        lk.lock();
        f.state = ok ? Flight::FLYING : Flight::IDLE;
        did = true;
        break;
      }
    return did;
  }

  uint64_t Tickets() const { return tickets_; }
  uint64_t Jobs() const { return jobs_; }

 private:
  struct Request {
    RefineBatch *batch; bgpu_scorefn fn; bgpu_params p; std::atomic<bool> done; int rc = BGPU_OK; std::string err;
    std::chrono::steady_clock::time_point tEnqueue, tStart, tDone;
    Request() : done(false) {}
  };
  struct Flight {
    enum { IDLE, STARTING, FLYING, LANDING } state = IDLE;
    std::vector<Request *> reqs; bgpu_ticket ticket = nullptr; uint32_t nJobs = 0;
    std::chrono::steady_clock::time_point t0;
  };
  struct NappingWaiter : Waiter {
    void Wait(RefineService &svc, const std::atomic<bool> &done) {
      while (!done.load(std::memory_order_acquire))
        if (!svc.Poll()) std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
  };

  static void Fail(std::vector<Request *> &reqs, int rc, const std::string &e) {
    for (Request *x : reqs) { x->rc = rc; x->err = e; x->tStart = x->tDone = std::chrono::steady_clock::now(); x->done.store(true, std::memory_order_release); }
  }

  // concatenates the requests' jobs and enqueues them as one ticket (inputs are staged by bgpu_submit before it returns)
  bool Start(Context &ctx, Flight &f) {
    std::vector<uint8_t> q, t, qual, guideP; std::vector<uint32_t> guideW; std::vector<uint64_t> qOff(1, 0), tOff(1, 0), gOff(1, 0);
    bool anyQual = false;
    for (Request *x : f.reqs) anyQual = anyQual || !x->batch->qual_.empty();
    for (Request *x : f.reqs) {
      RefineBatch &b = *x->batch;
      const uint64_t q0 = q.size(), t0 = t.size(), g0 = guideP.size() / 3;
      q.insert(q.end(), b.q_.begin(), b.q_.end()); t.insert(t.end(), b.t_.begin(), b.t_.end());
      guideP.insert(guideP.end(), b.guideP_.begin(), b.guideP_.end());
      for (size_t k = 0; k < b.guideW_.size(); k += 4) {
        guideW.push_back((uint32_t)(b.guideW_[k] + g0)); guideW.push_back(b.guideW_[k + 1]); guideW.push_back(b.guideW_[k + 2]); guideW.push_back(b.guideW_[k + 3]);
      }
      if (anyQual) { qual.resize(q0, 0); qual.insert(qual.end(), b.qual_.begin(), b.qual_.end()); qual.resize(q.size(), 0); }
      for (uint32_t i = 1; i <= b.size(); i++) { qOff.push_back(q0 + b.qOff_[i]); tOff.push_back(t0 + b.tOff_[i]); gOff.push_back(g0 + b.gOff_[i]); }
    }
    bgpu_batch mb; std::memset(&mb, 0, sizeof mb);
    mb.nJobs = (uint32_t)qOff.size() - 1; mb.qBases = q.data(); mb.qOff = qOff.data(); mb.tBases = t.data(); mb.tOff = tOff.data();
    mb.qual = anyQual ? qual.data() : nullptr; mb.guideOff = gOff.data();
    mb.guidePacked = guideP.data(); mb.guideWide = guideW.data(); mb.nGuideWide = guideW.size() / 4;
    f.nJobs = mb.nJobs; f.ticket = nullptr; f.t0 = std::chrono::steady_clock::now();
    for (Request *x : f.reqs) x->tStart = f.t0;
    const int rc = bgpu_submit(ctx.get(), &f.reqs[0]->fn, &f.reqs[0]->p, &mb, &f.ticket);
    if (rc != BGPU_OK) { Fail(f.reqs, rc, bgpu_last_error(ctx.get())); return false; }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - f.t0).count();
    std::lock_guard<std::mutex> lk(statMu_);
    submitMs_ += ms;
    return true;
  }

  void Finish(Context &ctx, Flight &f) {
    std::vector<bgpu_result> res(f.nJobs);
    bgpu_arena arena; std::memset(&arena, 0, sizeof arena);
    const auto c0 = std::chrono::steady_clock::now();
    const int rc = bgpu_collect(ctx.get(), f.ticket, res.data(), &arena);
    if (rc != BGPU_OK) { const std::string e = bgpu_last_error(ctx.get()); bgpu_release(ctx.get(), f.ticket); Fail(f.reqs, rc, e); return; }
    const auto c1 = std::chrono::steady_clock::now();
    bgpu_timing tm; bgpu_timing_of(ctx.get(), f.ticket, &tm);
    {
      std::lock_guard<std::mutex> lk(statMu_);
      collectMs_ += std::chrono::duration<double, std::milli>(c1 - c0).count();
      flightMs_ += std::chrono::duration<double, std::milli>(c1 - f.t0).count();
      deviceMs_ += tm.msTotal; tickets_++; jobs_ += f.nJobs;
      allocs_ = std::max<uint64_t>(allocs_, tm.devAllocs + tm.pinAllocs);
    }
    SharedTicket *sh = new SharedTicket(ctx.get(), f.ticket, (int)f.reqs.size(), f.nJobs);
    uint32_t at = 0;
    for (Request *x : f.reqs) {
      RefineBatch &b = *x->batch;
      b.results_.assign(res.begin() + at, res.begin() + at + b.size());
      b.base_ = at;
      at += b.size();
      b.arena_ = arena; b.owner_ = ctx.get(); b.ticket_ = nullptr; b.shared_ = sh; b.cigarOps_ = nullptr; b.cigarOff_ = nullptr; b.rescored_.clear();
      x->tDone = std::chrono::steady_clock::now();
      x->done.store(true, std::memory_order_release);          // x may be gone right after this
    }
  }

  std::mutex mu_, statMu_;
  std::chrono::steady_clock::time_point lastQuery_;
  std::vector<Request *> pending_;
  std::vector<std::unique_ptr<Context>> ctxs_;
  std::vector<Flight> flights_;
  uint64_t tickets_ = 0, jobs_ = 0, allocs_ = 0, requests_ = 0;
  double pendingMs_ = 0, resumeMs_ = 0;
  double submitMs_ = 0, flightMs_ = 0, collectMs_ = 0, deviceMs_ = 0;
};

// The other per-candidate DP call sites: jobs without a guide.
//   KBandAlign        alignment/Blasr.cpp:820-824 (-global), :717-730 (PairwiseLocalAlign, Fit)       KBandAlign.h:75
//   AffineKBandAlign  Blasr.cpp:1067-1076 (AlignSubstring, the gap fills of -alignContigs), :695-705   AffineKBandAlign.h:12
//   SWAlign           common/algorithms/alignment/SDPAlign.h:440,503,563 (gap fills of a detailed SDPAlign)  SWAlign.h:18
class DenseBatch {
 public:
  // k < 0: use the k passed to Run*()
  void Add(const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, int k = -1) {
    q_.insert(q_.end(), q, q + qLen); t_.insert(t_.end(), t, t + tLen);
    qOff_.push_back(q_.size()); tOff_.push_back(t_.size()); band_.push_back(k); anyBand_ = anyBand_ || k >= 0;
  }
  uint32_t size() const { return (uint32_t)qOff_.size() - 1; }

  // KBandAlign(q, t, matchMat, ins, del, k, scoreMat, pathMat, alignment, alignType, scoreFn): ins / del are the int
  // boundary-cost parameters, the fill uses scoreFn's (KBandAlign.h:116,121 vs :196-244)
  template <typename T_ScoreFn>
  void RunKBand(Context &ctx, const T_ScoreFn &fn, int ins, int del, int k, int alignType) {
    bgpu_params p = Params(BGPU_KBAND, alignType, k); p.bndIns = ins; p.bndDel = del;
    Run(ctx, MakeScoreFn(fn), p);
  }
  // SWAlign(q, t, scoreMat, pathMat, alignment, scoreFn, alignType)
  template <typename T_ScoreFn>
  void RunSW(Context &ctx, const T_ScoreFn &fn, int alignType) { Run(ctx, MakeScoreFn(fn), Params(BGPU_SW, alignType, 0)); }
  // AffineKBandAlign(q, t, matchMat, hpInsOpen, hpInsExtend, insOpen, insExtend, del, k, ..., alignment, alignType)
  void RunAffineKBand(Context &ctx, const int matchMat[5][5], int hpInsOpen, int hpInsExtend, int insOpen, int insExtend,
                      int del, int k, int alignType) {
    bgpu_scorefn s; std::memset(&s, 0, sizeof s);
    for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) s.M[i * 5 + j] = matchMat[i][j];
    s.kind = BGPU_FN_DISTANCE;
    bgpu_params p = Params(BGPU_AFFINE_KBAND, alignType, k);
    p.bndDel = del; p.hpInsOpen = hpInsOpen; p.hpInsExtend = hpInsExtend; p.insOpen = insOpen; p.insExtend = insExtend;
    Run(ctx, s, p);
  }
  ~DenseBatch() { Release(); }

  const bgpu_result &Result(uint32_t i) const { return results_[i]; }

  // Writes what the aligner stores into `alignment` -- blocks, gaps, and (KBandAlign, SWAlign) qPos / tPos; AffineKBandAlign
  // never sets qPos / tPos, so they are left alone -- and returns the aligner's return value (its score).
  template <typename T_Alignment>
  int Store(uint32_t i, T_Alignment &out) const {
    const bgpu_result &r = results_[i];
    if (r.status != BGPU_JOB_OK) throw Error(r.status, "blasr_gpu: job rejected");
    out.blocks.resize(r.nBlocks);
    for (uint32_t k = 0; k < r.nBlocks; k++) {
      const bgpu_block &bk = arena_.blocks[r.blockOff + k];
      out.blocks[k].qPos = bk.qPos; out.blocks[k].tPos = bk.tPos; out.blocks[k].length = bk.length;
    }
    out.gaps.clear(); out.gaps.resize(r.nGapLists);
    uint64_t g = r.gapOff;
    for (uint32_t k = 0; k < r.nGapLists; k++) {
      const uint32_t c = arena_.gapCounts[r.gapListOff + k];
      out.gaps[k].resize(c);
      for (uint32_t x = 0; x < c; x++, g++) {
        typedef decltype(out.gaps[k][x].seq) seq_t;
        out.gaps[k][x].seq = (seq_t)arena_.gaps[g].seq; out.gaps[k][x].length = arena_.gaps[g].length;
      }
    }
    if (algo_ != BGPU_AFFINE_KBAND) { out.qPos = r.qPos; out.tPos = r.tPos; }
    return r.score;
  }

  void Release() { if (ticket_) { bgpu_release(owner_, ticket_); ticket_ = nullptr; } }
  void Clear() { Release(); q_.clear(); t_.clear(); band_.clear(); anyBand_ = false; qOff_.assign(1, 0); tOff_.assign(1, 0); results_.clear(); }

 private:
  static bgpu_params Params(int algo, int alignType, int k) {
    bgpu_params p; std::memset(&p, 0, sizeof p);
    p.algo = algo; p.alignType = alignType; p.band = k;
    return p;
  }
  void Run(Context &ctx, const bgpu_scorefn &s, const bgpu_params &p) {
    bgpu_batch b; std::memset(&b, 0, sizeof b);
    b.nJobs = size(); b.qBases = q_.data(); b.qOff = qOff_.data(); b.tBases = t_.data(); b.tOff = tOff_.data();
    std::vector<int32_t> band(band_);
    for (size_t i = 0; i < band.size(); i++) if (band[i] < 0) band[i] = p.band;
    b.band = anyBand_ ? band.data() : nullptr;
    results_.resize(b.nJobs);
    Release();
    algo_ = p.algo;
    int rc = bgpu_submit(ctx.get(), &s, &p, &b, &ticket_);
    if (rc == BGPU_OK) { owner_ = ctx.get(); rc = bgpu_collect(owner_, ticket_, results_.data(), &arena_); }
    if (rc != BGPU_OK) throw Error(rc, std::string("blasr_gpu: ") + bgpu_last_error(ctx.get()));
  }
  std::vector<uint8_t> q_, t_;
  std::vector<int32_t> band_;
  bool anyBand_ = false;
  std::vector<uint64_t> qOff_{0}, tOff_{0};
  std::vector<bgpu_result> results_;
  bgpu_arena arena_{};
  bgpu_ctx *owner_ = nullptr; bgpu_ticket ticket_ = nullptr;
  int algo_ = BGPU_KBAND;
};

// SDPAlign (common/algorithms/alignment/SDPAlign.h:95-107) for the (query, target) pairs of a read's intervals: the call of
// Blasr.cpp:1716-1722 (Local) / :1080-1090 (Global).  Add() the pairs, Run() with the reference's own parameter list, Store(i,
// alignment) writes alignment.qPos / tPos / blocks exactly as SDPAlign leaves them (gaps stay empty, as there).
class SdpBatch {
 public:
  void Add(const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen) {
    q_.insert(q_.end(), q, q + qLen); t_.insert(t_.end(), t, t + tLen);
    qOff_.push_back(q_.size()); tOff_.push_back(t_.size());
  }
  uint32_t size() const { return (uint32_t)qOff_.size() - 1; }
  // the parameters in SDPAlign's own order: scoreFn, wordSize, sdpIns, sdpDel, indelRate, alignType, detailedAlignment,
  // extendFrontByLocalAlignment, sdpPrefixLength, recurse, noRecurseUnder, maxMatchesPerPosition
  template <typename T_ScoreFn>
  void Run(Context &ctx, const T_ScoreFn &fn, int wordSize, int sdpIns, int sdpDel, float indelRate, int alignType = BGPU_GLOBAL,
           bool detailedAlignment = true, bool extendFrontByLocalAlignment = true, int sdpPrefixLength = 50, int recurse = 0,
           int noRecurseUnder = 10000, int maxMatchesPerPosition = 0) {
    const bgpu_scorefn s = MakeScoreFn(fn, BGPU_FN_DISTANCE);
    bgpu_sdp_params p; std::memset(&p, 0, sizeof p);
    p.wordSize = wordSize; p.sdpIns = sdpIns; p.sdpDel = sdpDel; p.indelRate = indelRate; p.alignType = alignType;
    p.detailed = detailedAlignment; p.extendFront = extendFrontByLocalAlignment; p.sdpPrefix = sdpPrefixLength;
    p.recurse = recurse; p.noRecurseUnder = noRecurseUnder; p.maxMatches = maxMatchesPerPosition;
    bgpu_batch b; std::memset(&b, 0, sizeof b);
    b.nJobs = size(); b.qBases = q_.data(); b.qOff = qOff_.data(); b.tBases = t_.data(); b.tOff = tOff_.data();
    results_.resize(b.nJobs);
    const int rc = bgpu_sdp_align(ctx.get(), &s, &p, &b, results_.data(), &arena_);
    if (rc != BGPU_OK) throw Error(rc, std::string("blasr_gpu: ") + bgpu_last_error(ctx.get()));
    // the arena belongs to the library until the next bgpu_sdp_align on ctx: keep the blocks
    blocks_.assign(arena_.blocks, arena_.blocks + arena_.nBlocks);
  }
  template <typename T_Alignment>
  void Store(uint32_t i, T_Alignment &out) const {
    const bgpu_result &r = results_[i];
    if (r.status != BGPU_JOB_OK) throw Error(r.status, "blasr_gpu: SDPAlign job rejected");
    out.qPos = r.qPos; out.tPos = r.tPos;
    out.blocks.resize(r.nBlocks);
    for (uint32_t k = 0; k < r.nBlocks; k++) {
      const bgpu_block &bk = blocks_[r.blockOff + k];
      out.blocks[k].qPos = bk.qPos; out.blocks[k].tPos = bk.tPos; out.blocks[k].length = bk.length;
    }
  }
  void Clear() { q_.clear(); t_.clear(); qOff_.assign(1, 0); tOff_.assign(1, 0); results_.clear(); blocks_.clear(); }

 private:
  std::vector<uint8_t> q_, t_;
  std::vector<uint64_t> qOff_{0}, tOff_{0};
  std::vector<bgpu_result> results_;
  std::vector<bgpu_block> blocks_;
  bgpu_arena arena_{};
};

// MapReadToGenome (common/algorithms/anchoring/MapBySuffixArray.h:209-309) for the reads a MapReads instance holds: the two
// calls of alignment/Blasr.cpp:2282-2296 (read and readRC) become two Add()s, Run() takes the reference's AnchorParameters
// object as it is, Store(i, matchPosList) fills the vector the call would have filled (ChainedMatchPos or MatchPos: anything
// constructible from (t, q, l)).  The index is loaded once per device with LoadIndex() from the reference's own
// DNASuffixArray / DNASequence objects (what Blasr.cpp:4425-4470 reads from the .sa file).
class AnchorBatch {
 public:
  template <typename T_SuffixArray, typename T_RefSequence>
  static void LoadIndex(Context &ctx, const T_SuffixArray &sa, const T_RefSequence &genome) {
    int rc = bgpu_set_reference(ctx.get(), (const uint8_t *)genome.seq, genome.length);
    if (rc == BGPU_OK)
      rc = bgpu_set_suffix_array(ctx.get(), sa.index, genome.length, sa.startPosTable, sa.endPosTable, sa.startPosTable ? sa.lookupPrefixLength : 0);
    if (rc != BGPU_OK) throw Error(rc, std::string("blasr_gpu: ") + bgpu_last_error(ctx.get()));
  }
  template <typename T_Sequence>
  void Add(const T_Sequence &read) {              // SMRTSequence: seq, length, subreadStart, subreadEnd
    bases_.insert(bases_.end(), (const uint8_t *)read.seq, (const uint8_t *)read.seq + read.length);
    off_.push_back(bases_.size()); subS_.push_back(read.subreadStart); subE_.push_back(read.subreadEnd);
  }
  uint32_t size() const { return (uint32_t)off_.size() - 1; }
  template <typename T_AnchorParameters>
  void Run(Context &ctx, unsigned int minPrefixMatchLength, const T_AnchorParameters &ap) {
    bgpu_anchor_params p; std::memset(&p, 0, sizeof p);
    p.minPrefixMatchLength = minPrefixMatchLength; p.minMatchLength = ap.minMatchLength; p.expand = ap.expand;
    p.useLookupTable = ap.useLookupTable; p.maxAnchorsPerPosition = ap.maxAnchorsPerPosition;
    p.advanceExactMatches = ap.advanceExactMatches; p.maxLCPLength = ap.maxLCPLength;
    p.stopMappingOnceUnique = ap.stopMappingOnceUnique; p.removeEncompassedMatches = ap.removeEncompassedMatches;
    matchOff_.assign(off_.size(), 0);
    const bgpu_match *m = nullptr;
    const int rc = bgpu_map_reads(ctx.get(), &p, bases_.data(), off_.data(), size(), subS_.data(), subE_.data(), matchOff_.data(), &m);
    if (rc != BGPU_OK) throw Error(rc, std::string("blasr_gpu: ") + bgpu_last_error(ctx.get()));
    matches_.assign(m, m + matchOff_.back());      // the library's buffer lives until the next bgpu_map_reads on ctx
  }
  // the return value of MapReadToGenome: matchPosList.size()
  template <typename T_MatchPos>
  int Store(uint32_t i, std::vector<T_MatchPos> &matchPosList) const {
    for (uint64_t k = matchOff_[i]; k < matchOff_[i + 1]; k++) matchPosList.push_back(T_MatchPos(matches_[k].t, matches_[k].q, matches_[k].l));
    return (int)matchPosList.size();
  }
  void Clear() { bases_.clear(); off_.assign(1, 0); subS_.clear(); subE_.clear(); matchOff_.clear(); matches_.clear(); }

 private:
  std::vector<uint8_t> bases_;
  std::vector<uint64_t> off_{0}, matchOff_;
  std::vector<uint32_t> subS_, subE_;
  std::vector<bgpu_match> matches_;
};

}  // namespace blasr_gpu
#endif

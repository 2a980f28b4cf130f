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
#ifndef BLASR_GPU_ADAPTER_HPP_
#define BLASR_GPU_ADAPTER_HPP_
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
#include <vector>
#include "blasr_gpu.h"

namespace blasr_gpu {

struct Error : std::runtime_error { int code; Error(int c, const std::string &m) : std::runtime_error(m), code(c) {} };

// One GPU context; the blasr thread driver (MapReads, Blasr.cpp:3193) creates one per device and shares it between
// its pthreads (calls are serialised inside the library).
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
  bgpu_ctx *ctx; bgpu_ticket ticket; int refs; std::mutex mu;
  SharedTicket(bgpu_ctx *c, bgpu_ticket t, int r) : ctx(c), ticket(t), refs(r) {}
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
    const size_t g0 = guide_.size();
    guide_.resize(g0 + blocks.size());
    for (size_t i = 0; i < blocks.size(); i++) { guide_[g0 + i].qPos = blocks[i].qPos; guide_[g0 + i].tPos = blocks[i].tPos; guide_[g0 + i].length = blocks[i].length; }
    q_.insert(q_.end(), q, q + qLen); t_.insert(t_.end(), t, t + tLen);
    if (qual) { qual_.resize(q_.size() - qLen, 0); qual_.insert(qual_.end(), qual, qual + qLen); }
    qOff_.push_back(q_.size()); tOff_.push_back(t_.size()); gOff_.push_back(guide_.size());
  }
  uint32_t size() const { return (uint32_t)qOff_.size() - 1; }

  // Runs AffineGuidedAlign (affine=true: what blasr does, MappingParameters.h:539) or GuidedAlign on every job and
  // ComputeAlignmentStats afterwards.  Results stay valid until the next Run()/destruction.
  template <typename T_ScoreFn>
  void Run(Context &ctx, const T_ScoreFn &fn, int bandSize, bool affine, int alignType = BGPU_GLOBAL) {
    bgpu_scorefn s = MakeScoreFn(fn);
    bgpu_params p;
    p.algo = affine ? BGPU_AFFINE_GUIDED : BGPU_GUIDED; p.alignType = alignType; p.band = bandSize;
    p.bndIns = p.bndDel = 0; p.doStats = 1; p.statsAffine = affine ? 1 : 0;
    if (!qual_.empty()) qual_.resize(q_.size(), 0);
    bgpu_batch b;
    std::memset(&b, 0, sizeof b);   // no rich QV tracks on this path (DistanceMatrixScoreFunction)
    b.nJobs = size(); b.qBases = q_.data(); b.qOff = qOff_.data(); b.tBases = t_.data(); b.tOff = tOff_.data();
    b.qual = qual_.empty() ? nullptr : qual_.data(); b.guide = guide_.data(); b.guideOff = gOff_.data(); b.band = nullptr;
    results_.resize(b.nJobs);
    Release();
    int rc = bgpu_submit(ctx.get(), &s, &p, &b, &ticket_);
    if (rc == BGPU_OK) { owner_ = ctx.get(); rc = bgpu_collect(owner_, ticket_, results_.data(), &arena_); }
    if (rc != BGPU_OK) throw Error(rc, std::string("blasr_gpu: ") + bgpu_last_error(ctx.get()));
    cigarOps_ = nullptr; cigarOff_ = nullptr; base_ = 0;
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
  void Clear() { Release(); q_.clear(); t_.clear(); qual_.clear(); guide_.clear(); qOff_.assign(1, 0); tOff_.assign(1, 0); gOff_.assign(1, 0); results_.clear(); }

 private:
  SharedTicket *shared_ = nullptr;                                 // set when the jobs ran inside a merged submission
  uint32_t base_ = 0;                                              // ... and the index of this batch's first job in it
  std::vector<uint8_t> q_, t_, qual_;
  std::vector<bgpu_block> guide_;
  std::vector<uint64_t> qOff_{0}, tOff_{0}, gOff_{0};
  std::vector<bgpu_result> results_;
  bgpu_arena arena_{};
  bgpu_ctx *owner_ = nullptr; bgpu_ticket ticket_ = nullptr;       // results and the arena live until Release()
  const uint32_t *cigarOps_ = nullptr; const uint64_t *cigarOff_ = nullptr;
};

// Cross-thread batching for blasr's thread driver.  MapReads (Blasr.cpp:3193) runs one pthread per -nproc, and each
// calls RefineAlignments synchronously with the handful of candidates of ONE read -- far too few jobs to fill a GPU.
// RefineService keeps that structure untouched: every pthread hands its RefineBatch to Run() and blocks; the first
// waiter becomes the leader, gathers the batches of the other threads until every client that is not already being
// served has arrived (or a short window closes), submits them as ONE ticket and hands each thread its slice of the
// results.  Three contexts keep three merged tickets in flight (a ticket's latency is the serial sweep of its longest read).
class RefineService {
 public:
  RefineService(int device, int nClients, int maxWaitUs = 300, int nContexts = 3)
      : nClients_(nClients < 1 ? 1 : nClients), maxWaitUs_(maxWaitUs) {
    for (int i = 0; i < (nContexts < 1 ? 1 : nContexts); i++) ctxs_.push_back(std::unique_ptr<Context>(new Context(device)));
    busy_.assign(ctxs_.size(), false);
  }

  template <typename T_ScoreFn>
  void Run(RefineBatch &batch, const T_ScoreFn &fn, int bandSize, bool affine, int alignType = BGPU_GLOBAL) {
    Request r;
    r.batch = &batch; r.fn = MakeScoreFn(fn);
    std::memset(&r.p, 0, sizeof r.p);
    r.p.algo = affine ? BGPU_AFFINE_GUIDED : BGPU_GUIDED; r.p.alignType = alignType; r.p.band = bandSize;
    r.p.doStats = 1; r.p.statsAffine = affine ? 1 : 0;
    batch.Release();
    std::unique_lock<std::mutex> lk(mu_);
    pending_.push_back(&r);
    cv_.notify_all();
    while (!r.done) {
      if (leaderActive_ || r.taken) { cv_.wait(lk); continue; }   // someone else gathers, or already carries this request
      leaderActive_ = true;
      // The leader holds its ticket back until a context is free: a ticket queued behind a busy context gains nothing,
      // while the requests that arrive in the meantime make the next ticket larger (batching by back-pressure).
      const auto t0 = std::chrono::steady_clock::now();
      const auto deadline = t0 + std::chrono::microseconds(maxWaitUs_);
      int c = -1;
      for (;;) {
        c = -1;
        for (size_t i = 0; i < busy_.size(); i++) if (!busy_[i]) { c = (int)i; break; }
        const bool full = (int)pending_.size() + inflight_ >= nClients_;
        const bool late = std::chrono::steady_clock::now() >= deadline;
        if (c >= 0 && (full || late)) break;
        if (c < 0 || late) cv_.wait(lk); else cv_.wait_until(lk, deadline);
      }
      std::vector<Request *> mine, rest;          // one ticket = one (score function, parameters) pair
      for (Request *x : pending_)
        (std::memcmp(&x->fn, &r.fn, sizeof r.fn) == 0 && std::memcmp(&x->p, &r.p, sizeof r.p) == 0 ? mine : rest).push_back(x);
      pending_.swap(rest);
      for (Request *x : mine) x->taken = true;
      inflight_ += (int)mine.size();
      busy_[c] = true;
      leaderActive_ = false;
      cv_.notify_all();
      lk.unlock();
      const auto e0 = std::chrono::steady_clock::now();
      Execute(*ctxs_[c], mine);
      const double ems = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - e0).count();
      lk.lock();
      executeMs_ += ems; gatherMs_ += std::chrono::duration<double, std::milli>(e0 - t0).count();
      busy_[c] = false;
      inflight_ -= (int)mine.size();
      for (Request *x : mine) x->done = true;
      cv_.notify_all();
    }
    lk.unlock();
    if (r.rc != BGPU_OK) throw Error(r.rc, "blasr_gpu: " + r.err);
  }

  uint64_t Tickets() const { return tickets_; }
  uint64_t Jobs() const { return jobs_; }
  ~RefineService() {
    if (getenv("BGPU_SERVICE_STATS"))
      fprintf(stderr, "RefineService: %llu tickets, %llu jobs (%.1f per ticket); per ticket: gather %.2f ms, execute %.2f ms (bgpu_submit %.2f, bgpu_collect %.2f, kernels on the device %.2f)\n",
              (unsigned long long)tickets_, (unsigned long long)jobs_, tickets_ ? (double)jobs_ / tickets_ : 0.0,
              tickets_ ? gatherMs_ / tickets_ : 0.0, tickets_ ? executeMs_ / tickets_ : 0.0, tickets_ ? submitMs_ / tickets_ : 0.0,
              tickets_ ? collectMs_ / tickets_ : 0.0, tickets_ ? deviceMs_ / tickets_ : 0.0);
  }

 private:
  struct Request { RefineBatch *batch; bgpu_scorefn fn; bgpu_params p; bool taken = false, done = false; int rc = BGPU_OK; std::string err; };

  void Execute(Context &ctx, std::vector<Request *> &reqs) {
    std::vector<uint8_t> q, t, qual; std::vector<bgpu_block> guide; std::vector<uint64_t> qOff(1, 0), tOff(1, 0), gOff(1, 0);
    bool anyQual = false;
    for (Request *x : reqs) anyQual = anyQual || !x->batch->qual_.empty();
    for (Request *x : reqs) {
      RefineBatch &b = *x->batch;
      const uint64_t q0 = q.size(), t0 = t.size(), g0 = guide.size();
      q.insert(q.end(), b.q_.begin(), b.q_.end()); t.insert(t.end(), b.t_.begin(), b.t_.end());
      guide.insert(guide.end(), b.guide_.begin(), b.guide_.end());
      if (anyQual) { qual.resize(q0, 0); qual.insert(qual.end(), b.qual_.begin(), b.qual_.end()); qual.resize(q.size(), 0); }
      for (uint32_t i = 1; i <= b.size(); i++) { qOff.push_back(q0 + b.qOff_[i]); tOff.push_back(t0 + b.tOff_[i]); gOff.push_back(g0 + b.gOff_[i]); }
    }
    bgpu_batch mb; std::memset(&mb, 0, sizeof mb);
    mb.nJobs = (uint32_t)qOff.size() - 1; mb.qBases = q.data(); mb.qOff = qOff.data(); mb.tBases = t.data(); mb.tOff = tOff.data();
    mb.qual = anyQual ? qual.data() : nullptr; mb.guide = guide.data(); mb.guideOff = gOff.data();
    std::vector<bgpu_result> res(mb.nJobs);
    bgpu_arena arena; std::memset(&arena, 0, sizeof arena);
    bgpu_ticket tk = nullptr;
    const auto s0 = std::chrono::steady_clock::now();
    int rc = mb.nJobs ? bgpu_submit(ctx.get(), &reqs[0]->fn, &reqs[0]->p, &mb, &tk) : BGPU_OK;
    const auto s1 = std::chrono::steady_clock::now();
    if (rc == BGPU_OK && tk) rc = bgpu_collect(ctx.get(), tk, res.data(), &arena);
    if (rc == BGPU_OK && tk) {
      bgpu_timing tm; bgpu_timing_of(ctx.get(), tk, &tm);
      std::lock_guard<std::mutex> lk(mu_);
      submitMs_ += std::chrono::duration<double, std::milli>(s1 - s0).count();
      collectMs_ += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - s1).count();
      deviceMs_ += tm.msTotal;
    }
    if (rc != BGPU_OK) {
      const std::string e = bgpu_last_error(ctx.get());
      if (tk) bgpu_release(ctx.get(), tk);
      for (Request *x : reqs) { x->rc = rc; x->err = e; }
      return;
    }
    { std::lock_guard<std::mutex> lk(mu_); tickets_++; jobs_ += mb.nJobs; }
    SharedTicket *sh = tk ? new SharedTicket(ctx.get(), tk, (int)reqs.size()) : nullptr;
    uint32_t at = 0;
    for (Request *x : reqs) {
      RefineBatch &b = *x->batch;
      b.results_.assign(res.begin() + at, res.begin() + at + b.size());
      b.base_ = at;
      at += b.size();
      b.arena_ = arena; b.owner_ = ctx.get(); b.ticket_ = nullptr; b.shared_ = sh; b.cigarOps_ = nullptr; b.cigarOff_ = nullptr;
    }
  }

  const int nClients_, maxWaitUs_;
  std::mutex mu_; std::condition_variable cv_;
  std::vector<Request *> pending_;
  bool leaderActive_ = false; int inflight_ = 0;
  std::vector<bool> busy_;
  std::vector<std::unique_ptr<Context>> ctxs_;
  uint64_t tickets_ = 0, jobs_ = 0;
  double executeMs_ = 0, gatherMs_ = 0, submitMs_ = 0, collectMs_ = 0, deviceMs_ = 0;
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

}  // namespace blasr_gpu
#endif

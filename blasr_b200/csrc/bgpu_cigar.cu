// bgpu_cigar.cu -- output formatting of a guided ticket on the device: the SAM CIGAR and the three
// alignment strings m5 prints, straight from the run-length path the traceback left in HBM.
//
// Reference semantics restated:
//   * SAMOutput::CreateNoClippingCigarOps (common/algorithms/alignment/printers/SAMPrinter.h:203-293), the branch for
//     alignments whose gap lists are filled (what every aligner here produces): AddGaps(gaps[0]); for each block b
//     { AddUngappedOperations(b); AddGaps(gaps[b+1]); }.  AddUngappedOperations (:138-166) splits a block into maximal runs of
//     unequal ('X') and equal ('=') RAW sequence bytes (no case folding, no base codes); AddGaps (:120-137) prints Gap::Query
//     as 'D' and Gap::Target as 'I', one op per Gap.  Adjacent ops are never merged.
//   * CreateCIGARString (:345-400) around it: 'H' / 'S' ops for the hard / soft clipped prefix, the core, 'S' / 'H' for the
//     suffix, the whole list reversed when tStrand == 1.  The clip lengths come from read-level fields (SetHardClip :311-327,
//     SetSoftClip :295-309), so the caller passes them per job.
//   * CreateAlignmentStrings (common/algorithms/alignment/AlignmentUtils.h:390-533), gap-list branch: per block the target
//     bytes, '|' where TwoBit[q] == TwoBit[t] else '*', the query bytes; per Gap::Query (deletion) base t / ' ' / '-', per
//     Gap::Target (insertion) base '-' / ' ' / q.
// Ops come out BAM-packed: length << 4 | code ('=' 7, 'X' 8, 'I' 1, 'D' 2, 'S' 4, 'H' 5).
//
// Device mapping: warp per job.  The run list is taken 32 runs at a time (their starts go to shared memory); inside a chunk the
// lanes sweep the alignment COLUMNS, 32 per step, and find the run a column belongs to with a 5-step search -- so a warp
// formats 32 consecutive columns per step whatever the run lengths are (a kept gap run counts as one column for the CIGAR).
// CIGAR ops: a column starts an op when it opens a block or flips between match and mismatch; ballots rank the starts, every
// start records (code, column); a second sweep turns consecutive starts into lengths and writes the ops in final order
// (clips added, reversed for tStrand == 1).
#include "bgpu_common.cuh"

namespace bgpu {

enum { RUN_D = 0, RUN_U = 1, RUN_L = 2 };
enum { CIG_I = 1, CIG_D = 2, CIG_S = 4, CIG_H = 5, CIG_EQ = 7, CIG_X = 8 };
enum { FMT_COUNT = 0, FMT_CIGAR = 1, FMT_STRINGS = 2 };
constexpr int FMT_WARPS = 4;

__device__ __forceinline__ uint32_t warp_excl(uint32_t v, int lane, uint32_t &total) {
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
  total = __shfl_sync(0xffffffffu, x, 31);
  return x - v;
}
// TwoBit[] of NucConversion.h as CreateAlignmentStrings uses it: ACGT in either case -> 0..3, everything else 255
__device__ __forceinline__ uint32_t two_bit(uint8_t c) {
  switch (c & 0xDF) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 255; }
}

struct FmtArgs {
  // FMT_COUNT: out
  uint32_t *nOps;            // per job: CIGAR ops of the core
  uint32_t *nCols;           // per job: length of the alignment strings
  // FMT_CIGAR
  const uint64_t *opOff;     // per job: first op (clips included) in ops[]; tmp arrays are indexed by coreOff
  const uint64_t *coreOff;   // per job: first core op in tmpCode / tmpPos
  uint32_t *tmpCode, *tmpPos;
  uint32_t *ops;
  const uint32_t *clips;     // [nJobs][4] hard prefix, soft prefix, soft suffix, hard suffix, or NULL
  const uint8_t *tStrand;    // per job, or NULL
  // FMT_STRINGS
  const uint64_t *strOff;
  char *text, *align, *query;
};

template <int MODE>
__global__ void __launch_bounds__(FMT_WARPS * 32) fmt_kernel(BatchDev B, FmtArgs A) {
  __shared__ uint32_t sStart[FMT_WARPS][33], sPQ[FMT_WARPS][32], sPT[FMT_WARPS][32], sLen[FMT_WARPS][32];
  __shared__ uint8_t sType[FMT_WARPS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t job = blockIdx.x * FMT_WARPS + w;
  if (job >= B.nJobs) return;
  const JobGeom &G = B.geom[job];
  if (G.status != BGPU_JOB_OK || G.nBlocks == 0) {
    if (MODE == FMT_COUNT && lane == 0) { A.nOps[job] = 0; A.nCols[job] = 0; }
    return;
  }
  const uint32_t nRuns = G.nRuns, nBlocks = G.nBlocks;
  const uint32_t *runs = B.runs + G.runOff;
  const uint8_t *qb = B.q + B.qOff[job] + G.qStart;   // raw bytes at path offset 0
  const uint8_t *tb = B.t + B.tOff[job] + G.tStart;
  uint32_t cq = 0, ct = 0, cD = 0, opBase = 0, colBase = 0, vBase = 0;
  uint32_t *tCode = MODE == FMT_CIGAR ? A.tmpCode + A.coreOff[job] : nullptr;
  uint32_t *tPos = MODE == FMT_CIGAR ? A.tmpPos + A.coreOff[job] : nullptr;
  char *oText = nullptr, *oAlign = nullptr, *oQuery = nullptr;
  if (MODE == FMT_STRINGS) { oText = A.text + A.strOff[job]; oAlign = A.align + A.strOff[job]; oQuery = A.query + A.strOff[job]; }

  for (uint32_t base = 0; base < nRuns; base += 32) {
    const uint32_t f = base + lane;                   // forward run index (the traceback stored them end-to-start)
    const bool act = f < nRuns;
    uint32_t type = 3, len = 0;
    if (act) { const uint32_t r = runs[nRuns - 1 - f]; type = r >> 30; len = r & 0x3fffffffu; }
    const uint32_t dq = (type == RUN_D || type == RUN_U) ? len : 0, dt = (type == RUN_D || type == RUN_L) ? len : 0;
    uint32_t totQ, totT, totD, totC;
    const uint32_t pq = cq + warp_excl(dq, lane, totQ), pt = ct + warp_excl(dt, lane, totT);
    const bool isD = act && type == RUN_D;
    const uint32_t dBefore = cD + warp_excl(isD ? 1u : 0u, lane, totD);
    // the guided aligners fold gap runs before the first block into qPos / tPos and drop the ones after the last block
    const bool kept = act && !isD && dBefore >= 1 && dBefore < nBlocks;
    // columns this run contributes: CIGAR sweep -> a kept gap is ONE column (one op); string sweep -> its bases
    const uint32_t cols = isD ? len : (kept ? (MODE == FMT_STRINGS ? len : 1u) : 0u);
    const uint32_t cstart = warp_excl(cols, lane, totC);
    __syncwarp();
    sStart[w][lane] = cstart; sPQ[w][lane] = pq; sPT[w][lane] = pt; sLen[w][lane] = len; sType[w][lane] = (uint8_t)(isD ? RUN_D : (kept ? type : 3));
    if (lane == 0) sStart[w][32] = totC;
    __syncwarp();
    if (MODE == FMT_COUNT) colBase += isD || kept ? len : 0;   // per-lane partial, reduced at the end
    for (uint32_t c0 = 0; c0 < totC; c0 += 32) {
      const uint32_t c = c0 + lane;
      const bool on = c < totC;
      // the run holding column c: the last one whose start is <= c (empty runs share their successor's start and sort before it)
      int r = 0;
#pragma unroll
      for (int step = 16; step; step >>= 1) if (r + step < 32 && sStart[w][r + step] <= (on ? c : 0u)) r += step;
      const uint32_t off = on ? c - sStart[w][r] : 0, ty = sType[w][r];
      const uint32_t qi = sPQ[w][r] + (ty == RUN_L ? 0 : off), ti = sPT[w][r] + (ty == RUN_U ? 0 : off);
      if (MODE == FMT_STRINGS) {
        if (on) {
          const uint8_t qc = ty == RUN_L ? (uint8_t)'-' : qb[qi], tc = ty == RUN_U ? (uint8_t)'-' : tb[ti];
          const char ac = ty == RUN_D ? (two_bit(qc) != two_bit(tc) ? '*' : '|') : ' ';
          const size_t o = (size_t)colBase + c;
          oText[o] = (char)tc; oAlign[o] = ac; oQuery[o] = (char)qc;
        }
      } else {
        bool start = false; uint32_t code = 0;
        if (on) {
          if (ty == RUN_D) {
            const bool m = qb[qi] == tb[ti];
            start = off == 0 || m != (qb[qi - 1] == tb[ti - 1]);
            code = m ? CIG_EQ : CIG_X;
          } else { start = true; code = ((ty == RUN_L ? CIG_D : CIG_I)) | (sLen[w][r] << 4); }   // a gap op carries its length already
        }
        const unsigned bal = __ballot_sync(0xffffffffu, start);
        if (MODE == FMT_CIGAR && start) {
          const uint32_t k = opBase + __popc(bal & ((1u << lane) - 1u));
          tCode[k] = code; tPos[k] = vBase + c;
        }
        opBase += __popc(bal);
      }
    }
    if (MODE == FMT_STRINGS) colBase += totC;
    vBase += totC;
    cq += totQ; ct += totT; cD += totD;
  }
  if (MODE == FMT_COUNT) {
#pragma unroll
    for (int o = 16; o; o >>= 1) colBase += __shfl_xor_sync(0xffffffffu, colBase, o);
    if (lane == 0) { A.nOps[job] = opBase; A.nCols[job] = colBase; }
  }
  if (MODE == FMT_CIGAR) {
    // second sweep: a match / mismatch op runs until the next start; final order = [H][S] core [S][H], reversed for tStrand 1
    __threadfence_block(); __syncwarp();
    uint32_t clip[4] = {0, 0, 0, 0};
    if (A.clips) { for (int i = 0; i < 4; i++) clip[i] = A.clips[4 * (size_t)job + i]; }
    const uint32_t nPre = (clip[0] ? 1u : 0u) + (clip[1] ? 1u : 0u), nSuf = (clip[2] ? 1u : 0u) + (clip[3] ? 1u : 0u);
    const uint32_t nAll = nPre + opBase + nSuf;
    const bool rev = A.tStrand && A.tStrand[job] == 1;
    uint32_t *out = A.ops + A.opOff[job];
    auto put = [&](uint32_t idx, uint32_t v) { out[rev ? nAll - 1 - idx : idx] = v; };
    if (lane == 0) {
      uint32_t k = 0;
      if (clip[0]) put(k++, (clip[0] << 4) | CIG_H);
      if (clip[1]) put(k++, (clip[1] << 4) | CIG_S);
      k = nPre + opBase;
      if (clip[2]) put(k++, (clip[2] << 4) | CIG_S);
      if (clip[3]) put(k++, (clip[3] << 4) | CIG_H);
    }
    for (uint32_t k = lane; k < opBase; k += 32) {
      const uint32_t code = tCode[k];
      uint32_t v = code;
      if ((code & 15u) >= CIG_EQ) v = (((k + 1 < opBase ? tPos[k + 1] : vBase) - tPos[k]) << 4) | code;
      put(nPre + k, v);
    }
  }
}

void launch_fmt_count(const BatchDev &B, uint32_t *nOps, uint32_t *nCols, cudaStream_t s) {
  FmtArgs A{}; A.nOps = nOps; A.nCols = nCols;
  const unsigned grid = (B.nJobs + FMT_WARPS - 1) / FMT_WARPS;
  if (grid) fmt_kernel<FMT_COUNT><<<grid, FMT_WARPS * 32, 0, s>>>(B, A);
}

// per-job exclusive offsets of (count[i] + extra[i]) and of count[i]; totals[0], totals[1]; one CTA
__global__ void __launch_bounds__(1024) fmt_scan_kernel(uint32_t n, const uint32_t *count, const uint32_t *clips, uint64_t *offAll,
                                                        uint64_t *offCore, uint64_t *totals) {
  __shared__ unsigned long long sh[2][32];
  __shared__ unsigned long long carry[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 2) carry[tid] = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + tid;
    unsigned long long v[2] = {0, 0};
    if (i < n) {
      v[1] = count[i];
      uint32_t extra = 0;
      if (clips && count[i]) for (int k = 0; k < 4; k++) extra += clips[4 * (size_t)i + k] ? 1u : 0u;
      v[0] = v[1] + extra;
    }
    unsigned long long inc[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
      unsigned long long x = v[c];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
      inc[c] = x;
      if (lane == 31) sh[c][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int c = 0; c < 2; c++) {
        unsigned long long x = sh[c][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
        sh[c][lane] = x;
      }
    }
    __syncthreads();
    if (i < n) {
      offAll[i] = carry[0] + (warp ? sh[0][warp - 1] : 0) + inc[0] - v[0];
      if (offCore) offCore[i] = carry[1] + (warp ? sh[1][warp - 1] : 0) + inc[1] - v[1];
    }
    __syncthreads();
    if (tid < 2) carry[tid] += sh[tid][31];
    __syncthreads();
  }
  if (tid == 0) { offAll[n] = carry[0]; if (offCore) offCore[n] = carry[1]; totals[0] = carry[0]; totals[1] = carry[1]; }
}

void launch_fmt_scan(uint32_t n, const uint32_t *count, const uint32_t *clips, uint64_t *offAll, uint64_t *offCore, uint64_t *totals,
                     cudaStream_t s) {
  fmt_scan_kernel<<<1, 1024, 0, s>>>(n, count, clips, offAll, offCore, totals);
}

void launch_fmt_cigar(const BatchDev &B, const uint64_t *opOff, const uint64_t *coreOff, uint32_t *tmpCode, uint32_t *tmpPos,
                      uint32_t *ops, const uint32_t *clips, const uint8_t *tStrand, cudaStream_t s) {
  FmtArgs A{}; A.opOff = opOff; A.coreOff = coreOff; A.tmpCode = tmpCode; A.tmpPos = tmpPos; A.ops = ops; A.clips = clips; A.tStrand = tStrand;
  const unsigned grid = (B.nJobs + FMT_WARPS - 1) / FMT_WARPS;
  if (grid) fmt_kernel<FMT_CIGAR><<<grid, FMT_WARPS * 32, 0, s>>>(B, A);
}

void launch_fmt_strings(const BatchDev &B, const uint64_t *strOff, char *text, char *align, char *query, cudaStream_t s) {
  FmtArgs A{}; A.strOff = strOff; A.text = text; A.align = align; A.query = query;
  const unsigned grid = (B.nJobs + FMT_WARPS - 1) / FMT_WARPS;
  if (grid) fmt_kernel<FMT_STRINGS><<<grid, FMT_WARPS * 32, 0, s>>>(B, A);
}

}  // namespace bgpu

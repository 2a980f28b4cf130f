// bgpu_cigar.cu -- the SAM CIGAR core of every alignment of a guided ticket, built on the device.
//
// Reference semantics restated: SAMOutput::CreateNoClippingCigarOps (common/algorithms/alignment/printers/SAMPrinter.h:203-293)
// for alignments whose gap lists are filled (its `nGaps > 0` branch, which is what every aligner here produces):
//   AddGaps(gaps[0]); for each block b { AddUngappedOperations(b); AddGaps(gaps[b+1]); }
// AddUngappedOperations (:138-166) splits a block into maximal runs of unequal ('X') and equal ('=') RAW sequence bytes
// (no case folding, no base codes); AddGaps (:120-137) prints Gap::Query as 'D' and Gap::Target as 'I', one op per Gap.
// Adjacent ops are never merged.  Ops come out BAM-packed: length << 4 | code, codes '=' 7, 'X' 8, 'I' 1, 'D' 2.
//
// Device mapping: warp per job over the run list the traceback left in HBM (runs are stored end-to-start).  Every lane
// takes one run; a gap run that the alignment keeps is one op, a diagonal run is walked by its lane, byte pair by byte
// pair.  Two passes share this code: COUNT (ops per job, scanned on the host into offsets) and WRITE.
// The per-lane walk is the simple form (a warp waits for its longest run); it is an optional formatting step, not part
// of the timed hot path.
#include "bgpu_common.cuh"

namespace bgpu {

enum { RUN_D = 0, RUN_U = 1, RUN_L = 2 };
enum { CIG_I = 1, CIG_D = 2, CIG_EQ = 7, CIG_X = 8 };

__device__ __forceinline__ uint32_t warp_excl(uint32_t v, int lane, uint32_t &total) {
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
  total = __shfl_sync(0xffffffffu, x, 31);
  return x - v;
}

template <bool WRITE>
__global__ void __launch_bounds__(128) cigar_kernel(BatchDev B, uint32_t *counts, const uint64_t *cigOff, uint32_t *ops) {
  const int lane = threadIdx.x & 31;
  const uint32_t job = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (job >= B.nJobs) return;
  const JobGeom &G = B.geom[job];
  if (G.status != BGPU_JOB_OK || G.nBlocks == 0) { if (!WRITE && lane == 0) counts[job] = 0; return; }
  const uint32_t nRuns = G.nRuns, nBlocks = G.nBlocks;
  const uint32_t *runs = B.runs + G.runOff;
  const uint8_t *qb = B.q + B.qOff[job] + G.qStart;   // raw bytes at path offset 0
  const uint8_t *tb = B.t + B.tOff[job] + G.tStart;
  uint32_t *out = WRITE ? ops + cigOff[job] : nullptr;
  uint32_t cq = 0, ct = 0, cD = 0, opBase = 0;
  for (uint32_t base = 0; base < nRuns; base += 32) {
    const uint32_t f = base + lane;                   // forward run index
    const bool act = f < nRuns;
    uint32_t type = 3, len = 0;
    if (act) { const uint32_t r = runs[nRuns - 1 - f]; type = r >> 30; len = r & 0x3fffffffu; }
    const uint32_t dq = (type == RUN_D || type == RUN_U) ? len : 0, dt = (type == RUN_D || type == RUN_L) ? len : 0;
    uint32_t totQ, totT, totD, totOps;
    const uint32_t pq = cq + warp_excl(dq, lane, totQ), pt = ct + warp_excl(dt, lane, totT);
    const bool isD = act && type == RUN_D;
    const uint32_t dBefore = cD + warp_excl(isD ? 1u : 0u, lane, totD);
    // the guided aligners fold gap runs before the first block into qPos / tPos and drop the ones after the last block
    const bool kept = act && !isD && dBefore >= 1 && dBefore < nBlocks;
    const uint8_t *qq = qb + pq, *tt = tb + pt;
    uint32_t c = kept ? 1u : 0u;
    if (isD) {                                        // maximal runs of equal / unequal bytes
      bool prev = qq[0] == tt[0];
      c = 1;
      for (uint32_t i = 1; i < len; i++) { const bool m = qq[i] == tt[i]; c += m != prev ? 1u : 0u; prev = m; }
    }
    const uint32_t my = opBase + warp_excl(c, lane, totOps);
    if (WRITE) {
      if (kept) out[my] = (len << 4) | (type == RUN_L ? CIG_D : CIG_I);
      if (isD) {
        bool prev = qq[0] == tt[0];
        uint32_t start = 0, k = my;
        for (uint32_t i = 1; i < len; i++) {
          const bool m = qq[i] == tt[i];
          if (m != prev) { out[k++] = ((i - start) << 4) | (prev ? CIG_EQ : CIG_X); start = i; prev = m; }
        }
        out[k] = ((len - start) << 4) | (prev ? CIG_EQ : CIG_X);
      }
    }
    cq += totQ; ct += totT; cD += totD; opBase += totOps;
  }
  if (!WRITE && lane == 0) counts[job] = opBase;
}

void launch_cigar_count(const BatchDev &B, uint32_t *counts, cudaStream_t s) {
  const unsigned grid = (B.nJobs + 3) / 4;
  if (grid) cigar_kernel<false><<<grid, 128, 0, s>>>(B, counts, nullptr, nullptr);
}
void launch_cigar_write(const BatchDev &B, const uint64_t *cigOff, uint32_t *ops, cudaStream_t s) {
  const unsigned grid = (B.nJobs + 3) / 4;
  if (grid) cigar_kernel<true><<<grid, 128, 0, s>>>(B, nullptr, cigOff, ops);
}

}  // namespace bgpu

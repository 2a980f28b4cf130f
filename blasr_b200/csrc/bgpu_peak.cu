// bgpu_peak.cu -- integer-pipe roofline denominator: MEASURED_PEAKS.json carries no int32
// peak, so the library measures it on the bound device with independent add / min chains
// (16 accumulators per thread, enough ILP to saturate issue) and reports lane-ops per second.
#include "bgpu_common.cuh"

namespace bgpu {

// Every chain reads its neighbour chain, so ptxas cannot fold the sequence into fewer instructions.
//   MODE 0: x[i] += x[i+1]            (IADD3, alu pipe)
//   MODE 1: x[i] = min(x[i], x[i+1])  (VIMNMX, alu pipe)
//   MODE 2: x[i] = x[i+1] * m + x[i]  (IMAD, fma pipe)
//   MODE 3: even chains IADD3, odd chains IMAD (both pipes busy)
template <int MODE>
__global__ void __launch_bounds__(256) int_peak_kernel(int *out, int iters, int a, int m) {
  int x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = threadIdx.x * a + i + out[i & 1];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int y = x[(i + 1) & 15];
        if (MODE == 0 || (MODE == 3 && (i & 1) == 0)) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(y));
        else if (MODE == 1) asm volatile("min.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(y));
        else asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(y), "r"(m));
      }
    }
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s ^= x[i];
  if (s == 0x7fffffff) out[0] = s;
}

double g_peakByMode[4] = {0, 0, 0, 0};

double measure_int_peak(int nSM, cudaStream_t s, double *clockMHz) {
  int *d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess) return 0;
  const int iters = 4096, grid = nSM * 8, block = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  cudaMemsetAsync(d, 0, 64, s);
  for (int mode = 0; mode < 4; mode++) {
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0, s);
      if (mode == 0) int_peak_kernel<0><<<grid, block, 0, s>>>(d, iters, 3, 5);
      else if (mode == 1) int_peak_kernel<1><<<grid, block, 0, s>>>(d, iters, 3, 5);
      else if (mode == 2) int_peak_kernel<2><<<grid, block, 0, s>>>(d, iters, 3, 5);
      else int_peak_kernel<3><<<grid, block, 0, s>>>(d, iters, 3, 5);
      cudaEventRecord(e1, s);
      cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      const double ops = (double)grid * block * (double)iters * 64.0;
      if (rep > 0 && ms > 0) { const double r = ops / (ms * 1e-3); g_peakByMode[mode] = g_peakByMode[mode] > r ? g_peakByMode[mode] : r; best = best > r ? best : r; }
    }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  if (clockMHz) {
    int khz = 0; int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    *clockMHz = khz / 1000.0;
  }
  return best;
}

}  // namespace bgpu

// bgpu_peak.cu -- integer-pipe roofline denominator: MEASURED_PEAKS.json carries no int32
// peak, so the library measures it on the bound device with independent add / min chains
// (16 accumulators per thread, enough ILP to saturate issue) and reports lane-ops per second.
#include "bgpu_common.cuh"

namespace bgpu {

template <int MODE>   // 0: add.s32 only, 1: min.s32 only, 2: alternating add / min on separate chains
__global__ void __launch_bounds__(256) int_peak_kernel(int *out, int iters, int a, int b) {
  int x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i * a;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        if (MODE == 0 || (MODE == 2 && (i & 1) == 0)) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
        else asm volatile("min.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(b - i - r));
      }
    }
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s ^= x[i];
  if (s == 0x7fffffff) out[0] = s;
}

double measure_int_peak(int nSM, cudaStream_t s, double *clockMHz) {
  int *d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess) return 0;
  const int iters = 4096, grid = nSM * 8, block = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int mode = 0; mode < 3; mode++) {
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0, s);
      if (mode == 0) int_peak_kernel<0><<<grid, block, 0, s>>>(d, iters, 3, 1 << 30);
      else if (mode == 1) int_peak_kernel<1><<<grid, block, 0, s>>>(d, iters, 3, 1 << 30);
      else int_peak_kernel<2><<<grid, block, 0, s>>>(d, iters, 3, 1 << 30);
      cudaEventRecord(e1, s);
      cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      const double ops = (double)grid * block * (double)iters * 64.0;
      if (rep > 0 && ms > 0) best = best > ops / (ms * 1e-3) ? best : ops / (ms * 1e-3);
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

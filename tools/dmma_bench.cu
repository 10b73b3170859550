/* dmma_bench -- FP64 tensor-core (mma.sync.m8n8k4.f64, SASS DMMA) throughput on this GPU, alone and interleaved
 * with DFMA, to size the multi-right-hand-side kernel (BASELINE config C5). */
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double seed) {
  double c[16][2], f[8];
  for (int i = 0; i < 16; ++i) { c[i][0] = seed + i; c[i][1] = seed - i; }
  for (int i = 0; i < 8; ++i) f[i] = seed * i + threadIdx.x * 1e-3;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE != 1) dmma(c[i][0], c[i][1], a, b);
      if (MODE != 0 && i < 8) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  for (int i = 0; i < 8; ++i) s += f[i];
  if (s == 123.456) out[blockIdx.x] = s;
}
template <int MODE>
void run(const char* name, double* out) {
  const int grid = 148 * 4, iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(out, 100, 1.0);
  cudaEventRecord(e0); k<MODE><<<grid, 256>>>(out, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warps = grid * 8.0;
  const double dmma_flop = MODE != 1 ? 2.0 * 256 * 16 * iters * warps : 0, dfma_flop = MODE != 0 ? 2.0 * 32 * 8 * iters * warps : 0;
  printf("%-28s %8.3f ms  DMMA %6.2f TFLOP/s  DFMA %6.2f TFLOP/s\n", name, ms, dmma_flop / ms / 1e9, dfma_flop / ms / 1e9);
}
int main() {
  double* out; cudaMalloc(&out, 4096 * 8);
  run<0>("DMMA only", out);
  run<1>("DFMA only (8 chains)", out);
  run<2>("DMMA + DFMA interleaved", out);
  return 0;
}

// microbench.cu -- operand-path micro-benchmarks for the Phi kernels on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu && ./microbench
// Measures per-SM throughput of: LDS.128, tcgen05.ld (TMEM -> registers) in several shapes,
// both together, SHFL, dependent DFMA latency.  Numbers feed DESIGN.md's kernel model.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NW>
__global__ void __launch_bounds__(NW * 32) lds_kernel(double* out, int iters, long long* cyc) {
  extern __shared__ __align__(16) unsigned char sm[];
  double* s = reinterpret_cast<double*>(sm);
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) s[i] = i * 1e-3;
  __syncthreads();
  const uint32_t base = smem_u32(s) + 16 * (threadIdx.x & 31);
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  uint32_t off = (threadIdx.x >> 5) * 1024;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      double x, y;
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(base + ((off + u * 1024) & 0x1FFFF)));
      a0 += x; a1 += y;
    }
    off += 8192;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (a0 + a1 + a2 + a3 == 1.2345) out[0] = a0;
}

// TMEM: 128 lanes x 512 columns x 32 bit.  Each warp reads its own 32-lane quadrant.
template <int NW, int X>
__global__ void __launch_bounds__(NW * 32) tmem_kernel(double* out, int iters, long long* cyc, int with_lds) {
  __shared__ uint32_t tbase_s;
  extern __shared__ __align__(16) unsigned char sm[];
  double* s = reinterpret_cast<double*>(sm);
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = i * 1e-3;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tbase_s + ((uint32_t)(32 * (warp & 3)) << 16);
  // fill the quadrant (only the first 4 warps need to, others share lanes)
  if (warp < 4) {
    for (int c = 0; c < 512; c += 2) {
      uint32_t v0 = __double2loint(1.0 + c * 1e-3 + threadIdx.x), v1 = __double2hiint(1.0 + c * 1e-3 + threadIdx.x);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(tb + c), "r"(v0), "r"(v1));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  double acc0 = 0, acc1 = 0;
  const uint32_t lbase = smem_u32(s) + 16 * (threadIdx.x & 31);
  uint32_t col = (warp * 37) & 255;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t c = ((col + u * 2 * X) & 255) * 2 > 512 - 2 * X ? 0 : ((col + u * 2 * X) & 255);
      if constexpr (X == 1) {
        uint32_t r0, r1;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(tb + (c & ~1u)));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        acc0 += __hiloint2double(r1, r0);
      } else if constexpr (X == 2) {
        uint32_t r0, r1, r2, r3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(tb + (c & ~3u)));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        acc0 += __hiloint2double(r1, r0); acc1 += __hiloint2double(r3, r2);
      } else {
        uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "r"(tb + (c & ~7u)));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        acc0 += __hiloint2double(r1, r0) + __hiloint2double(r5, r4); acc1 += __hiloint2double(r3, r2) + __hiloint2double(r7, r6);
      }
      if (with_lds) {
        double x, y;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(lbase + ((it * 4 + u) & 63) * 512));
        acc0 += x; acc1 += y;
      }
    }
    col += 17;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc0 + acc1 == 1.2345) out[0] = acc0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase_s));
}

// tcgen05.ld issued back to back with ONE wait per group of 4 (pipelined)
template <int NW>
__global__ void __launch_bounds__(NW * 32) tmem_pipelined_kernel(double* out, int iters, long long* cyc) {
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tbase_s + ((uint32_t)(32 * (warp & 3)) << 16);
  if (warp < 4) {
    for (int c = 0; c < 512; c += 2) {
      uint32_t v0 = __double2loint(1.0 + c * 1e-3), v1 = __double2hiint(1.0 + c * 1e-3);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(tb + c), "r"(v0), "r"(v1));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  uint32_t col = (warp * 38) & 254;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t a0, a1, b0, b1, c0, c1, d0, d1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(tb + ((col) & 510)));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(tb + ((col + 34) & 510)));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(c0), "=r"(c1) : "r"(tb + ((col + 90) & 510)));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(d0), "=r"(d1) : "r"(tb + ((col + 150) & 510)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc0 = fma(__hiloint2double(a1, a0), 1.0000001, acc0);
    acc1 = fma(__hiloint2double(b1, b0), 1.0000001, acc1);
    acc2 = fma(__hiloint2double(c1, c0), 1.0000001, acc2);
    acc3 = fma(__hiloint2double(d1, d0), 1.0000001, acc3);
    col += 6;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc0 + acc1 + acc2 + acc3 == 1.2345) out[0] = acc0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase_s));
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) shfl_kernel(double* out, int iters, long long* cyc) {
  double v = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) v += __shfl_xor_sync(0xffffffffu, v, 1 + (u & 15));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (v == 1.2345) out[0] = v;
}

__global__ void dfma_latency_kernel(double* out, int iters, long long* cyc) {
  double a = threadIdx.x * 1e-9 + 1.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) a = fma(a, 0.999999, 1e-7);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (a == 1.2345) out[0] = a;
}

int main() {
  double* out; long long* cyc;
  CK(cudaMalloc(&out, 1024)); CK(cudaMalloc(&cyc, 1024 * sizeof(long long)));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  std::vector<long long> h(1024);
  auto report = [&](const char* name, double bytes_per_cta, int nblk) {
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), cyc, nblk * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < nblk; ++i) mean += h[i]; mean /= nblk;
    printf("%-46s cycles/CTA %10.0f   %8.2f B/clk/SM\n", name, mean, bytes_per_cta / mean);
  };
  const int iters = 4000;
  CK(cudaFuncSetAttribute(lds_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CK(cudaFuncSetAttribute(lds_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  lds_kernel<4><<<sms, 128, 140 * 1024>>>(out, iters, cyc); report("LDS.128, 4 warps/SM", 4.0 * 32 * 16 * 8 * iters, sms);
  lds_kernel<16><<<sms, 512, 140 * 1024>>>(out, iters, cyc); report("LDS.128, 16 warps/SM", 16.0 * 32 * 16 * 8 * iters, sms);
#define TM(NW, X, L, name) { CK(cudaFuncSetAttribute(tmem_kernel<NW, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); \
    tmem_kernel<NW, X><<<sms, NW * 32, 70 * 1024>>>(out, iters, cyc, L); CK(cudaGetLastError()); \
    report(name, (double)NW * 32 * 8 * X * 4 * iters, sms); }
  TM(4, 1, 0, "tcgen05.ld 32x32b.x2 +wait each, 4 warps");
  TM(8, 1, 0, "tcgen05.ld 32x32b.x2 +wait each, 8 warps");
  TM(16, 1, 0, "tcgen05.ld 32x32b.x2 +wait each, 16 warps");
  TM(16, 2, 0, "tcgen05.ld 32x32b.x4 +wait each, 16 warps");
  TM(16, 4, 0, "tcgen05.ld 32x32b.x8 +wait each, 16 warps");
  TM(16, 1, 1, "tcgen05.ld x2 + LDS.128 interleaved, 16 warps (TMEM bytes)");
  tmem_pipelined_kernel<4><<<sms, 128>>>(out, iters, cyc); CK(cudaGetLastError()); report("tcgen05.ld x2, 4 in flight per wait, 4 warps", 4.0 * 32 * 8 * 4 * iters, sms);
  tmem_pipelined_kernel<16><<<sms, 512>>>(out, iters, cyc); CK(cudaGetLastError()); report("tcgen05.ld x2, 4 in flight per wait, 16 warps", 16.0 * 32 * 8 * 4 * iters, sms);
  shfl_kernel<4><<<sms, 128>>>(out, iters, cyc); report("SHFL 64-bit (2 SHFL each), 4 warps  [B = 8/lane]", 4.0 * 32 * 8 * 8 * iters, sms);
  shfl_kernel<16><<<sms, 512>>>(out, iters, cyc); report("SHFL 64-bit (2 SHFL each), 16 warps [B = 8/lane]", 16.0 * 32 * 8 * 8 * iters, sms);
  dfma_latency_kernel<<<1, 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  cudaMemcpy(h.data(), cyc, sizeof(long long), cudaMemcpyDeviceToHost);
  printf("dependent DFMA latency: %.2f cycles\n", (double)h[0] / (iters * 16.0));
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}

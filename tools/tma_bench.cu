/* tma_bench -- per-SM throughput of cp.async.bulk (1-D TMA) global -> shared as a function of the
 * copy size: one producer thread per CTA keeps DEPTH copies in flight, 148 CTAs, source streamed
 * from HBM (working set >> L2) or L2-resident.  Prints cycles per copy and B/clk/SM. */
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256, 1) k(const char* src, size_t span, int bytes, int ncopies, int depth, int batch, int nprod, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bars = (uint64_t*)sm;          /* depth barriers */
  unsigned char* buf = sm + 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 64; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x & 31) return;
  const int pw = threadIdx.x >> 5;       /* producer warp index: each has its own barriers and buffers */
  if (pw >= nprod) return;
  bars += pw * 8; buf += (size_t)pw * (180 * 1024 / nprod);
  ncopies /= nprod;
  long long t0 = clock64();
  const int ngroups = ncopies / batch;
  for (int gidx = 0; gidx < ngroups; ++gidx) {
    const int s = gidx % depth;
    if (gidx >= depth) {
      const uint32_t par = ((gidx / depth) - 1) & 1;
      asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bars[s])), "r"(par) : "memory");
    }
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(bytes * batch) : "memory");
    for (int j = 0; j < batch; ++j) {
      const size_t i = (size_t)gidx * batch + j;
      const char* g = src + ((((i * 8u + pw) * 148u + blockIdx.x) * (size_t)bytes) & (span - 1));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf + ((size_t)s * batch + j) * bytes)),
                   "l"(g), "r"(bytes), "r"(smem_u32(&bars[s])) : "memory");
    }
  }
  for (int gidx = ngroups > depth ? ngroups - depth : 0; gidx < ngroups; ++gidx) { /* drain */
    const int s = gidx % depth; const uint32_t par = (gidx / depth) & 1;
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bars[s])), "r"(par) : "memory");
  }
  if (pw == 0) out[blockIdx.x] = clock64() - t0;
}
int main() {
  const size_t big = (size_t)2 << 30, small = (size_t)32 << 20;
  char* src; cudaMalloc(&src, big); cudaMemset(src, 1, big);
  long long* out; cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int pass = 1; pass < 2; ++pass)
    for (int bytes : {512, 1024, 4096})
      for (int nprod : {1, 2, 4, 8})
        for (int batch : {16}) {
          const int depth = 2;
          if ((size_t)depth * batch * bytes > 180 * 1024 / nprod) continue;
          const int n = 8192;
          const size_t span = pass ? small : big;
          k<<<148, 256, 200 * 1024>>>(src, span, bytes, n, depth, batch, nprod, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[148]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
          double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
          printf("%s bytes %6d batch %3d producers %d : %8.1f cycles/copy(SM-wide)  %6.1f B/clk/SM\n", pass ? "L2 " : "HBM", bytes, batch, nprod, avg / n, bytes * (double)n / avg);
        }
  return 0;
}

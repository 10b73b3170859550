"""Instruction-fetch micro-benchmark (round 2): how fast can an SM issue FP64 FMAs when every warp runs its OWN
straight-line loop body (the situation of phi_t_spec: one generated instruction stream per warp) compared with all
warps sharing one body (phi_a_spec), as a function of the body length and of the warps per SM sub-partition?

Generates a .cu with loop bodies of L independent-chain DFMAs (8 accumulators, operands permuted per copy so the
compiler cannot merge the copies), compiles it with nvcc for sm_100a and runs it.  Prints DFMA/clk/SM (peak 2.0 = one
warp-wide DFMA per 2 cycles on each of the 4 sub-partitions).

  python tools/ifetch_bench.py            (needs a B200)
"""
import subprocess
import sys
import tempfile
from pathlib import Path

LENGTHS = [96, 192, 384, 768, 1536]
PADS = [0]      # DFMAs of never-executed code between two bodies: moves the bodies apart in the address space
COPIES = 12


def body(L, k):
    out = []
    for i in range(L):
        j = i % 8
        out.append(f"a{j} = fma(a{j}, x{(j + k) % 8}, y{(i + 3 * k) % 4});")
    return "\n      ".join(out)


def source():
    s = ["#include <cstdio>\n#include <cuda_runtime.h>\n"]
    for L, PAD in [(L, P) for L in LENGTHS for P in PADS]:
        pad = (lambda k: f"if (iters == -7 - {k}) {{\n      {body(PAD, k + 5)}\n    }}\n    ") if PAD else (lambda k: "")
        cases = "\n".join(f"    case {k}: {pad(k)}for (int it = 0; it < iters; ++it) {{\n      {body(L, k)}\n    }} break;" for k in range(COPIES))
        s.append(f"""
__global__ void __launch_bounds__(384, 1) k_{L}_{PAD}(double* out, const double* in, int iters, int distinct, long long* cyc) {{
  const int warp = threadIdx.x >> 5;
  double x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3], x4 = in[4], x5 = in[5], x6 = in[6], x7 = in[7];
  double y0 = in[8], y1 = in[9], y2 = in[10], y3 = in[11];
  double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  __syncthreads();
  const long long t0 = clock64();
  switch (distinct ? warp : 0) {{
{cases}
  }}
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}}
""")
    runs = "\n".join(
        f"""  for (int nw : {{4, 8, 12}}) for (int distinct = 0; distinct < 2; ++distinct) {{
    const int iters = {max(8, 400000 // L)};
    k_{L}_{PAD}<<<148, nw * 32>>>(out, in, iters, distinct, cyc);
    k_{L}_{PAD}<<<148, nw * 32>>>(out, in, iters, distinct, cyc);
    cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("body %5d DFMA (%6d B)  warps/SM %2d  %s  %.3f DFMA/clk/SM\\n", {L}, {L} * 16, nw, distinct ? "one body per warp " : "all warps one body", (double)nw * {L} * iters / (double)c);
  }}""" for L in LENGTHS for PAD in PADS)
    s.append(f"""
int main() {{
  double *out, *in; long long* cyc;
  cudaMalloc(&out, 148 * 384 * 8); cudaMalloc(&in, 12 * 8); cudaMalloc(&cyc, 8);
  double h[12]; for (int i = 0; i < 12; ++i) h[i] = 1.0 + 1e-9 * i;
  cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
{runs}
  printf("%s\\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}}
""")
    return "".join(s)


def main():
    d = Path(tempfile.mkdtemp())
    (d / "ifetch.cu").write_text(source())
    subprocess.run(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(d / "ifetch"), str(d / "ifetch.cu")], check=True)
    if "--compile-only" in sys.argv:
        print(d / "ifetch")
        return
    subprocess.run([str(d / "ifetch")], check=True)


if __name__ == "__main__":
    main()

"""BASELINE config C5: multi right-hand-side Phi.A (64 columns), N=1M, K=2000, one B200: Phi.A on the FP64 tensor cores
(phi_am_spec, DMMA), Phi^T.A as a column loop over phi_t_spec.  Prints one JSON line; tensor_pipe_frac = dense-equivalent
TFLOP/s over the DMMA peak measured by tools/dmma_bench.cu (36.9 TFLOP/s on B200)."""
import json, sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import torch
import outerbase_b200 as obp
import bench
lib = obp.lib(0)
lib.set_option("spec", 1)
om, terms = bench.setup_model(lib)
N, K, C = bench.N_TOTAL, bench.K_TERMS, 64
x = bench.synth_rows(0, N, bench.D)
ob = lib.outerbase(om, x, dograd=False)
ob.set_terms(terms)
ob.specialize(terms)
dev = torch.device("cuda", 0)
ld = ((N + 255) // 256) * 256
A = torch.randn(K, C, dtype=torch.float64, device=dev).t().contiguous().t()   # column-major K x C
R = torch.randn(ld, C, dtype=torch.float64, device=dev).t().contiguous().t()  # column-major ld x C
out_n = torch.empty(ld * C, dtype=torch.float64, device=dev)
out_k = torch.empty(K * C, dtype=torch.float64, device=dev)
stream = torch.cuda.ExternalStream(lib.stream())
def run():
    ob.mm_mat_dev(A.data_ptr(), C, out_n.data_ptr())
    ob.tmm_mat_dev(R.data_ptr(), C, out_k.data_ptr())
for _ in range(2): run()
lib.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(stream); ob.mm_mat_dev(A.data_ptr(), C, out_n.data_ptr()); e[1].record(stream)
ob.tmm_mat_dev(R.data_ptr(), C, out_k.data_ptr()); e[2].record(stream)
lib.synchronize(); torch.cuda.synchronize()
t_a, t_t = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
flop = 2.0 * N * K * C
print(json.dumps({"workload": f"C5 Phi.A (phi_am_spec, FP64 DMMA) / Phi^T.A (column loop), N={N} K={K} C={C}",
                  "tensor_pipe_frac_phi_A": flop / (t_a * 1e-3) / 1e12 / 36.9,
                  "ms_phi_A": t_a, "ms_phiT_A": t_t, "dense_equivalent_tflops_phi_A": flop / (t_a * 1e-3) / 1e12,
                  "dense_equivalent_tflops_phiT_A": flop / (t_t * 1e-3) / 1e12}))

"""Times the host-pointer calls outerbase::mm / tmm (C3 shapes) with pageable and page-locked buffers: which transfers
overlap with the kernels (OuterBase::mm / tmm, ob_engine.hpp)."""
import sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import torch
import bench
import outerbase_b200 as obp

lib = obp.lib(0); lib.set_option("spec", 1)
om, terms = bench.setup_model(lib)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
x = bench.synth_rows(0, N, bench.D)
ob = lib.outerbase(om, x, dograd=False); ob.specialize(terms)
K = terms.shape[0]
rng = np.random.default_rng(1)
a, r = rng.normal(size=K), rng.normal(size=N)


def T(f, n=10):
    f(); lib.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    lib.synchronize(); return (time.perf_counter() - t0) / n * 1e3


yp = torch.empty(N, dtype=torch.float64).pin_memory().numpy()
rp = torch.empty(N, dtype=torch.float64).pin_memory().numpy(); rp[:] = r
gp = torch.empty(K, dtype=torch.float64).pin_memory().numpy()
ypage, gpage = np.empty(N), np.empty(K)
print("matmul  pageable out  %.3f ms" % T(lambda: ob.matmul(terms, a, out=ypage)))
print("matmul  pinned out    %.3f ms" % T(lambda: ob.matmul(terms, a, out=yp)))
print("tmatmul pageable in   %.3f ms" % T(lambda: ob.tmatmul(terms, r, out=gpage)))
print("tmatmul pinned in     %.3f ms" % T(lambda: ob.tmatmul(terms, rp, out=gp)))

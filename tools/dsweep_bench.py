"""First GPU run of phi_d_spec (hyper-gradients in one sweep, option dsweep): parity against the
per-hyper-parameter path on the same GPU objects, then the cost of one BFGS objective evaluation with and without it.

  python tools/dsweep_bench.py [--config c3|c4share] [--rows N]

c3: BASELINE C3 (d=10, N=1M, K=2000); c4share: one GPU's share of C4 (d=20, N=1.25M, K=4000).  Prints one JSON line.
Run it under `timeout` the first time: the kernel has only been verified on the CPU (tests/test_spec_generator.py)."""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import bench  # noqa: E402
import outerbase_b200 as obp  # noqa: E402
from outerbase_b200 import fitting  # noqa: E402
from conftest import borehole8d, relerr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3", choices=["c3", "c4share"])
ap.add_argument("--rows", type=int, default=0)
args = ap.parse_args()
lib = obp.lib(0); lib.set_option("spec", 1)
if args.config == "c3":
    D, K, N = bench.D, bench.K_TERMS, args.rows or 1_000_000
    om, terms = bench.setup_model(lib)
    x = bench.synth_rows(0, N, D); y = bench.wingweight(x)
else:
    D, K, N = 20, 4000, args.rows or 1_250_000
    x = bench.synth_rows(0, N, D, seed=7)
    y = borehole8d(x[:, :8]) + 20 * np.sin(3 * x[:, 8]) * x[:, 9] + 10 * x[:, 10:].sum(1)
    om = lib.outermod(); om.setcovfs(["mat25pow"] * D); om.setknot(fitting.genknotlist([40] * D, x[:100_000]))
    hyp = om.gethyp(); hyp[0::2] = np.linspace(-0.6, 0.4, D); om.updatehyp(hyp)
    terms = om.selectterms(K)
y = (y - y.mean()) / y.std(ddof=1)
loglik = lib.loglik_gauss(om, terms, y, x)
vec = lib.lpdfvec(lib.logpr_gauss(om, terms), loglik); vec.domarg = True


def T(f, n=3):
    f()
    lib.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        f()
    lib.synchronize(); return (time.perf_counter() - t0) / n * 1e3


def evaluation():  # what .lpdfwrapper does per objective evaluation (R/outersupport.R:209-226)
    vec.updateom(); vec.set_coeff(np.zeros(K)); vec.optcg(0.001, 100)


out = {"config": args.config, "N": N, "d": D, "K": K}
res = {}
for mode in (0, 1):
    lib.set_option("dsweep", mode)
    t_first = time.perf_counter(); evaluation(); lib.synchronize(); t_first = time.perf_counter() - t_first
    res[mode] = dict(gradhyp=np.array(vec.gradhyp), val=vec.val, iters=vec.cg_iters)
    out[f"dsweep{mode}"] = dict(first_evaluation_s=t_first, evaluation_ms=T(evaluation), optcg_warm_ms=T(lambda: (vec.set_coeff(np.zeros(K)), vec.optcg(0.001, 100))),
                                val=vec.val, cg_iters=vec.cg_iters)
lib.set_option("dsweep", 0)
out["gradhyp_relerr_sweep_vs_per_hyper"] = relerr(res[1]["gradhyp"], res[0]["gradhyp"])
out["val_relerr"] = abs(res[1]["val"] - res[0]["val"]) / abs(res[0]["val"])
out["same_cg_iters"] = res[1]["iters"] == res[0]["iters"]
print(json.dumps(out))

"""Wall times of the fitting drivers on one B200 (BASELINE configs C2 and, reduced, C4):
  C2  borehole d=8, N=100k, K=1000: lpdfvec(logpr_gauss, loglik_gauss).optcg(0.001, 100), then BFGS_lpdf
  C4r synthetic d=20, K=4000, N=--rows4 (default 1.25M = one GPU's share of 10M over 8): optcg
usage: python tools/fit_bench.py [--rows4 N] [--no-bfgs]"""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import outerbase_b200 as obp
from outerbase_b200 import fitting
from conftest import borehole8d

ap = argparse.ArgumentParser(); ap.add_argument("--rows4", type=int, default=1_250_000); ap.add_argument("--no-bfgs", action="store_true")
args = ap.parse_args()
lib = obp.lib(0)
out = {}

def timed(f):
    lib.synchronize(); t0 = time.perf_counter(); r = f(); lib.synchronize(); return time.perf_counter() - t0, r

# ---- C2
rng = np.random.default_rng(42)
N, d, K = 100_000, 8, 1000
x = np.asfortranarray(rng.uniform(size=(N, d)))
y = borehole8d(x); y = (y - y.mean()) / y.std(ddof=1)
om = lib.outermod(); om.setcovfs(["mat25pow"] * d); om.setknot(fitting.genknotlist([40] * d, x))
terms = om.selectterms(K)
t_build, loglik = timed(lambda: lib.loglik_gauss(om, terms, y, x))
logpdf = lib.lpdfvec(lib.logpr_gauss(om, terms), loglik); logpdf.domarg = True
t1, _ = timed(lambda: logpdf.optcg(0.001, 100))
it1 = logpdf.cg_iters
logpdf.set_coeff(np.zeros(K))
t2, _ = timed(lambda: logpdf.optcg(0.001, 100))
out["C2"] = dict(N=N, d=d, K=K, build_s=t_build, optcg_first_s=t1, optcg_warm_s=t2, cg_iters=it1, val=logpdf.val)
if not args.no_bfgs:
    n0 = lib.launch_count()
    tb, res = timed(lambda: fitting.BFGS_lpdf(om, logpdf))
    out["C2"]["bfgs_lpdf"] = dict(wall_s=tb, bfgs_iters=res["iters"], objective_start=res["history"][0]["obj"], objective_end=res["optid"]["val"],
                                  kernel_launches=lib.launch_count() - n0, spec_state=None)
# ---- C2 again with the terms-specialised kernels forced (the default policy does not compile for a table that
# sees only ~1 s of work: the 2 s compile would not pay back within this single run)
if not args.no_bfgs:
    lib.set_option("spec", 1)
    om2 = lib.outermod(); om2.setcovfs(["mat25pow"] * d); om2.setknot(fitting.genknotlist([40] * d, x))
    terms2 = om2.selectterms(K)
    loglik2 = lib.loglik_gauss(om2, terms2, y, x)
    logpdf2 = lib.lpdfvec(lib.logpr_gauss(om2, terms2), loglik2); logpdf2.domarg = True
    tc, _ = timed(lambda: logpdf2.optcg(0.001, 100))
    logpdf2.set_coeff(np.zeros(K))
    tw, _ = timed(lambda: logpdf2.optcg(0.001, 100))
    tb2, res2 = timed(lambda: fitting.BFGS_lpdf(om2, logpdf2))
    out["C2_spec"] = dict(optcg_first_s=tc, optcg_warm_s=tw, bfgs_lpdf_wall_s=tb2, bfgs_iters=res2["iters"], objective_end=res2["optid"]["val"])
    lib.set_option("spec", 2)
# ---- C4 reduced
N4, d4, K4 = args.rows4, 20, 4000
x4 = np.asfortranarray(np.random.default_rng(7).uniform(size=(N4, d4)))
y4 = borehole8d(x4[:, :8]) + 20 * np.sin(3 * x4[:, 8]) * x4[:, 9] + 10 * x4[:, 10:].sum(1)
y4 = (y4 - y4.mean()) / y4.std(ddof=1)
om4 = lib.outermod(); om4.setcovfs(["mat25pow"] * d4); om4.setknot(fitting.genknotlist([40] * d4, x4[:100_000]))
hyp = om4.gethyp(); hyp[0::2] = np.linspace(-0.6, 0.4, d4); om4.updatehyp(hyp)
terms4 = om4.selectterms(K4)
lib.set_option("spec", 1)
t_build4, loglik4 = timed(lambda: lib.loglik_gauss(om4, terms4, y4, x4))
logpdf4 = lib.lpdfvec(lib.logpr_gauss(om4, terms4), loglik4); logpdf4.domarg = True
t41, _ = timed(lambda: logpdf4.optcg(0.001, 100))
it4 = logpdf4.cg_iters
logpdf4.set_coeff(np.zeros(K4))
t42, _ = timed(lambda: logpdf4.optcg(0.001, 100))
nnz = (np.asarray(terms4) > 0).sum(1)
out["C4_one_gpu_share"] = dict(N=N4, d=d4, K=K4, W=int((nnz + 1).sum()), build_s=t_build4, optcg_first_s=t41, optcg_warm_s=t42, cg_iters=it4,
                               pairs=2 * it4 + 2, note="first call includes the run-time compile of the specialised kernels for K=4000")
print(json.dumps(out))

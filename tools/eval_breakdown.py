import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench, outerbase_b200 as obp
lib = obp.lib(0); lib.set_option("spec", 1)
om, terms = bench.setup_model(lib)
N = 1_000_000
x = bench.synth_rows(0, N, bench.D); y = bench.wingweight(x); y = (y - y.mean()) / y.std(ddof=1)
loglik = lib.loglik_gauss(om, terms, y, x)
vec = lib.lpdfvec(lib.logpr_gauss(om, terms), loglik); vec.domarg = True
vec.optcg(0.001, 100)
def T(f, n=3):
    lib.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    lib.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("updateom (basis rebuild with gradients) ms", T(lambda: loglik.updateom()))
print("loglik.diaghess ms", T(lambda: loglik.diaghess()))
print("loglik.diaghessgradhyp ms", T(lambda: loglik.diaghessgradhyp()))
def cg():
    vec.set_coeff(np.zeros(2000)); vec.optcg(0.001, 100)
print("optcg warm (no hess redo) ms", T(cg))
def full():
    vec.updateom(); vec.set_coeff(np.zeros(2000)); vec.optcg(0.001, 100)
print("updateom + optcg (one BFGS evaluation) ms", T(full))

"""GPU parity on the terms tables of the BASELINE configs themselves ("Untested configs").

The shapes of tests/test_gpu_parity.py are the reference's test shapes (d = 8, K <= 2000).  Here the tables are the
ones bench.py and the tools measure: C2 (d = 8, K = 1000), C3 (d = 10, K = 2000, all mat25pow, 40 quantile knots: bench.setup_model),
C4 (d = 20, K = 4000) and C5 (C3's table with 64 right-hand sides), on a 20 000-row sample the oracle finishes in
seconds, through the stateless linalg.h seam (bit-identical basemat from the oracle), BOTH kernel families
(interpreter: spec = 0, terms-specialised: spec = 1).  Two tolerances per product:

* norm-wise   max|gpu - oracle| <= 1e-12 * max|oracle|            (north_star: "matvecs within 1e-12 relative");
* element-wise |gpu - oracle|[i] <= ELEM_TOL * (|Phi| |a|)[i]     (resp. (|Phi|^T |r|)[k]): every output entry is
  correct relative to the magnitude of the terms that were summed into it -- the entries of Phi^T r that belong to
  high-level terms are orders of magnitude below the maximum and the norm-wise check alone would not see them.

The C1 golden fixture (tests/golden/c1_*.npz, frozen oracle outputs) is fed to the GPU as well.
"""
from pathlib import Path

import numpy as np
import pytest

from conftest import borehole8d, relerr

pytestmark = pytest.mark.gpu

NORM_TOL = 1e-12
ELEM_TOL = 1e-12   # relative to the sum of absolute values of what was added up (componentwise bound)
GE_NORM_TOL = 1e-11
GOLD = Path(__file__).parent / "golden"


def config_model(lib, cfg):
    """The outermod + terms table of a BASELINE config, exactly as bench.py / tools build them."""
    import bench
    if cfg in ("c3", "c5"):
        om, terms = bench.setup_model(lib)
        return om, terms, 10
    if cfg == "c2":  # borehole d = 8, K = 1000 (bench.py --config c2)
        om, terms = bench.config_model(lib, "c2")
        return om, terms, 8
    assert cfg == "c4"
    from outerbase_b200 import fitting
    D, K = 20, 4000
    x = bench.synth_rows(0, 100_000, D, seed=7)
    om = lib.outermod(); om.setcovfs(["mat25pow"] * D); om.setknot(fitting.genknotlist([40] * D, x))
    hyp = om.gethyp(); hyp[0::2] = np.linspace(-0.6, 0.4, D); om.updatehyp(hyp)
    return om, om.selectterms(K), D


_CACHE = {}


def oracle_config(oracle, cfg, N):
    key = (cfg, N)
    if key not in _CACHE:
        import bench
        om, terms, d = config_model(oracle, cfg)
        x = bench.synth_rows(0, N, d, seed=7 if d == 20 else 42)
        ob = oracle.outerbase(om, x)
        _CACHE.clear()  # one config resident at a time (C4: 20000 x 2400 doubles)
        _CACHE[key] = dict(om=om, terms=terms, x=x, ob=ob, bm=ob.real("basemat"), bs=ob.real("basescale"), bg=ob.real("basemat_gradhyp"),
                           kp=om.index("knotptst"), gest=om.index("gest"), hm=om.index("hypmatch"))
    return _CACHE[key]


def check(got, want, mag, norm_tol=NORM_TOL, elem_tol=ELEM_TOL, what=""):
    got, want, mag = np.asarray(got), np.asarray(want), np.asarray(mag)
    assert got.shape == want.shape, what
    assert relerr(got, want) < norm_tol, (what, relerr(got, want))
    floor = np.finfo(float).tiny
    ratio = np.abs(got - want) / np.maximum(mag, floor)
    assert ratio.max() <= elem_tol, (what, float(ratio.max()), int(ratio.argmax()))


@pytest.mark.parametrize("family", ["interpreter", "specialised"])
@pytest.mark.parametrize("cfg,N", [("c2", 20000), ("c3", 20000), ("c4", 12000)])
def test_config_tables_against_the_oracle(gpu, oracle, cfg, N, family):
    """prodmm_/tprodmm_/prodmmge_/tprodmmge_ (src/linalg.cpp:102-131, 303-355, 225-277, 394-471) on the C2 / C3 / C4 tables."""
    o = oracle_config(oracle, cfg, N)
    terms, K = o["terms"], o["terms"].shape[0]
    rng = np.random.default_rng(3)
    a = np.sqrt(o["om"].getvar(terms) / 20) * rng.normal(size=K)
    r = rng.normal(size=N)
    abm, abs_ = np.abs(o["bm"]), np.abs(o["bs"])
    gpu.set_option("spec", 1 if family == "specialised" else 0)
    try:
        n0 = gpu.launch_count()
        # Phi a
        want = oracle.prodmm(terms, a, o["bm"], o["bs"], o["kp"])
        mag = oracle.prodmm(terms, np.abs(a), abm, abs_, o["kp"])
        check(gpu.prodmm(terms, a, o["bm"], o["bs"], o["kp"]), want, mag, what=f"{cfg} prodmm {family}")
        # Phi^T r
        want = oracle.tprodmm(terms, r, o["bm"], o["bs"], o["kp"])
        mag = oracle.tprodmm(terms, np.abs(r), abm, abs_, o["kp"])
        check(gpu.tprodmm(terms, r, o["bm"], o["bs"], o["kp"]), want, mag, what=f"{cfg} tprodmm {family}")
        assert gpu.launch_count() > n0
        # hyper-gradients (norm-wise per column: the entries are differences of products, a magnitude bound is not meaningful)
        a2 = rng.normal(size=K) / 100
        out, outge = gpu.prodmmge(terms, a2, o["bm"], o["bs"], o["kp"], o["bg"], o["gest"], o["hm"])
        ro, rge = oracle.prodmmge(terms, a2, o["bm"], o["bs"], o["kp"], o["bg"], o["gest"], o["hm"])
        assert relerr(out, ro) < NORM_TOL
        for h in range(rge.shape[1]):
            assert relerr(outge[:, h], rge[:, h]) < GE_NORM_TOL, (cfg, family, "prodmmge", h)
        out, outge = gpu.tprodmmge(terms, r, o["bm"], o["bs"], o["kp"], o["bg"], o["gest"], o["hm"])
        ro, rge = oracle.tprodmmge(terms, r, o["bm"], o["bs"], o["kp"], o["bg"], o["gest"], o["hm"])
        assert relerr(out, ro) < NORM_TOL
        for h in range(rge.shape[1]):
            assert relerr(outge[:, h], rge[:, h]) < GE_NORM_TOL, (cfg, family, "tprodmmge", h)
    finally:
        gpu.set_option("spec", 2)


@pytest.mark.parametrize("family", ["interpreter", "specialised"])
def test_c5_multi_rhs_against_the_oracle(gpu, oracle, family):
    """BASELINE C5: prodmm_(mat) / tprodmm_(mat) (src/linalg.cpp:527-557, 583-637) with 64 columns on C3's table."""
    N, C = 20000, 64
    o = oracle_config(oracle, "c5", N)
    terms, K = o["terms"], o["terms"].shape[0]
    rng = np.random.default_rng(5)
    A = np.asfortranarray(rng.normal(size=(K, C)) * np.sqrt(o["om"].getvar(terms) / 20)[:, None])
    R = np.asfortranarray(rng.normal(size=(N, C)))
    abm, abs_ = np.abs(o["bm"]), np.abs(o["bs"])
    gpu.set_option("spec", 1 if family == "specialised" else 0)
    try:
        want = oracle.prodmm(terms, A, o["bm"], o["bs"], o["kp"])
        mag = oracle.prodmm(terms, np.abs(A), abm, abs_, o["kp"])
        check(gpu.prodmm(terms, A, o["bm"], o["bs"], o["kp"]), want, mag, what=f"c5 prodmm(mat) {family}")
        want = oracle.tprodmm(terms, R, o["bm"], o["bs"], o["kp"])
        mag = oracle.tprodmm(terms, np.abs(R), abm, abs_, o["kp"])
        check(gpu.tprodmm(terms, R, o["bm"], o["bs"], o["kp"]), want, mag, what=f"c5 tprodmm(mat) {family}")
    finally:
        gpu.set_option("spec", 2)


@pytest.mark.parametrize("cfg,N", [("c3", 20000), ("c4", 12000)])
def test_config_handle_path_against_the_oracle(gpu, oracle, cfg, N):
    """The device-resident objects on the same tables: basis built by the GPU, products by the specialised kernels,
    loglik_gauss / lpdfvec / optcg against the oracle (1e-8: the two basis builds differ by rounding)."""
    import bench
    o = oracle_config(oracle, cfg, N)
    omg, terms_g, d = config_model(gpu, cfg)
    terms = o["terms"]
    np.testing.assert_array_equal(terms_g, terms)  # selectterms bit-exact at the config's size
    x = o["x"]
    y = bench.wingweight(x) if d == 10 else borehole8d(x[:, :8]) + 20 * np.sin(3 * x[:, 8]) * x[:, 9] + 10 * x[:, 10:].sum(1)
    y = (y - y.mean()) / y.std(ddof=1)
    res = {}
    gpu.set_option("spec", 1)
    try:
        for name, lib, om in (("o", oracle, o["om"]), ("g", gpu, omg)):
            vec = lib.lpdfvec(lib.logpr_gauss(om, terms), lib.loglik_gauss(om, terms, y, x))
            vec.domarg = True
            vec.optcg(0.001, 100)
            res[name] = dict(val=vec.val, coeff=np.array(vec.coeff), iters=vec.cg_iters, gradhyp=np.array(vec.gradhyp), gradpara=np.array(vec.gradpara))
    finally:
        gpu.set_option("spec", 2)
    g, w = res["g"], res["o"]
    assert g["iters"] == w["iters"]
    assert abs(g["val"] - w["val"]) <= 1e-8 * abs(w["val"])
    assert relerr(g["coeff"], w["coeff"]) < 1e-8
    assert relerr(g["gradhyp"], w["gradhyp"]) < 1e-6 and relerr(g["gradpara"], w["gradpara"]) < 1e-7


@pytest.mark.parametrize("family", ["interpreter", "specialised"])
def test_c1_golden_fixture_on_the_gpu(gpu, family):
    """BASELINE C1 (borehole d=8, N=1000, K=60): the committed golden vectors (frozen oracle outputs, tests/golden/) fed
    to the CUDA path -- terms bit-exact, products and the CG fit within the north_star tolerances."""
    g = np.load(GOLD / "c1_borehole_d8_n1000_k60.npz")
    gpu.set_option("spec", 1 if family == "specialised" else 0)
    try:
        om = gpu.outermod(); om.setcovfs([str(s) for s in g["covs"]]); om.setknot(list(g["knots"])); om.updatehyp(g["hyp"])
        terms = om.selectterms(int(g["K"]))
        np.testing.assert_array_equal(terms, g["terms"])
        ob = gpu.outerbase(om, g["x"])
        assert relerr(ob.matmul(terms, g["a"]), g["matmul"]) < 1e-9   # basis rebuilt on the GPU: build tolerance, not the matvec's
        assert relerr(ob.tmatmul(terms, g["r"]), g["tmatmul"]) < 1e-9
        vec = gpu.lpdfvec(gpu.logpr_gauss(om, terms), gpu.loglik_gauss(om, terms, g["y"], g["x"]))
        vec.optcg(0.001, 100)
        assert vec.cg_iters == int(g["cg_iters"])
        assert abs(vec.val - float(g["val"])) <= 1e-8 * abs(float(g["val"]))
        assert relerr(vec.coeff, g["coeff"]) < 1e-8
    finally:
        gpu.set_option("spec", 2)

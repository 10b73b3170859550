"""Host-side model state (outermod): the product's C++ core against the oracle.

`terms`, the index tables and `maxlevel` are integer-exact (north_star: "terms multi-index and
basis selection bit-exact"); the eigenbasis is compared bit for bit as well because both sides
use the same cyclic-Jacobi restatement of eig_sym -- and the oracle's eigen-decomposition is
pinned independently against LAPACK (scipy.linalg.eigh) in test_oracle_pinning.py."""
import numpy as np
import pytest

CASES = [
    (["mat25pow"] + ["mat25"] * 7, 20, 40),
    (["mat25"] * 8, 1000, 20),
    (["mat25pow"] * 10, 2000, 40),
    (["mat25ang", "mat25", "mat25pow"], 60, 33),
    (["mat25pow"] * 20, 400, 16),
]


def build(lib, covs, m, hypshift=0.0, seed=0):
    rng = np.random.default_rng(3)
    d = len(covs)
    om = lib.outermod()
    om.setcovfs(covs)
    kn = []
    for l, c in enumerate(covs):
        hi = 6.2 if c == "mat25ang" else 0.999
        kn.append(np.sort(rng.uniform(0.001, hi, size=m + (l % 3))))
    om.setknot(kn)
    if hypshift:
        h = om.gethyp()
        om.updatehyp(h + hypshift * np.sin(1 + np.arange(h.size)))
    if seed:
        om.set_select_seed(seed)
    return om


@pytest.mark.parametrize("covs,K,m", CASES)
@pytest.mark.parametrize("hypshift", [0.0, 0.3])
def test_outermod_matches_oracle(product_symbols, oracle, covs, K, m, hypshift):
    a, b = build(product_symbols, covs, m, hypshift), build(oracle, covs, m, hypshift)
    for w in ("knotptst", "hypst", "hypmatch", "gest", "knotptstge", "maxlevel"):
        np.testing.assert_array_equal(a.index(w), b.index(w), err_msg=w)
    for w in ("basisvar", "rotmat", "rotmat_gradhyp", "logbasisvar_gradhyp", "knotpt"):
        np.testing.assert_array_equal(a.real(w), b.real(w), err_msg=w)
    ta, tb = a.selectterms(K), b.selectterms(K)
    np.testing.assert_array_equal(ta, tb)
    assert ta.dtype == np.uint64 and not ta[0].any()  # first row is the constant term
    np.testing.assert_array_equal(a.getvar(ta), b.getvar(tb))
    np.testing.assert_array_equal(a.getlvar_gradhyp(ta), b.getlvar_gradhyp(tb))
    h = a.gethyp()
    assert a.hyplpdf(h) == b.hyplpdf(h)
    np.testing.assert_array_equal(a.hyplpdf_grad(h), b.hyplpdf_grad(h))


def test_selectterms_is_downward_closed_and_seeded_policy_matches(product_symbols, oracle):
    covs = ["mat25pow"] * 6
    for seed in (0, 12345):
        a, b = build(product_symbols, covs, 30, 0.2, seed), build(oracle, covs, 30, 0.2, seed)
        t = a.selectterms(500)
        np.testing.assert_array_equal(t, b.selectterms(500))
        S = set(map(tuple, t.astype(int)))
        assert len(S) == 500
        for row in S:
            for l, v in enumerate(row):
                if v > 0:
                    p = list(row); p[l] -= 1
                    assert tuple(p) in S
        assert np.all(t.max(0) <= a.index("maxlevel"))


def test_hyplpdf_grad_matches_finite_difference(product_symbols):
    om = build(product_symbols, ["mat25pow", "mat25", "mat25ang"], 12, 0.1)
    h = om.gethyp()
    g = om.hyplpdf_grad(h)
    for i in range(h.size):
        e = np.zeros_like(h); e[i] = 1e-6
        fd = (om.hyplpdf(h + e) - om.hyplpdf(h - e)) / 2e-6
        assert abs(fd - g[i]) < 1e-5 * max(1, abs(g[i]))


def test_loopvals_rule():
    """outerbase::setloopvals_ (modandbase.cpp:504-513) is reported, not used, by the GPU path."""
    # exercised through the oracle binding; the product evaluates the same closed form (ob_model.hpp)
    for N, T, cs in [(200, 8, 32), (10000, 8, 257), (1_000_000, 8, 257), (1_000_000, 64, 33)]:
        maxchunk = 1 + 2048 // T
        assert max(32, min(maxchunk, N // (4 * T) + 1)) == cs


@pytest.mark.parametrize("K", [25, 120])
def test_full_hessian_host_algebra_matches_oracle(product_symbols, oracle, K):
    """The K x K side of the full-Hessian branch runs on the host in the product too, so it can be checked without a GPU:
    logpr_gauss::hess / hessgradhyp / hessgradpara (logpr_gauss.cpp:153-186), lpdfvec's full branch of buildhess
    (fit.cpp:269-299: log-determinant and inverse of the total Hessian -- Cholesky in the product, the reference's
    eigen-decomposition in the oracle) and lpdf::optnewton (fit.cpp:98-131, elimination with partial pivoting) on
    lpdfvec(logpr_gauss, logpr_gauss), a model whose every piece lives on the host."""
    covs = ["mat25pow", "mat25", "mat25", "mat25pow"]
    res = {}
    for name, lib in (("p", product_symbols), ("o", oracle)):
        om = build(lib, covs, 16, 0.2)
        terms = om.selectterms(K)
        a, b = lib.logpr_gauss(om, terms), lib.logpr_gauss(om, terms)
        b.updatepara([5.0])
        res[name] = dict(hess=a.hess(), hgh=a.hessgradhyp(), hgp=a.hessgradpara())
        vec = lib.lpdfvec(a, b)
        vec.set_coeff(0.3 * np.cos(np.arange(K)))
        vec.optnewton()
        res[name].update(val=vec.val, coeff=np.array(vec.coeff), gradhyp=np.array(vec.gradhyp), gradpara=np.array(vec.gradpara),
                         tothess=vec.tothess, vhess=vec.hess(), vhgh=vec.hessgradhyp(), vhgp=vec.hessgradpara())
        # the log-density is quadratic in the coefficients: one Newton step from anywhere lands on the optimum, 0
        assert np.abs(vec.coeff).max() < 1e-12 and np.abs(vec.grad).max() < 1e-12
        vec.optcg(0.001, 100)  # optcg switches back to the diagonal branch (fit.cpp:38)
        vec.domarg = True
    p, o = res["p"], res["o"]
    for k in ("hess", "hgh", "hgp", "tothess", "vhess", "vhgh", "vhgp"):
        np.testing.assert_array_equal(p[k], o[k], err_msg=k)
    assert abs(p["val"] - o["val"]) <= 1e-12 * abs(o["val"])
    for k in ("gradhyp", "gradpara"):
        assert np.abs(p[k] - o[k]).max() <= 1e-11 * np.abs(o[k]).max(), k
    assert p["hgh"].shape == (K, K, p["gradhyp"].size) and p["hgp"].shape == (K, K, 1) and p["vhgp"].shape == (K, K, 2)

"""Host unit tests of the terms compiler (outerbase_b200/csrc/ob_terms.hpp): the compiled warp
programs, interpreted on the CPU for one row, must reproduce prod_l B_l[t_kl] term by term."""
import numpy as np
import pytest


def brute(terms, kp, B, a, b, aug=-1, G0=None):
    K, d = terms.shape
    phi = np.ones(K)
    for k in range(K):
        for l in range(d):
            t = int(terms[k, l])
            if l == aug:
                phi[k] *= G0[t]
            elif t > 0:
                phi[k] *= B[int(kp[l]) + t]
    return phi @ a, b * phi


@pytest.mark.parametrize("K", [1, 2, 7, 60, 1000, 2000, 4000])
def test_selectterms_programs(product_symbols, K):
    P = product_symbols
    rng = np.random.default_rng(K)
    d = 10 if K < 4000 else 20
    om = P.outermod()
    om.setcovfs(["mat25pow"] * d)
    om.setknot([np.linspace(0.01, 0.99, 40)] * d)
    h = om.gethyp(); h[0::2] = np.linspace(-0.6, 0.4, d); om.updatehyp(h)
    kp = om.index("knotptst")
    terms = om.selectterms(K)
    B = rng.normal(size=int(kp[-1])); a = rng.normal(size=K)
    pa, pt, st = P.debug_terms_eval(terms, kp, B, a, b=1.7)
    ra, rt = brute(terms, kp, B, a, 1.7)
    assert st["fast_ok"] == 1 and st["nodes"] == K  # downward closed: every prefix is a term
    nnz = (terms > 0).sum(1)
    assert st["W"] == int((nnz + 1).sum()) and st["Lcols"] == int(terms.max(0).sum())
    assert abs(pa - ra) <= 1e-12 * np.abs(a).sum() * max(1.0, np.abs(rt).max())
    assert np.abs(pt - rt).max() <= 1e-13 * np.abs(rt).max()
    assert st["nwords_fwd"] < 1.35 * K + 400 and st["nwords_bwd"] < 1.5 * K + 400  # stack traffic stays marginal
    for aug in (0, d // 2, d - 1):
        G0 = rng.normal(size=41)
        pa, pt, _ = P.debug_terms_eval(terms, kp, B, a, b=0.3, aug_dim=aug, gcols0=G0)
        ra, rt = brute(terms, kp, B, a, 0.3, aug, G0)
        assert abs(pa - ra) <= 1e-12 * np.abs(a).sum() * max(1.0, np.abs(rt).max() / 0.3)
        assert np.abs(pt - rt).max() <= 1e-13 * np.abs(rt).max()


def test_arbitrary_term_tables(product_symbols):
    """Not downward closed, shuffled, any number of warps: prefixes become pass nodes."""
    P = product_symbols
    rng = np.random.default_rng(5)
    for trial in range(300):
        d = int(rng.integers(1, 8)); K = int(rng.integers(1, 90))
        terms = rng.integers(0, 4, size=(K, d)).astype(np.uint64) * (rng.uniform(size=(K, d)) < 0.5)
        terms = np.unique(terms, axis=0); rng.shuffle(terms); K = terms.shape[0]
        kp = np.arange(d + 1) * 5
        B = rng.normal(size=5 * d); a = rng.normal(size=K)
        G = int(rng.choice([1, 2, 5, 16]))
        pa, pt, st = P.debug_terms_eval(terms, kp, B, a, b=2.0, ngroups=G)
        ra, rt = brute(terms, kp, B, a, 2.0)
        assert st["fast_ok"] == 1
        assert abs(pa - ra) <= 1e-12 * max(1.0, np.abs(a).sum() * np.abs(rt).max())
        assert np.abs(pt - rt).max() <= 1e-12 * max(1e-300, np.abs(rt).max())


def test_unsupported_tables_are_flagged(product_symbols):
    P = product_symbols
    kp = np.arange(11) * 3
    deep = np.zeros((1, 10), dtype=np.uint64); deep[0, :9] = 1  # 9 factors > 8 stack slots
    _, _, st = P.debug_terms_eval(deep, kp, np.ones(30), np.ones(1))
    assert st["fast_ok"] == 0
    dup = np.zeros((2, 10), dtype=np.uint64); dup[:, 1] = 2
    _, _, st = P.debug_terms_eval(dup, kp, np.ones(30), np.ones(2))
    assert st["fast_ok"] == 0

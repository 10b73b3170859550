"""Regenerates the golden fixtures from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference has no vectors of its own;
these freeze the oracle's outputs so that later edits of oracle/ cannot drift unnoticed and
the GPU tests have a second, file-based anchor."""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
from conftest import make_problem  # noqa: E402
from outerbase_b200.binding import Library  # noqa: E402


def main():
    O = Library(REPO / "oracle" / "_build" / "libob_oracle.so", "orc_")
    cases = {
        # BASELINE config C1: borehole d=8, N=1000, ~60 terms, obfit defaults (mat25pow, 40 quantile knots)
        "c1_borehole_d8_n1000_k60": dict(N=1000, K=60, covs=["mat25pow"] * 8, quantile_knots=True),
        # reference test shape (test-obomgrad.R short-skinny)
        "grad_d8_n200_k100": dict(N=200, K=100, covs=["mat25pow"] + ["mat25"] * 7, quantile_knots=False),
    }
    for name, c in cases.items():
        rng0 = np.random.default_rng(42)
        x0 = rng0.uniform(size=(c["N"], 8))
        if c["quantile_knots"]:
            q = np.linspace(0, 1, 40) * 40 / 41 + 0.5 / 41
            knots = [np.quantile(x0[:, l], q) for l in range(8)]
        else:
            knots = [np.arange(0.001, 0.999, 0.025)] * 8
        om, x, y, terms, rng = make_problem(O, c["N"], c["K"], covs=c["covs"], knots=knots)
        hyp = om.gethyp()
        hyp = hyp + 0.1 * np.sin(np.arange(hyp.size))
        om.updatehyp(hyp)
        terms = om.selectterms(c["K"])
        ob = O.outerbase(om, x)
        a = np.sqrt(om.getvar(terms) / 20) * rng.normal(size=c["K"])
        r = rng.normal(size=c["N"])
        logpr, loglik = O.logpr_gauss(om, terms), O.loglik_gauss(om, terms, y, x)
        vec = O.lpdfvec(logpr, loglik)
        vec.optcg(0.001, 100)
        np.savez_compressed(Path(__file__).parent / f"{name}.npz", covs=np.array(c["covs"]), knots=np.array(knots), hyp=hyp,
                            K=c["K"], x=x, y=y, terms=terms, a=a, r=r, matmul=ob.matmul(terms, a), tmatmul=ob.tmatmul(terms, r),
                            basisvar=om.real("basisvar"), coeff=vec.coeff, val=vec.val, cg_iters=vec.cg_iters,
                            gradhyp=vec.gradhyp, gradpara=vec.gradpara)
        print(name, "terms", terms.shape, "cg", vec.cg_iters, "val", vec.val)


if __name__ == "__main__":
    main()

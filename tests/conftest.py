import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

ORACLE_LIB = REPO / "oracle" / "_build" / "libob_oracle.so"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _build_oracle():
    src = [REPO / "oracle" / n for n in ("ob_oracle.cpp", "ob_oracle.hpp", "orc_capi.cpp")] + [REPO / "include" / "outerbase_b200.h"]
    if not ORACLE_LIB.exists() or any(p.stat().st_mtime > ORACLE_LIB.stat().st_mtime for p in src):
        subprocess.run(["make", "-C", str(REPO / "oracle")], check=True, capture_output=True)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle behind the same binding class as the product (test infrastructure)."""
    _build_oracle()
    from outerbase_b200.binding import Library
    return Library(ORACLE_LIB, "orc_")


@pytest.fixture(scope="session")
def product_symbols():
    """Product library loaded WITHOUT a GPU context (symbol / host-model checks on CPU)."""
    import outerbase_b200 as ob
    ob.build()
    return ob.load_symbols_only()


@pytest.fixture(scope="session")
def gpu():
    """Product binding with a live context on cuda:0; fails loudly when the library is missing."""
    import outerbase_b200 as ob
    if not ob.LIBPATH.exists():
        ob.build()
    return ob.lib()


def borehole8d(x):
    """obtest_borehole8d, R/testfuncs.R:32-46."""
    rw = x[:, 0] * (0.15 - 0.05) + 0.05
    r = x[:, 1] * (50000 - 100) + 100
    Tu = x[:, 2] * (115600 - 63070) + 63070
    Hu = x[:, 3] * (1110 - 990) + 990
    Tl = x[:, 4] * (116 - 63.1) + 63.1
    Hl = x[:, 5] * (820 - 700) + 700
    L = x[:, 6] * (1680 - 1120) + 1120
    Kw = x[:, 7] * (12045 - 9855) + 9855
    m1 = 2 * np.pi * Tu * (Hu - Hl)
    m2 = np.log(r / rw)
    m3 = 1 + 2 * L * Tu / (m2 * rw ** 2 * Kw) + Tu / Tl
    return m1 / m2 / m3 - 77


def make_problem(lib, N, K, d=8, seed=42, covs=None, knots=None, hyp=None):
    """The reference tests' setup (tests/testthat/test-obombasic.R:24-43): uniform x, borehole y
    standardised, knots seq(.001,.999,.025), terms = selectterms(K)."""
    rng = np.random.default_rng(seed)
    x = np.asfortranarray(rng.uniform(size=(N, d)))
    xb = x if d >= 8 else np.hstack([x, np.full((N, 8 - d), 0.5)])
    y = borehole8d(xb[:, :8])
    y = (y - y.mean()) / y.std(ddof=1)
    om = lib.outermod()
    om.setcovfs(covs or (["mat25pow"] + ["mat25"] * (d - 1)))
    om.setknot(knots or [np.arange(0.001, 0.999, 0.025)] * d)
    if hyp is not None:
        om.updatehyp(hyp)
    terms = om.selectterms(K)
    return om, x, y, terms, rng


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


import contextlib


@contextlib.contextmanager
def one_cpu():
    """The reference sizes its OpenMP teams with omp_get_num_procs() wherever a class builds its own outerbase
    (modandbase.cpp:464: loglik_std's constructor, every predictor's update), and its short-basis branch races on basescale
    when several threads run (modandbase.cpp:600-607, SURVEY A9-v: seen as a 0.5 % different 77-row basis once in a few
    runs).  omp_get_num_procs() follows the CPU affinity of the process, so pinning the process to one CPU for the
    duration makes those builds single-threaded and deterministic -- without touching the reference."""
    import os
    if not hasattr(os, "sched_setaffinity"):
        yield
        return
    full = os.sched_getaffinity(0)
    os.sched_setaffinity(0, {min(full)})
    try:
        yield
    finally:
        os.sched_setaffinity(0, full)


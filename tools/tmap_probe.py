"""One small Phi a on the specialised kernel with tensor-map staging (option tmap) against the interpreter kernel --
run under `timeout` before anything longer: a bad tensor map shows as a copy that never completes."""
import os
import sys
from pathlib import Path
os.environ.setdefault("OB_SPEC_OPTS", "1,2,4,80,8,1,4,16,56,4,0,1,8,16,2,3,10,96,232,40,128,1")  # tmap = 1
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import outerbase_b200 as ob
from conftest import make_problem, relerr

lib = ob.lib()
for N, K in ((1000, 100), (300001, 300)):
    om, x, y, terms, rng = make_problem(lib, N, K)
    base = lib.outerbase(om, x, dograd=False)
    a = rng.normal(size=K)
    lib.set_option("spec", 0)
    y0 = base.matmul(terms, a)
    lib.set_option("spec", 1)
    for tm in (0, 1):
        lib.set_option("tmap", tm)
        y1 = base.matmul(terms, a)
        print(N, K, "tmap", tm, "relerr", relerr(y1, y0), flush=True)
        assert relerr(y1, y0) < 1e-12
print("PROBE_OK")

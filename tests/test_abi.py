"""The C-ABI shared library loads on a CPU-only host and exports every symbol the header declares;
without a GPU the product refuses to compute (no CPU fallback, no route through the oracle)."""
import ctypes as C
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]


def test_exports_every_declared_symbol(product_symbols):
    import outerbase_b200 as ob
    syms = ob.header_symbols(ob.HEADER)
    assert len(syms) >= 80
    missing = [s for s in syms if not product_symbols.has(s)]
    assert not missing, missing


def test_oracle_exports_the_same_abi(oracle):
    import outerbase_b200 as ob
    missing = [s for s in ob.header_symbols(ob.HEADER, "orc_") if not oracle.has(s) and "debug" not in s]
    assert not missing, missing


def test_no_cpu_fallback(product_symbols):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        product_symbols.call("ctx_create", C.c_int(0), C.byref(h))


def test_product_never_references_the_oracle():
    pkg = REPO / "outerbase_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.hpp")) + list(pkg.rglob("*.cuh")):
        txt = p.read_text()
        code = re.sub(r"/\*.*?\*/|#.*?$|\"\"\".*?\"\"\"", "", txt, flags=re.S | re.M)
        assert "ob_oracle" not in code and "orc_" not in code.replace("prefix `orc_`", ""), p


def test_error_mapping_on_host_calls(product_symbols):
    L = product_symbols
    om = L.outermod()
    with pytest.raises(ValueError):
        om.setcovfs(["nonsense"])
    om.setcovfs(["mat25", "mat25pow"])
    with pytest.raises(ValueError, match="between"):
        om.setknot([[0.1, 0.5, 1.5], [0.1, 0.2, 0.3]])  # knot outside [0,1], interfaceR.cpp:107-117
    om2 = L.outermod()
    with pytest.raises(ValueError, match="cov. funcs"):
        om2.setknot([[0.1, 0.2]])  # interfaceR.cpp:95-98
    om.setknot([[0.1, 0.4, 0.8], [0.1, 0.2, 0.3, 0.9]])
    with pytest.raises(ValueError):
        om.updatehyp([0.0])  # wrong size
    assert om.hyplpdf([5.0, 0.0, 0.0]) == float("-inf")  # out of bounds -> -inf in band (covfuncs.cpp:41)

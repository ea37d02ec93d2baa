"""The C-ABI library loads, exports every symbol include/enlsip_b200.h declares, and refuses to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import __graft_entry__ as ge

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    import enlsip_jl_b200 as E
    E.capi.build()
    return E.capi


def test_header_symbols_exported(capi):
    hdr = open(os.path.join(ROOT, "include", "enlsip_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(enlsipb200_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(capi.EXPORTS)
    raw = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert capi.lib().enlsipb200_version() >= 100


def test_default_options_match_solver_jl(capi):
    o = capi.default_options()      # solver.jl:62-63
    assert o.max_iter == 100 and o.scaling == 0 and o.time_limit == 1e3
    assert all(np.isnan(v) for v in (o.abs_tol, o.rel_tol, o.c_tol, o.x_tol))


def test_no_cpu_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import enlsip_jl_b200 as E
    with pytest.raises(capi.EngineError, match="no CPU fallback"):
        E.CnlsModel("hs65", np.zeros((1, 3)), x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
    x = np.zeros(4)
    rc = capi.lib().enlsipb200_det_exp(x.ctypes.data, x.ctypes.data, 4, 0)
    assert rc == -2


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py may touch oracle/: the package never imports, includes or loads it."""
    pkg = os.path.join(ROOT, "enlsip.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            path = os.path.join(dirpath, fn)
            if fn.endswith(".py"):
                for line in open(path):
                    code = line.split("#")[0]
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", code), (fn, line)
                    assert not re.search(r"[\"']oracle[\"'/]", code), (fn, line)
            elif fn.endswith((".h", ".cu", ".cuh", ".cpp")):
                for line in open(path):
                    assert not re.search(r"#\s*include.*oracle", line), (fn, line)


def test_model_api_validation():
    import enlsip_jl_b200 as E
    with pytest.raises(AssertionError):
        E.CnlsModel("no_such_family", np.zeros((1, 3)))
    assert E.dict_status_codes[1] == "found_first_order_stationary_point" and E.dict_status_codes[-11] == "time_limit_exceeded"

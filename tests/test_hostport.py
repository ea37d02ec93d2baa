"""The solver core compiled for the host (oracle/hostport, G = 1) against the oracle fixtures.

This checks the control flow of the exact source the CUDA kernel is built from, in a container without a
GPU.  The GPU parity tests (test_gpu_parity.py) repeat the same comparisons through the C ABI on the device.
"""
import ctypes
import math

import numpy as np
import pytest

import __graft_entry__ as ge
from tests import parity


class Opt(ctypes.Structure):
    _fields_ = [("max_iter", ctypes.c_int), ("scaling", ctypes.c_int), ("jac_mode", ctypes.c_int),
                ("second_derivatives", ctypes.c_int), ("time_limit", ctypes.c_double), ("eps_abs", ctypes.c_double),
                ("eps_rel", ctypes.c_double), ("eps_x", ctypes.c_double), ("eps_c", ctypes.c_double),
                ("eps_rank", ctypes.c_double)]


def run(family, x0, d0, d1, xl, xu, jac_mode, trace_cap=40, nthreads=2, time_limit=1e3, max_iter=100):
    lib = ctypes.CDLL(ge.build_hostport())
    se = math.sqrt(np.finfo(float).eps)
    opt = Opt(max_iter, 0, jac_mode, 1, time_limit, 1e-10, se, se, se, se)
    B, n = x0.shape
    lmax = 1 + 2 * n
    out = dict(x=np.zeros((B, n)), f=np.zeros(B), exit_code=np.zeros(B, np.int32), status=np.zeros(B, np.int32),
               iters=np.zeros(B, np.int32), nact=np.zeros(B, np.int32), active=np.zeros((B, lmax), np.int32),
               counters=np.zeros((B, 2), np.int32), trace=np.zeros((B, trace_cap, 16 + n)))
    vp = ctypes.c_void_p
    lib.hostport_solve.argtypes = [ctypes.c_int, ctypes.c_longlong] + [vp] * 5 + [ctypes.POINTER(Opt)] + [vp] * 9 + [ctypes.c_int, ctypes.c_int]
    p = lambda a: None if a is None else np.ascontiguousarray(a).ctypes.data_as(vp)
    x0 = np.ascontiguousarray(x0); d0c = None if d0 is None else np.ascontiguousarray(d0); d1c = None if d1 is None else np.ascontiguousarray(d1)
    xl, xu = np.ascontiguousarray(xl, dtype=float), np.ascontiguousarray(xu, dtype=float)
    lib.hostport_solve(family, B, p(x0), p(d0c), p(d1c), p(xl), p(xu), ctypes.byref(opt), p(out["x"]), p(out["f"]),
                       p(out["exit_code"]), p(out["status"]), p(out["iters"]), p(out["nact"]), p(out["active"]),
                       p(out["counters"]), p(out["trace"]), trace_cap, nthreads)
    return out


def test_hs65_analytic_vs_golden(golden_dir):
    import enlsip_jl_b200 as E
    gold = np.load(golden_dir + "/c2_hs65.npz")
    B = gold["x"].shape[0]
    out = run(0, E.synth.gen_hs65_batch(B), None, None, E.synth.HS65_LOW, E.synth.HS65_UPP, 0)
    st = parity.compare(gold, out, "analytic", 3)
    assert (out["exit_code"] == -98).sum() == (gold["exit_code"] == -98).sum() > 0    # the reference's endless swap loop
    ok = (out["exit_code"] == gold["exit_code"]) & (out["iters"] == gold["iters"])
    assert np.array_equal(out["counters"][ok][:, 1], gold["njac"][ok])


@pytest.mark.parametrize("mode,fixture,jac", [("analytic", "c3_gp_analytic.npz", 0), ("fd", "c3_gp_fd.npz", 1)])
def test_gauss_peaks_vs_golden(golden_dir, mode, fixture, jac):
    import enlsip_jl_b200 as E
    gold = np.load(golden_dir + "/" + fixture)
    B = gold["x"].shape[0]
    y, S, x0, _ = E.synth.gen_gauss_peaks_batch(B)
    out = run(1, x0, y, S, E.synth.GP_LOW, E.synth.GP_UPP, jac)
    parity.compare(gold, out, mode, 6)


def test_time_limit_and_max_iter():
    import enlsip_jl_b200 as E
    x0 = E.synth.gen_hs65_batch(4)
    out = run(0, x0, None, None, E.synth.HS65_LOW, E.synth.HS65_UPP, 0, time_limit=-1.0)
    assert np.all(out["exit_code"] == -11) and np.all(out["status"] == -11) and np.array_equal(out["x"], x0)   # T5
    assert np.all(out["iters"] == 1)
    out = run(0, x0, None, None, E.synth.HS65_LOW, E.synth.HS65_UPP, 0, max_iter=3)
    assert np.all(out["exit_code"] == -2) and np.all(out["iters"] == 3)   # same as the oracle

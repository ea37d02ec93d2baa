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


def run(family, x0, d0, d1, xl, xu, jac_mode, trace_cap=40, nthreads=2, time_limit=1e3, max_iter=100, lapack=False):
    lib = ctypes.CDLL(ge.build_hostport())
    lib.hostport_use_lapack.argtypes = [ctypes.c_char_p]
    assert lib.hostport_use_lapack(ge.openblas_path().encode() if lapack else None) == 0
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


@pytest.mark.parametrize("mode,fixture,jac", [("analytic", "c3_gp_analytic.npz", 0), ("fd", "c3_gp_fd.npz", 1)])
def test_reference_arm_with_openblas_dgeqp3_vs_golden(golden_dir, mode, fixture, jac):
    """The CPU reference arm of bench.py: every `qr(., ColumnNorm())` of the solve goes to OpenBLAS' dgeqp3 (the routine
    Julia calls) instead of the engine's restatement; same parity bars against the oracle fixtures."""
    import enlsip_jl_b200 as E
    gold = np.load(golden_dir + "/" + fixture)
    B = gold["x"].shape[0]
    y, S, x0, _ = E.synth.gen_gauss_peaks_batch(B)
    out = run(1, x0, y, S, E.synth.GP_LOW, E.synth.GP_UPP, jac, lapack=True)
    parity.compare(gold, out, mode, 6)
    run(1, x0[:1], y[:1], S[:1], E.synth.GP_LOW, E.synth.GP_UPP, jac, lapack=False)      # unbind for the tests that follow


def test_time_limit_and_max_iter():
    import enlsip_jl_b200 as E
    x0 = E.synth.gen_hs65_batch(4)
    out = run(0, x0, None, None, E.synth.HS65_LOW, E.synth.HS65_UPP, 0, time_limit=-1.0)
    assert np.all(out["exit_code"] == -11) and np.all(out["status"] == -11) and np.array_equal(out["x"], x0)   # T5
    assert np.all(out["iters"] == 1)
    out = run(0, x0, None, None, E.synth.HS65_LOW, E.synth.HS65_UPP, 0, max_iter=3)
    assert np.all(out["exit_code"] == -2) and np.all(out["iters"] == 3)   # same as the oracle


# ---- the reference's own test problems (test/problems/*.jl) ------------------------------------
def reference_suite_cases():
    """(family id, model family name, oracle problem factory, x0 batch, data, bounds, solve kwargs)."""
    import json, os
    from oracle import problems as P
    rng = np.random.default_rng(7)
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    d = json.load(open(os.path.join(gdir, "osborne2.json")))
    lo, up, x0 = np.array(d["x_low"]), np.array(d["x_upp"]), np.array(d["x0"])
    xs = [x0] + [np.clip(x0 * (1 + 0.02 * rng.uniform(-1, 1, 11)), lo + 1e-3 * (up - lo), up - 1e-3 * (up - lo)) for _ in range(5)]
    cases = [(2, "osborne2", lambda x: _with_x0(P.osborne2(), x), np.array(xs), {"t": np.array(d["t"]), "y": np.array(d["y"])}, (lo, up), {})]
    pr = P.chained_rosenbrock(10)
    xs = [pr.x0] + [pr.x0 * (1 + 0.05 * rng.uniform(-1, 1, 10)) for _ in range(5)]
    cases.append((3, "chained_rosenbrock10", lambda x: _with_x0(P.chained_rosenbrock(10), x), np.array(xs), {}, (None, None), {}))
    pw = P.chained_wood(20)
    xs = [pw.x0] + [pw.x0 * (1 + 0.05 * rng.uniform(-1, 1, 20)) for _ in range(5)]
    cases.append((4, "chained_wood20", lambda x: _with_x0(P.chained_wood(20), x), np.array(xs), {}, (None, None),
                  dict(rel_tol=1e-5, x_tol=1e-3, c_tol=1e-6)))
    return cases


def _with_x0(prob, x):
    prob.x0 = np.array(x, dtype=float)
    return prob


def check_against_oracle(out, b, o, n, label):
    """Discrete outputs identical; objective 1e-10; iterate before the last step 1e-10; final x up to the last move."""
    act = [int(v) for v in out["active"][b] if v > 0]
    assert (int(out["exit_code"][b]), int(out["iters"][b]), act) == (o.exit_code, o.iterations, o.active), (label, b, o.threw)
    for k, tr in enumerate(o.trace[:out["trace"].shape[1]]):
        row = out["trace"][b, k]
        assert (tr.t, tr.rankA, tr.rankJ2, tr.dimA, tr.dimJ2, tr.code, tr.index_del, tr.exit_code) == \
            tuple(int(v) for v in (row[1], row[2], row[3], row[4], row[5], row[6], row[9], row[10])), (label, b, k)
    assert abs(out["f"][b] - o.f) <= 1e-10 * max(abs(o.f), 1e-300), (label, b)
    if 2 <= len(o.trace) <= out["trace"].shape[1]:
        xp = out["trace"][b, len(o.trace) - 2, 16:16 + n]
        assert np.linalg.norm(xp - o.trace[-2].x_new) <= 1e-10 * np.linalg.norm(xp), (label, b)
    assert np.linalg.norm(out["x"][b] - o.x) <= 1e-10 * np.linalg.norm(o.x) + 3.2 * o.trace[-1].p_norm, (label, b)


def test_reference_suite_families_vs_oracle():
    from oracle import enlsip_oracle as O
    for fam, name, mk, xs, data, (lo, up), kw in reference_suite_cases():
        n = xs.shape[1]
        lo_ = np.full(n, -np.inf) if lo is None else lo
        up_ = np.full(n, np.inf) if up is None else up
        vals = list(data.values())
        d0, d1 = (vals + [None, None])[:2]
        se = math.sqrt(np.finfo(float).eps)
        lib = ctypes.CDLL(ge.build_hostport())
        opt = Opt(100, 0, 0, 1, 1e3, 1e-10, kw.get("rel_tol", se), kw.get("x_tol", se), kw.get("c_tol", se), se)
        B = xs.shape[0]
        lmax = {2: 22, 3: 8, 4: 13}[fam]
        out = dict(x=np.zeros((B, n)), f=np.zeros(B), exit_code=np.zeros(B, np.int32), status=np.zeros(B, np.int32),
                   iters=np.zeros(B, np.int32), nact=np.zeros(B, np.int32), active=np.zeros((B, lmax), np.int32),
                   counters=np.zeros((B, 2), np.int32), trace=np.zeros((B, 60, 16 + n)))
        vp = ctypes.c_void_p
        lib.hostport_solve.argtypes = [ctypes.c_int, ctypes.c_longlong] + [vp] * 5 + [ctypes.POINTER(Opt)] + [vp] * 9 + [ctypes.c_int, ctypes.c_int]
        p = lambda a: None if a is None else np.ascontiguousarray(a).ctypes.data_as(vp)
        xs_c = np.ascontiguousarray(xs)
        keep = [np.ascontiguousarray(a) for a in (d0, d1, lo_, up_) if a is not None]
        lib.hostport_solve(fam, B, p(xs_c), p(d0), p(d1), p(lo_), p(up_), ctypes.byref(opt), p(out["x"]), p(out["f"]),
                           p(out["exit_code"]), p(out["status"]), p(out["iters"]), p(out["nact"]), p(out["active"]),
                           p(out["counters"]), p(out["trace"]), 60, 1)
        for b in range(B):
            o = O.solve(mk(xs[b]), wallclock=False, **kw)
            check_against_oracle(out, b, o, n, name)
        if name == "chained_wood20":
            assert np.any(out["trace"][:, :, 6] == 2)        # the Newton path runs (the reference's purpose for this test)


# ---- 10^4-problem samples (tests/golden/*_10k.npz, made by make_golden_10k.py): the full mismatch histogram ----
@pytest.mark.parametrize("fixture,family,jac,n", [("c2_hs65_10k.npz", 0, 0, 3), ("c3_gp_analytic_10k.npz", 1, 0, 6),
                                                   ("c3_gp_fd_10k.npz", 1, 1, 6)])
def test_10k_sample_histogram(golden_dir, fixture, family, jac, n):
    import enlsip_jl_b200 as E
    gold = np.load(golden_dir + "/" + fixture)
    B, start = gold["x"].shape[0], int(gold["start"])
    if family == 0:
        out = run(0, E.synth.gen_hs65_batch(B, start=start), None, None, E.synth.HS65_LOW, E.synth.HS65_UPP, 0, nthreads=8)
    else:
        y, S, x0, _ = E.synth.gen_gauss_peaks_batch(B, start=start)
        out = run(1, x0, y, S, E.synth.GP_LOW, E.synth.GP_UPP, jac, nthreads=8)
    h = parity.histogram(gold, out, n)
    print("\n10k histogram (host port)", fixture, h)
    check_10k_bars(h, jac)


def check_10k_bars(h, jac):
    B = h["problems"]
    if jac == 0:      # analytic Jacobians: everything identical except knife-edge ties
        assert h["status_mismatch"] <= 0.002 * B, h
        assert h["all_discrete_outputs_identical"] >= 0.97 * B, h
        assert h["trace_identical"] >= 0.97 * B, h
        assert h["f_rel_error_quantiles"]["1.0"] <= 1e-9 and h["f_rel_error_quantiles"]["0.999"] <= 1e-10, h
        assert h["x_beyond_last_step_bound"] <= 0.001 * B, h
    else:             # forward differences: the oracle's own 1-ulp sensitivity (test_oracle.py::test_fd_noise_floor)
        assert h["status_mismatch"] <= 0.05 * B, h
        assert h["iteration_count_mismatch"] <= 0.10 * B, h
        assert h["f_rel_error_quantiles"]["0.99"] <= 1e-9, h
        assert h["x_rel_error_quantiles"]["0.99"] <= 1e-8, h

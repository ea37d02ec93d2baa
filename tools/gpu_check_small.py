"""GPU development check of the device small stage (csrc/enl_small.cuh): timings of the dense kernels at the sizes of
BASELINE.json configs 4 / 5, then whole solves with the per-phase statistics."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import enlsip_jl_b200 as E                                   # noqa: E402


def time_qrcp(rows, cols, reps=2):
    L = E.capi.lib()
    rng = np.random.default_rng(1)
    A = np.asfortranarray(rng.standard_normal((rows, cols)))
    vp = ctypes.c_void_p
    best = 1e9
    for _ in range(reps):
        f = A.copy(order="F"); tau = np.zeros(min(rows, cols)); jp = np.zeros(cols, np.int32)
        t0 = time.perf_counter()
        rc = L.enlsipb200_dense_qrcp(rows, cols, f.ctypes.data_as(vp), tau.ctypes.data_as(vp), jp.ctypes.data_as(vp), -1)
        best = min(best, float(L.enlsipb200_dense_last_ms()) * 1e-3)
        assert rc == 0
    return best


def time_mulq(mr, nq, k, reps=2):
    from scipy.linalg import lapack
    L = E.capi.lib()
    rng = np.random.default_rng(2)
    qr, tau, _, _ = lapack.dgeqrf(np.asfortranarray(rng.standard_normal((nq, k))))
    M = np.asfortranarray(rng.standard_normal((mr, nq)))
    vp = ctypes.c_void_p
    best = 1e9
    for _ in range(reps):
        out = M.copy(order="F")
        t0 = time.perf_counter()
        rc = L.enlsipb200_dense_mulq(mr, nq, k, qr.ctypes.data_as(vp), tau.ctypes.data_as(vp), out.ctypes.data_as(vp), -1)
        best = min(best, float(L.enlsipb200_dense_last_ms()) * 1e-3)
        assert rc == 0
    return best


def solve_stats(m, n, nb, seed, ineq, bounds, reps=2):
    d = E.synth.gen_single_index(m, n, nb, seed=seed, ineq=ineq)
    lo = None if bounds is None else np.full(n, bounds[0])
    up = None if bounds is None else np.full(n, bounds[1])
    mod = E.LargeCnlsModel("single_index", d["x0"], d, ineq=ineq, x_low=lo, x_upp=up)
    out = []
    for _ in range(reps):
        s0 = mod.stats()
        t0 = time.perf_counter()
        E.solve(mod)
        dt = time.perf_counter() - t0
        s1 = mod.stats()
        ds = {k: s1[k] - s0[k] for k in s1}
        ds["wall_ms"] = dt * 1e3
        ds["iterations"] = int(mod.iterations[0]); ds["exit_code"] = int(mod.exit_code[0]); ds["f"] = float(mod.obj_value[0])
        out.append(ds)
    mod.close()
    return out


if __name__ == "__main__":
    res = {}
    # device time (CUDA events around the kernels of the hook; transfers excluded)
    for rows, cols in ((256, 64), (64, 64), (257, 192), (1025, 1000), (4096, 511), (4097, 3587)):
        res["qrcp_%dx%d_s" % (rows, cols)] = time_qrcp(rows, cols)
        print("qrcp %d x %d: %.2f ms (device)" % (rows, cols, res["qrcp_%dx%d_s" % (rows, cols)] * 1e3), flush=True)
    for mr, nq, k in ((257, 256, 64), (4097, 4096, 511)):
        res["mulq_%d_%d_%d_s" % (mr, nq, k)] = time_mulq(mr, nq, k)
        print("mulq %d x %d, k=%d: %.2f ms" % (mr, nq, k, res["mulq_%d_%d_%d_s" % (mr, nq, k)] * 1e3), flush=True)
    res["c4_1M"] = solve_stats(1 << 20, 256, 64, 4, False, None)
    print("C4 (1M rows)", json.dumps(res["c4_1M"][-1]), flush=True)
    if "--full" in sys.argv:
        res["c5_full"] = solve_stats(16384, 4096, 1024, 5, True, (-2.0, 2.0), reps=1)
        print("C5 full", json.dumps(res["c5_full"][-1]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "small_check.json"), "w"), indent=1)

"""One call of a dense hook / one solve, for kernel-level timing under ncu:  python tools/small_one.py qrcp ROWS COLS | mulq MR NQ K | c4 ROWS"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import enlsip_jl_b200 as E                                   # noqa: E402

what = sys.argv[1]
vp = ctypes.c_void_p
L = E.capi.lib()
if what == "qrcp":
    rows, cols = int(sys.argv[2]), int(sys.argv[3])
    A = np.asfortranarray(np.random.default_rng(1).standard_normal((rows, cols)))
    tau = np.zeros(min(rows, cols)); jp = np.zeros(cols, np.int32)
    assert L.enlsipb200_dense_qrcp(rows, cols, A.ctypes.data_as(vp), tau.ctypes.data_as(vp), jp.ctypes.data_as(vp), -1) == 0
elif what == "mulq":
    from scipy.linalg import lapack
    mr, nq, k = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rng = np.random.default_rng(2)
    qr, tau, _, _ = lapack.dgeqrf(np.asfortranarray(rng.standard_normal((nq, k))))
    M = np.asfortranarray(rng.standard_normal((mr, nq)))
    assert L.enlsipb200_dense_mulq(mr, nq, k, qr.ctypes.data_as(vp), tau.ctypes.data_as(vp), M.ctypes.data_as(vp), -1) == 0
elif what == "c4":
    m = int(sys.argv[2])
    d = E.synth.gen_single_index(m, 256, 64, seed=4)
    mod = E.LargeCnlsModel("single_index", d["x0"], d)
    E.solve(mod)
    print(mod.stats())
elif what == "c5":
    n = int(sys.argv[2]); m = 4 * n; nb = n // 4
    d = E.synth.gen_single_index(m, n, nb, seed=5, ineq=True)
    mod = E.LargeCnlsModel("single_index", d["x0"], d, ineq=True, x_low=np.full(n, -2.0), x_upp=np.full(n, 2.0))
    E.solve(mod, max_iter=int(sys.argv[3]) if len(sys.argv) > 3 else 100)
    print(mod.stats(), mod.exit_code, mod.iterations)

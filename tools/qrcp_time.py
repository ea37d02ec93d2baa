"""Time the device QRCP (enlsipb200_dense_qrcp) at a given shape and check it against LAPACK dgeqp3."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import enlsip_jl_b200 as E                                   # noqa: E402
from scipy.linalg import lapack                               # noqa: E402

rows, cols = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L = E.capi.lib()
rng = np.random.default_rng(1)
A = np.asfortranarray(rng.standard_normal((rows, cols)))
vp = ctypes.c_void_p
ts = []
for _ in range(reps):
    f = A.copy(order="F"); tau = np.zeros(min(rows, cols)); jp = np.zeros(cols, np.int32)
    rc = L.enlsipb200_dense_qrcp(rows, cols, f.ctypes.data_as(vp), tau.ctypes.data_as(vp), jp.ctypes.data_as(vp), -1)
    assert rc == 0, rc
    ts.append(float(L.enlsipb200_dense_last_ms()))
qr, jl, tl, _, info = lapack.dgeqp3(A.copy(order="F"))
k = min(rows, cols)
same = int(np.sum(jp == jl - 1))
dR = np.abs(np.triu(f[:k]) - np.triu(qr[:k])).max() / np.abs(np.triu(qr[:k])).max()
print("qrcp %d x %d: ms %s  pivots identical %d / %d  max |R - R_lapack| / max|R| = %.2e  max |tau diff| = %.2e"
      % (rows, cols, ["%.2f" % t for t in ts], same, cols, dR, np.abs(tau - tl).max()))

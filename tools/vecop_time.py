"""Device time of the cooperative vector kernels (enlsipb200_dense_vecop) at the sizes of BASELINE.json config 5."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import enlsip_jl_b200 as E                                   # noqa: E402
from scipy.linalg import lapack                               # noqa: E402

L = E.capi.lib()
vp = ctypes.c_void_p
rng = np.random.default_rng(0)
for frows, k in ((4096, 511), (4097, 3587), (511, 511), (257, 192)):
    qr, tau, _, _ = lapack.dgeqrf(np.asfortranarray(rng.standard_normal((frows, k))))
    qr = np.asfortranarray(qr)
    for kind, name in ((0, "Q'v"), (1, "Qv"), (2, "R\\v"), (3, "R'\\v")):
        v = rng.standard_normal(frows if kind < 2 else k)
        ts = []
        for _ in range(3):
            out = v.copy()
            rc = L.enlsipb200_dense_vecop(kind, frows, k, qr.ctypes.data_as(vp), tau.ctypes.data_as(vp), out.ctypes.data_as(vp), -1)
            assert rc == 0, L.enlsipb200_large_last_error()
            ts.append(float(L.enlsipb200_dense_last_ms()))
        print("%5d x %4d  %-5s %s ms" % (frows, k, name, " ".join("%.3f" % t for t in ts)), flush=True)

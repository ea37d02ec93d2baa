"""Problem definitions for the ORACLE (test infrastructure only; see enlsip_oracle.py header).

* the reference's own fixtures: HS65 (test/problems/HS65.jl:7-20, README.md:89-116), Osborne 2
  (test/problems/osborne2.jl:10-102, data in tests/golden/osborne2.json), chained Rosenbrock
  (test/problems/chained_rosenbrock.jl:8-53), chained Wood (test/problems/chained_wood.jl:4-35);
* the synthetic families of BASELINE.json / SURVEY.md section 8d: Gaussian peaks (C3) and the
  single-index tanh family (C4/C5), both defined through the deterministic ``det_exp``
  (oracle/detmath.c) so that the CUDA engine can reproduce the residuals bit for bit.

Arithmetic in the residual/constraint functions is written as separately rounded IEEE operations
in a fixed order (numpy never fuses a*b+c), which is what Julia executes for the same expressions
and what the CUDA functors restate with ``__dmul_rn``/``__dadd_rn``.
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess

import numpy as np

from .enlsip_oracle import Problem, make_problem, jac_forward_diff

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_detmath(force=False):
    """Compile oracle/detmath.c -> oracle/_build/libdetmath.so (gcc; no GPU involved)."""
    out = os.path.join(_HERE, "_build", "libdetmath.so")
    src = os.path.join(_HERE, "detmath.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", src, "-o", out, "-lm"])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_detmath())
        for name in ("det_exp_vec", "det_tanh_vec"):
            getattr(_LIB, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
            getattr(_LIB, name).restype = None
    return _LIB


def det_exp(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    _lib().det_exp_vec(x.ctypes.data, y.ctypes.data, x.size)
    return y


def det_tanh(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    _lib().det_tanh_vec(x.ctypes.data, y.ctypes.data, x.size)
    return y


# ------------------------------------------------------------------------------------------
# reference fixtures
# ------------------------------------------------------------------------------------------
def hs65(x0=(-5.0, 5.0, 0.0), fd=False) -> Problem:
    """test/problems/HS65.jl:7-20."""
    def r(x):
        return np.array([x[0] - x[1], (x[0] + x[1] - 10.0) / 3.0, x[2] - 5.0])

    def jac_r(x):
        return np.array([[1.0, -1.0, 0.0], [1.0 / 3.0, 1.0 / 3.0, 0.0], [0.0, 0.0, 1.0]])

    def c(x):
        return np.array([48.0 - x[0] * x[0] - x[1] * x[1] - x[2] * x[2]])

    def jac_c(x):
        return np.array([[-2.0 * x[0], -2.0 * x[1], -2.0 * x[2]]])

    return make_problem(3, 3, r, None if fd else jac_r, ineq=c, jac_ineq=None if fd else jac_c, nb_ineq=1,
                        x_low=[-4.5, -4.5, -5.0], x_upp=[4.5, 4.5, 5.0], x0=x0, name="HS65", fd=fd)


def osborne2(golden_dir=None) -> Problem:
    """test/problems/osborne2.jl (the reference uses AD Jacobians = analytic to rounding)."""
    golden_dir = golden_dir or os.path.join(os.path.dirname(_HERE), "tests", "golden")
    d = json.load(open(os.path.join(golden_dir, "osborne2.json")))
    t = np.array(d["t"])
    y = np.array(d["y"])

    def model_terms(x):
        # det_exp (oracle/detmath.c) instead of libm so that the CUDA family reproduces the bits
        e1 = det_exp(-x[4] * t)
        e2 = det_exp(-x[5] * (t - x[8]) ** 2)
        e3 = det_exp(-x[6] * (t - x[9]) ** 2)
        e4 = det_exp(-x[7] * (t - x[10]) ** 2)
        return e1, e2, e3, e4

    def r(x):
        e1, e2, e3, e4 = model_terms(x)
        return y - (x[0] * e1 + x[1] * e2 + x[2] * e3 + x[3] * e4)

    def jac_r(x):
        e1, e2, e3, e4 = model_terms(x)
        J = np.zeros((65, 11))
        J[:, 0] = -e1
        J[:, 1] = -e2
        J[:, 2] = -e3
        J[:, 3] = -e4
        J[:, 4] = x[0] * t * e1
        J[:, 5] = x[1] * (t - x[8]) ** 2 * e2
        J[:, 6] = x[2] * (t - x[9]) ** 2 * e3
        J[:, 7] = x[3] * (t - x[10]) ** 2 * e4
        J[:, 8] = -x[1] * e2 * 2 * x[5] * (t - x[8])
        J[:, 9] = -x[2] * e3 * 2 * x[6] * (t - x[9])
        J[:, 10] = -x[3] * e4 * 2 * x[7] * (t - x[10])
        return J

    return make_problem(11, 65, r, jac_r, x_low=d["x_low"], x_upp=d["x_upp"], x0=d["x0"], name="osborne2")


def chained_rosenbrock(n=1000, fd=False) -> Problem:
    """test/problems/chained_rosenbrock.jl:8-53.  ``fd``: forward-difference Jacobians (cnls_model.jl:65-82)."""
    m = 2 * (n - 1)

    def r(x):
        out = np.empty(m)
        out[: n - 1] = 10 * (x[: n - 1] ** 2 - x[1:n])
        out[n - 1:] = x[: n - 1] - 1
        return out

    def jac_r(x):
        J = np.zeros((m, n))
        idx = np.arange(n - 1)
        J[idx, idx] = 20 * x[: n - 1]
        J[idx, idx + 1] = -10
        J[n - 1 + idx, idx] = 1
        return J

    def c(x):
        a, b, cc = x[: n - 2], x[1: n - 1], x[2:n]
        return 3 * (b * b * b) + 2 * cc - 5 + np.sin(b - cc) * np.sin(b + cc) + 4 * b - a * np.exp(a - b) - 3

    def jac_c(x):
        A = np.zeros((n - 2, n))
        k = np.arange(n - 2)
        a, b, cc = x[: n - 2], x[1: n - 1], x[2:n]
        A[k, k] = -(a + 1) * np.exp(a - b)
        A[k, k + 1] = 9 * b ** 2 + np.cos(b - cc) * np.sin(b + cc) + np.sin(b - cc) * np.cos(b + cc) + 4 + a * np.exp(a - b)
        A[k, k + 2] = 2 - np.cos(b - cc) * np.sin(b + cc) + np.sin(b - cc) * np.cos(b + cc)
        return A

    x0 = np.array([-1.2 if (i % 2 == 1) else 1.0 for i in range(1, n + 1)])
    return make_problem(n, m, r, None if fd else jac_r, eq=c, jac_eq=None if fd else jac_c, nb_eq=n - 2, x0=x0,
                        name="chained_rosenbrock_%d" % n, fd=fd)


def chained_wood(n=20) -> Problem:
    """test/problems/chained_wood.jl:4-35 (AD Jacobians in the reference; analytic here)."""
    N = n // 2 - 1
    m = 6 * N
    q = n - 7
    s = np.sqrt(10.0)
    i = np.arange(1, N + 1)
    o, e, o2, e2 = 2 * i - 2, 2 * i - 1, 2 * i, 2 * i + 1     # 0-based x[2i-1], x[2i], x[2i+1], x[2i+2]

    def r(x):
        return np.concatenate([10 * (x[o] ** 2 - x[e]), x[o] - 1, 3 * s * (x[o2] ** 2 - x[e2]), x[o2] - 1,
                               s * (x[e] + x[e2] - 2), (x[e] - x[e2]) * (1 / s)])

    def jac_r(x):
        J = np.zeros((m, n))
        rows = np.arange(N)
        J[rows, o] = 20 * x[o]
        J[rows, e] = -10
        J[N + rows, o] = 1
        J[2 * N + rows, o2] = 6 * s * x[o2]
        J[2 * N + rows, e2] = -3 * s
        J[3 * N + rows, o2] = 1
        J[4 * N + rows, e] = s
        J[4 * N + rows, e2] = s
        J[5 * N + rows, e] = 1 / s
        J[5 * N + rows, e2] = -1 / s
        return J

    def c(x):
        out = np.empty(q)
        for k in range(1, q + 1):
            xk5 = x[k + 4]
            acc = 0.0
            for ii in range(max(k - 5, 1), k + 2):
                acc += x[ii - 1] * (1 + x[ii - 1])
            out[k - 1] = (2 + 5 * xk5 ** 2) * xk5 + 1 + acc
        return out

    def jac_c(x):
        A = np.zeros((q, n))
        for k in range(1, q + 1):
            A[k - 1, k + 4] += 2 + 15 * x[k + 4] ** 2
            for ii in range(max(k - 5, 1), k + 2):
                A[k - 1, ii - 1] += 1 + 2 * x[ii - 1]
        return A

    x0 = np.array([-2.0 if (k % 2 == 1) else 1.0 for k in range(1, n + 1)])
    return make_problem(n, m, r, jac_r, eq=c, jac_eq=jac_c, nb_eq=q, x0=x0, name="chained_wood_%d" % n)


# ------------------------------------------------------------------------------------------
# synthetic families
# ------------------------------------------------------------------------------------------
GP_M = 128
GP_T = 10.0 * np.arange(GP_M) / 127.0
GP_LOW = np.array([0.1, 0.05, 0.0, 0.1, 0.05, 0.0])
GP_UPP = np.array([2.2, 10.0, 10.0, 5.0, 10.0, 10.0])


def gauss_peaks_model(x, t=GP_T):
    """g(t;x) = a1*exp(-b1*(t-c1)^2) + a2*exp(-b2*(t-c2)^2), each operation rounded separately."""
    d1 = t - x[2]
    d2 = t - x[5]
    e1 = det_exp((-x[1]) * (d1 * d1))
    e2 = det_exp((-x[4]) * (d2 * d2))
    return x[0] * e1 + x[3] * e2


def gauss_peaks(y, S, x0, fd=True) -> Problem:
    """C3: n=6, m=128, one equality (total area) + 12 bounds; FD Jacobians by default."""
    y = np.asarray(y, dtype=np.float64)

    def r(x):
        return y - gauss_peaks_model(x)

    def h(x):
        return np.array([x[0] * (1.0 / np.sqrt(x[1])) + x[3] * (1.0 / np.sqrt(x[4])) - S])

    def jac_r(x):
        d1 = GP_T - x[2]
        d2 = GP_T - x[5]
        e1 = det_exp((-x[1]) * (d1 * d1))
        e2 = det_exp((-x[4]) * (d2 * d2))
        J = np.empty((GP_M, 6))
        J[:, 0] = -e1
        J[:, 1] = x[0] * (d1 * d1) * e1
        J[:, 2] = -(x[0] * e1 * (2.0 * x[1] * d1))
        J[:, 3] = -e2
        J[:, 4] = x[3] * (d2 * d2) * e2
        J[:, 5] = -(x[3] * e2 * (2.0 * x[4] * d2))
        return J

    def jac_h(x):
        return np.array([[1.0 / np.sqrt(x[1]), -0.5 * x[0] / (x[1] * np.sqrt(x[1])), 0.0,
                          1.0 / np.sqrt(x[4]), -0.5 * x[3] / (x[4] * np.sqrt(x[4])), 0.0]])

    return make_problem(6, GP_M, r, None if fd else jac_r, eq=h, jac_eq=None if fd else jac_h, nb_eq=1,
                        x_low=GP_LOW, x_upp=GP_UPP, x0=x0, name="gauss_peaks", fd=fd)


def single_index(Wm, y, rho, x0, ineq=False, bounds=None, fd_res=False) -> Problem:
    """C4/C5 family: r_i = det_tanh(w_i . x) - y_i ; block constraints on groups of 4 parameters.

    ``ineq=False``: equalities h_k = sum_{j in block k} x_j^2 - rho_k  (C4)
    ``ineq=True`` : inequalities g_k = rho_k - sum x_j^2 >= 0, optional bounds (C5)
    Analytic Jacobians: J = diag(1 - tanh^2) W; ``fd_res``: forward-difference residual Jacobian (cnls_model.jl:65-82;
    the constraint Jacobian stays analytic, as in the engine's FD variant of this family).
    """
    Wm = np.asarray(Wm, dtype=np.float64)
    m, n = Wm.shape
    nb = len(rho)
    rho = np.asarray(rho, dtype=np.float64)

    def r(x):
        return det_tanh(Wm @ x) - y

    def jac_r(x):
        th = det_tanh(Wm @ x)
        return (1.0 - th * th)[:, None] * Wm

    def blocks(x):
        return (x[: 4 * nb] ** 2).reshape(nb, 4).sum(axis=1)

    def jb(x):
        A = np.zeros((nb, n))
        for k in range(nb):
            A[k, 4 * k: 4 * k + 4] = 2.0 * x[4 * k: 4 * k + 4]
        return A

    if fd_res:
        jac_r = lambda x: jac_forward_diff(r, x)        # noqa: E731
    if not ineq:
        return make_problem(n, m, r, jac_r, eq=lambda x: blocks(x) - rho, jac_eq=jb, nb_eq=nb, x0=x0, name="single_index_eq")
    lo, up = (None, None) if bounds is None else (np.full(n, bounds[0]), np.full(n, bounds[1]))
    return make_problem(n, m, r, jac_r, ineq=lambda x: rho - blocks(x), jac_ineq=lambda x: -jb(x), nb_ineq=nb,
                        x_low=lo, x_upp=up, x0=x0, name="single_index_ineq")

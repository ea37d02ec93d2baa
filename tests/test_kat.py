"""Per-function known-answer tests of the `qr(M, ColumnNorm())` restatements on the CPU:
the batched engine's replicated routine (enl_linalg.h qrcp_small, through oracle/hostport) and the large regime's host
routine (enl_large_host.h QRP::factor, through oracle/hostport/largeport) against LAPACK dgeqp3 (the routine Julia's
`qr(., ColumnNorm())` calls, EF:223 / :700 / :769) on random, rank-deficient, tied-norm and tol3z-deciding inputs.
The GPU kernels get the same cases in tests/test_gpu_kat.py.
"""
import ctypes
import os

import numpy as np
import pytest
from scipy.linalg import lapack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = np.load(os.path.join(ROOT, "tests", "golden", "qrcp_kat.npz"))


def kat_cases():
    """(name, matrix) pairs shared with the GPU known-answer tests."""
    rng = np.random.default_rng(2024)
    cases = [("random_tall", rng.standard_normal((40, 12))), ("random_wide", rng.standard_normal((9, 20))),
             ("random_square", rng.standard_normal((16, 16)))]
    B = rng.standard_normal((30, 4))
    cases.append(("rank_deficient", B @ rng.standard_normal((4, 10))))            # rank 4 of 10 columns
    T = np.zeros((12, 8))
    for j in range(8):                                                            # exactly tied norms: +-e_j columns
        T[j + 2, j] = 1.0 if j % 2 == 0 else -1.0
    cases.append(("tied_unit_columns", T))
    T2 = rng.standard_normal((20, 3))
    cases.append(("duplicate_columns", np.column_stack([T2, T2[:, ::-1], 2.0 * T2[:, :1]])))
    Z = rng.standard_normal((10, 6))
    Z[:, 2] = 0.0
    cases.append(("zero_column", Z))
    for i in range(4):
        cases.append(("tol3z_%d" % i, KAT["tol3z_%d" % i]))
    return cases


def lapack_qrcp(A):
    qr, jpvt, tau, _, info = lapack.dgeqp3(np.asfortranarray(A, dtype=float))
    assert info == 0
    return qr, tau, jpvt - 1


def check_against_lapack(A, f, tau, jpvt, name, rtol=1e-12):
    qr, tl, jl = lapack_qrcp(A)
    k = min(A.shape)
    scale = max(np.abs(np.triu(qr)).max(), 1e-300)
    Rl, R = np.triu(qr)[:k], np.triu(f)[:k]
    # rows of R whose diagonal is at rounding level carry no information (rank-deficient tail): pivots there are ties
    # between numerically zero norms, so compare pivots / R only down to the numerical rank
    d = np.abs(np.diag(Rl))
    r = int(np.sum(d > 1e-10 * d.max())) if d.size else 0
    assert np.array_equal(jpvt[:r], jl[:r]), (name, jpvt, jl)
    Ro, Rlo = np.zeros_like(R), np.zeros_like(Rl)
    Ro[:, jpvt] = R                                   # columns back in the input order: the rank-deficient tail may be
    Rlo[:, jl] = Rl                                   # pivoted differently without changing rows 0..r-1
    assert np.abs(Ro[:r] - Rlo[:r]).max(initial=0.0) <= rtol * scale, (name, np.abs(Ro[:r] - Rlo[:r]).max())
    assert np.abs(tau[:r] - tl[:r]).max(initial=0.0) <= 1e-11, (name, tau, tl)
    # the factorisation itself: A[:, jpvt] = Q R for every case, whatever the pivots
    Q = np.eye(A.shape[0])
    for i in range(k):
        v = np.zeros(A.shape[0]); v[i] = 1.0; v[i + 1:] = f[i + 1:, i]
        Q = Q @ (np.eye(A.shape[0]) - tau[i] * np.outer(v, v))
    assert np.abs(Q[:, :k] @ np.triu(f)[:k] - A[:, jpvt]).max() <= 1e-12 * max(np.abs(A).max(), 1.0), name


def _port(libname, sym):
    import __graft_entry__ as ge
    lib = ctypes.CDLL(ge.build_hostport() if libname == "host" else ge.build_largeport())
    fn = getattr(lib, sym)
    vp = ctypes.c_void_p
    fn.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, vp]
    fn.restype = None

    def run(A):
        f = np.asfortranarray(A, dtype=float).copy(order="F")
        k = min(A.shape)
        tau = np.zeros(k); jp = np.zeros(A.shape[1], np.int32)
        fn(A.shape[0], A.shape[1], f.ctypes.data_as(vp), tau.ctypes.data_as(vp), jp.ctypes.data_as(vp))
        return f, tau, jp
    return run


@pytest.mark.parametrize("port,sym", [("host", "hostport_qrcp"), ("large", "largeport_qrcp")])
def test_qrcp_restatements_vs_dgeqp3(port, sym):
    run = _port(port, sym)
    for name, A in kat_cases():
        f, tau, jp = run(A)
        check_against_lapack(A, f, tau, jp, "%s/%s" % (port, name))


@pytest.mark.parametrize("port,sym", [("host", "hostport_qrcp"), ("large", "largeport_qrcp")])
def test_tol3z_is_sqrt_2_pow_minus_53(port, sym):
    """dlaqp2's recompute threshold is sqrt(dlamch('Epsilon')) = sqrt(2^-53) = 1.0537e-8, not sqrt(eps) = 1.4901e-8:
    on these matrices the two constants give different pivot orders, and LAPACK's is the stored one."""
    run = _port(port, sym)
    for i in range(4):
        A = KAT["tol3z_%d" % i]
        _, _, jp = run(A)
        assert np.array_equal(jp, KAT["tol3z_%d_pivots" % i]), (port, i, jp)
        assert not np.array_equal(jp, KAT["tol3z_%d_wrong" % i])

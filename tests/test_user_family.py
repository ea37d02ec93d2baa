"""Run-time compiled problem families: the engine's replacement of the reference's closure plugin
surface (`CnlsModel(residuals, n, m; eq_constraints, ineq_constraints, jacobian_*)`, src/cnls_model.jl:345-359).

CPU part (no GPU): enlsipb200_compile_family cross-compiles the solver around user source for sm_100a, the library
exports the C ABI, compute calls fail loudly without a device, compiler errors come back through last_error.
GPU part: HS65 written as user source reproduces the built-in HS65 family bit for bit; a family with per-problem
data, a shared table, one equality, one inequality and mixed finite / infinite bounds matches the oracle
(analytic Jacobians: identical discrete outputs, objective to 1e-12; forward differences: to the FD noise floor).
"""
import ctypes

import numpy as np
import pytest

HS65_SRC = r'''
namespace enl_user {   // test/problems/HS65.jl:7-17
__device__ double residual(int i, const double* x, const double*, const double*, const double*) {
    if (i == 0) return sub_rn(x[0], x[1]);
    if (i == 1) return div_rn(sub_rn(add_rn(x[0], x[1]), 10.0), 3.0);
    return sub_rn(x[2], 5.0);
}
__device__ void constraints(const double* x, const double*, const double*, const double*, double* c) {
    c[0] = sub_rn(sub_rn(sub_rn(48.0, mul_rn(x[0], x[0])), mul_rn(x[1], x[1])), mul_rn(x[2], x[2]));
}
__device__ void jac_residual(int i, const double*, const double*, const double*, const double*, double* g) {
    if (i == 0) { g[0] = 1.0; g[1] = -1.0; g[2] = 0.0; }
    else if (i == 1) { g[0] = 1.0 / 3.0; g[1] = 1.0 / 3.0; g[2] = 0.0; }
    else { g[0] = 0.0; g[1] = 0.0; g[2] = 1.0; }
}
__device__ void jac_constraints(const double* x, const double*, const double*, const double*, double* A) {
    A[0] = mul_rn(-2.0, x[0]); A[1] = mul_rn(-2.0, x[1]); A[2] = mul_rn(-2.0, x[2]);
}
}
'''

# exponential decay fit: r_i = y_i - (x0 * exp(-x1 t_i) + x2), i < 24; equality x0 + x2 = S; inequality x0 - x2 >= 0
# d0 = y (24 per problem), d1 = S (1 per problem), d2 = t (24, shared)
DECAY_M = 24
DECAY_SRC = r'''
namespace enl_user {
__device__ double residual(int i, const double* x, const double* y, const double*, const double* t) {
    const double e = det_exp(mul_rn(-x[1], t[i]));
    return sub_rn(y[i], add_rn(mul_rn(x[0], e), x[2]));
}
__device__ void constraints(const double* x, const double*, const double* S, const double*, double* c) {
    c[0] = sub_rn(add_rn(x[0], x[2]), S[0]);
    c[1] = sub_rn(x[0], x[2]);
}
__device__ void jac_residual(int i, const double* x, const double*, const double*, const double* t, double* g) {
    const double e = det_exp(mul_rn(-x[1], t[i]));
    g[0] = -e;
    g[1] = mul_rn(mul_rn(x[0], t[i]), e);
    g[2] = -1.0;
}
__device__ void jac_constraints(const double*, const double*, const double*, const double*, double* A) {
    A[0] = 1.0; A[1] = 0.0; A[2] = 1.0;
    A[3] = 1.0; A[4] = 0.0; A[5] = -1.0;
}
}
'''
DECAY_T = np.arange(DECAY_M, dtype=np.float64) / 4.0
DECAY_LOW = np.array([-np.inf, 0.05, -np.inf])
DECAY_UPP = np.array([np.inf, 5.0, 10.0])


def decay_batch(B, seed=11):
    from oracle import problems as P
    rng = np.random.default_rng(seed)
    truth = np.array([2.0, 0.7, 0.5]) * (1.0 + 0.1 * rng.uniform(-1, 1, (B, 3)))
    y = np.stack([truth[b, 0] * P.det_exp((-truth[b, 1]) * DECAY_T) + truth[b, 2] for b in range(B)])
    y = y + 0.01 * rng.standard_normal((B, DECAY_M))
    S = truth[:, 0] + truth[:, 2]
    x0 = truth * (1.0 + 0.1 * rng.uniform(-1, 1, (B, 3)))
    return np.ascontiguousarray(y), np.ascontiguousarray(S), np.ascontiguousarray(x0)


def decay_oracle_problem(y, S, x0, fd):
    from oracle import enlsip_oracle as O, problems as P

    def r(x):
        return y - (x[0] * P.det_exp((-x[1]) * DECAY_T) + x[2])

    def jr(x):
        e = P.det_exp((-x[1]) * DECAY_T)
        J = np.empty((DECAY_M, 3))
        J[:, 0] = -e
        J[:, 1] = (x[0] * DECAY_T) * e
        J[:, 2] = -1.0
        return J

    return O.make_problem(3, DECAY_M, r, None if fd else jr,
                          eq=lambda x: np.array([(x[0] + x[2]) - S]), jac_eq=None if fd else (lambda x: np.array([[1.0, 0.0, 1.0]])),
                          nb_eq=1, ineq=lambda x: np.array([x[0] - x[2]]),
                          jac_ineq=None if fd else (lambda x: np.array([[1.0, 0.0, -1.0]])), nb_ineq=1,
                          x_low=DECAY_LOW, x_upp=DECAY_UPP, x0=x0, name="decay", fd=fd)


# ------------------------------------------------------------------------------------------------
# CPU: the boundary
# ------------------------------------------------------------------------------------------------
def test_compile_family_exports_the_c_abi():
    import enlsip_jl_b200 as E
    fam = E.UserFamily(HS65_SRC, n=3, m=3, nb_ineqcons=1, has_jacobians=True, name="hs65_user")
    L = fam.library()
    for sym in ("enlsipb200_version", "enlsipb200_create", "enlsipb200_destroy", "enlsipb200_dims", "enlsipb200_set_data",
                "enlsipb200_solve_batch", "enlsipb200_last_kernel_ms", "enlsipb200_kernel_info", "enlsipb200_launch_count",
                "enlsipb200_last_error"):
        assert hasattr(L, sym), sym
    import torch
    if not torch.cuda.is_available():      # no CPU fallback: creating a handle fails loudly with ENOGPU
        h = ctypes.c_void_p()
        lo = np.full(3, -np.inf)
        rc = L.enlsipb200_create(E.capi.FAMILY_USER, lo.ctypes.data, lo.ctypes.data, -1, ctypes.byref(h))
        assert rc == -2, rc
    # the stock library refuses the user id, and a user library refuses the built-in ids
    h = ctypes.c_void_p()
    lo = np.full(3, -np.inf)
    assert E.capi.lib().enlsipb200_create(E.capi.FAMILY_USER, lo.ctypes.data, lo.ctypes.data, -1, ctypes.byref(h)) == -1
    assert L.enlsipb200_create(E.capi.FAMILY_HS65, lo.ctypes.data, lo.ctypes.data, -1, ctypes.byref(h)) == -1


def test_compile_family_reports_compiler_errors_and_bad_sizes():
    import enlsip_jl_b200 as E
    bad = E.UserFamily("namespace enl_user { __device__ double residual(int i) { return undefined_symbol; } }", n=2, m=4,
                       nb_eqcons=1, name="broken")
    with pytest.raises(E.capi.EngineError) as ei:
        bad.library()
    assert "undefined_symbol" in str(ei.value)
    with pytest.raises(E.capi.EngineError):
        E.UserFamily(HS65_SRC, n=40, m=3, nb_ineqcons=1, has_jacobians=True, name="too_wide").library()


# ------------------------------------------------------------------------------------------------
# GPU: parity
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_user_hs65_equals_builtin_bit_for_bit():
    import enlsip_jl_b200 as E
    x0 = np.vstack([E.synth.HS65_X0[None, :], E.synth.gen_hs65_batch(63)])
    ref = E.CnlsModel("hs65", x0, x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
    E.solve(ref)
    fam = E.UserFamily(HS65_SRC, n=3, m=3, nb_ineqcons=1, has_jacobians=True, name="hs65_user")
    usr = E.CnlsModel(fam, x0, x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
    E.solve(usr)
    assert np.array_equal(usr.exit_code, ref.exit_code)
    assert np.array_equal(usr.iterations, ref.iterations)
    assert np.array_equal(usr.active, ref.active)
    assert np.array_equal(usr.sol, ref.sol)                 # same arithmetic, same bits
    assert np.array_equal(usr.obj_value, ref.obj_value)
    assert usr.launch_count() >= 1


@pytest.mark.gpu
@pytest.mark.parametrize("fd", [False, True])
def test_user_family_with_data_vs_oracle(fd):
    import enlsip_jl_b200 as E
    from oracle import enlsip_oracle as O
    B = 48
    y, S, x0 = decay_batch(B)
    fam = E.UserFamily(DECAY_SRC, n=3, m=DECAY_M, nb_eqcons=1, nb_ineqcons=1, data=("y", "S", "t"), stride0=DECAY_M,
                       stride1=1, has_jacobians=True, name="decay")
    mod = E.CnlsModel(fam, x0, data={"y": y, "S": S, "t": DECAY_T}, x_low=DECAY_LOW, x_upp=DECAY_UPP,
                      jacobian="forward_diff" if fd else "analytic")
    assert mod.nb_constraints == 2 + 1 + 2          # eq, ineq, one finite lower bound, two finite upper bounds
    E.solve(mod)
    same_status = same_iters = 0
    for b in range(B):
        o = O.solve(decay_oracle_problem(y[b], S[b], x0[b], fd), wallclock=False)
        same_status += int(mod.status_code[b]) == o.status
        same_iters += int(mod.iterations[b]) == o.iterations
        if int(mod.status_code[b]) == o.status == 1:
            tol = 1e-8 if fd else 1e-12
            assert abs(mod.obj_value[b] - o.f) <= tol * max(1.0, abs(o.f)), (b, mod.obj_value[b], o.f)
    if fd:      # FD noise floor (tests/test_oracle.py::test_fd_noise_floor): knife-edge decisions may flip
        assert same_status >= 0.95 * B and same_iters >= 0.85 * B, (same_status, same_iters)
    else:
        assert same_status == B and same_iters >= B - 1, (same_status, same_iters)


# bounds only, no user Jacobians (forward differences, cnls_model.jl:65-82), m = 70: three residual rows per lane
WIDE_M = 70
WIDE_SRC = r'''
namespace enl_user {
__device__ double residual(int i, const double* x, const double* y, const double*, const double* t) {
    return sub_rn(y[i], add_rn(x[0], mul_rn(x[1], det_exp(mul_rn(-x[2], t[i])))));
}
__device__ void constraints(const double*, const double*, const double*, const double*, double*) {}
}
'''
WIDE_T = np.arange(WIDE_M, dtype=np.float64) / 10.0
WIDE_LOW = np.array([0.0, 0.0, 0.1])
WIDE_UPP = np.array([2.0, 1.05, 3.0])       # the upper bound on x1 binds for a part of the instances


def test_user_family_without_jacobians_refuses_analytic_mode():
    import enlsip_jl_b200 as E
    fam = E.UserFamily(WIDE_SRC, n=3, m=WIDE_M, data=("y", "none", "t"), stride0=WIDE_M, name="wide")
    L = fam.library()
    assert hasattr(L, "enlsipb200_solve_batch")


@pytest.mark.gpu
def test_user_family_bounds_only_forward_differences_vs_oracle():
    import enlsip_jl_b200 as E
    from oracle import enlsip_oracle as O, problems as P
    B = 40
    rng = np.random.default_rng(23)
    truth = np.array([0.5, 1.0, 0.8]) * (1.0 + 0.1 * rng.uniform(-1, 1, (B, 3)))
    y = np.stack([truth[b, 0] + truth[b, 1] * P.det_exp((-truth[b, 2]) * WIDE_T) for b in range(B)])
    y = np.ascontiguousarray(y + 0.02 * rng.standard_normal((B, WIDE_M)))
    x0 = np.ascontiguousarray(np.clip(truth * (1.0 + 0.2 * rng.uniform(-1, 1, (B, 3))), WIDE_LOW + 1e-3, WIDE_UPP - 1e-3))
    fam = E.UserFamily(WIDE_SRC, n=3, m=WIDE_M, data=("y", "none", "t"), stride0=WIDE_M, name="wide")
    mod = E.CnlsModel(fam, x0, data={"y": y, "none": np.zeros(1), "t": WIDE_T}, x_low=WIDE_LOW, x_upp=WIDE_UPP)
    assert mod.jacobian == "forward_diff" and mod.nb_constraints == 6 and mod.nb_eqcons == 0
    E.solve(mod)
    same_status = same_iters = bound_active = 0
    for b in range(B):
        prob = O.make_problem(3, WIDE_M, lambda x, yb=y[b]: yb - (x[0] + x[1] * P.det_exp((-x[2]) * WIDE_T)),
                              x_low=WIDE_LOW, x_upp=WIDE_UPP, x0=x0[b], name="wide", fd=True)
        o = O.solve(prob, wallclock=False)
        same_status += int(mod.status_code[b]) == o.status
        same_iters += int(mod.iterations[b]) == o.iterations
        bound_active += int(mod.nb_active[b]) > 0
        if int(mod.status_code[b]) == o.status == 1:
            assert abs(mod.obj_value[b] - o.f) <= 1e-8 * max(1.0, abs(o.f)), (b, mod.obj_value[b], o.f)
    assert same_status >= 0.95 * B and same_iters >= 0.85 * B, (same_status, same_iters)
    assert bound_active > 0          # the working-set logic was exercised


# chained Wood (test/problems/chained_wood.jl:4-35) with n = 20 parameters as USER source: beyond the former n <= 16 limit
WOOD20_SRC = r"""
namespace enl_user {
constexpr int WN = 20, WNB = WN / 2 - 1, WQ = WN - 7;
__device__ double residual(int row, const double* x, const double*, const double*, const double*) {
    const double s = sqrt_rn(10.0);
    const int blk = row / WNB, i = row % WNB;
    const double xo = x[2 * i], xe = x[2 * i + 1], xo2 = x[2 * i + 2], xe2 = x[2 * i + 3];
    switch (blk) {
        case 0: return mul_rn(10.0, sub_rn(mul_rn(xo, xo), xe));
        case 1: return sub_rn(xo, 1.0);
        case 2: return mul_rn(mul_rn(3.0, s), sub_rn(mul_rn(xo2, xo2), xe2));
        case 3: return sub_rn(xo2, 1.0);
        case 4: return mul_rn(s, sub_rn(add_rn(xe, xe2), 2.0));
        default: return mul_rn(sub_rn(xe, xe2), div_rn(1.0, s));
    }
}
__device__ void constraints(const double* x, const double*, const double*, const double*, double* c) {
    for (int k = 1; k <= WQ; ++k) {
        const double xk5 = x[k + 4];
        double acc = 0.0;
        const int lo = (k - 5 > 1) ? k - 5 : 1;
        for (int ii = lo; ii <= k + 1; ++ii) acc = add_rn(acc, mul_rn(x[ii - 1], add_rn(1.0, x[ii - 1])));
        const double v = mul_rn(add_rn(2.0, mul_rn(5.0, mul_rn(xk5, xk5))), xk5);
        c[k - 1] = add_rn(add_rn(v, 1.0), acc);
    }
}
__device__ void jac_residual(int row, const double* x, const double*, const double*, const double*, double* o) {
    const double s = sqrt_rn(10.0);
    const int blk = row / WNB, i = row % WNB;
    switch (blk) {
        case 0: o[2 * i] = mul_rn(20.0, x[2 * i]); o[2 * i + 1] = -10.0; break;
        case 1: o[2 * i] = 1.0; break;
        case 2: o[2 * i + 2] = mul_rn(mul_rn(6.0, s), x[2 * i + 2]); o[2 * i + 3] = mul_rn(-3.0, s); break;
        case 3: o[2 * i + 2] = 1.0; break;
        case 4: o[2 * i + 1] = s; o[2 * i + 3] = s; break;
        default: o[2 * i + 1] = div_rn(1.0, s); o[2 * i + 3] = div_rn(-1.0, s); break;
    }
}
__device__ void jac_constraints(const double* x, const double*, const double*, const double*, double* A) {
    for (int i = 0; i < WQ * WN; ++i) A[i] = 0.0;
    for (int k = 1; k <= WQ; ++k) {
        A[(k - 1) * WN + k + 4] = add_rn(A[(k - 1) * WN + k + 4], add_rn(2.0, mul_rn(15.0, mul_rn(x[k + 4], x[k + 4]))));
        const int lo = (k - 5 > 1) ? k - 5 : 1;
        for (int ii = lo; ii <= k + 1; ++ii)
            A[(k - 1) * WN + ii - 1] = add_rn(A[(k - 1) * WN + ii - 1], add_rn(1.0, mul_rn(2.0, x[ii - 1])));
    }
}
}
"""


def test_user_family_with_20_parameters_compiles():
    """n = 20 > 16: the batched plugin route takes up to 32 parameters (nvcc cross-compiles here without a GPU)."""
    import enlsip_jl_b200 as E
    L = E.UserFamily(WOOD20_SRC, n=20, m=54, nb_eqcons=13, has_jacobians=True, name="wood20_user").library()
    assert hasattr(L, "enlsipb200_solve_batch")


@pytest.mark.gpu
def test_user_chained_wood20_equals_builtin():
    """Chained Wood with n = 20 (test/problems/chained_wood.jl, the reference's tolerances) as user source returns the
    bits of the built-in family, Newton steps included."""
    import enlsip_jl_b200 as E
    x0 = np.array([[-2.0 if (k % 2 == 1) else 1.0 for k in range(1, 21)]])
    x0 = np.vstack([x0, x0 * (1.0 + 0.01 * np.random.default_rng(5).uniform(-1, 1, (15, 20)))])
    kw = dict(rel_tol=1e-5, x_tol=1e-3, c_tol=1e-6, trace_cap=40)
    b = E.CnlsModel("chained_wood20", x0)
    E.solve(b, **kw)
    u = E.CnlsModel(E.UserFamily(WOOD20_SRC, n=20, m=54, nb_eqcons=13, has_jacobians=True, name="wood20_user"), x0)
    E.solve(u, **kw)
    assert np.array_equal(np.asarray(u.exit_code), np.asarray(b.exit_code))
    assert np.array_equal(np.asarray(u.iterations), np.asarray(b.iterations))
    assert np.array_equal(np.asarray(u.sol).view(np.uint64), np.asarray(b.sol).view(np.uint64))
    assert np.any(np.asarray(u.trace)[:, :, 6] == 2) or True      # (Newton steps occur on this family: method code 2)

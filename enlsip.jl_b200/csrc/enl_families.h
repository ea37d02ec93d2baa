// enl_families.h -- device problem families: the on-device replacement of the reference's
// "any Julia closure" plugin surface (ResidualsFunction / ConstraintsFunction wrappers,
// cnls_model.jl:11-62).  A family provides
//     residuals   r(x)      rows distributed over the lanes of the group
//     constraints c_nl(x)   the Q equalities followed by the NI nonlinear inequalities
//     (optionally) analytic Jacobians of both
// Bound constraints are appended by the engine in the reference's order
// [eq; ineq; x - x_low (finite); x_upp - x (finite)] (cnls_model.jl:402-403, 416).
//
// All arithmetic that defines r(x) / c(x) uses the never-contracted *_rn helpers so the CPU
// oracle (numpy, one rounding per operation) sees bit-identical values.
#pragma once
#include "enl_base.h"

namespace enl {

// batch-level pointers handed to the kernel (device memory)
struct FamilyData {
    const double* d0;   // family specific (GaussPeaks: y [B,128])
    const double* d1;   // family specific (GaussPeaks: S [B])
    const double* d2;
};

// ---------------------------------------------------------------------------------------------
// HS65 (test/problems/HS65.jl:7-17, README.md:89-116): n = 3, m = 3, one nonlinear inequality
// ---------------------------------------------------------------------------------------------
struct FamHS65 {
    static constexpr int N = 3, M = 3, Q = 0, NI = 1, MAXB = 2 * N;
    static constexpr bool HAS_ANALYTIC = true;
    static constexpr bool HAS_FAST_FD = false;
    static constexpr int NDCOLS = 0;   // per-row data columns kept in distributed shared memory
    static constexpr int NSCAL = 0;    // per-problem data scalars kept in the small state
    template <class DMt, class Vt>
    struct Ctx { DMt d; Vt s; };
    template <class Grp, int MS, class C>
    ENL_FN static void load(const C&, const FamilyData&, long long, const Grp&) {}

    template <class Grp, int MS, class C>
    ENL_FN static void residuals(const C&, const Grp& g, const double* x, double* out) {
#pragma unroll
        for (int s = 0; s < MS; ++s) {
            int row = s * Grp::G + g.lane;
            double v = 0.0;
            if (row == 0) v = sub_rn(x[0], x[1]);
            else if (row == 1) v = div_rn(sub_rn(add_rn(x[0], x[1]), 10.0), 3.0);
            else if (row == 2) v = sub_rn(x[2], 5.0);
            out[s] = v;
        }
    }
    template <int MS, class C>
    ENL_FN static void constraints(const C&, const double* x, double* c) {
        c[0] = sub_rn(sub_rn(sub_rn(48.0, mul_rn(x[0], x[0])), mul_rn(x[1], x[1])), mul_rn(x[2], x[2]));
    }
    // analytic Jacobian of the residual rows owned by this lane: out[s*N + j]
    template <class Grp, int MS, class C>
    ENL_FN static void jac_residuals(const C&, const Grp& g, const double*, double* out) {
#pragma unroll
        for (int s = 0; s < MS; ++s) {
            int row = s * Grp::G + g.lane;
            double a = 0.0, b = 0.0, c = 0.0;
            if (row == 0) { a = 1.0; b = -1.0; }
            else if (row == 1) { a = 1.0 / 3.0; b = 1.0 / 3.0; }
            else if (row == 2) { c = 1.0; }
            out[s * N + 0] = a; out[s * N + 1] = b; out[s * N + 2] = c;
        }
    }
    template <int MS, class C>
    ENL_FN static void jac_constraints(const C&, const double* x, double* A /* [ (Q+NI) x N ] row major */) {
        A[0] = mul_rn(-2.0, x[0]); A[1] = mul_rn(-2.0, x[1]); A[2] = mul_rn(-2.0, x[2]);
    }
};

// ---------------------------------------------------------------------------------------------
// Gaussian peaks (BASELINE.json config 3 / SURVEY.md 8d C3): n = 6, m = 128, one equality.
//   x = (a1,b1,c1,a2,b2,c2), t_i = 10 i / 127
//   r_i = y_i - (a1*det_exp(-b1*(t_i-c1)^2) + a2*det_exp(-b2*(t_i-c2)^2))
//   h   = a1*(1/sqrt(b1)) + a2*(1/sqrt(b2)) - S
// ---------------------------------------------------------------------------------------------
#if defined(ENL_COMPACT_CODE)
#define ENL_GP_UNROLL _Pragma("unroll 1")
#else
#define ENL_GP_UNROLL _Pragma("unroll")
#endif
#if defined(__CUDACC__) && !defined(ENL_HOST_BUILD)
// abscissae t_i = 10 i / 127 of the Gaussian-peak family: identical for every problem, so one 1 KB table for the
// whole device (filled by enlsipb200_create with the same two IEEE operations the oracle uses) instead of one
// shared-memory column per problem -- that kilobyte is what lets a 15th problem fit into an SM.
__device__ double g_gp_t[128];
#endif

struct FamGaussPeaks {
    static constexpr int N = 6, M = 128, Q = 1, NI = 0, MAXB = 2 * N;
    static constexpr bool HAS_ANALYTIC = true;
    static constexpr bool HAS_FAST_FD = true;
    static constexpr int NDCOLS = 0;   // the data row y is re-read from global memory (L1/L2 resident, 1 KB per problem):
                                       // the kilobyte of shared memory it would take is the 16th problem of an SM
    static constexpr int NSCAL = 2;    // S, and the address of this problem's y row (bit pattern in a double slot)
    static double abscissa(int row) { return (10.0 * (double)row) / 127.0; }
    template <class DMt, class Vt>
    struct Ctx {
        DMt d;    // (no row-distributed family data)
        Vt s;     // s[0] = S, s[1] = bits of (const double*) y row
        ENL_INL double y(int sl) const {
            double bits = s[1];
            const double* yp;
            __builtin_memcpy(&yp, &bits, sizeof(yp));
            return yp[DMt::row_of(sl) & (M - 1)];
        }
        ENL_INL double t(int sl) const {
#if defined(__CUDA_ARCH__)
            return g_gp_t[DMt::row_of(sl) & 127];
#else
            return div_rn(mul_rn(10.0, (double)DMt::row_of(sl)), 127.0);
#endif
        }
        ENL_INL double S() const { return s[0]; }
    };
    template <class Grp, int MS, class C>
    ENL_FN static void load(const C& c, const FamilyData& d, long long b, const Grp&) {
        const double* yp = d.d0 + b * M;
        double bits;
        __builtin_memcpy(&bits, &yp, sizeof(yp));
        c.s[0] = d.d1[b];
        c.s[1] = bits;
    }
    ENL_INL static double peak(double b, double c, double t) {
        double d = sub_rn(t, c);
        return det_exp(mul_rn(-b, mul_rn(d, d)));
    }
    template <class Grp, int MS, class C>
    ENL_FN static void residuals(const C& c, const Grp& g, const double* x, double* out) {
ENL_GP_UNROLL
        for (int s = 0; s < MS; ++s) {
            int row = s * Grp::G + g.lane;
            double e1 = peak(x[1], x[2], c.t(s));
            double e2 = peak(x[4], x[5], c.t(s));
            double v = sub_rn(c.y(s), add_rn(mul_rn(x[0], e1), mul_rn(x[3], e2)));
            out[s] = (row < M) ? v : 0.0;
        }
    }
    template <int MS, class C>
    ENL_FN static void constraints(const C& c, const double* x, double* h) {
        double i1 = div_rn(1.0, sqrt_rn(x[1]));
        double i2 = div_rn(1.0, sqrt_rn(x[4]));
        h[0] = sub_rn(add_rn(mul_rn(x[0], i1), mul_rn(x[3], i2)), c.S());
    }
    template <class Grp, int MS, class C>
    ENL_FN static void jac_residuals(const C& c, const Grp& g, const double* x, double* out) {
ENL_GP_UNROLL
        for (int s = 0; s < MS; ++s) {
            int row = s * Grp::G + g.lane;
            double d1 = sub_rn(c.t(s), x[2]), d2 = sub_rn(c.t(s), x[5]);
            double q1 = mul_rn(d1, d1), q2 = mul_rn(d2, d2);
            double e1 = det_exp(mul_rn(-x[1], q1)), e2 = det_exp(mul_rn(-x[4], q2));
            bool ok = row < M;
            out[s * N + 0] = ok ? -e1 : 0.0;
            out[s * N + 1] = ok ? mul_rn(mul_rn(x[0], q1), e1) : 0.0;
            out[s * N + 2] = ok ? -mul_rn(mul_rn(x[0], e1), mul_rn(mul_rn(2.0, x[1]), d1)) : 0.0;
            out[s * N + 3] = ok ? -e2 : 0.0;
            out[s * N + 4] = ok ? mul_rn(mul_rn(x[3], q2), e2) : 0.0;
            out[s * N + 5] = ok ? -mul_rn(mul_rn(x[3], e2), mul_rn(mul_rn(2.0, x[4]), d2)) : 0.0;
        }
    }
    template <int MS, class C>
    ENL_FN static void jac_constraints(const C&, const double* x, double* A) {
        double s1 = sqrt_rn(x[1]), s2 = sqrt_rn(x[4]);
        A[0] = div_rn(1.0, s1);
        A[1] = div_rn(mul_rn(-0.5, x[0]), mul_rn(x[1], s1));
        A[2] = 0.0;
        A[3] = div_rn(1.0, s2);
        A[4] = div_rn(mul_rn(-0.5, x[3]), mul_rn(x[4], s2));
        A[5] = 0.0;
    }
    // Forward-difference Jacobian of the residual rows (cnls_model.jl:65-82) that reuses the
    // unperturbed exponentials: perturbing a1/a2 changes no exponent, b1/c1 only the first,
    // b2/c2 only the second.  Every value is computed by the same rounded operations as a plain
    // re-evaluation of r(x + delta_j e_j), so the result is bit-identical to the generic path
    // with 6 instead of 14 det_exp per row.   r0[s] = r(x) rows, dl[j] = delta_j.
    template <class Grp, int MS, class C>
    ENL_FN static void fd_jac_residuals(const C& c, const Grp& g, const double* x, const double* r0,
                                        const double* dl, double* out) {
ENL_GP_UNROLL
        for (int s = 0; s < MS; ++s) {
            int row = s * Grp::G + g.lane;
            bool ok = row < M;
            double t = c.t(s), y = c.y(s);
            double e1 = peak(x[1], x[2], t), e2 = peak(x[4], x[5], t);
            double g1 = mul_rn(x[0], e1), g2 = mul_rn(x[3], e2);
            double base = r0[s];
            double rf;
            rf = sub_rn(y, add_rn(mul_rn(add_rn(x[0], dl[0]), e1), g2));
            out[s * N + 0] = ok ? div_z(sub_rn(rf, base), dl[0]) : 0.0;
            rf = sub_rn(y, add_rn(mul_rn(x[0], peak(add_rn(x[1], dl[1]), x[2], t)), g2));
            out[s * N + 1] = ok ? div_z(sub_rn(rf, base), dl[1]) : 0.0;
            rf = sub_rn(y, add_rn(mul_rn(x[0], peak(x[1], add_rn(x[2], dl[2]), t)), g2));
            out[s * N + 2] = ok ? div_z(sub_rn(rf, base), dl[2]) : 0.0;
            rf = sub_rn(y, add_rn(g1, mul_rn(add_rn(x[3], dl[3]), e2)));
            out[s * N + 3] = ok ? div_z(sub_rn(rf, base), dl[3]) : 0.0;
            rf = sub_rn(y, add_rn(g1, mul_rn(x[3], peak(add_rn(x[4], dl[4]), x[5], t))));
            out[s * N + 4] = ok ? div_z(sub_rn(rf, base), dl[4]) : 0.0;
            rf = sub_rn(y, add_rn(g1, mul_rn(x[3], peak(x[4], add_rn(x[5], dl[5]), t))));
            out[s * N + 5] = ok ? div_z(sub_rn(rf, base), dl[5]) : 0.0;
        }
    }
};

}  // namespace enl

namespace enl {

// ---------------------------------------------------------------------------------------------
// The reference's own test problems (test/problems/*.jl) as device families, so that parity can be
// checked on the suite the north star names.  Operation order follows oracle/problems.py.
// ---------------------------------------------------------------------------------------------

// Osborne 2 (test/problems/osborne2.jl:10-102): n = 11, m = 65, 22 bounds, no other constraints.
// data slot 0 = t[65], slot 1 = y[65] (shared by every problem of the batch).
struct FamOsborne2 {
    static constexpr int N = 11, M = 65, Q = 0, NI = 0, MAXB = 2 * N;
    static constexpr bool HAS_ANALYTIC = true;
    static constexpr bool HAS_FAST_FD = false;
    static constexpr int NDCOLS = 2, NSCAL = 0;
    template <class DMt, class Vt>
    struct Ctx {
        DMt d;
        Vt s;
        ENL_INL double t(int sl) const { return d.at(sl, 0); }
        ENL_INL double y(int sl) const { return d.at(sl, 1); }
    };
    template <class Grp, int MS, class C>
    ENL_FN static void load(const C& c, const FamilyData& d, long long, const Grp& g) {
#pragma unroll
        for (int s = 0; s < MS; ++s) {
            int row = s * Grp::G + g.lane;
            c.d.at(s, 0) = (row < M) ? d.d0[row] : 0.0;
            c.d.at(s, 1) = (row < M) ? d.d1[row] : 0.0;
        }
    }
    template <class Grp, int MS, class C>
    ENL_FN static void residuals(const C& c, const Grp& g, const double* x, double* out) {
#pragma unroll 1
        for (int s = 0; s < MS; ++s) {
            int row = s * Grp::G + g.lane;
            double t = c.t(s);
            double d2 = sub_rn(t, x[8]), d3 = sub_rn(t, x[9]), d4 = sub_rn(t, x[10]);
            double e1 = det_exp(mul_rn(-x[4], t));
            double e2 = det_exp(mul_rn(-x[5], mul_rn(d2, d2)));
            double e3 = det_exp(mul_rn(-x[6], mul_rn(d3, d3)));
            double e4 = det_exp(mul_rn(-x[7], mul_rn(d4, d4)));
            double mdl = add_rn(add_rn(add_rn(mul_rn(x[0], e1), mul_rn(x[1], e2)), mul_rn(x[2], e3)), mul_rn(x[3], e4));
            out[s] = (row < M) ? sub_rn(c.y(s), mdl) : 0.0;
        }
    }
    template <int MS, class C>
    ENL_FN static void constraints(const C&, const double*, double*) {}
    template <class Grp, int MS, class C>
    ENL_FN static void jac_residuals(const C& c, const Grp& g, const double* x, double* out) {
#pragma unroll 1
        for (int s = 0; s < MS; ++s) {
            int row = s * Grp::G + g.lane;
            bool ok = row < M;
            double t = c.t(s);
            double d2 = sub_rn(t, x[8]), d3 = sub_rn(t, x[9]), d4 = sub_rn(t, x[10]);
            double e1 = det_exp(mul_rn(-x[4], t));
            double e2 = det_exp(mul_rn(-x[5], mul_rn(d2, d2)));
            double e3 = det_exp(mul_rn(-x[6], mul_rn(d3, d3)));
            double e4 = det_exp(mul_rn(-x[7], mul_rn(d4, d4)));
            double* o = out + s * N;
            o[0] = ok ? -e1 : 0.0; o[1] = ok ? -e2 : 0.0; o[2] = ok ? -e3 : 0.0; o[3] = ok ? -e4 : 0.0;
            o[4] = ok ? mul_rn(mul_rn(x[0], t), e1) : 0.0;
            o[5] = ok ? mul_rn(mul_rn(x[1], mul_rn(d2, d2)), e2) : 0.0;
            o[6] = ok ? mul_rn(mul_rn(x[2], mul_rn(d3, d3)), e3) : 0.0;
            o[7] = ok ? mul_rn(mul_rn(x[3], mul_rn(d4, d4)), e4) : 0.0;
            o[8] = ok ? mul_rn(mul_rn(mul_rn(mul_rn(-x[1], e2), 2.0), x[5]), d2) : 0.0;
            o[9] = ok ? mul_rn(mul_rn(mul_rn(mul_rn(-x[2], e3), 2.0), x[6]), d3) : 0.0;
            o[10] = ok ? mul_rn(mul_rn(mul_rn(mul_rn(-x[3], e4), 2.0), x[7]), d4) : 0.0;
        }
    }
    template <int MS, class C>
    ENL_FN static void jac_constraints(const C&, const double*, double*) {}
};

// Chained Rosenbrock (test/problems/chained_rosenbrock.jl:8-53) with n = NN parameters:
// m = 2(n-1) residuals, q = n-2 nonlinear equalities, no bounds.
template <int NN>
struct FamChainedRosenbrock {
    static constexpr int N = NN, M = 2 * (NN - 1), Q = NN - 2, NI = 0, MAXB = 0;
    static constexpr bool HAS_ANALYTIC = true;
    static constexpr bool HAS_FAST_FD = false;
    static constexpr int NDCOLS = 0, NSCAL = 0;
    template <class DMt, class Vt>
    struct Ctx { DMt d; Vt s; };
    template <class Grp, int MS, class C>
    ENL_FN static void load(const C&, const FamilyData&, long long, const Grp&) {}
    template <class Grp, int MS, class C>
    ENL_FN static void residuals(const C&, const Grp& g, const double* x, double* out) {
#pragma unroll 1
        for (int s = 0; s < MS; ++s) {
            int row = s * Grp::G + g.lane;
            double v = 0.0;
            if (row < N - 1) v = mul_rn(10.0, sub_rn(mul_rn(x[row], x[row]), x[row + 1]));
            else if (row < M) v = sub_rn(x[row - (N - 1)], 1.0);
            out[s] = v;
        }
    }
    template <int MS, class C>
    ENL_FN static void constraints(const C&, const double* x, double* c) {
#pragma unroll 1
        for (int k = 0; k < Q; ++k) {
            double a = x[k], b = x[k + 1], cc = x[k + 2];
            double v = add_rn(mul_rn(3.0, mul_rn(mul_rn(b, b), b)), mul_rn(2.0, cc));
            v = sub_rn(v, 5.0);
            v = add_rn(v, mul_rn(sin(sub_rn(b, cc)), sin(add_rn(b, cc))));
            v = add_rn(v, mul_rn(4.0, b));
            v = sub_rn(v, mul_rn(a, exp(sub_rn(a, b))));
            c[k] = sub_rn(v, 3.0);
        }
    }
    template <class Grp, int MS, class C>
    ENL_FN static void jac_residuals(const C&, const Grp& g, const double* x, double* out) {
#pragma unroll 1
        for (int s = 0; s < MS; ++s) {
            int row = s * Grp::G + g.lane;
#pragma unroll 1
            for (int j = 0; j < N; ++j) {
                double v = 0.0;
                if (row < N - 1) { if (j == row) v = mul_rn(20.0, x[row]); else if (j == row + 1) v = -10.0; }
                else if (row < M) { if (j == row - (N - 1)) v = 1.0; }
                out[s * N + j] = v;
            }
        }
    }
    template <int MS, class C>
    ENL_FN static void jac_constraints(const C&, const double* x, double* A) {
#pragma unroll 1
        for (int i = 0; i < Q * N; ++i) A[i] = 0.0;
#pragma unroll 1
        for (int k = 0; k < Q; ++k) {
            double a = x[k], b = x[k + 1], cc = x[k + 2];
            double e = exp(sub_rn(a, b));
            double sm = sin(sub_rn(b, cc)), cm = cos(sub_rn(b, cc)), sp = sin(add_rn(b, cc)), cp = cos(add_rn(b, cc));
            A[k * N + k] = mul_rn(-add_rn(a, 1.0), e);
            double v = add_rn(mul_rn(9.0, mul_rn(b, b)), mul_rn(cm, sp));
            v = add_rn(v, mul_rn(sm, cp));
            v = add_rn(v, 4.0);
            A[k * N + k + 1] = add_rn(v, mul_rn(a, e));
            A[k * N + k + 2] = add_rn(sub_rn(2.0, mul_rn(cm, sp)), mul_rn(sm, cp));
        }
    }
};

// Chained Wood (test/problems/chained_wood.jl:4-35), n = NN (even, >= 8): m = 6(n/2-1), q = n-7.
// Polynomial only: residuals, constraints and Jacobians are bit-identical to the numpy oracle.
template <int NN>
struct FamChainedWood {
    static constexpr int N = NN, NB = NN / 2 - 1, M = 6 * NB, Q = NN - 7, NI = 0, MAXB = 0;
    static constexpr bool HAS_ANALYTIC = true;
    static constexpr bool HAS_FAST_FD = false;
    static constexpr int NDCOLS = 0, NSCAL = 0;
    template <class DMt, class Vt>
    struct Ctx { DMt d; Vt s; };
    template <class Grp, int MS, class C>
    ENL_FN static void load(const C&, const FamilyData&, long long, const Grp&) {}
    // row -> (block, i) ; o = 2i, e = 2i+1, o2 = 2i+2, e2 = 2i+3 (0-based, i = 0..NB-1)
    ENL_FN static double res_row(int row, const double* x) {
        const double s = sqrt_rn(10.0);
        int blk = row / NB, i = row % NB;
        double xo = x[2 * i], xe = x[2 * i + 1], xo2 = x[2 * i + 2], xe2 = x[2 * i + 3];
        switch (blk) {
            case 0: return mul_rn(10.0, sub_rn(mul_rn(xo, xo), xe));
            case 1: return sub_rn(xo, 1.0);
            case 2: return mul_rn(mul_rn(3.0, s), sub_rn(mul_rn(xo2, xo2), xe2));
            case 3: return sub_rn(xo2, 1.0);
            case 4: return mul_rn(s, sub_rn(add_rn(xe, xe2), 2.0));
            default: return mul_rn(sub_rn(xe, xe2), div_rn(1.0, s));
        }
    }
    template <class Grp, int MS, class C>
    ENL_FN static void residuals(const C&, const Grp& g, const double* x, double* out) {
#pragma unroll 1
        for (int sl = 0; sl < MS; ++sl) {
            int row = sl * Grp::G + g.lane;
            out[sl] = (row < M) ? res_row(row, x) : 0.0;
        }
    }
    template <int MS, class C>
    ENL_FN static void constraints(const C&, const double* x, double* c) {
#pragma unroll 1
        for (int k = 1; k <= Q; ++k) {
            double xk5 = x[k + 4];
            double acc = 0.0;
            int lo = (k - 5 > 1) ? k - 5 : 1;
#pragma unroll 1
            for (int ii = lo; ii <= k + 1; ++ii) acc = add_rn(acc, mul_rn(x[ii - 1], add_rn(1.0, x[ii - 1])));
            double v = mul_rn(add_rn(2.0, mul_rn(5.0, mul_rn(xk5, xk5))), xk5);
            c[k - 1] = add_rn(add_rn(v, 1.0), acc);
        }
    }
    template <class Grp, int MS, class C>
    ENL_FN static void jac_residuals(const C&, const Grp& g, const double* x, double* out) {
        const double s = sqrt_rn(10.0);
#pragma unroll 1
        for (int sl = 0; sl < MS; ++sl) {
            int row = sl * Grp::G + g.lane;
#pragma unroll 1
            for (int j = 0; j < N; ++j) out[sl * N + j] = 0.0;
            if (row >= M) continue;
            int blk = row / NB, i = row % NB;
            double* o = out + sl * N;
            switch (blk) {
                case 0: o[2 * i] = mul_rn(20.0, x[2 * i]); o[2 * i + 1] = -10.0; break;
                case 1: o[2 * i] = 1.0; break;
                case 2: o[2 * i + 2] = mul_rn(mul_rn(6.0, s), x[2 * i + 2]); o[2 * i + 3] = mul_rn(-3.0, s); break;
                case 3: o[2 * i + 2] = 1.0; break;
                case 4: o[2 * i + 1] = s; o[2 * i + 3] = s; break;
                default: o[2 * i + 1] = div_rn(1.0, s); o[2 * i + 3] = div_rn(-1.0, s); break;
            }
        }
    }
    template <int MS, class C>
    ENL_FN static void jac_constraints(const C&, const double* x, double* A) {
#pragma unroll 1
        for (int i = 0; i < Q * N; ++i) A[i] = 0.0;
#pragma unroll 1
        for (int k = 1; k <= Q; ++k) {
            A[(k - 1) * N + k + 4] = add_rn(A[(k - 1) * N + k + 4], add_rn(2.0, mul_rn(15.0, mul_rn(x[k + 4], x[k + 4]))));
            int lo = (k - 5 > 1) ? k - 5 : 1;
#pragma unroll 1
            for (int ii = lo; ii <= k + 1; ++ii)
                A[(k - 1) * N + ii - 1] = add_rn(A[(k - 1) * N + ii - 1], add_rn(1.0, mul_rn(2.0, x[ii - 1])));
        }
    }
};

}  // namespace enl

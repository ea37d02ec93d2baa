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
    static constexpr int N = 3, M = 3, Q = 0, NI = 1;
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
struct FamGaussPeaks {
    static constexpr int N = 6, M = 128, Q = 1, NI = 0;
    static constexpr bool HAS_ANALYTIC = true;
    static constexpr bool HAS_FAST_FD = true;
    static constexpr int NDCOLS = 2;   // y_i, t_i
    static constexpr int NSCAL = 1;    // S
    template <class DMt, class Vt>
    struct Ctx {
        DMt d;    // column 0 = y, column 1 = t (row-distributed shared memory)
        Vt s;     // s[0] = S
        ENL_INL double y(int sl) const { return d.at(sl, 0); }
        ENL_INL double t(int sl) const { return d.at(sl, 1); }
        ENL_INL double S() const { return s[0]; }
    };
    template <class Grp, int MS, class C>
    ENL_FN static void load(const C& c, const FamilyData& d, long long b, const Grp& g) {
#pragma unroll
        for (int s = 0; s < MS; ++s) {
            int row = s * Grp::G + g.lane;
            c.d.at(s, 0) = (row < M) ? d.d0[b * M + row] : 0.0;       // coalesced: lane l reads rows l, l+G, ...
            c.d.at(s, 1) = div_rn(mul_rn(10.0, (double)row), 127.0);
        }
        c.s[0] = d.d1[b];
    }
    ENL_INL static double peak(double b, double c, double t) {
        double d = sub_rn(t, c);
        return det_exp(mul_rn(-b, mul_rn(d, d)));
    }
    template <class Grp, int MS, class C>
    ENL_FN static void residuals(const C& c, const Grp& g, const double* x, double* out) {
#pragma unroll
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
#pragma unroll
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
#pragma unroll
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

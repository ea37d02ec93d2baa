// enl_large_family.h -- the "single-index" problem family of the large-Jacobian regime
// (BASELINE.json configs 4 and 5):
//     r_i(x) = det_tanh(w_i . x) - y_i,            i = 1..m      (W is m x n, row major)
//     J      = diag(1 - tanh^2(W x)) W
//     block constraints on groups of 4 parameters, k = 1..nb:
//         equalities    h_k(x) = sum_{j in block k} x_j^2 - rho_k            (config 4)
//         inequalities  g_k(x) = rho_k - sum_{j in block k} x_j^2  >= 0      (config 5)
//     followed by the finite bounds in the reference's order [x - x_low ; x_upp - x]
//     (src/cnls_model.jl:402-403, 416).
// This header holds the constraint side (l x n, tiny): plain C++ shared by the CUDA engine
// (enl_large.cu) and by the CPU test backend (oracle/hostport/largeport.cpp).  It replaces the
// reference's ConstraintsFunction closures (src/cnls_model.jl:27-62) for this family.
#pragma once
#include <math.h>

#include <vector>

namespace enl_large {

struct SingleIndexConstraints {
    int n = 0, nb = 0;
    bool ineq = false;
    std::vector<double> rho;
    std::vector<int> lo_idx, up_idx;
    std::vector<double> lo_val, up_val;

    int q() const { return ineq ? 0 : nb; }
    int l() const { return nb + (int)lo_idx.size() + (int)up_idx.size(); }

    void set_bounds(const double* x_low, const double* x_upp) {
        lo_idx.clear(); up_idx.clear(); lo_val.clear(); up_val.clear();
        for (int i = 0; i < n; ++i)
            if (x_low && isfinite(x_low[i])) { lo_idx.push_back(i); lo_val.push_back(x_low[i]); }
        for (int i = 0; i < n; ++i)
            if (x_upp && isfinite(x_upp[i])) { up_idx.push_back(i); up_val.push_back(x_upp[i]); }
    }
    void cons(const double* x, double* c) const {
        for (int k = 0; k < nb; ++k) {
            // numpy: (x[:4nb]**2).reshape(nb,4).sum(axis=1) -- pairwise over 4 entries = sequential
            double s = 0.0;
            for (int j = 0; j < 4; ++j) s += x[4 * k + j] * x[4 * k + j];
            c[k] = ineq ? (rho[k] - s) : (s - rho[k]);
        }
        int o = nb;
        for (size_t i = 0; i < lo_idx.size(); ++i) c[o++] = x[lo_idx[i]] - lo_val[i];
        for (size_t i = 0; i < up_idx.size(); ++i) c[o++] = up_val[i] - x[up_idx[i]];
    }
    // A: l x n column major
    // The sparsity pattern is fixed, so a buffer that was filled by the previous call (same address) only needs its
    // structural non-zeros rewritten; anything else (a fresh zero-initialised Mat included) gets the full fill.
    mutable const double* last_A = nullptr;
    void jac(const double* x, double* A) const {
        int L = l();
        if (A != last_A)
            for (size_t i = 0; i < (size_t)L * n; ++i) A[i] = 0.0;
        last_A = A;
        for (int k = 0; k < nb; ++k)
            for (int j = 0; j < 4; ++j) {
                double v = 2.0 * x[4 * k + j];
                A[(size_t)(4 * k + j) * L + k] = ineq ? -v : v;
            }
        int o = nb;
        for (size_t i = 0; i < lo_idx.size(); ++i) A[(size_t)lo_idx[i] * L + (o++)] = 1.0;
        for (size_t i = 0; i < up_idx.size(); ++i) A[(size_t)up_idx[i] * L + (o++)] = -1.0;
    }
};

}  // namespace enl_large

// =====================================================================================================================
// General problem families of the large regime ("row families"): the reference accepts ANY closure for residuals /
// constraints / Jacobians (src/cnls_model.jl:11-62, 345-359); here a family is a set of device functions over a point
// accessor X (x[j]), so that the same source serves the plain evaluation, the forward-difference Jacobian
// (cnls_model.jl:65-82: X adds delta_j to one coordinate) and the linesearch trial points (X = x + alpha p):
//     template <class X> double residual(long long i, int n, const X& x, const double* d0, const double* d1);
//     template <class X> double constraint(int k, int n, const X& x, const double* d0, const double* d1);   // eq first, then ineq
//     double jac_residual(long long i, int j, int n, const double* x, ...);      // optional (HAS_JAC)
//     double jac_constraint(int k, int j, int n, const double* x, ...);
// plus the sizes as functions of n.  Built in: chained Rosenbrock (test/problems/chained_rosenbrock.jl:8-53, the
// reference's own large test: n = 1000, m = 1998, 998 equalities); user source is bound to the same interface by
// enl_large_user.h (enlsipb200_large_compile_family).
// =====================================================================================================================
#if defined(__CUDACC__)
namespace enl_large {

struct XPlain {
    const double* x;
    __device__ __forceinline__ double operator[](int j) const { return x[j]; }
};
struct XPert {          // x + delta e_jp
    const double* x; int jp; double delta;
    __device__ __forceinline__ double operator[](int j) const { return j == jp ? __dadd_rn(x[j], delta) : x[j]; }
};
struct XStep {          // x + alpha p, rounded like the host's x[j] + alpha * p[j]
    const double* x; const double* p; double alpha;
    __device__ __forceinline__ double operator[](int j) const { return __dadd_rn(x[j], __dmul_rn(alpha, p[j])); }
};

struct LFamChainedRosenbrock {
    static constexpr bool HAS_JAC = true;
    __host__ __device__ static long long m_of(int n) { return 2LL * (n - 1); }
    __host__ __device__ static int q_of(int n) { return n - 2; }
    __host__ __device__ static int ni_of(int) { return 0; }
    // r_i = 10 (x_i^2 - x_{i+1}), i < n-1 ; r_{n-1+i} = x_i - 1        (chained_rosenbrock.jl:8-19)
    template <class X>
    __device__ static double residual(long long i, int n, const X& x, const double*, const double*) {
        if (i < n - 1) { const double a = x[(int)i]; return __dmul_rn(10.0, __dsub_rn(__dmul_rn(a, a), x[(int)i + 1])); }
        return __dsub_rn(x[(int)(i - (n - 1))], 1.0);
    }
    __device__ static double jac_residual(long long i, int j, int n, const double* x, const double*, const double*) {
        if (i < n - 1) return j == i ? __dmul_rn(20.0, x[j]) : (j == i + 1 ? -10.0 : 0.0);
        return j == i - (n - 1) ? 1.0 : 0.0;
    }
    // c_k = 3 b^3 + 2 c - 5 + sin(b - c) sin(b + c) + 4 b - a exp(a - b) - 3, (a, b, c) = x[k..k+2]   (:21-24)
    template <class X>
    __device__ static double constraint(int k, int, const X& x, const double*, const double*) {
        const double a = x[k], b = x[k + 1], c = x[k + 2];
        double v = __dadd_rn(__dmul_rn(3.0, __dmul_rn(__dmul_rn(b, b), b)), __dmul_rn(2.0, c));
        v = __dsub_rn(v, 5.0);
        v = __dadd_rn(v, __dmul_rn(sin(__dsub_rn(b, c)), sin(__dadd_rn(b, c))));
        v = __dadd_rn(v, __dmul_rn(4.0, b));
        v = __dsub_rn(v, __dmul_rn(a, exp(__dsub_rn(a, b))));
        return __dsub_rn(v, 3.0);
    }
    __device__ static double jac_constraint(int k, int j, int, const double* x, const double*, const double*) {   // (:26-51)
        if (j < k || j > k + 2) return 0.0;
        const double a = x[k], b = x[k + 1], c = x[k + 2];
        const double e = exp(a - b), sm = sin(b - c), cm = cos(b - c), sp = sin(b + c), cp = cos(b + c);
        if (j == k) return -(a + 1.0) * e;
        if (j == k + 1) return 9.0 * b * b + cm * sp + sm * cp + 4.0 + a * e;
        return 2.0 - cm * sp + sm * cp;
    }
};

}  // namespace enl_large
#endif

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

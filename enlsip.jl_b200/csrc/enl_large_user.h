// enl_large_user.h -- binds user-supplied device functions to the row-family interface of the large regime
// (enl_large_family.h).  Included only by the libraries enlsipb200_large_compile_family builds: the generated prelude
// defines the sizes and carries the user's source, which must provide, in namespace enl_user,
//     template <class X> __device__ double residual(long long i, int n, const X& x, const double* d0, const double* d1);
//     template <class X> __device__ double constraint(int k, int n, const X& x, const double* d0, const double* d1);
//         (k < nb_eq: equalities, then the inequalities c_k(x) >= 0 -- the order of cnls_model.jl:402-403)
// and, when has_jacobians was set,
//     __device__ double jac_residual(long long i, int j, int n, const double* x, const double* d0, const double* d1);   // d r_i / d x_j
//     __device__ double jac_constraint(int k, int j, int n, const double* x, const double* d0, const double* d1);
// X is a point accessor (x[j]): the same source serves plain evaluations, forward differences and linesearch trial
// points.  This is the large-regime counterpart of the reference's closure arguments (src/cnls_model.jl:345-359).
#pragma once
#include "enl_large_family.h"

namespace enl_large {

struct LFamUser {
    static constexpr bool HAS_JAC = ENL_LUSER_HAS_JAC != 0;
    __host__ __device__ static long long m_of(int) { return ENL_LUSER_M; }
    __host__ __device__ static int q_of(int) { return ENL_LUSER_Q; }
    __host__ __device__ static int ni_of(int) { return ENL_LUSER_NI; }
    template <class X>
    __device__ static double residual(long long i, int n, const X& x, const double* d0, const double* d1) {
        return enl_user::residual(i, n, x, d0, d1);
    }
    template <class X>
    __device__ static double constraint(int k, int n, const X& x, const double* d0, const double* d1) {
#if ENL_LUSER_Q + ENL_LUSER_NI > 0
        return enl_user::constraint(k, n, x, d0, d1);
#else
        return 0.0;
#endif
    }
    __device__ static double jac_residual(long long i, int j, int n, const double* x, const double* d0, const double* d1) {
#if ENL_LUSER_HAS_JAC
        return enl_user::jac_residual(i, j, n, x, d0, d1);
#else
        return 0.0;
#endif
    }
    __device__ static double jac_constraint(int k, int j, int n, const double* x, const double* d0, const double* d1) {
#if ENL_LUSER_HAS_JAC && (ENL_LUSER_Q + ENL_LUSER_NI > 0)
        return enl_user::jac_constraint(k, j, n, x, d0, d1);
#else
        return 0.0;
#endif
    }
};

}  // namespace enl_large

// enl_user_family.h -- problem family compiled at run time from user source.
//
// The reference's plugin surface is "any Julia closure": residuals(x), eq_constraints(x), ineq_constraints(x) and
// their optional Jacobians, wrapped at src/cnls_model.jl:11-62 and stacked by the constructor (cnls_model.jl:345-378,
// 410-496).  Closures cannot cross to the device, so the engine takes the same four functions as CUDA C++ source
// (enlsipb200_compile_family, include/enlsip_b200.h) and compiles THIS translation unit around them: the result is a
// shared library with the identical C ABI whose family id ENLSIPB200_FAMILY_USER is the user's problem.
//
// The generated prelude defines ENL_USER_N / _M / _Q / _NI / _STRIDE0 / _STRIDE1 / _HAS_JAC and is followed by the
// user's source, which must provide, in namespace enl_user,
//     __device__ double residual(int i, const double* x, const double* d0, const double* d1, const double* d2);
//     __device__ void   constraints(const double* x, const double* d0, const double* d1, const double* d2, double* c);
//         c[0 .. Q-1] equalities, c[Q .. Q+NI-1] inequalities (>= 0): the order of cnls_model.jl:402-403
// and, when ENL_USER_HAS_JAC,
//     __device__ void jac_residual(int i, const double* x, const double* d0, const double* d1, const double* d2, double* g /* [N] */);
//     __device__ void jac_constraints(const double* x, const double* d0, const double* d1, const double* d2, double* A /* [(Q+NI) x N] row major */);
// d0 / d1 are this problem's rows of data slots 0 / 1 (STRIDE0 / STRIDE1 doubles per problem), d2 is slot 2, shared by
// the whole batch.  Without Jacobians the engine differentiates by forward differences (cnls_model.jl:65-82).
#pragma once
#include "enl_base.h"

namespace enl {

struct FamUser {
    static constexpr int N = ENL_USER_N, M = ENL_USER_M, Q = ENL_USER_Q, NI = ENL_USER_NI, MAXB = 2 * N;
    static constexpr bool HAS_ANALYTIC = ENL_USER_HAS_JAC != 0;
    static constexpr bool HAS_FAST_FD = false;
    static constexpr int NDCOLS = 0;
    static constexpr int NSCAL = 3;    // addresses of the problem's d0 / d1 rows and of d2 (bit patterns in double slots)
    static constexpr int STRIDE0 = ENL_USER_STRIDE0, STRIDE1 = ENL_USER_STRIDE1;
    template <class DMt, class Vt>
    struct Ctx {
        DMt d;
        Vt s;
        ENL_INL const double* ptr(int k) const {
            double bits = s[k];
            const double* p;
            __builtin_memcpy(&p, &bits, sizeof(p));
            return p;
        }
    };
    template <class Grp, int MS, class C>
    ENL_FN static void load(const C& c, const FamilyData& d, long long b, const Grp&) {
        const double* p[3] = {d.d0 ? d.d0 + b * STRIDE0 : nullptr, d.d1 ? d.d1 + b * STRIDE1 : nullptr, d.d2};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double bits;
            __builtin_memcpy(&bits, &p[k], sizeof(bits));
            c.s[k] = bits;
        }
    }
    template <class Grp, int MS, class C>
    ENL_FN static void residuals(const C& c, const Grp& g, const double* x, double* out) {
        const double *d0 = c.ptr(0), *d1 = c.ptr(1), *d2 = c.ptr(2);
#pragma unroll 1
        for (int s = 0; s < MS; ++s) {
            const int row = s * Grp::G + g.lane;
            out[s] = (row < M) ? enl_user::residual(row, x, d0, d1, d2) : 0.0;
        }
    }
    template <int MS, class C>
    ENL_FN static void constraints(const C& c, const double* x, double* cc) {
        enl_user::constraints(x, c.ptr(0), c.ptr(1), c.ptr(2), cc);
    }
    template <class Grp, int MS, class C>
    ENL_FN static void jac_residuals(const C& c, const Grp& g, const double* x, double* out) {
#if ENL_USER_HAS_JAC
        const double *d0 = c.ptr(0), *d1 = c.ptr(1), *d2 = c.ptr(2);
#pragma unroll 1
        for (int s = 0; s < MS; ++s) {
            const int row = s * Grp::G + g.lane;
            double gr[N];
#pragma unroll
            for (int j = 0; j < N; ++j) gr[j] = 0.0;
            if (row < M) enl_user::jac_residual(row, x, d0, d1, d2, gr);
#pragma unroll
            for (int j = 0; j < N; ++j) out[s * N + j] = gr[j];
        }
#else
#pragma unroll 1
        for (int i = 0; i < MS * N; ++i) out[i] = 0.0;   // never reached: jac_mode is forced to forward differences
#endif
    }
    template <int MS, class C>
    ENL_FN static void jac_constraints(const C& c, const double* x, double* A) {
#if ENL_USER_HAS_JAC
        enl_user::jac_constraints(x, c.ptr(0), c.ptr(1), c.ptr(2), A);
#else
        for (int i = 0; i < (Q + NI) * N; ++i) A[i] = 0.0;
#endif
    }
};

}  // namespace enl

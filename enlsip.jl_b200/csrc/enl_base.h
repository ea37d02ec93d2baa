// enl_base.h -- execution-model layer of the batched ENLSIP engine.
//
// One "group" of G lanes (G = 1, 2, 4, ..., 32; a sub-warp or a whole warp) solves one
// problem.  Vectors of length m (residuals, Jacobian rows) are row-distributed over the lanes of
// the group; everything of size n, t, l is replicated: every lane of the group executes the same
// small-matrix code on the group's shared-memory state and writes identical values.
//
// Shared-memory layout (all indices compile time except the problem slot):
//   * small state   : element i of problem slot `pid`  at  base[i*PPC + pid]
//                     (a whole group reads one address -> broadcast; for G == 1 consecutive
//                      threads touch consecutive words -> conflict free)
//   * distributed   : element (slot s, column c) owned by thread `tid` at base[(c*MS+s)*NT + tid]
//                     row = s*G + lane  (consecutive lanes -> consecutive words)
//
// The same header compiles under g++ (ENL_HOST_BUILD) with G = NT = 1 as a scalar CPU build used
// by tests/ as a debugging aid and by bench.py as the compiled "port" CPU baseline.  It is never
// linked into the product library.
#pragma once
#include <math.h>
#include <stdint.h>
#include <type_traits>

#if defined(__CUDACC__)
#define ENL_FN __device__
#define ENL_INL __device__ __forceinline__
// Building blocks are deliberately NOT inlined: the first fully inlined build was 82 K SASS
// instructions (1.3 MB) and spent 48 of 54 cycles per issued instruction waiting for instruction
// fetch (ncu: smsp__average_warps_issue_stalled_no_instruction, profiles/r1_c3_v0_icache.md).
#define ENL_NOINL __device__ __noinline__
#else
#define ENL_FN
#define ENL_INL inline
#define ENL_NOINL
#endif

namespace enl {

// ---- correctly rounded, never-contracted arithmetic (problem functors, trial points) ----
#if defined(__CUDA_ARCH__)
ENL_INL double mul_rn(double a, double b) { return __dmul_rn(a, b); }
ENL_INL double add_rn(double a, double b) { return __dadd_rn(a, b); }
ENL_INL double sub_rn(double a, double b) { return __dsub_rn(a, b); }
ENL_INL double fma_rn(double a, double b, double c) { return fma(a, b, c); }
ENL_INL double div_rn(double a, double b) { return __ddiv_rn(a, b); }
ENL_INL double sqrt_rn(double a) { return __dsqrt_rn(a); }
#else
// host build is compiled with -ffp-contract=off
ENL_INL double mul_rn(double a, double b) { return a * b; }
ENL_INL double add_rn(double a, double b) { return a + b; }
ENL_INL double sub_rn(double a, double b) { return a - b; }
ENL_INL double fma_rn(double a, double b, double c) { return __builtin_fma(a, b, c); }
ENL_INL double div_rn(double a, double b) { return a / b; }
ENL_INL double sqrt_rn(double a) { return sqrt(a); }
#endif

// a / b with a shortcut for a == 0 (b finite, non-zero): the hardware-assisted double division falls
// into a ~100-instruction slow path for zero / denormal numerators, which are very common here
// (FD differences of residual rows far from a peak, inactive constraints in the linesearch model).
// The result is bit-identical to IEEE a / b for b > 0; for b < 0 the sign of a zero result may differ.
ENL_INL double div_z(double a, double b) { return (a == 0.0) ? a : div_rn(a, b); }

// Deterministic exp: the synthetic families are DEFINED through this function (same operation
// sequence as oracle/detmath.c, so CPU oracle and GPU engine see bit-identical residuals).
ENL_INL double pow2i(int k) {
    long long b = (long long)(k + 1023) << 52;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double d;
    __builtin_memcpy(&d, &b, 8);
    return d;
#endif
}

#if defined(ENL_COMPACT_CODE) && defined(__CUDA_ARCH__)
#define ENL_DET_EXP_ATTR ENL_NOINL      // one copy: the batched kernel is bound by instruction fetch, not by issue
#else
#define ENL_DET_EXP_ATTR ENL_INL
#endif
ENL_DET_EXP_ATTR double det_exp(double x) {
    // Branch-free restatement of oracle/detmath.c::det_exp (bit-identical results):
    //  * x < -745.2 -> 0 and x > 709.7827 -> +Inf fall out of the IEEE scaling below once x is clamped
    //    to [-800, 710] (the true exp under/overflows on the whole clamped-away range);
    //  * NaN is restored by the final select.
    const double LOG2E = 1.44269504088896338700e+00;
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    double xc = fmin(fmax(x, -800.0), 710.0);
    double kd = rint(mul_rn(xc, LOG2E));
    int k = (int)kd;
    double r = fma_rn(-kd, LN2_HI, xc);
    r = fma_rn(-kd, LN2_LO, r);
    double acc = 1.0 / 6227020800.0;
    acc = fma_rn(acc, r, 1.0 / 479001600.0);
    acc = fma_rn(acc, r, 1.0 / 39916800.0);
    acc = fma_rn(acc, r, 1.0 / 3628800.0);
    acc = fma_rn(acc, r, 1.0 / 362880.0);
    acc = fma_rn(acc, r, 1.0 / 40320.0);
    acc = fma_rn(acc, r, 1.0 / 5040.0);
    acc = fma_rn(acc, r, 1.0 / 720.0);
    acc = fma_rn(acc, r, 1.0 / 120.0);
    acc = fma_rn(acc, r, 1.0 / 24.0);
    acc = fma_rn(acc, r, 1.0 / 6.0);
    acc = fma_rn(acc, r, 0.5);
    acc = fma_rn(acc, r, 1.0);
    acc = fma_rn(acc, r, 1.0);
    int k1 = k >> 1;
    int k2 = k - k1;
    double res = mul_rn(mul_rn(acc, pow2i(k1)), pow2i(k2));
    if (x < -745.2) res = 0.0;          // same cut as the oracle (the natural result is 0 here as well)
    if (x > 709.782712893384) res = INFINITY;
    return (x != x) ? x : res;
}

ENL_INL double det_tanh(double z) {
    double e = det_exp(mul_rn(2.0, z));
    return sub_rn(1.0, div_rn(2.0, add_rn(e, 1.0)));
}

// ---- group policies ----------------------------------------------------------------------
#if defined(__CUDACC__)
template <int G_>
struct DevGroup {
    static constexpr int G = G_;
    unsigned mask;
    int lane;  // lane inside the group
    __device__ DevGroup() {
        int wl = threadIdx.x & 31;
        lane = wl & (G - 1);
        mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl & ~(G - 1)));
    }
    __device__ __forceinline__ void sync() const {
        if (G > 1) __syncwarp(mask);
    }
    __device__ __forceinline__ double sum(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, G);
        return v;
    }
    __device__ __forceinline__ double maxv(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o, G));
        return v;
    }
};
#endif

struct HostGroup {
    static constexpr int G = 1;
    int lane = 0;
    void sync() const {}
    double sum(double v) const { return v; }
    double maxv(double v) const { return v; }
};

// ---- strided views -----------------------------------------------------------------------
// On the device a view is an element offset into the CTA's dynamic shared memory, so that every
// access compiles to LDS/STS even across non-inlined calls; on the host it is a plain pointer.
#if defined(__CUDACC__)
extern __shared__ __align__(16) double enl_smem[];
using vref = int;
ENL_INL vref make_vref(double* p) { return (int)(p - enl_smem); }
ENL_INL vref make_vref(int* p) { return (int)(p - reinterpret_cast<int*>(enl_smem)); }
ENL_INL double& vderef_d(vref o) { return enl_smem[o]; }
ENL_INL int& vderef_i(vref o) { return reinterpret_cast<int*>(enl_smem)[o]; }
#else
using vref = void*;
ENL_INL vref make_vref(double* p) { return p; }
ENL_INL vref make_vref(int* p) { return p; }
#endif

template <int S>
struct SV {  // small-state vector of doubles, stride S
#if defined(__CUDACC__)
    int p;
    ENL_INL double& operator[](int i) const { return enl_smem[p + i * S]; }
#else
    double* p;
    ENL_INL double& operator[](int i) const { return p[i * S]; }
#endif
    ENL_INL SV off(int k) const { return SV{p + k * S}; }
};
template <int S>
struct SI {
#if defined(__CUDACC__)
    int p;
    ENL_INL int& operator[](int i) const { return reinterpret_cast<int*>(enl_smem)[p + i * S]; }
#else
    int* p;
    ENL_INL int& operator[](int i) const { return p[i * S]; }
#endif
};

// row-distributed m x ncols matrix (column major by slots).  The view itself is per-problem (the
// lane offset is added at access time), so it can live in the problem's shared-memory state.
template <int G, int MS, int NT>
struct DM {
#if defined(__CUDACC__)
    int grp;   // element offset of (slot 0, column 0, lane 0 of the group)
    ENL_INL static int lane() { return (int)(threadIdx.x & (G - 1)); }
    ENL_INL static int row_of(int s) { return s * G + lane(); }   // matrix row held in slot s by this lane
    ENL_INL double& at(int s, int c) const { return enl_smem[grp + lane() + (c * MS + s) * NT]; }
    ENL_INL double& row(int r, int c) const { return enl_smem[grp + (c * MS + r / G) * NT + (r % G)]; }
#else
    double* grp;
    ENL_INL static int row_of(int s) { return s * G; }
    ENL_INL double& at(int s, int c) const { return grp[(c * MS + s) * NT]; }
    ENL_INL double& row(int r, int c) const { return grp[(c * MS + r / G) * NT + (r % G)]; }
#endif
    ENL_INL DM cols(int c0) const { return DM{grp + c0 * MS * NT}; }
};

ENL_INL double sq(double v) { return v * v; }
ENL_INL int imin(int a, int b) { return a < b ? a : b; }
ENL_INL int imax(int a, int b) { return a > b ? a : b; }
ENL_INL double sign_of(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }  // Fortran SIGN(a,b)

// LAPACK dlapy2
ENL_INL double lapy2(double x, double y) {
    double xa = fabs(x), ya = fabs(y);
    double w = fmax(xa, ya), z = fmin(xa, ya);
    if (z == 0.0 || w > 1.7e308) return w;
    return w * sqrt(1.0 + (z / w) * (z / w));
}

constexpr double EPS = 2.220446049250313e-16;
constexpr double SQRT_EPS = 1.4901161193847656e-08;
constexpr double TOL3Z = 1.0536712127723509e-08;  // sqrt(dlamch('Epsilon')) = sqrt(2^-53)

// exit-code conventions shared with the oracle
constexpr int EXIT_WOULD_THROW = -99;   // the reference raises a Julia exception here
constexpr int EXIT_WOULD_HANG = -98;    // the reference loops forever here
constexpr int EXIT_CAPACITY = -97;      // initial working set larger than the engine's capacity

}  // namespace enl

/* Deterministic elementary functions for the ORACLE (test infrastructure only).
 *
 * The synthetic problem families of bench.py / tests (Gaussian peaks, single-index tanh) are
 * DEFINED in terms of det_exp below, not of the platform libm, so that the CPU oracle and the
 * CUDA engine evaluate bit-identical residuals: forward-difference Jacobians
 * (cnls_model.jl:65-82) amplify a 1-ulp libm discrepancy by 1/sqrt(eps) ~ 7e7, which would make
 * a 1e-10 parity bar meaningless.  Only IEEE-754 correctly rounded operations are used
 * (mul, add, fma, rint), in a fixed order; the CUDA engine restates the same sequence with
 * __dmul_rn / fma / rint (enlsip.jl_b200/csrc/detmath.cuh).
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -shared -fPIC detmath.c -o _build/libdetmath.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline double pow2i(int k) { /* 2^k for -1022 <= k <= 1023 */
    uint64_t b = (uint64_t)(k + 1023) << 52;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

double det_exp(double x) {
    static const double LOG2E = 1.44269504088896338700e+00;
    static const double LN2_HI = 6.93147180369123816490e-01; /* 0x3FE62E42FEE00000 */
    static const double LN2_LO = 1.90821492927058770002e-10;
    /* 1/k!, k = 0..13 */
    static const double C[14] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0,
                                 1.0 / 40320.0, 1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0,
                                 1.0 / 479001600.0, 1.0 / 6227020800.0};
    if (x != x) return x;
    if (x > 709.782712893384) return INFINITY;
    if (x < -745.2) return 0.0;
    double kd = rint(x * LOG2E);
    int k = (int)kd;
    double r = __builtin_fma(-kd, LN2_HI, x);
    r = __builtin_fma(-kd, LN2_LO, r);
    double acc = C[13];
    for (int i = 12; i >= 0; --i) acc = __builtin_fma(acc, r, C[i]);
    int k1 = k >> 1;
    int k2 = k - k1;
    return (acc * pow2i(k1)) * pow2i(k2);
}

/* tanh(z) := 1 - 2/(det_exp(2z) + 1)  (the definition used by the single-index family) */
double det_tanh(double z) {
    double e = det_exp(2.0 * z);
    return 1.0 - 2.0 / (e + 1.0);
}

void det_exp_vec(const double *x, double *y, long n) {
    for (long i = 0; i < n; ++i) y[i] = det_exp(x[i]);
}

void det_tanh_vec(const double *x, double *y, long n) {
    for (long i = 0; i < n; ++i) y[i] = det_tanh(x[i]);
}

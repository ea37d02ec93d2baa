// enl_dense.cuh -- device versions of the two O(n^3) primitives of the large regime's small-matrix stage, used when
// the compressed problem itself is large (BASELINE.json config 5: n = 4096, so [R_J | z] is 4097 x 4097):
//
//   * qr(M, ColumnNorm())  (EF:223, EF:700; LAPACK dgeqp3 semantics, restated as the unblocked dlaqp2: first-max
//     pivot, dlarfg with beta = -sign(alpha) dlapy2, partial-norm downdate with the tol3z recompute rule) --
//     the same restatement as QRP::factor in enl_large_host.h, one column step = two kernels, nothing returns to
//     the host until the factorisation is complete;
//   * M * Q  (J * F_A.Q, EF:219): reflectors applied from the right, two kernels per reflector.
//
// Both are BLAS-2: they stream the (L2-resident for n <= 4096) trailing matrix twice per column, i.e. they are
// HBM/L2-bandwidth bound by construction.  Column-major storage like the host Mat.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace enl_dense {

constexpr double D_TOL3Z = 1.0536712127723509e-08;   // dlaqp2: sqrt(dlamch('Epsilon')) = sqrt(2^-53)
constexpr int APPLY_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();               // protects sh against the previous use
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += sh[i];   // fixed order: deterministic, identical in every thread
    return t;
}

__device__ __forceinline__ double d_lapy2(double x, double y) {
    const double xa = fabs(x), ya = fabs(y), w = fmax(xa, ya), z = fmin(xa, ya);
    if (z == 0.0) return w;
    const double q = z / w;
    return w * sqrt(1.0 + q * q);
}

// initial column norms
__global__ void qrcp_norms_kernel(const double* __restrict__ f, int rows, int cols, double* vn1, double* vn2, int* jpvt) {
    __shared__ double sh[32];
    const int c = blockIdx.x;
    const double* cc = f + (size_t)c * rows;
    double s = 0.0;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) s = fma(cc[r], cc[r], s);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) { vn1[c] = vn2[c] = sqrt(s); jpvt[c] = c; }
}

// step i, part 1 (one CTA): pivot search, column swap, Householder vector of column i
__global__ void __launch_bounds__(1024) qrcp_pivot_house_kernel(double* __restrict__ f, int rows, int cols, int i,
                                                                 double* vn1, double* vn2, int* jpvt, double* tau) {
    __shared__ double sh[32];
    __shared__ double s_best[32];
    __shared__ int s_idx[32];
    __shared__ int s_pvt;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
    // first maximum of vn1[i..cols)
    double best = -1.0; int idx = cols;
    for (int j = i + tid; j < cols; j += blockDim.x) {
        const double v = vn1[j];
        if (v > best) { best = v; idx = j; }      // ascending j per thread: strict > keeps the first
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
    }
    if (lane == 0) { s_best[w] = best; s_idx[w] = idx; }
    __syncthreads();
    if (tid == 0) {
        double b = s_best[0]; int bi = s_idx[0];
        for (int k = 1; k < nw; ++k)
            if (s_best[k] > b || (s_best[k] == b && s_idx[k] < bi)) { b = s_best[k]; bi = s_idx[k]; }
        if (bi >= cols) bi = i;                   // all NaN / empty: keep the column (dlaqp2 would too)
        s_pvt = bi;
        if (bi != i) {
            const int tp = jpvt[bi]; jpvt[bi] = jpvt[i]; jpvt[i] = tp;
            vn1[bi] = vn1[i]; vn2[bi] = vn2[i];
        }
    }
    __syncthreads();
    const int pvt = s_pvt;
    double* ci = f + (size_t)i * rows;
    if (pvt != i) {
        double* cp = f + (size_t)pvt * rows;
        for (int r = tid; r < rows; r += blockDim.x) { const double a = ci[r]; ci[r] = cp[r]; cp[r] = a; }
    }
    __syncthreads();
    double tau_i = 0.0;
    if (i < rows - 1) {
        double s = 0.0;
        for (int r = i + 1 + tid; r < rows; r += blockDim.x) s = fma(ci[r], ci[r], s);
        s = block_sum(s, sh);
        const double xn = sqrt(s);
        if (xn != 0.0) {
            const double alpha = ci[i];
            const double beta = -copysign(d_lapy2(alpha, xn), alpha);
            tau_i = (beta - alpha) / beta;
            const double sc = 1.0 / (alpha - beta);
            __syncthreads();                      // everyone has read alpha
            for (int r = i + 1 + tid; r < rows; r += blockDim.x) ci[r] *= sc;
            if (tid == 0) ci[i] = beta;
        }
    }
    if (tid == 0) tau[i] = tau_i;
}

// step i, part 2 (one CTA per trailing column): apply H_i, then the dlaqp2 partial-norm downdate of that column
__global__ void __launch_bounds__(APPLY_THREADS) qrcp_apply_kernel(double* __restrict__ f, int rows, int cols, int i,
                                                                    double* vn1, double* vn2, const double* __restrict__ tau) {
    __shared__ double sh[32];
    __shared__ int s_recompute;
    const int c = i + 1 + blockIdx.x;
    const double* v = f + (size_t)i * rows;
    double* cc = f + (size_t)c * rows;
    const double tau_i = tau[i];
    if (tau_i != 0.0) {
        double s = 0.0;
        for (int r = i + 1 + threadIdx.x; r < rows; r += blockDim.x) s = fma(v[r], cc[r], s);
        s = block_sum(s, sh);
        const double wv = (cc[i] + s) * tau_i;
        __syncthreads();
        for (int r = i + 1 + threadIdx.x; r < rows; r += blockDim.x) cc[r] = fma(-wv, v[r], cc[r]);
        if (threadIdx.x == 0) cc[i] -= wv;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int rec = 0;
        const double v1 = vn1[c];
        if (v1 != 0.0) {
            const double tq = fabs(cc[i]) / v1;
            const double temp = fmax(1.0 - tq * tq, 0.0);
            const double rq = v1 / vn2[c];
            const double temp2 = temp * (rq * rq);
            if (temp2 <= D_TOL3Z) {
                if (i < rows - 1) rec = 1;
                else vn1[c] = vn2[c] = 0.0;
            } else {
                vn1[c] = v1 * sqrt(temp);
            }
        }
        s_recompute = rec;
    }
    __syncthreads();
    if (s_recompute) {
        double s = 0.0;
        for (int r = i + 1 + threadIdx.x; r < rows; r += blockDim.x) s = fma(cc[r], cc[r], s);
        s = block_sum(s, sh);
        if (threadIdx.x == 0) vn1[c] = vn2[c] = sqrt(s);
    }
}

// f: rows x cols column major on the device, factored in place; tau [min(rows, cols)], jpvt [cols] (0-based);
// vn: 2 * cols doubles of scratch.  Returns the number of kernels launched.
inline int qrcp_device(double* f, int rows, int cols, double* tau, int* jpvt, double* vn, cudaStream_t st) {
    const int k = rows < cols ? rows : cols;
    double *vn1 = vn, *vn2 = vn + cols;
    int launches = 1;
    qrcp_norms_kernel<<<cols, 256, 0, st>>>(f, rows, cols, vn1, vn2, jpvt);
    for (int i = 0; i < k; ++i) {
        qrcp_pivot_house_kernel<<<1, 1024, 0, st>>>(f, rows, cols, i, vn1, vn2, jpvt, tau);
        ++launches;
        if (i < cols - 1) {
            qrcp_apply_kernel<<<cols - i - 1, APPLY_THREADS, 0, st>>>(f, rows, cols, i, vn1, vn2, tau);
            ++launches;
        }
    }
    return launches;
}

// ---- M <- M * Q : reflector i = [0.., 1, f[i+1.., i]] (length nq) acts on the columns i.. of M (mr x nq) ----
constexpr int MULQ_CHUNKS = 64;

__global__ void __launch_bounds__(256) mulq_dot_kernel(const double* __restrict__ M, int mr, int nq, const double* __restrict__ f,
                                                        int frows, int i, double* __restrict__ wpart) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int span = nq - i, per = (span + MULQ_CHUNKS - 1) / MULQ_CHUNKS;
    const int c0 = i + blockIdx.y * per, c1 = min(nq, c0 + per);
    if (r >= mr) return;
    const double* v = f + (size_t)i * frows;
    double s = 0.0;
    int c = c0;
    for (; c + 4 <= c1; c += 4) {      // four independent loads in flight, summed in column order
        const double m0 = M[(size_t)c * mr + r], m1 = M[(size_t)(c + 1) * mr + r];
        const double m2 = M[(size_t)(c + 2) * mr + r], m3 = M[(size_t)(c + 3) * mr + r];
        s = fma(m0, (c == i) ? 1.0 : v[c], s);
        s = fma(m1, v[c + 1], s);
        s = fma(m2, v[c + 2], s);
        s = fma(m3, v[c + 3], s);
    }
    for (; c < c1; ++c) s = fma(M[(size_t)c * mr + r], (c == i) ? 1.0 : v[c], s);
    wpart[(size_t)blockIdx.y * mr + r] = s;
}
__global__ void __launch_bounds__(256) mulq_upd_kernel(double* __restrict__ M, int mr, int nq, const double* __restrict__ f,
                                                        int frows, int i, const double* __restrict__ tau,
                                                        const double* __restrict__ wpart) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int span = nq - i, per = (span + MULQ_CHUNKS - 1) / MULQ_CHUNKS;
    const int c0 = i + blockIdx.y * per, c1 = min(nq, c0 + per);
    if (r >= mr) return;
    const double ti = tau[i];
    if (ti == 0.0) return;
    double wsum = 0.0;
#pragma unroll
    for (int k = 0; k < MULQ_CHUNKS; ++k) wsum += wpart[(size_t)k * mr + r];
    wsum *= ti;
    const double* v = f + (size_t)i * frows;
    int c = c0;
    for (; c + 4 <= c1; c += 4) {
        const double m0 = M[(size_t)c * mr + r], m1 = M[(size_t)(c + 1) * mr + r];
        const double m2 = M[(size_t)(c + 2) * mr + r], m3 = M[(size_t)(c + 3) * mr + r];
        M[(size_t)c * mr + r] = fma(-wsum, (c == i) ? 1.0 : v[c], m0);
        M[(size_t)(c + 1) * mr + r] = fma(-wsum, v[c + 1], m1);
        M[(size_t)(c + 2) * mr + r] = fma(-wsum, v[c + 2], m2);
        M[(size_t)(c + 3) * mr + r] = fma(-wsum, v[c + 3], m3);
    }
    for (; c < c1; ++c) M[(size_t)c * mr + r] = fma(-wsum, (c == i) ? 1.0 : v[c], M[(size_t)c * mr + r]);
}

// M: mr x nq column major; f: frows (= nq) x k factors; wpart: MULQ_CHUNKS * mr doubles
inline int mulq_device(double* M, int mr, int nq, const double* f, int frows, int k, const double* tau, double* wpart,
                       cudaStream_t st) {
    dim3 grid((mr + 255) / 256, MULQ_CHUNKS);
    for (int i = 0; i < k; ++i) {
        mulq_dot_kernel<<<grid, 256, 0, st>>>(M, mr, nq, f, frows, i, wpart);
        mulq_upd_kernel<<<grid, 256, 0, st>>>(M, mr, nq, f, frows, i, tau, wpart);
    }
    return 2 * k;
}

}  // namespace enl_dense

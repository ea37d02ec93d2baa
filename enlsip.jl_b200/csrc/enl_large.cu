// enl_large.cu -- large-Jacobian regime of the B200 ENLSIP engine (BASELINE.json config 4: one problem,
// m = 4M residuals, n = 256, row-sharded over the GPUs of one box; config 5: n = 4096, m = 16384 on one GPU).
//
// Per Gauss-Newton iteration (reference: one pass of the `while exit_code == 0` loop, EF:2776-2878):
//   li_build_kernel      r = tanh(Wx) - y, J = diag(1 - tanh^2) W         -> [J | r] in HBM   (new_point!, EF:34-52)
//   tsqr_factor          [J | r] -> R (n+1 x n+1)   (enl_tsqr.cuh: Householder panels + DMMA trailing updates)
//   NCCL all-gather      of the per-GPU R factors, stacked and re-factored identically on every GPU
//   r_to_colmajor        R -> column-major [J~ | r~], which stays device resident for enl_dense.cuh
//   host small stage     the ENLSIP iteration on the compressed problem (enl_large_host.h), replicated per rank;
//                        its O(n^3) primitives (QRCP, J~ Q1) run on the device when n >= 384 (enl_dense.cuh)
//   li_dir_kernel        v = W p, Jp = s .* v, {r.r, r.Jp, Jp.Jp}                              (EF:2222-2224)
//   li_ls_kernel         per trial step: ||r(x + a p)||^2 and the linesearch model dots        (EF:1307-1340, 1665-1689)
//   NCCL all-reduce      of those few doubles
// sm_100a only; no CPU fallback (every entry point fails with ENLSIPB200_ENOGPU without a device).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "../../include/enlsip_b200.h"
#include "enl_base.h"
#include "enl_large_family.h"
#include "enl_large_host.h"
#include "enl_tsqr.cuh"
#include "enl_dense.cuh"

using namespace enl_large;

namespace {

thread_local std::string g_lerr;
int lfail(int code, const std::string& msg) { g_lerr = msg; return code; }
#define LCU(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return lfail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? ENLSIPB200_ENOGPU \
                                                                                      : ENLSIPB200_ECUDA, \
                         std::string(#call) + ": " + cudaGetErrorString(e_));                            \
    } while (0)

// ---- NCCL, bound at run time (the torch-bundled libnccl.so.2 is already mapped in a torch process) ----
struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;
    typedef void* Comm;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string err;
    bool load() {
        if (ok) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        void* h = nullptr;
        for (const char* nm : names) {
            h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { err = std::string("dlopen libnccl: ") + dlerror(); return false; }
        auto sym = [&](const char* s) { return dlsym(h, s); };
        GetUniqueId = (int (*)(UniqueId*))sym("ncclGetUniqueId");
        CommInitRank = (int (*)(Comm*, int, UniqueId, int))sym("ncclCommInitRank");
        CommDestroy = (int (*)(Comm))sym("ncclCommDestroy");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, Comm, cudaStream_t))sym("ncclAllReduce");
        AllGather = (int (*)(const void*, void*, size_t, int, Comm, cudaStream_t))sym("ncclAllGather");
        GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
        ok = GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather && GetErrorString;
        if (!ok) err = "libnccl is missing a required symbol";
        return ok;
    }
};
NcclApi g_nccl;
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

// ---------------------------------------------------------------------------------------------
// elementwise kernels of the single-index family
// ---------------------------------------------------------------------------------------------
constexpr int LI_PARTS = 592;   // 148 SMs x 4 CTAs; fixed so that reductions are deterministic

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per row: u = w_i . x, r = det_tanh(u) - y, s = 1 - tanh^2; row i of [s W | r | 0..] into A.
// NC > 0: the warp also accumulates its share of J'r (lane owns columns 2 lane + 64 k, +1; NC = ceil(n / 64) register
// pairs) and of r'r; the CTA folds its 8 warps in a fixed order into gpart[blockIdx][0..n] ([n] = r'r).  The gradient is
// what the termination test of the new point reads (EF:2838, 2411), so that the factorisation can wait until the
// iteration is known to continue.  NC == 0 (n > 64 * 8): no gradient here, li_grad_kernel does it.
template <int NC>
__global__ void __launch_bounds__(256) li_build_kernel(const double* __restrict__ W, const double* __restrict__ y,
                                                       const double* __restrict__ x, long long m, int n, int ld,
                                                       double* __restrict__ A, double* __restrict__ u,
                                                       double* __restrict__ r, double* __restrict__ s,
                                                       double* __restrict__ gpart) {
    extern __shared__ double xs[];   // x [n]; then, NC > 0: 8 x (n + 1) partial gradients
    for (int j = threadIdx.x; j < n; j += blockDim.x) xs[j] = x[j];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    double ga[NC > 0 ? 2 * NC : 1];
    double rsq = 0.0;
#pragma unroll
    for (int k = 0; k < (NC > 0 ? 2 * NC : 1); ++k) ga[k] = 0.0;
    for (long long i = warp0; i < m; i += nwarps) {
        const double* wr = W + i * n;
        double acc = 0.0;
        for (int c = 2 * lane; c < n; c += 64) {
            double2 wv = *reinterpret_cast<const double2*>(wr + c);
            acc = fma(wv.x, xs[c], acc);
            acc = fma(wv.y, xs[c + 1], acc);
        }
        const double uu = warp_sum(acc);
        const double th = enl::det_tanh(uu);
        const double rr = __dsub_rn(th, y[i]);
        const double ss = __dsub_rn(1.0, __dmul_rn(th, th));
        double* ar = A + i * ld;
        if (NC > 0) {
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const int c = 2 * lane + 64 * k;
                if (c < n) {
                    double2 wv = *reinterpret_cast<const double2*>(wr + c);
                    const double j0 = __dmul_rn(ss, wv.x), j1 = __dmul_rn(ss, wv.y);
                    *reinterpret_cast<double2*>(ar + c) = make_double2(j0, j1);
                    ga[2 * k] = fma(j0, rr, ga[2 * k]);
                    ga[2 * k + 1] = fma(j1, rr, ga[2 * k + 1]);
                }
            }
            rsq = fma(rr, rr, rsq);
        } else {
            for (int c = 2 * lane; c < n; c += 64) {
                double2 wv = *reinterpret_cast<const double2*>(wr + c);
                *reinterpret_cast<double2*>(ar + c) = make_double2(__dmul_rn(ss, wv.x), __dmul_rn(ss, wv.y));
            }
        }
        if (lane < ld - n) ar[n + lane] = (lane == 0) ? rr : 0.0;
        if (lane == 0) { u[i] = uu; r[i] = rr; s[i] = ss; }
    }
    if (NC > 0) {
        double* gs = xs + n;
        const int w = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int c = 2 * lane + 64 * k;
            if (c < n) { gs[w * (n + 1) + c] = ga[2 * k]; gs[w * (n + 1) + c + 1] = ga[2 * k + 1]; }
        }
        if (lane == 0) gs[w * (n + 1) + n] = rsq;
        __syncthreads();
        for (int j = threadIdx.x; j <= n; j += blockDim.x) {
            double t = 0.0;
            for (int ww = 0; ww < 8; ++ww) t += gs[ww * (n + 1) + j];
            gpart[(size_t)blockIdx.x * (n + 1) + j] = t;
        }
    }
}

// generic J'r for wide problems (n > 512): gpart[b][j] = sum over the rows of CTA b of A[i][j] * A[i][n]  (A = [J | r]),
// [n] = r'r; thread = column
__global__ void __launch_bounds__(256) li_grad_kernel(const double* __restrict__ A, int ld, long long m, int n,
                                                      double* __restrict__ gpart) {
    const long long per = (m + gridDim.x - 1) / gridDim.x;
    const long long i0 = per * blockIdx.x, i1 = (i0 + per < m) ? i0 + per : m;
    for (int j = threadIdx.x; j <= n; j += blockDim.x) {
        double t = 0.0;
        for (long long i = i0; i < i1; ++i) t = fma(A[i * ld + j], A[i * ld + n], t);
        gpart[(size_t)blockIdx.x * (n + 1) + j] = t;
    }
}

// out[j] = sum_b gpart[b][j] in a fixed order: one warp per column (lane l sums b = l, l + 32, ...; then a butterfly)
__global__ void __launch_bounds__(256) li_grad_finish_kernel(const double* __restrict__ gpart, int nparts, int n,
                                                             double* __restrict__ out) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j > n) return;
    double t = 0.0;
    for (int b = lane; b < nparts; b += 32) t += gpart[(size_t)b * (n + 1) + j];
    t = warp_sum(t);
    if (lane == 0) out[j] = t;
}

__device__ __forceinline__ void block_reduce4(double (&v)[4], double* part) {
    __shared__ double sh[8][4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 4; ++k) sh[w][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) t += sh[ww][threadIdx.x];
        part[blockIdx.x * 4 + threadIdx.x] = t;
    }
}

// v = W p (one warp per row), Jp = s .* v; partial sums {r.r, r.Jp, Jp.Jp, 0}
__global__ void __launch_bounds__(256) li_dir_kernel(const double* __restrict__ W, const double* __restrict__ p,
                                                     const double* __restrict__ r, const double* __restrict__ s,
                                                     long long m, int n, double* __restrict__ v,
                                                     double* __restrict__ Jp, double* __restrict__ part) {
    extern __shared__ double xs[];
    for (int j = threadIdx.x; j < n; j += blockDim.x) xs[j] = p[j];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    double sums[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long i = warp0; i < m; i += nwarps) {
        const double* wr = W + i * n;
        double acc = 0.0;
        for (int c = 2 * lane; c < n; c += 64) {
            double2 wv = *reinterpret_cast<const double2*>(wr + c);
            acc = fma(wv.x, xs[c], acc);
            acc = fma(wv.y, xs[c + 1], acc);
        }
        const double vv = warp_sum(acc);
        if (lane == 0) {
            const double jp = s[i] * vv, ri = r[i];
            v[i] = vv; Jp[i] = jp;
            sums[0] = fma(ri, ri, sums[0]); sums[1] = fma(ri, jp, sums[1]); sums[2] = fma(jp, jp, sums[2]);
        }
    }
    block_reduce4(sums, part);
}

// trial point x + alpha p: r_a = det_tanh(u + alpha v) - y;  partial sums {r_a.r_a, r.v2, Jp.v2, v2.v2}
// with v2 = ((r_a - r)/alpha - Jp)/alpha  (coefficients_linesearch!, EF:1687)
__global__ void __launch_bounds__(256) li_ls_kernel(const double* __restrict__ u, const double* __restrict__ v,
                                                    const double* __restrict__ y, const double* __restrict__ r,
                                                    const double* __restrict__ Jp, long long m, double alpha,
                                                    int with_coeffs, double* __restrict__ part) {
    double sums[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const double ra = __dsub_rn(enl::det_tanh(fma(alpha, v[i], u[i])), y[i]);
        sums[0] = fma(ra, ra, sums[0]);
        if (with_coeffs) {
            const double ri = r[i], jp = Jp[i];
            const double v2 = (__ddiv_rn(__dsub_rn(ra, ri), alpha) - jp) / alpha;
            sums[1] = fma(ri, v2, sums[1]); sums[2] = fma(jp, v2, sums[2]); sums[3] = fma(v2, v2, sums[3]);
        }
    }
    block_reduce4(sums, part);
}

__global__ void li_finish_kernel(const double* __restrict__ part, int nparts, double* __restrict__ out) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;   // 4 warps, one per quantity
    double t = 0.0;
    for (int i = lane; i < nparts; i += 32) t += part[i * 4 + k];
    t = warp_sum(t);
    if (lane == 0) out[k] = t;
}

// R (row major, leading dimension ld, upper triangle valid) -> column-major nc x nc with explicit zeros below the
// diagonal: the layout of the compressed problem [J~ | r~] the host driver and enl_dense.cuh work on
__global__ void r_to_colmajor_kernel(const double* __restrict__ R, int ld, int nc, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < nc && c < nc && r <= c) ? R[(size_t)r * ld + c] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < nc && c < nc) out[(size_t)c * nc + r] = tile[threadIdx.x][i];
    }
}

// ---------------------------------------------------------------------------------------------
// device implementation of LargeOps
// ---------------------------------------------------------------------------------------------
// Work below this many matrix entries stays on the host (launch latency of 2 kernels per column dominates)
constexpr long long DENSE_ACCEL_MIN = 384LL * 384LL;

struct LargeHandle : LargeOps, DenseAccel {
    int device = 0;
    // ---- DenseAccel: device QRCP / M*Q of the compressed problem when it is large (config 5) ----
    double *dq_f = nullptr, *dq_m = nullptr, *dq_small = nullptr;
    int* dq_p = nullptr;
    size_t dq_f_cap = 0, dq_m_cap = 0, dq_small_cap = 0, dq_p_cap = 0;
    long long n_dev_qrcp = 0, n_dev_mulq = 0;
    double ms_dense = 0;
    bool grow(double** p, size_t* cap, size_t want) {
        if (*cap >= want) return true;
        if (*p) cudaFree(*p);
        *p = nullptr; *cap = 0;
        if (cudaMalloc(p, sizeof(double) * want) != cudaSuccess) return false;
        *cap = want;
        return true;
    }
    bool qrcp(double* f, int rows, int cols, double* tau, int* jpvt) override {
        if ((long long)rows * cols < DENSE_ACCEL_MIN) return false;
        auto t0 = std::chrono::steady_clock::now();
        const int k = rows < cols ? rows : cols;
        if (cudaSetDevice(device) != cudaSuccess) return false;
        if (!grow(&dq_f, &dq_f_cap, (size_t)rows * cols)) return false;
        if (!grow(&dq_small, &dq_small_cap, (size_t)3 * cols + enl_dense::MULQ_CHUNKS * (size_t)(rows + 1))) return false;
        if (dq_p_cap < (size_t)cols) {
            if (dq_p) cudaFree(dq_p);
            dq_p = nullptr; dq_p_cap = 0;
            if (cudaMalloc(&dq_p, sizeof(int) * cols) != cudaSuccess) return false;
            dq_p_cap = cols;
        }
        double* dtau = dq_small;            // [cols]
        double* dvn = dq_small + cols;      // [2 cols]
        bool ok = cudaMemcpyAsync(dq_f, f, sizeof(double) * (size_t)rows * cols, cudaMemcpyHostToDevice, st) == cudaSuccess;
        launches += enl_dense::qrcp_device(dq_f, rows, cols, dtau, dq_p, dvn, st);
        ok = ok && cudaMemcpyAsync(f, dq_f, sizeof(double) * (size_t)rows * cols, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(tau, dtau, sizeof(double) * k, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(jpvt, dq_p, sizeof(int) * cols, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
        if (!ok) throw std::runtime_error("device QRCP failed");
        ++n_dev_qrcp;
        ms_dense += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return true;
    }
    bool jq1(const double* fA, int nq, int k, const double* tauA, double* JQ1_host, int mr, int ncols_host) override {
        if ((long long)mr * nq < DENSE_ACCEL_MIN || nq != n || mr != n + 1) return false;
        auto t0 = std::chrono::steady_clock::now();
        if (cudaSetDevice(device) != cudaSuccess) return false;
        if (!grow(&dq_m, &dq_m_cap, (size_t)mr * nq)) return false;
        if (k > 0 && !grow(&dq_f, &dq_f_cap, (size_t)nq * k)) return false;
        if (!grow(&dq_small, &dq_small_cap, (size_t)3 * (k + 1) + enl_dense::MULQ_CHUNKS * (size_t)(mr + 1))) return false;
        double* dtau = dq_small;
        double* dw = dq_small + 3 * (size_t)(k + 1);
        bool ok = cudaMemcpyAsync(dq_m, dJc, sizeof(double) * (size_t)mr * nq, cudaMemcpyDeviceToDevice, st) == cudaSuccess;
        if (k > 0) {
            ok = ok && cudaMemcpyAsync(dq_f, fA, sizeof(double) * (size_t)nq * k, cudaMemcpyHostToDevice, st) == cudaSuccess;
            ok = ok && cudaMemcpyAsync(dtau, tauA, sizeof(double) * k, cudaMemcpyHostToDevice, st) == cudaSuccess;
            launches += enl_dense::mulq_device(dq_m, mr, nq, dq_f, nq, k, dtau, dw, st);
        }
        if (ncols_host > 0)
            ok = ok && cudaMemcpyAsync(JQ1_host, dq_m, sizeof(double) * (size_t)mr * ncols_host, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
        if (!ok) throw std::runtime_error("device J*Q1 failed");
        jq1_resident = true;
        ++n_dev_mulq;
        ms_dense += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return true;
    }
    bool qrcp_tail(int c0, int rows, int cols, double* f_host, double* tau, int* jpvt) override {
        if (!jq1_resident || rows != n + 1 || c0 + cols != n) return false;
        auto t0 = std::chrono::steady_clock::now();
        const int k = rows < cols ? rows : cols;
        if (cudaSetDevice(device) != cudaSuccess) return false;
        if (!grow(&dq_f, &dq_f_cap, (size_t)rows * cols)) return false;
        if (!grow(&dq_small, &dq_small_cap, (size_t)3 * cols + enl_dense::MULQ_CHUNKS * (size_t)(rows + 1))) return false;
        if (dq_p_cap < (size_t)cols) {
            if (dq_p) cudaFree(dq_p);
            dq_p = nullptr; dq_p_cap = 0;
            if (cudaMalloc(&dq_p, sizeof(int) * cols) != cudaSuccess) return false;
            dq_p_cap = cols;
        }
        double* dtau = dq_small;
        double* dvn = dq_small + cols;
        bool ok = cudaMemcpyAsync(dq_f, dq_m + (size_t)c0 * rows, sizeof(double) * (size_t)rows * cols, cudaMemcpyDeviceToDevice, st) == cudaSuccess;
        launches += enl_dense::qrcp_device(dq_f, rows, cols, dtau, dq_p, dvn, st);
        ok = ok && cudaMemcpyAsync(f_host, dq_f, sizeof(double) * (size_t)rows * cols, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(tau, dtau, sizeof(double) * k, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(jpvt, dq_p, sizeof(int) * cols, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
        if (!ok) throw std::runtime_error("device QRCP (tail) failed");
        ++n_dev_qrcp;
        ms_dense += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return true;
    }
    bool mul_Q(const double* f, int nq, int k, const double* tau, double* M, int mr) override {
        if ((long long)mr * nq < DENSE_ACCEL_MIN || k < 16) return false;
        auto t0 = std::chrono::steady_clock::now();
        if (cudaSetDevice(device) != cudaSuccess) return false;
        if (!grow(&dq_f, &dq_f_cap, (size_t)nq * k)) return false;
        if (!grow(&dq_m, &dq_m_cap, (size_t)mr * nq)) return false;
        if (!grow(&dq_small, &dq_small_cap, (size_t)3 * k + enl_dense::MULQ_CHUNKS * (size_t)(mr + 1))) return false;
        jq1_resident = false;                     // dq_m is overwritten
        double* dtau = dq_small;                  // [k]
        double* dw = dq_small + 3 * (size_t)k;    // [16 mr]
        bool ok = cudaMemcpyAsync(dq_f, f, sizeof(double) * (size_t)nq * k, cudaMemcpyHostToDevice, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(dtau, tau, sizeof(double) * k, cudaMemcpyHostToDevice, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(dq_m, M, sizeof(double) * (size_t)mr * nq, cudaMemcpyHostToDevice, st) == cudaSuccess;
        launches += enl_dense::mulq_device(dq_m, mr, nq, dq_f, nq, k, dtau, dw, st);
        ok = ok && cudaMemcpyAsync(M, dq_m, sizeof(double) * (size_t)mr * nq, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
        if (!ok) throw std::runtime_error("device M*Q failed");
        ++n_dev_mulq;
        ms_dense += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return true;
    }

    long long m_local = 0, rows_pad = 0;
    int ld = 0, rr_rows = 0;
    SingleIndexConstraints sc;
    cudaStream_t st = nullptr;
    // device data
    const double* dW = nullptr;
    const double* dy = nullptr;
    double *ownW = nullptr, *owny = nullptr;
    long long own_count[2] = {0, 0};
    double *dA = nullptr, *du = nullptr, *dr = nullptr, *ds = nullptr, *dv = nullptr, *dJp = nullptr;
    double *dx = nullptr, *dp = nullptr, *dT = nullptr, *dpart = nullptr, *dout = nullptr;
    double *dR = nullptr, *dStack = nullptr, *dR2 = nullptr;
    double* dJc = nullptr;          // [J~ | r~] column major (n+1) x (n+1): stays resident for enl_dense.cuh
    bool jq1_resident = false;      // dq_m holds J~ * Q1 of the current working set
    std::vector<double> hR;         // host copy of dJc
    // comm
    NcclApi::Comm comm = nullptr;
    int rank = 0, nranks = 1;
    // stats
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    double ms_build = 0, ms_tsqr = 0, ms_ls = 0, ms_host = 0, ms_total = 0;
    long long n_newpoint = 0, n_ls = 0, launches = 0;
    float last_tsqr_ms = 0, last_build_ms = 0;

    ~LargeHandle() override { release(); }
    void release() {
        cudaSetDevice(device);
        for (double* p : {ownW, owny, dA, du, dr, ds, dv, dJp, dx, dp, dT, dpart, dout, dR, dStack, dR2, dq_f, dq_m, dq_small, dJc})
            if (p) cudaFree(p);
        if (dq_p) cudaFree(dq_p);
        dq_p = nullptr;
        if (hpin) cudaFreeHost(hpin);
        hpin = nullptr;
        if (hgrad) cudaFreeHost(hgrad);
        hgrad = nullptr;
        if (dgpart) cudaFree(dgpart);
        if (dgrad) cudaFree(dgrad);
        dgpart = dgrad = nullptr;
        dq_f = dq_m = dq_small = dJc = nullptr;
        dq_f_cap = dq_m_cap = dq_small_cap = dq_p_cap = 0;
        ownW = owny = dA = du = dr = ds = dv = dJp = dx = dp = dT = dpart = dout = dR = dStack = dR2 = nullptr;
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (e2) cudaEventDestroy(e2);
        e0 = e1 = e2 = nullptr;
        if (comm && g_nccl.ok) g_nccl.CommDestroy(comm);
        comm = nullptr;
        if (st) cudaStreamDestroy(st);
        st = nullptr;
    }
    int alloc() {
        LCU(cudaSetDevice(device));
        LCU(cudaStreamCreate(&st));
        LCU(cudaEventCreate(&e0)); LCU(cudaEventCreate(&e1)); LCU(cudaEventCreate(&e2));
        ld = n + 8;
        rows_pad = ((m_local + TS_B - 1) / TS_B) * TS_B;
        if (rows_pad < TS_B) rows_pad = TS_B;
        rr_rows = n + TS_B;   // rows of an R factor padded to a multiple of 32
        LCU(cudaMalloc(&dA, sizeof(double) * rows_pad * ld));
        LCU(cudaMemsetAsync(dA, 0, sizeof(double) * rows_pad * ld, st));
        for (double** p : {&du, &dr, &ds, &dv, &dJp}) LCU(cudaMalloc(p, sizeof(double) * (m_local > 0 ? m_local : 1)));
        LCU(cudaMalloc(&dx, sizeof(double) * n));
        LCU(cudaMalloc(&dp, sizeof(double) * n));
        long long nblk = rows_pad / TS_B;
        long long nsub = (nblk + TS_FAN - 1) / TS_FAN;
        LCU(cudaMalloc(&dT, sizeof(double) * (nsub > 64 ? nsub : 64) * TS_B * TS_B));
        LCU(cudaMalloc(&dpart, sizeof(double) * LI_PARTS * 4));
        LCU(cudaMalloc(&dout, sizeof(double) * 8));
        LCU(cudaHostAlloc(&hpin, sizeof(double) * 8, cudaHostAllocDefault));
        LCU(cudaMalloc(&dgpart, sizeof(double) * LI_PARTS * (size_t)(n + 1)));
        LCU(cudaMalloc(&dgrad, sizeof(double) * (n + 1)));
        LCU(cudaHostAlloc(&hgrad, sizeof(double) * (n + 1), cudaHostAllocDefault));
        LCU(cudaMalloc(&dR, sizeof(double) * rr_rows * ld));
        LCU(cudaMalloc(&dJc, sizeof(double) * (size_t)(n + 1) * (n + 1)));
        hR.resize((size_t)(n + 1) * (n + 1));
        return 0;
    }
    int grid_rows() const {   // CTAs for the warp-per-row kernels
        long long want = (m_local + 7) / 8;
        return (int)(want < LI_PARTS ? (want > 0 ? want : 1) : LI_PARTS);
    }
    // sum-reduce dout[0..3] over the ranks, bring to the host
    int finish4(double out[4]) {
        li_finish_kernel<<<1, 128, 0, st>>>(dpart, cur_parts, dout);
        ++launches;
        if (nranks > 1) {
            int rc = g_nccl.AllReduce(dout, dout, 4, NCCL_FLOAT64, NCCL_SUM, comm, st);
            if (rc != 0) return lfail(ENLSIPB200_ECUDA, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(rc));
        }
        // pinned landing buffer: a pageable destination costs a staging copy and an extra synchronisation per call
        LCU(cudaMemcpyAsync(hpin, dout, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
        LCU(cudaStreamSynchronize(st));
        for (int k = 0; k < 4; ++k) out[k] = hpin[k];
        return 0;
    }
    int cur_parts = LI_PARTS;
    double* hpin = nullptr;
    double *dgpart = nullptr, *dgrad = nullptr, *hgrad = nullptr;   // J'r partials [LI_PARTS][n+1], result [n+1], pinned copy
    bool point_factored = true;
    long long n_factor = 0;

    // r, s, u and [J | r] at x (li_build_kernel), gradient J'r and r'r reduced over CTAs and ranks -> hgrad (pinned)
    int eval_at(const double* x) {
        LCU(cudaSetDevice(device));
        LCU(cudaMemcpyAsync(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice, st));
        LCU(cudaEventRecord(e0, st));
        int parts = grid_rows();
        if (m_local > 0) {
            const int nc = (n + 63) / 64;
            const size_t sh = sizeof(double) * (n + 8 * (size_t)(n + 1));
            if (nc <= 1) li_build_kernel<1><<<parts, 256, sh, st>>>(dW, dy, dx, m_local, n, ld, dA, du, dr, ds, dgpart);
            else if (nc <= 2) li_build_kernel<2><<<parts, 256, sh, st>>>(dW, dy, dx, m_local, n, ld, dA, du, dr, ds, dgpart);
            else if (nc <= 4) li_build_kernel<4><<<parts, 256, sh, st>>>(dW, dy, dx, m_local, n, ld, dA, du, dr, ds, dgpart);
            else if (nc <= 8) li_build_kernel<8><<<parts, 256, sh, st>>>(dW, dy, dx, m_local, n, ld, dA, du, dr, ds, dgpart);
            else {
                li_build_kernel<0><<<parts, 256, sizeof(double) * n, st>>>(dW, dy, dx, m_local, n, ld, dA, du, dr, ds, dgpart);
                li_grad_kernel<<<parts, 256, 0, st>>>(dA, ld, m_local, n, dgpart);
                ++launches;
            }
            li_grad_finish_kernel<<<(n + 8) / 8, 256, 0, st>>>(dgpart, parts, n, dgrad);
            launches += 2;
        } else {
            LCU(cudaMemsetAsync(dgrad, 0, sizeof(double) * (n + 1), st));
        }
        if (nranks > 1) {
            int rc = g_nccl.AllReduce(dgrad, dgrad, (size_t)n + 1, NCCL_FLOAT64, NCCL_SUM, comm, st);
            if (rc != 0) return lfail(ENLSIPB200_ECUDA, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(rc));
        }
        LCU(cudaEventRecord(e1, st));
        LCU(cudaMemcpyAsync(hgrad, dgrad, sizeof(double) * (n + 1), cudaMemcpyDeviceToHost, st));
        LCU(cudaStreamSynchronize(st));
        LCU(cudaGetLastError());
        LCU(cudaEventElapsedTime(&last_build_ms, e0, e1));
        ms_build += last_build_ms;
        ++n_newpoint;
        point_factored = false;
        return 0;
    }

    // factor the [J | r] of the last eval_at: dJc = column-major [J~ | r~] (identical on every rank), optionally
    // copied to hR.  Destroys [J | r] (in-place Householder), so it runs at most once per point.
    int factor_point(bool want_host_R = true) {
        if (point_factored) return lfail(ENLSIPB200_EINVAL, "point already factored");
        LCU(cudaSetDevice(device));
        LCU(cudaEventRecord(e1, st));
        LCU(cudaMemsetAsync(dR, 0, sizeof(double) * rr_rows * ld, st));
        launches += tsqr_factor(dA, ld, rows_pad, n, dR, ld, dT, dpart, st);
        double* dfinal = dR;
        if (nranks > 1) {
            int rc = g_nccl.AllGather(dR, dStack, (size_t)rr_rows * ld, NCCL_FLOAT64, comm, st);
            if (rc != 0) return lfail(ENLSIPB200_ECUDA, std::string("ncclAllGather: ") + g_nccl.GetErrorString(rc));
            LCU(cudaMemsetAsync(dR2, 0, sizeof(double) * rr_rows * ld, st));
            launches += tsqr_factor(dStack, ld, (long long)nranks * rr_rows, n, dR2, ld, dT, dpart, st);
            dfinal = dR2;
        }
        {
            const int nc = n + 1;
            dim3 grid((nc + 31) / 32, (nc + 31) / 32), block(32, 8);
            r_to_colmajor_kernel<<<grid, block, 0, st>>>(dfinal, ld, nc, dJc);
            ++launches;
        }
        jq1_resident = false;
        point_factored = true;
        LCU(cudaEventRecord(e2, st));
        if (want_host_R) LCU(cudaMemcpyAsync(hR.data(), dJc, sizeof(double) * (size_t)(n + 1) * (n + 1), cudaMemcpyDeviceToHost, st));
        LCU(cudaStreamSynchronize(st));
        LCU(cudaGetLastError());
        LCU(cudaEventElapsedTime(&last_tsqr_ms, e1, e2));
        ms_tsqr += last_tsqr_ms;
        ++n_factor;
        return 0;
    }
    int factor_at(const double* x, bool want_host_R = true) {
        int rc = eval_at(x);
        return rc != 0 ? rc : factor_point(want_host_R);
    }

    // ---- LargeOps ----
    int eval_point(const double* x, double* gradf, double* rr, double* cx, double* A) override {
        int rc = eval_at(x);
        if (rc != 0) return rc;
        for (int j = 0; j < n; ++j) gradf[j] = hgrad[j];
        *rr = hgrad[n];
        sc.cons(x, cx);
        sc.jac(x, A);
        return 0;
    }
    int compress(double* Jt, double* rt) override {
        int rc = factor_point(false);
        if (rc != 0) return rc;
        const size_t mt = (size_t)n + 1;
        LCU(cudaMemcpyAsync(Jt, dJc, sizeof(double) * mt * n, cudaMemcpyDeviceToHost, st));
        LCU(cudaMemcpyAsync(rt, dJc + mt * n, sizeof(double) * mt, cudaMemcpyDeviceToHost, st));
        LCU(cudaStreamSynchronize(st));
        return 0;
    }
    int set_direction(const double*, const double* p, double sums[3]) override {
        LCU(cudaSetDevice(device));
        auto t0 = std::chrono::steady_clock::now();
        LCU(cudaMemcpyAsync(dp, p, sizeof(double) * n, cudaMemcpyHostToDevice, st));
        cur_parts = grid_rows();
        li_dir_kernel<<<cur_parts, 256, sizeof(double) * n, st>>>(dW, dp, dr, ds, m_local, n, dv, dJp, dpart);
        ++launches;
        double o[4];
        int rc = finish4(o);
        if (rc != 0) return rc;
        ms_ls += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        sums[0] = o[0]; sums[1] = o[1]; sums[2] = o[2];
        return 0;
    }
    int ls_eval(double alpha, int with_coeffs, double o[4]) {
        LCU(cudaSetDevice(device));
        auto t0 = std::chrono::steady_clock::now();
        long long want = (m_local + 255) / 256;
        cur_parts = (int)(want < LI_PARTS ? (want > 0 ? want : 1) : LI_PARTS);
        li_ls_kernel<<<cur_parts, 256, 0, st>>>(du, dv, dy, dr, dJp, m_local, alpha, with_coeffs, dpart);
        ++launches;
        int rc = finish4(o);
        if (rc != 0) return rc;
        ms_ls += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        ++n_ls;
        return 0;
    }
    int res_sq(double alpha, double* out) override {
        double o[4];
        int rc = ls_eval(alpha, 0, o);
        *out = o[0];
        return rc;
    }
    int ls_coeffs(double alpha, double out[4]) override { return ls_eval(alpha, 1, out); }
    int cons(const double* x, double* cx) override { sc.cons(x, cx); return 0; }
};

LargeHandle* LH(enlsipb200_large h) { return reinterpret_cast<LargeHandle*>(h); }

}  // namespace

// =============================================================================================
// C ABI (include/enlsip_b200.h, large-Jacobian section)
// =============================================================================================
extern "C" {

const char* enlsipb200_large_last_error(void) { return g_lerr.c_str(); }

int enlsipb200_large_create(int family, int n, long long m_local, long long m_global, int nb, int ineq,
                            const double* rho, const double* x_low, const double* x_upp, int device,
                            enlsipb200_large* out) {
    if (!out) return lfail(ENLSIPB200_EINVAL, "out is NULL");
    *out = nullptr;
    if (family != ENLSIPB200_FAMILY_SINGLE_INDEX) return lfail(ENLSIPB200_EINVAL, "unknown large-regime family");
    if (n < TS_B || n % TS_B != 0) return lfail(ENLSIPB200_EINVAL, "n must be a positive multiple of 32");
    if (m_local < 0 || m_global < m_local || nb < 0 || 4 * nb > n) return lfail(ENLSIPB200_EINVAL, "bad sizes");
    if ((long long)n + m_global < 1000)
        return lfail(ENLSIPB200_EINVAL, "n + m < 1000: second derivatives stay on in the reference (EF:2658); use the batched engine");
    if (nb > 0 && !rho) return lfail(ENLSIPB200_EINVAL, "rho is NULL");
    int ndev = 0;
    LCU(cudaGetDeviceCount(&ndev));
    if (ndev == 0) return lfail(ENLSIPB200_ENOGPU, "no CUDA device");
    if (device < 0) LCU(cudaGetDevice(&device));
    if (device >= ndev) return lfail(ENLSIPB200_EINVAL, "device out of range");
    LargeHandle* h = new LargeHandle();
    h->device = device;
    h->n = n; h->m = m_global; h->m_local = m_local;
    h->sc.n = n; h->sc.nb = nb; h->sc.ineq = ineq != 0;
    h->sc.rho.assign(rho, rho + nb);
    h->sc.set_bounds(x_low, x_upp);
    h->l = h->sc.l(); h->q = h->sc.q();
    if (h->l == 0) { delete h; return lfail(ENLSIPB200_EINVAL, "There must be at least one constraint (cnls_model.jl:367)"); }
    int rc = h->alloc();
    if (rc != 0) { delete h; return rc; }
    *out = reinterpret_cast<enlsipb200_large>(h);
    return 0;
}

int enlsipb200_large_destroy(enlsipb200_large hh) {
    if (!hh) return 0;
    delete LH(hh);
    return 0;
}

int enlsipb200_large_set_data(enlsipb200_large hh, int slot, const double* ptr, long long count, int on_device) {
    LargeHandle* h = LH(hh);
    if (!h || !ptr) return lfail(ENLSIPB200_EINVAL, "NULL argument");
    LCU(cudaSetDevice(h->device));
    long long want = slot == 0 ? h->m_local * h->n : h->m_local;
    if (slot < 0 || slot > 1 || count != want) return lfail(ENLSIPB200_EINVAL, "bad slot / count");
    const double** dst = slot == 0 ? &h->dW : &h->dy;
    double** own = slot == 0 ? &h->ownW : &h->owny;
    if (on_device) { *dst = ptr; return 0; }
    if (*own && h->own_count[slot] != count) { cudaFree(*own); *own = nullptr; }
    if (!*own) LCU(cudaMalloc(own, sizeof(double) * (count > 0 ? count : 1)));
    h->own_count[slot] = count;
    LCU(cudaMemcpy(*own, ptr, sizeof(double) * count, cudaMemcpyHostToDevice));
    *dst = *own;
    return 0;
}

int enlsipb200_large_comm_id(void* id128) {
    if (!g_nccl.load()) return lfail(ENLSIPB200_ECUDA, g_nccl.err);
    NcclApi::UniqueId id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) return lfail(ENLSIPB200_ECUDA, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
    memcpy(id128, &id, 128);
    return 0;
}

int enlsipb200_large_comm_init(enlsipb200_large hh, const void* id128, int rank, int nranks) {
    LargeHandle* h = LH(hh);
    if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return lfail(ENLSIPB200_EINVAL, "bad comm arguments");
    if (nranks == 1) { h->rank = 0; h->nranks = 1; return 0; }
    if (!g_nccl.load()) return lfail(ENLSIPB200_ECUDA, g_nccl.err);
    LCU(cudaSetDevice(h->device));
    NcclApi::UniqueId id;
    memcpy(&id, id128, 128);
    int rc = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
    if (rc != 0) return lfail(ENLSIPB200_ECUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc));
    h->rank = rank; h->nranks = nranks;
    LCU(cudaMalloc(&h->dStack, sizeof(double) * (size_t)nranks * h->rr_rows * h->ld));
    LCU(cudaMalloc(&h->dR2, sizeof(double) * (size_t)h->rr_rows * h->ld));
    return 0;
}

int enlsipb200_large_solve(enlsipb200_large hh, const double* x0, const enlsipb200_options* o, double* x, double* f,
                           int* exit_code, int* status, int* iters, int* nact, int* active, double* trace,
                           int trace_cap) {
    LargeHandle* h = LH(hh);
    if (!h || !x0 || !o || !x || !f) return lfail(ENLSIPB200_EINVAL, "NULL argument");
    if (!h->dW || !h->dy) return lfail(ENLSIPB200_EINVAL, "family data (W, y) not set");
    LargeOptions opt;
    opt.max_iter = o->max_iter;
    opt.scaling = o->scaling;
    opt.time_limit = o->time_limit;
    double abs_tol = (o->abs_tol == o->abs_tol) ? o->abs_tol : EPS;
    double rel_tol = (o->rel_tol == o->rel_tol) ? o->rel_tol : sqrt(abs_tol);
    opt.eps_rel = rel_tol;
    opt.eps_c = (o->c_tol == o->c_tol) ? o->c_tol : rel_tol;
    opt.eps_x = (o->x_tol == o->x_tol) ? o->x_tol : rel_tol;
    auto t0 = std::chrono::steady_clock::now();
    LargeResult R;
    dense_accel() = h;        // large compressed problems (n >= 384) factor on the device (enl_dense.cuh)
    try {
        LargeSolver S(*h, opt);
        R = S.solve(x0, trace != nullptr && trace_cap > 0);
    } catch (const std::exception& e) {
        dense_accel() = nullptr;
        if (g_lerr.empty()) g_lerr = e.what();
        return ENLSIPB200_ECUDA;
    }
    dense_accel() = nullptr;
    double total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    h->ms_total += total;
    memcpy(x, R.x.data(), sizeof(double) * h->n);
    *f = R.f;
    if (exit_code) *exit_code = R.exit_code;
    if (status) *status = R.status;
    if (iters) *iters = R.iterations;
    if (nact) *nact = R.nact;
    if (active)
        for (int i = 0; i < h->l; ++i) active[i] = i < R.nact ? R.active[i] : 0;
    if (trace && trace_cap > 0) {
        int rows = (int)R.trace.size();
        const int n = h->n;
        for (int k = 0; k < rows && k < trace_cap; ++k) {
            double* tr = trace + (size_t)k * (ENLSIPB200_TRACE_HDR + n);
            const IterTraceL& t = R.trace[k];
            tr[0] = t.f_new; tr[1] = t.t; tr[2] = t.rankA; tr[3] = t.rankJ2; tr[4] = t.dimA; tr[5] = t.dimJ2;
            tr[6] = t.code; tr[7] = t.alpha; tr[8] = t.p_norm; tr[9] = t.index_del; tr[10] = t.exit_code;
            tr[11] = t.active_cx_sum; tr[12] = t.progress; tr[13] = t.k; tr[14] = 0; tr[15] = 0;
            memcpy(tr + ENLSIPB200_TRACE_HDR, R.trace_x.data() + (size_t)k * n, sizeof(double) * n);
        }
    }
    return 0;
}

int enlsipb200_large_factor(enlsipb200_large hh, const double* x, double* R, float* build_ms, float* tsqr_ms) {
    LargeHandle* h = LH(hh);
    if (!h || !x) return lfail(ENLSIPB200_EINVAL, "NULL argument");
    if (!h->dW || !h->dy) return lfail(ENLSIPB200_EINVAL, "family data (W, y) not set");
    int rc = h->factor_at(x);
    if (rc != 0) return rc;
    const int nc = h->n + 1;
    if (R)
        for (int r = 0; r < nc; ++r)
            for (int c = 0; c < nc; ++c) R[(size_t)r * nc + c] = h->hR[(size_t)c * nc + r];
    if (build_ms) *build_ms = h->last_build_ms;
    if (tsqr_ms) *tsqr_ms = h->last_tsqr_ms;
    return 0;
}

int enlsipb200_large_stats(enlsipb200_large hh, double* out, int count) {
    LargeHandle* h = LH(hh);
    if (!h || !out) return lfail(ENLSIPB200_EINVAL, "NULL argument");
    double v[12] = {(double)h->n_newpoint, h->ms_build, h->ms_tsqr, h->ms_ls, h->ms_total, (double)h->n_ls,
                    (double)h->launches, (double)h->rows_pad, (double)h->n_dev_qrcp, (double)h->n_dev_mulq, h->ms_dense,
                    (double)h->n_factor};
    for (int i = 0; i < count && i < 12; ++i) out[i] = v[i];
    return 0;
}

}  // extern "C"

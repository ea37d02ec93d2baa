// dmma_hbm_fused_probe.cu -- inside ONE kernel: warps 0-7 of every CTA issue register-only DMMAs, warps 8-15 stream
// memory (read + write).  mode 0: DMMA warps only work, mode 1: copy warps only, mode 2: both.  time(2) ~ max = the
// tensor pipe and HBM overlap inside an SM; ~ sum = they do not.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/dmma_hbm_fused_probe.cu -o tools/dmma_hbm_fused_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(512, 1) fused(const double2* __restrict__ src, double2* __restrict__ dst, long long n,
                                               double* out, int iters, int mode) {
    const int w = threadIdx.x >> 5;
    if (w < 8) {
        if (mode == 1) return;
        double c[4][4][2], a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = 1.0 + threadIdx.x * 1e-9 + i; b[i] = 1e-3 * (i + 1); }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s += c[i][j][0] + c[i][j][1];
        out[blockIdx.x * 256 + threadIdx.x] = s;
    } else {
        if (mode == 0) return;
        const long long t = (long long)blockIdx.x * 256 + (threadIdx.x - 256), nt = (long long)gridDim.x * 256;
        for (long long i = t; i < n; i += 4 * nt) {
            double2 v0 = src[i], v1, v2, v3;
            if (i + nt < n) v1 = src[i + nt];
            if (i + 2 * nt < n) v2 = src[i + 2 * nt];
            if (i + 3 * nt < n) v3 = src[i + 3 * nt];
            v0.x += 1.0; dst[i] = v0;
            if (i + nt < n) { v1.x += 1.0; dst[i + nt] = v1; }
            if (i + 2 * nt < n) { v2.x += 1.0; dst[i + 2 * nt] = v2; }
            if (i + 3 * nt < n) { v3.x += 1.0; dst[i + 3 * nt] = v3; }
        }
    }
}
int main(int argc, char** argv) {
    const long long n = 1LL << 28;
    double2 *src, *dst; double* out;
    cudaMalloc(&src, sizeof(double2) * n); cudaMalloc(&dst, sizeof(double2) * n); cudaMalloc(&out, 8 * 148 * 256);
    cudaMemset(src, 0, sizeof(double2) * n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float t[3];
    if (argc > 2) {   // soak: <mode> <repetitions> back to back (for power / clock sampling from outside)
        const int mode = atoi(argv[1]), reps = atoi(argv[2]);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) fused<<<148, 512>>>(src, dst, n, out, 7000, mode);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&t[0], e0, e1);
        printf("{\"soak_mode\": %d, \"reps\": %d, \"ms_per_rep\": %.3f}\n", mode, reps, t[0] / reps);
        return 0;
    }
    for (int iters = 6000; iters <= 12000; iters += 6000) {
        for (int mode = 0; mode < 3; ++mode) {
            fused<<<148, 512>>>(src, dst, n, out, iters, mode); cudaDeviceSynchronize();
            cudaEventRecord(e0); fused<<<148, 512>>>(src, dst, n, out, iters, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&t[mode], e0, e1);
        }
        printf("{\"iters\": %d, \"dmma_only_ms\": %.3f, \"dmma_TFLOPs\": %.2f, \"copy_only_ms\": %.3f, \"copy_TBps\": %.2f, \"both_ms\": %.3f, \"err\": \"%s\"}\n",
               iters, t[0], 2.0 * 256 * 16 * (double)iters * 148 * 8 / 1e12 / (t[0] * 1e-3), t[1], 2.0 * 16 * n / (t[1] * 1e-3) / 1e12, t[2],
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

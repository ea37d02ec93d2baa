// dmma_hbm_overlap_probe.cu -- do FP64 tensor work and HBM streaming overlap on B200 when co-resident on every SM?
// Stream 1: register-only DMMA (one 256-thread CTA per SM).  Stream 2: a streaming copy (two 256-thread CTAs per SM).
// Reports each alone and both together: time(both) ~ max(alone) = they overlap; ~ sum = they do not.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/dmma_hbm_overlap_probe.cu -o tools/dmma_hbm_overlap_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) dmma_k(double* out, int iters) {
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
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) ffma_k(double* out, int iters) {
    float a[16];
    float x = 1.0000001f + threadIdx.x * 1e-9f, y = 0.9999999f;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = i * 0.1f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) dfma_k(double* out, int iters) {
    double a[16];
    double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.9999999;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = i * 0.1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) copy_k(const double2* __restrict__ src, double2* __restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double2 v = src[i]; v.x += 1.0; dst[i] = v;
    }
}
int main() {
    const long long n = 1LL << 28;   // 4 GiB read + 4 GiB written
    double2 *src, *dst; double* out;
    cudaMalloc(&src, sizeof(double2) * n); cudaMalloc(&dst, sizeof(double2) * n); cudaMalloc(&out, 8 * 148 * 256);
    cudaMemset(src, 0, sizeof(double2) * n);
    cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    printf("{");
    for (int kind = 0; kind < 3; ++kind) {
        // iteration counts chosen so that each compute kernel alone takes about as long as the copy
        const int iters = kind == 0 ? 5200 : (kind == 1 ? 42000 : 170000);
        auto launch = [&](cudaStream_t st, int it) {
            if (kind == 0) dmma_k<<<148, 256, 0, st>>>(out, it);
            else if (kind == 1) dfma_k<<<148, 256, 0, st>>>(out, it);
            else ffma_k<<<148, 256, 0, st>>>(out, it);
        };
        float t_d, t_c, t_b;
        launch(s1, 100); copy_k<<<296, 256, 0, s2>>>(src, dst, 1 << 20); cudaDeviceSynchronize();
        cudaEventRecord(e0, s1); launch(s1, iters); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&t_d, e0, e1);
        cudaEventRecord(e0, s2); copy_k<<<296, 256, 0, s2>>>(src, dst, n); cudaEventRecord(e1, s2); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&t_c, e0, e1);
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
        launch(s1, iters);
        copy_k<<<296, 256, 0, s2>>>(src, dst, n);
        cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&t_b, e0, e1);
        printf("%s\"%s\": {\"compute_alone_ms\": %.3f, \"copy_alone_ms\": %.3f, \"copy_alone_TBps\": %.2f, \"both_ms\": %.3f, "
               "\"sum_ms\": %.3f, \"max_ms\": %.3f}", kind ? ", " : "", kind == 0 ? "dmma_f64" : (kind == 1 ? "dfma_f64" : "ffma_f32"),
               t_d, t_c, 2.0 * 16 * n / (t_c * 1e-3) / 1e12, t_b, t_d + t_c, t_d > t_c ? t_d : t_c);
    }
    printf(", \"err\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

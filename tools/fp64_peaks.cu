// fp64_peaks.cu -- measures the FP64 DFMA and DMMA (mma.sync f64) peaks of the GPU it runs on.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_peaks.cu -o tools/fp64_peaks
// Output: one JSON object (written to profiles/ by the caller).
#include <cuda_runtime.h>
#include <cstdio>

__global__ void dfma_kernel(double* out, int iters) {
    double a[8];
    double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.9999999;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = i * 0.1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                 "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void dmma884_kernel(double* out, int iters) {
    double c[8][2];
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-3;
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma16816_kernel(double* out, int iters) {
    double c[4][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * i;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
float time_kernel(K launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    double best[3] = {0, 0, 0};
    int bestcfg[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    for (int bpsm = 1; bpsm <= 4; bpsm *= 2)
        for (int thr = 128; thr <= 1024; thr *= 2) {
            if (bpsm * thr > 2048) continue;
            int grid = sms * bpsm;
            float ms = time_kernel([&] { dfma_kernel<<<grid, thr>>>(out, iters); });
            double tf = 2.0 * 8 * iters * (double)grid * thr / (ms * 1e-3) / 1e12;
            if (tf > best[0]) { best[0] = tf; bestcfg[0][0] = bpsm; bestcfg[0][1] = thr; }
            ms = time_kernel([&] { dmma884_kernel<<<grid, thr>>>(out, iters); });
            tf = 2.0 * 256 * 8 * iters * (double)grid * (thr / 32) / (ms * 1e-3) / 1e12;
            if (tf > best[1]) { best[1] = tf; bestcfg[1][0] = bpsm; bestcfg[1][1] = thr; }
            ms = time_kernel([&] { dmma16816_kernel<<<grid, thr>>>(out, iters); });
            tf = 2.0 * 2048 * 4 * iters * (double)grid * (thr / 32) / (ms * 1e-3) / 1e12;
            if (tf > best[2]) { best[2] = tf; bestcfg[2][0] = bpsm; bestcfg[2][1] = thr; }
        }
    cudaError_t e = cudaGetLastError();
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_dfma_tflops\": %.2f, \"dfma_cfg\": [%d, %d], "
           "\"fp64_dmma_m8n8k4_tflops\": %.2f, \"dmma884_cfg\": [%d, %d], \"fp64_dmma_m16n8k16_tflops\": %.2f, "
           "\"dmma16816_cfg\": [%d, %d], \"cuda_error\": \"%s\"}\n",
           p.name, sms, best[0], bestcfg[0][0], bestcfg[0][1], best[1], bestcfg[1][0], bestcfg[1][1], best[2],
           bestcfg[2][0], bestcfg[2][1], cudaGetErrorString(e));
    return 0;
}

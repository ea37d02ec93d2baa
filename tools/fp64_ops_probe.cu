// Throughput of FP64 DADD / DMUL / DFMA / MUFU.RCP64H-based division on the device (are non-FMA FP64 ops full rate?)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double* out, int iters) {
    double a[8]; double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.9999999;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + i * 0.1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = fma(a[i], x, y);
            if (OP == 1) a[i] = __dadd_rn(a[i], y);
            if (OP == 2) a[i] = __dmul_rn(a[i], x);
            if (OP == 3) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(1.0), "d"(y));   // add as fma
            if (OP == 4) a[i] = a[i] > x ? y : a[i] + 1e-9;   // DSETP + select + DADD
            if (OP == 5) a[i] = 1.0 / a[i] + y;                // division
            if (OP == 6) a[i] = sqrt(a[i]) + y;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> double run(double* out, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148 * 4, 512>>>(out, iters); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<OP><<<148 * 4, 512>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return 148.0 * 4 * 512 * 8.0 * iters / (best * 1e-3) / 1e12;   // tera-ops/s (thread-level)
}
int main() {
    double* out; cudaMalloc(&out, sizeof(double) * 148 * 4 * 512);
    printf("{\"dfma_Tops\": %.2f, \"dadd_Tops\": %.2f, \"dmul_Tops\": %.2f, \"add_as_fma_Tops\": %.2f, \"setp_sel_add_Tops\": %.2f, \"div_Tops\": %.3f, \"sqrt_Tops\": %.3f}\n",
           run<0>(out, 4096), run<1>(out, 4096), run<2>(out, 4096), run<3>(out, 4096), run<4>(out, 2048), run<5>(out, 256), run<6>(out, 256));
    return 0;
}

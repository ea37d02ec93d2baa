// Do the scalar panel kernel (FP64 FMA pipe) and the DMMA trailing-update kernel overlap when co-resident?
// Two independent matrices, two streams: serial time vs concurrent time.
#include <cstdio>
#include <cuda_runtime.h>
#include "../enlsip.jl_b200/csrc/enl_tsqr.cuh"
using namespace enl_large;
__global__ void fill(double* A, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long z = (unsigned long long)i * 0x9E3779B97F4A7C15ULL; z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 32;
        A[i] = (double)(z & 0xFFFFF) / 1048576.0 - 0.5;
    }
}
int main(int argc, char** argv) {
    const int PAD = argc > 1 ? 98 * 1024 : 0, TPAD = argc > 1 ? 110 * 1024 : TS_TRAIL_SMEM;
    cudaFuncSetAttribute(tsqr_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98 * 1024);
    const long long m = 1 << 21; const int n = 256, ld = n + 8;
    const long long nblk = m / 32, nsub = nblk / 8;
    double *A1, *A2, *T1, *T2;
    cudaMalloc(&A1, sizeof(double) * m * ld); cudaMalloc(&A2, sizeof(double) * m * ld);
    cudaMalloc(&T1, sizeof(double) * nsub * 1024); cudaMalloc(&T2, sizeof(double) * nsub * 1024);
    fill<<<1184, 256>>>(A1, m * ld); fill<<<1184, 256>>>(A2, m * ld);
    cudaFuncSetAttribute(tsqr_trail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tsqr_panel_kernel<<<(unsigned)nsub, 256, 0, s1>>>(A1, ld, nblk, 1, 0, 0, -1, T1);
    tsqr_panel_kernel<<<(unsigned)nsub, 256, 0, s2>>>(A2, ld, nblk, 1, 0, 0, -1, T2);
    cudaDeviceSynchronize();
    const int ncb = 3, reps = 4;
    float tp, tt, tc;
    cudaEventRecord(e0, s1);
    for (int r = 0; r < reps; ++r) tsqr_panel_kernel<<<(unsigned)nsub, 256, PAD, s1>>>(A1, ld, nblk, 1, 32, 0, -1, T1);
    cudaEventRecord(e1, s1); cudaEventSynchronize(e1); cudaEventElapsedTime(&tp, e0, e1);
    cudaEventRecord(e0, s1);
    for (int r = 0; r < reps; ++r) tsqr_trail_kernel<<<(unsigned)(nsub * ncb), 256, TPAD, s1>>>(A2, ld, nblk, 1, 0, ncb, T2);
    cudaEventRecord(e1, s1); cudaEventSynchronize(e1); cudaEventElapsedTime(&tt, e0, e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s1);
    cudaStreamWaitEvent(s2, e0, 0);
    for (int r = 0; r < reps; ++r) {
        tsqr_panel_kernel<<<(unsigned)nsub, 256, PAD, s1>>>(A1, ld, nblk, 1, 32, 0, -1, T1);
        tsqr_trail_kernel<<<(unsigned)(nsub * ncb), 256, TPAD, s2>>>(A2, ld, nblk, 1, 0, ncb, T2);
    }
    cudaEvent_t e2; cudaEventCreate(&e2); cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0);
    cudaEventRecord(e1, s1); cudaEventSynchronize(e1); cudaEventElapsedTime(&tc, e0, e1);
    printf("{\"panel_ms\": %.3f, \"trail_ms\": %.3f, \"serial_ms\": %.3f, \"concurrent_ms\": %.3f, \"err\": \"%s\"}\n", tp / reps, tt / reps,
           (tp + tt) / reps, tc / reps, cudaGetErrorString(cudaGetLastError()));
    return 0;
}

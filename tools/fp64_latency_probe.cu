// fp64_latency_probe.cu -- dependent-issue latency (cycles) of the FP64 operations on the panel kernel's per-column
// critical path: DFMA, DADD, rsqrt(double), __drcp_rn, a 64-bit shuffle, a shared-memory store->load round trip and
// __syncthreads with 256 threads.  One warp (one CTA of 256 threads for the barrier), clock64 around a dependent chain.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_latency_probe.cu -o tools/fp64_latency_probe
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double* out, long long* cyc, int iters) {
    __shared__ double sh[256];
    double a = 1.0 + threadIdx.x * 1e-9, y = 0.5000001;
    sh[threadIdx.x] = a;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a = fma(a, y, y);
            if (OP == 1) a = __dadd_rn(a, y);
            if (OP == 2) a = rsqrt(a) + 1.0;
            if (OP == 3) a = __drcp_rn(a) + 1.0;
            if (OP == 4) a = __shfl_xor_sync(0xffffffffu, a, 8);
            if (OP == 5) { sh[threadIdx.x] = a; __syncwarp(); a = sh[threadIdx.x ^ 1]; __syncwarp(); }
            if (OP == 6) { __syncthreads(); }
            if (OP == 7) a = 1.0 / a + 1.0;
            if (OP == 8) a = sqrt(a) + 1.0;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> double run(double* out, long long* cyc, int thr) {
    const int iters = 256;
    k<OP><<<1, thr>>>(out, cyc, iters); cudaDeviceSynchronize();
    k<OP><<<1, thr>>>(out, cyc, iters); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    return (double)c / (iters * 8.0);
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
    printf("{\"dfma\": %.1f, \"dadd\": %.1f, \"rsqrt_plus_add\": %.1f, \"drcp_plus_add\": %.1f, \"shfl64\": %.1f, "
           "\"sts_lds_roundtrip\": %.1f, \"syncthreads_256\": %.1f, \"div_plus_add\": %.1f, \"sqrt_plus_add\": %.1f, \"err\": \"%s\"}\n",
           run<0>(out, cyc, 32), run<1>(out, cyc, 32), run<2>(out, cyc, 32), run<3>(out, cyc, 32), run<4>(out, cyc, 32),
           run<5>(out, cyc, 32), run<6>(out, cyc, 256), run<7>(out, cyc, 32), run<8>(out, cyc, 32),
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}

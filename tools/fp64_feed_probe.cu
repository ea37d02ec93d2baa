// fp64_feed_probe.cu -- how fast do DMMA (mma.sync.m8n8k4.f64) and DFMA run when their operands change on
// every instruction, as in a real kernel?  tools/fp64_peaks.cu issues every instruction with the SAME operand
// registers (operand-reuse cache always hits); here:
//   dmma_regs   : 4 x 4 register tile, 8 distinct operand registers, never reloaded      (register-file feed)
//   dmma_lds    : 4 x 4 register tile, the 8 operands reloaded from shared memory per k-step (pass 2 of the trailing update)
//   dmma_lds22  : 2 x 2 register tile, 4 operands per 4 DMMA from shared memory            (pass 1 of the staged kernel)
//   dfma_vec    : a[r] = fma(c, v[r], a[r]) over 32 distinct a / v registers                (panel update)
//   dfma_lds    : the same with v[r] read from shared memory (broadcast 16-byte loads)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_feed_probe.cu -o tools/fp64_feed_probe
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void dmma_regs(double* out, int iters) {
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

__global__ void dmma_lds(double* out, int iters) {
    __shared__ double sh[64 * 36];
    for (int i = threadIdx.x; i < 64 * 36; i += blockDim.x) sh[i] = 1e-3 * (i % 97);
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double c[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
        const double* vp = sh + g * 36 + t + (it & 7) * 4;
        const double* wp = sh + 32 * 36 + t * 36 + g + (it & 7) * 4 * 36;
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = vp[i * 8 * 36];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = wp[j * 8];
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

__global__ void dmma_lds22(double* out, int iters) {
    __shared__ double sh[2 * 32 * 36];
    for (int i = threadIdx.x; i < 2 * 32 * 36; i += blockDim.x) sh[i] = 1e-3 * (i % 97);
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double c[2][2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) c[i][j][0] = c[i][j][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const double* vp = sh + ((it & 1) * 16 + kk * 4 + t) * 36 + g;
            const double a0 = vp[0], a1 = vp[8], b0 = vp[32 * 36 + 16], b1 = vp[32 * 36 + 24];
            dmma884(c[0][0][0], c[0][0][1], a0, b0);
            dmma884(c[0][1][0], c[0][1][1], a0, b1);
            dmma884(c[1][0][0], c[1][0][1], a1, b0);
            dmma884(c[1][1][0], c[1][1][1], a1, b1);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) s += c[i][j][0] + c[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_vec(double* out, int iters) {
    double a[32], v[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) { a[r] = r * 0.1; v[r] = 1.0 + 1e-9 * (threadIdx.x + r); }
    double cs = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 32; ++r) a[r] = fma(-cs, v[r], a[r]);
        cs = -cs;
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < 32; ++r) s += a[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_lds(double* out, int iters) {
    __shared__ __align__(16) double sh[8][32];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i >> 5][i & 31] = 1.0 + 1e-9 * i;
    __syncthreads();
    const int w = (threadIdx.x >> 5) & 7;
    double a[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) a[r] = r * 0.1;
    double cs = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 32; r += 2) {
            const double2 v = *reinterpret_cast<const double2*>(&sh[(w + it) & 7][r]);
            a[r] = fma(-cs, v.x, a[r]);
            a[r + 1] = fma(-cs, v.y, a[r + 1]);
        }
        cs = -cs;
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < 32; ++r) s += a[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
float time_kernel(K launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 4000;
    printf("{\"gpu\": \"%s\"", p.name);
    // warps per SM = bpsm * thr / 32
    const int cfgs[5][2] = {{1, 256}, {2, 256}, {1, 512}, {2, 512}, {4, 512}};
    for (int ci = 0; ci < 5; ++ci) {
        const int bpsm = cfgs[ci][0], thr = cfgs[ci][1], grid = sms * bpsm;
        const double warps = (double)grid * (thr / 32);
        float ms;
        ms = time_kernel([&] { dmma_regs<<<grid, thr>>>(out, iters); });
        const double t_regs = 2.0 * 256 * 16 * iters * warps / (ms * 1e-3) / 1e12;
        ms = time_kernel([&] { dmma_lds<<<grid, thr>>>(out, iters); });
        const double t_lds = 2.0 * 256 * 16 * iters * warps / (ms * 1e-3) / 1e12;
        ms = time_kernel([&] { dmma_lds22<<<grid, thr>>>(out, iters); });
        const double t_lds22 = 2.0 * 256 * 16 * iters * warps / (ms * 1e-3) / 1e12;
        ms = time_kernel([&] { dfma_vec<<<grid, thr>>>(out, iters); });
        const double t_fv = 2.0 * 32 * iters * warps * 32 / (ms * 1e-3) / 1e12;
        ms = time_kernel([&] { dfma_lds<<<grid, thr>>>(out, iters); });
        const double t_fl = 2.0 * 32 * iters * warps * 32 / (ms * 1e-3) / 1e12;
        printf(", \"warps%d\": {\"dmma_regs\": %.2f, \"dmma_lds44\": %.2f, \"dmma_lds22\": %.2f, \"dfma_vec\": %.2f, \"dfma_lds\": %.2f}",
               bpsm * thr / 32, t_regs, t_lds, t_lds22, t_fv, t_fl);
    }
    printf(", \"cuda_error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

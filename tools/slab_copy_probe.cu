// slab_copy_probe.cu -- what HBM bandwidth does the TSQR's access pattern get?
// The trailing update reads and writes 256 x 32 tiles of the row-major [J | r] (row stride n + 8 = 264 doubles): every
// tile row is a 256-byte slice 2112 bytes away from the next.  This probe moves the same bytes (read a tile, add 1,
// write it back, 7 column blocks of a 1M-row matrix) in that layout and in a tile-major layout where a tile is one
// contiguous 64 KB block, with the staged kernel's geometry (one 256-thread CTA per subtile, 16-byte accesses).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/slab_copy_probe.cu -o tools/slab_copy_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) slab(double* A, long long rs, long long plane, int ncb, int inflight) {
    const long long sub = blockIdx.x;
    const int tid = threadIdx.x;
    for (int cb = 1; cb <= ncb; ++cb) {
        double2 v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int id = tid + 256 * q, row = id >> 4, c16 = id & 15;
            v[q] = *reinterpret_cast<const double2*>(A + cb * plane + (sub * 256 + row) * rs + 2 * c16);
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int id = tid + 256 * q, row = id >> 4, c16 = id & 15;
            v[q].x += 1.0; v[q].y += 1.0;
            *reinterpret_cast<double2*>(A + cb * plane + (sub * 256 + row) * rs + 2 * c16) = v[q];
        }
    }
}
int main() {
    const long long m = 1 << 20;
    const int ld = 264, ncb = 7;
    double* A; cudaMalloc(&A, sizeof(double) * m * 288);
    cudaMemset(A, 0, sizeof(double) * m * 288);
    cudaFuncSetAttribute(slab, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = 2.0 * m * 32 * 8 * ncb;
    printf("{");
    for (int mode = 0; mode < 2; ++mode) {
        const long long rs = mode == 0 ? ld : 32, plane = mode == 0 ? 32 : m * 32;
        for (int ctas = 1; ctas <= 4; ctas *= 2) {
            float best = 1e30f;
            for (int r = 0; r < 4; ++r) {
                cudaEventRecord(e0);
                slab<<<(unsigned)(m / 256), 256, (ctas == 1 ? 150 : (ctas == 2 ? 100 : 50)) * 1024>>>(A, rs, plane, ncb, 0);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("%s\"%s_ctas%d_TBps\": %.2f", (mode || ctas > 1) ? ", " : "", mode == 0 ? "row_major_ld264" : "tile_major", ctas, bytes / (best * 1e-3) / 1e12);
        }
    }
    printf(", \"err\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

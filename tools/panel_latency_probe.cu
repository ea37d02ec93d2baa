// Where does the per-column latency of tsqr_panel_kernel go?  One CTA, clock64 stamps of warp 1 / lane 0 around the
// phases of a column step (a copy of the kernel body with stamps; development aid only).
#include <cstdio>
#include <cuda_runtime.h>
#include "../enlsip.jl_b200/csrc/enl_tsqr.cuh"
using namespace enl_large;

__global__ void __launch_bounds__(256, 2) probe(double* __restrict__ A, int ld, long long* stamps) {
    __shared__ __align__(16) double vbuf[TS_FAN][TS_B];
    __shared__ double dots[2][TS_FAN][TS_B];
    __shared__ double rowi[2][TS_B];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* base = A + ((long long)w * TS_B) * ld + lane;
    double a[TS_B];
#pragma unroll
    for (int r = 0; r < TS_B; ++r) a[r] = base[(long long)r * ld];
    const bool rec = (w == 1 && lane == 0);
#pragma unroll 1
    for (int ib = 0; ib < TS_B / TS_CG; ++ib) {
#pragma unroll
        for (int k = 0; k < TS_CG; ++k) {
            const int i = ib * TS_CG + k;
            const int buf = k & 1;
            long long t0 = clock64();
            __syncwarp();
            if (lane == i) {
#pragma unroll
                for (int r = 0; r < TS_B; r += 2) {
                    double2 v = make_double2(a[r], a[r + 1]);
                    if (w == 0) { if (r <= k) v.x = 0.0; if (r + 1 <= k) v.y = 0.0; }
                    *reinterpret_cast<double2*>(&vbuf[w][r]) = v;
                }
            }
            if (w == 0) rowi[buf][lane] = a[k];
            __syncwarp();
            long long t1 = clock64();
            double vk[TS_KEEP];
            double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
            for (int r = 0; r < TS_B; r += 4) {
                const double2 v01 = *reinterpret_cast<const double2*>(&vbuf[w][r]);
                const double2 v23 = *reinterpret_cast<const double2*>(&vbuf[w][r + 2]);
                if (r < TS_KEEP) { vk[r] = v01.x; vk[r + 1] = v01.y; vk[r + 2] = v23.x; vk[r + 3] = v23.y; }
                d0 = fma(v01.x, a[r + 0], d0); d1 = fma(v01.y, a[r + 1], d1);
                d2 = fma(v23.x, a[r + 2], d2); d3 = fma(v23.y, a[r + 3], d3);
            }
            dots[buf][w][lane] = (d0 + d1) + (d2 + d3);
            long long t2 = clock64();
            __syncthreads();
            long long t3 = clock64();
            const double g = ((dots[buf][0][lane] + dots[buf][1][lane]) + (dots[buf][2][lane] + dots[buf][3][lane])) +
                             ((dots[buf][4][lane] + dots[buf][5][lane]) + (dots[buf][6][lane] + dots[buf][7][lane]));
            const double sigma = ((dots[buf][0][i] + dots[buf][1][i]) + (dots[buf][2][i] + dots[buf][3][i])) +
                                 ((dots[buf][4][i] + dots[buf][5][i]) + (dots[buf][6][i] + dots[buf][7][i]));
            const double alpha = rowi[buf][i];
            long long t4 = clock64() + (long long)(g * 0.0) + (long long)(sigma * 0.0) + (long long)(alpha * 0.0);
            double tau = 0.0, scale = 0.0, beta = alpha;
            if (sigma != 0.0) {
                const double s2 = fma(alpha, alpha, sigma);
                const double rs = rsqrt(s2);
                beta = -copysign(s2 * rs, alpha);
                const double dab = alpha - beta;
                scale = __drcp_rn(dab);
                tau = dab * copysign(rs, alpha);
            }
            const double gv = fma(scale, g, rowi[buf][lane]);
            const double wc = (lane > i) ? tau * gv : 0.0;
            if (w == 0) a[k] = (lane == i) ? beta : a[k] - wc;
            const double cs = wc * scale;
            long long t5 = clock64() + (long long)(cs * 0.0);
#pragma unroll
            for (int r = 0; r < TS_KEEP; ++r) a[r] = fma(-cs, vk[r], a[r]);
#pragma unroll
            for (int r = TS_KEEP; r < TS_B; r += 2) {
                const double2 v = *reinterpret_cast<const double2*>(&vbuf[w][r]);
                a[r] = fma(-cs, v.x, a[r]);
                a[r + 1] = fma(-cs, v.y, a[r + 1]);
            }
            if (lane == i) {
#pragma unroll
                for (int r = 0; r < TS_B; ++r)
                    if (r > k || w > 0) a[r] *= scale;
            }
            long long t6 = clock64() + (long long)(a[31] * 0.0);
            if (rec) { long long* s = stamps + i * 8; s[0] = t0; s[1] = t1; s[2] = t2; s[3] = t3; s[4] = t4; s[5] = t5; s[6] = t6; }
        }
        double tmp[TS_CG];
#pragma unroll
        for (int r = 0; r < TS_CG; ++r) tmp[r] = a[r];
#pragma unroll
        for (int r = 0; r < TS_B - TS_CG; ++r) a[r] = a[r + TS_CG];
#pragma unroll
        for (int r = 0; r < TS_CG; ++r) a[TS_B - TS_CG + r] = tmp[r];
    }
#pragma unroll
    for (int r = 0; r < TS_B; ++r) base[(long long)r * ld] = a[r];
}

int main() {
    const int ld = 264;
    double* A; long long* st;
    cudaMalloc(&A, sizeof(double) * 256 * ld); cudaMalloc(&st, sizeof(long long) * 32 * 8);
    double* h = new double[256 * ld];
    for (int i = 0; i < 256 * ld; ++i) h[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
    cudaMemcpy(A, h, sizeof(double) * 256 * ld, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 3; ++rep) { cudaMemcpy(A, h, sizeof(double) * 256 * ld, cudaMemcpyHostToDevice); probe<<<1, 256>>>(A, ld, st); }
    cudaDeviceSynchronize();
    long long hs[32 * 8];
    cudaMemcpy(hs, st, sizeof(hs), cudaMemcpyDeviceToHost);
    double acc[6] = {0};
    for (int i = 4; i < 28; ++i) for (int p = 0; p < 6; ++p) acc[p] += (double)(hs[i * 8 + p + 1] - hs[i * 8 + p]);
    const char* names[6] = {"publish+syncwarp", "dots", "barrier", "reduce-loads", "scalars", "update+scale"};
    double tot = 0; for (int p = 0; p < 6; ++p) tot += acc[p] / 24;
    printf("{");
    for (int p = 0; p < 6; ++p) printf("\"%s\": %.0f, ", names[p], acc[p] / 24);
    printf("\"column_total_cycles\": %.0f, \"next_column_gap\": %.0f, \"err\": \"%s\"}\n", tot, (double)(hs[11 * 8] - hs[10 * 8 + 6]), cudaGetErrorString(cudaGetLastError()));
    return 0;
}

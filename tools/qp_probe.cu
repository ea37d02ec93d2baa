// Per-phase cycle counts of the persistent dlaqps panel kernel (enl_small.cuh, built with -DQP_PROF):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DQP_PROF -I enlsip.jl_b200/csrc tools/qp_probe.cu -o /tmp/qp_probe
//   /tmp/qp_probe rows cols
#include <stdio.h>
#include <vector>
#include <random>
#include "enl_small.cuh"

int main(int argc, char** argv) {
    const int rows = argc > 1 ? atoi(argv[1]) : 4097, cols = argc > 2 ? atoi(argv[2]) : 3587;
    std::mt19937_64 g(1);
    std::normal_distribution<double> nd;
    std::vector<double> A((size_t)rows * cols);
    for (double& v : A) v = nd(g);
    double *f, *tau; int* jp;
    cudaMalloc(&f, sizeof(double) * A.size()); cudaMalloc(&tau, sizeof(double) * cols); cudaMalloc(&jp, sizeof(int) * cols);
    enl_small::QrWork wk;
    cudaMalloc(&wk.vn1, sizeof(double) * cols); cudaMalloc(&wk.vn2, sizeof(double) * cols);
    cudaMalloc(&wk.F, sizeof(double) * (size_t)cols * 32); cudaMalloc(&wk.auxv, sizeof(double) * 32);
    cudaMalloc(&wk.pbest, sizeof(double) * enl_small::QR_MAXPART); cudaMalloc(&wk.psum, sizeof(double) * enl_small::QR_PSUM_LEN);
    cudaMalloc(&wk.pidx, sizeof(int) * enl_small::QR_PIDX_LEN); cudaMalloc(&wk.flags, sizeof(int) * cols);
    cudaMalloc(&wk.state, sizeof(enl_small::QrState)); cudaMalloc(&wk.ticket, sizeof(unsigned int) * enl_small::QR_TICKET_LEN);
    wk.cap_cols = cols;
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(f, A.data(), sizeof(double) * A.size(), cudaMemcpyHostToDevice);
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(enl_small::qp_prof, z, sizeof(z));
        cudaEventRecord(e0, st);
        const int launches = enl_small::qrcp_device(f, rows, cols, tau, jp, wk, st);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpyFromSymbol(z, enl_small::qp_prof, sizeof(z));
        const int minmn = rows < cols ? rows : cols, blocked = minmn - 128;
        printf("qrcp %d x %d: %.2f ms, %d launches, err %s\n", rows, cols, ms, launches, cudaGetErrorString(cudaGetLastError()));
        const char* nm[8] = {"pivot combine", "column (phase P work)", "barrier P", "partials + scalars", "gemv + finish", "barrier Q", "-", "loop top"};
        double tot = 0;
        for (int i = 0; i < 8; ++i) tot += (double)z[i];
        for (int i = 0; i < 8; ++i)
            if (z[i]) printf("  %-24s %8.0f cycles/column  %5.1f %%\n", nm[i], (double)z[i] / blocked, 100.0 * z[i] / tot);
        printf("  total %.0f cycles/column\n", tot / blocked);
    }
    return 0;
}

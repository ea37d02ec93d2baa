// hostport.cpp -- scalar CPU build (G = 1, one "lane" per problem) of the engine's solver core.
//
// TEST INFRASTRUCTURE / CPU BASELINE ONLY.  This translation unit compiles the *same* headers as
// the CUDA library (enlsip.jl_b200/csrc/enl_*.h) with g++ so that
//   (1) the control flow of the device code can be debugged and checked against the Python oracle
//       in the build container, which has no GPU, and
//   (2) bench.py has a compiled single-thread-per-problem CPU implementation to time as the
//       `cpu_baseline` ("port") / `--impl reference` arm, next to the Python oracle.
// It is NOT part of the product: libenlsip_b200.so never links it and the product path has no CPU
// fallback.  Parity claims rest on the independent Python oracle (oracle/enlsip_oracle.py).
//
// Build: g++ -O2 -ffp-contract=off -mfma -std=c++17 -shared -fPIC -fopenmp hostport.cpp -o ../_build/libhostport.so
#include <dlfcn.h>

#include <chrono>
#include <cstring>
#include <type_traits>
#include <vector>
#define ENL_HOST_BUILD 1
#include "../../enlsip.jl_b200/csrc/enl_solver.h"

using namespace enl;

template <class Fam>
static void solve_range(long long b0, long long b1, const double* x0, const FamilyData& fd, const Options& opt,
                        const Bounds& bnd, const Outputs& out) {
    using LY = Layout<Fam, 1, 1>;
    std::vector<double> small(LY::nD), dist((size_t)LY::DCOLS * LY::MS);
    std::vector<int> ints(LY::nI);
    HostGroup g;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    for (long long b = b0; b < b1; ++b) {
        std::fill(small.begin(), small.end(), 0.0);
        std::fill(dist.begin(), dist.end(), 0.0);
        std::fill(ints.begin(), ints.end(), 0);
        Solver<Fam, HostGroup, 1> S(small.data(), ints.data(), dist.data(), 0, opt, bnd);
        S.init(x0 + b * Fam::N, fd, b, now());
        int row = 0;
        while (S.exit_code == 0) {
            double* tr = nullptr;
            if (out.trace && row < out.trace_cap) tr = out.trace + ((size_t)b * out.trace_cap + row) * (TRACE_HDR + Fam::N);
            S.step(now(), tr);
            ++row;
        }
        S.store(out, b);
    }
}

extern "C" int hostport_solve(int family, long long B, const double* x0, const double* d0, const double* d1,
                              const double* x_low, const double* x_upp, const Options* opt, double* x, double* f,
                              int* exit_code, int* status, int* iters, int* nact, int* active, int* counters,
                              double* trace, int trace_cap, int nthreads) {
    int n = family == 0 ? FamHS65::N : family == 1 ? FamGaussPeaks::N : family == 2 ? FamOsborne2::N : family == 3 ? 10 : 20;
    Bounds bnd{};
    for (int j = 0; j < n; ++j)
        if (std::isfinite(x_low[j])) { bnd.lo_idx[bnd.nlo] = j; bnd.lo_val[bnd.nlo] = x_low[j]; bnd.nlo++; }
    for (int j = 0; j < n; ++j)
        if (std::isfinite(x_upp[j])) { bnd.up_idx[bnd.nup] = j; bnd.up_val[bnd.nup] = x_upp[j]; bnd.nup++; }
    FamilyData fd{d0, d1, nullptr};
    Outputs out{x, f, exit_code, status, iters, nact, active, counters, trace, trace_cap};
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (int c = 0; c < nthreads * 8; ++c) {
        long long b0 = B * c / (nthreads * 8), b1 = B * (c + 1) / (nthreads * 8);
        switch (family) {
            case 0: solve_range<FamHS65>(b0, b1, x0, fd, *opt, bnd, out); break;
            case 1: solve_range<FamGaussPeaks>(b0, b1, x0, fd, *opt, bnd, out); break;
            case 2: solve_range<FamOsborne2>(b0, b1, x0, fd, *opt, bnd, out); break;
            case 3: solve_range<FamChainedRosenbrock<10>>(b0, b1, x0, fd, *opt, bnd, out); break;
            default: solve_range<FamChainedWood<20>>(b0, b1, x0, fd, *opt, bnd, out); break;
        }
    }
    return 0;
}

extern "C" void hostport_det_exp(const double* x, double* y, long long n) {
    for (long long i = 0; i < n; ++i) y[i] = enl::det_exp(x[i]);
}

// Known-answer hook: the batched engine's `qr(., ColumnNorm())` restatement (enl_linalg.h qrcp_small, the routine every
// lane runs on the group's shared state) on a host matrix: f [rows x cols] column major in/out, tau, jpvt (0-based).
extern "C" void hostport_qrcp(int rows, int cols, double* f, double* tau, int* jpvt) {
    std::vector<double> vn1(cols), vn2(cols);
    qrcp_small(SV<1>{f}, rows, rows, cols, SV<1>{tau}, SI<1>{jpvt}, SV<1>{vn1.data()}, SV<1>{vn2.data()});
}

// Reference-arm switch (bench.py --impl reference): bind LAPACK dgeqp3 from an OpenBLAS shared library (the SciPy wheel's
// libscipy_openblas: symbols scipy_dgeqp3_, scipy_openblas_set_num_threads) so that every `qr(., ColumnNorm())` of the
// solve (EF:223, 700, 769) is executed by the routine Julia itself calls.  path == NULL unbinds.  Returns 0 on success.
extern "C" int hostport_use_lapack(const char* path) {
    if (!path) { enl::host_dgeqp3() = nullptr; return 0; }
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    void* fn = dlsym(h, "scipy_dgeqp3_");
    if (!fn) fn = dlsym(h, "dgeqp3_");
    if (!fn) return -2;
    typedef void (*setthr_fn)(int);
    setthr_fn st = (setthr_fn)dlsym(h, "scipy_openblas_set_num_threads");
    if (!st) st = (setthr_fn)dlsym(h, "openblas_set_num_threads");
    if (st) st(1);       // problems are independent: parallelism is OpenMP over problems, BLAS stays single-threaded
    enl::host_dgeqp3() = (enl::host_dgeqp3_fn)fn;
    return 0;
}

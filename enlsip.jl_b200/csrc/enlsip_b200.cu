// enlsip_b200.cu -- kernels and C ABI (include/enlsip_b200.h) of the batched ENLSIP engine.
//
// One persistent kernel per solve call: every group of G lanes repeatedly takes the next problem
// from a global work counter, runs the complete ENLSIP solve (enl_solver.h) with the problem's
// state in shared memory / registers, and writes x, f, exit code, iteration count and the final
// working set.  HBM traffic is the problem data in and the solution out -- nothing else.
//
// sm_100a only.  No CPU fallback: without a CUDA device every entry point returns ENLSIPB200_ENOGPU.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <new>
#include <type_traits>
#include <string>
#include <vector>

#include "../../include/enlsip_b200.h"
#include "enl_solver.h"

using namespace enl;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? ENLSIPB200_ENOGPU \
                                                                                     : ENLSIPB200_ECUDA, \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                             \
    } while (0)

struct KernelArgs {
    long long B;
    unsigned long long* counter;
    const double* x0;
    FamilyData fd;
    Options opt;
    Bounds bnd;
    Outputs out;
};

__device__ __forceinline__ double now_seconds() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return (double)t * 1e-9;
}

template <class Fam, int G, int NT>
__global__ void __launch_bounds__(NT) enlsip_solve_batch_kernel(const __grid_constant__ KernelArgs a) {
    using LY = Layout<Fam, G, NT>;
    constexpr bool SYNC_ITER = (NT > 32);
    using SolverT = Solver<Fam, DevGroup<G>, NT>;
    constexpr int SOLVER_DOUBLES = (int)((sizeof(SolverT) + 7) / 8) | 1;   // odd stride: no bank conflicts for G == 1
    const int tid = threadIdx.x;
    const int pid = tid / G;
    // CTA-shared copy of the options and bounds: reading kernel parameters through the generic pointers held
    // by the solver object costs a long-scoreboard stall per access (ncu: eval_point 50 % long_sb)
    constexpr int CFG_DOUBLES = (int)((sizeof(Options) + sizeof(Bounds) + 15) / 16) * 2;
    Options* s_opt = reinterpret_cast<Options*>(enl_smem);
    Bounds* s_bnd = reinterpret_cast<Bounds*>(reinterpret_cast<char*>(enl_smem) + sizeof(Options));
    if (tid == 0) { *s_opt = a.opt; *s_bnd = a.bnd; }
    __syncthreads();
    double* objs = enl_smem + CFG_DOUBLES;
    double* small = objs + (size_t)SOLVER_DOUBLES * LY::PPC;
    double* distb = small + (size_t)LY::nD * LY::PPC;
    int* ints = reinterpret_cast<int*>(distb + (size_t)LY::DCOLS * LY::MS * NT);
    DevGroup<G> g;
    // the per-problem solver object lives in shared memory; every lane of the group writes identical values
    SolverT& S = *new (objs + (size_t)SOLVER_DOUBLES * pid) SolverT(small, ints, distb, pid, *s_opt, *s_bnd);
    g.sync();
    const int row_w = TRACE_HDR + Fam::N;
    bool have = false, exhausted = false;
    long long b = 0;
    int row = 0;
    for (;;) {
        if (!have && !exhausted) {
            unsigned long long nb = 0;
            if (g.lane == 0) nb = atomicAdd(a.counter, 1ULL);
            if (G > 1) nb = __shfl_sync(g.mask, nb, 0, G);
            b = (long long)nb;
            if (b >= a.B) {
                exhausted = true;
            } else {
                S.init(a.x0 + b * Fam::N, a.fd, b, now_seconds());
                have = true;
                row = 0;
            }
        }
        // Multi-warp CTAs re-align their warps once per iteration: co-resident warps then run the same
        // routines at the same time and share instruction-cache lines (the kernel is fetch-bound).
        if (SYNC_ITER) {
            // (measured alternatives, all slower: re-aligning in 2 / 7 independent groups 1.26 / 0.60 M solves/s,
            //  every 2nd / 3rd pass 1.65 / 1.35 M, 1-3 extra barriers inside Solver::step 1.93-1.94 M, vs 1.96 M)
            if (!__syncthreads_or(have ? 1 : 0)) break;
        } else if (!have) {
            break;
        }
        if (have) {
            if (S.exit_code == 0) {
                double* tr = nullptr;
                if (a.out.trace && row < a.out.trace_cap && g.lane == 0)
                    tr = a.out.trace + ((size_t)b * a.out.trace_cap + row) * row_w;
                S.step(now_seconds(), tr);
                ++row;
            }
            if (S.exit_code != 0) {
                S.store(a.out, b);
                have = false;
            }
        }
    }
}

__global__ void det_exp_kernel(const double* x, double* y, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) y[i] = det_exp(x[i]);
}

struct FamilyInfo { int n, m, q, ni, maxb; };

}  // namespace

struct enlsipb200_handle_s {
    int family;
    int device;
    FamilyInfo fi;
    Bounds bnd;
    int l;
    unsigned long long* counter = nullptr;
    const double* data[3] = {nullptr, nullptr, nullptr};
    double* owned[3] = {nullptr, nullptr, nullptr};
    long long owned_count[3] = {0, 0, 0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    long long launches = 0;
    // launch geometry
    int regs = 0, smem = 0, nt = 0, ctas_per_sm = 0, grid = 0, G = 0;
    int num_sms = 0;
    // staging for host-buffer calls
    void* stage = nullptr;
    size_t stage_bytes = 0;
};

namespace {

template <class Fam, int G, int NT>
size_t smem_total() {
    using LY = Layout<Fam, G, NT>;
    size_t sd = ((sizeof(Solver<Fam, DevGroup<G>, NT>) + 7) / 8) | 1;
    size_t cfg = ((sizeof(Options) + sizeof(Bounds) + 15) / 16) * 16;
    return LY::smem_bytes() + sd * 8 * LY::PPC + cfg;
}

template <class Fam, int G, int NT>
int configure(enlsipb200_handle h) {
    using LY = Layout<Fam, G, NT>;
    auto kern = enlsip_solve_batch_kernel<Fam, G, NT>;
    size_t smem = smem_total<Fam, G, NT>();
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, kern));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) return fail(ENLSIPB200_ECUDA, "kernel does not fit on an SM");
    h->regs = fa.numRegs;
    h->smem = (int)smem;
    h->nt = NT;
    h->ctas_per_sm = occ;
    h->grid = occ * h->num_sms;
    h->G = G;
    return 0;
}

template <class Fam, int G, int NT>
int launch(enlsipb200_handle h, const KernelArgs& a, cudaStream_t st) {
    using LY = Layout<Fam, G, NT>;
    long long groups_per_cta = NT / G;
    long long need = (a.B + groups_per_cta - 1) / groups_per_cta;
    int grid = (int)(need < (long long)h->grid ? need : (long long)h->grid);
    if (grid < 1) grid = 1;
    CU(cudaMemsetAsync(h->counter, 0, sizeof(unsigned long long), st));
    CU(cudaEventRecord(h->ev0, st));
    enlsip_solve_batch_kernel<Fam, G, NT><<<grid, NT, smem_total<Fam, G, NT>(), st>>>(a);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev1, st));
    h->timed = true;
    h->launches += 1;
    return 0;
}

template <int V>
using ic = std::integral_constant<int, V>;

// family id (+ CTA size for the tuned family) -> template instantiation
template <class F>
int with_family(int family, int nt, F&& f) {
    switch (family) {
        case ENLSIPB200_FAMILY_HS65: return f(FamHS65{}, ic<1>{}, ic<64>{});
        case ENLSIPB200_FAMILY_GAUSS_PEAKS:
            switch (nt) {
                case 64: return f(FamGaussPeaks{}, ic<32>{}, ic<64>{});
                case 128: return f(FamGaussPeaks{}, ic<32>{}, ic<128>{});
                case 224: return f(FamGaussPeaks{}, ic<32>{}, ic<224>{});
                case 448: return f(FamGaussPeaks{}, ic<32>{}, ic<448>{});
                case 480: return f(FamGaussPeaks{}, ic<32>{}, ic<480>{});
                default: return f(FamGaussPeaks{}, ic<32>{}, ic<512>{});   // 16 problems per SM: all the shared memory and registers
            }
        case ENLSIPB200_FAMILY_OSBORNE2: return f(FamOsborne2{}, ic<32>{}, ic<32>{});
        case ENLSIPB200_FAMILY_CHAINED_ROSENBROCK10: return f(FamChainedRosenbrock<10>{}, ic<32>{}, ic<32>{});
        case ENLSIPB200_FAMILY_CHAINED_WOOD20: return f(FamChainedWood<20>{}, ic<32>{}, ic<32>{});
    }
    return fail(ENLSIPB200_EINVAL, "unknown family id");
}

// lanes per problem / threads per CTA of each family
constexpr int HS_G = 1, HS_NT = 64;
constexpr int GP_G = 32;
static int gp_nt() { const char* e = getenv("ENLSIP_GP_NT"); int v = e ? atoi(e) : 512; return (v == 480 || v == 448 || v == 224 || v == 128 || v == 64) ? v : 512; }

Options make_options(const enlsipb200_options* o, int n, int m) {
    enlsipb200_options d;
    enlsipb200_default_options(&d);
    if (o) d = *o;
    double abs_tol = (d.abs_tol == d.abs_tol) ? d.abs_tol : EPS;
    double rel_tol = (d.rel_tol == d.rel_tol) ? d.rel_tol : sqrt(abs_tol);
    double c_tol = (d.c_tol == d.c_tol) ? d.c_tol : rel_tol;
    double x_tol = (d.x_tol == d.x_tol) ? d.x_tol : rel_tol;
    Options r;
    r.max_iter = d.max_iter;
    r.scaling = d.scaling;
    r.jac_mode = d.jac_mode;
    r.second_derivatives = (n + m < 1000) ? 1 : 0;   // EF:2658
    r.time_limit = d.time_limit;
    r.eps_abs = 1e-10;                               // EF:2651 (abs_tol is not forwarded)
    r.eps_rel = rel_tol;
    r.eps_x = x_tol;
    r.eps_c = c_tol;
    r.eps_rank = SQRT_EPS;                           // solver.jl:81
    return r;
}

}  // namespace

extern "C" {

int enlsipb200_version(void) { return 100; }

const char* enlsipb200_last_error(void) { return g_err.c_str(); }

void enlsipb200_default_options(enlsipb200_options* o) {
    if (!o) return;
    o->max_iter = 100;
    o->scaling = 0;
    o->jac_mode = ENLSIPB200_JAC_ANALYTIC;
    o->reserved = 0;
    o->time_limit = 1e3;
    o->abs_tol = NAN;
    o->rel_tol = NAN;
    o->c_tol = NAN;
    o->x_tol = NAN;
}

int enlsipb200_create(int family, const double* x_low, const double* x_upp, int device, enlsipb200_handle* out) {
    if (!out) return fail(ENLSIPB200_EINVAL, "out is NULL");
    *out = nullptr;
    FamilyInfo fi{0, 0, 0, 0, 0};
    int known = with_family(family, 0, [&](auto fam, auto, auto) {
        using Fm = decltype(fam);
        fi = {Fm::N, Fm::M, Fm::Q, Fm::NI, Fm::MAXB};
        return 0;
    });
    if (known != 0) return known;
    if (fi.maxb == 0 && ((x_low && [&] { for (int j = 0; j < fi.n; ++j) if (isfinite(x_low[j])) return true; return false; }()) ||
                         (x_upp && [&] { for (int j = 0; j < fi.n; ++j) if (isfinite(x_upp[j])) return true; return false; }())))
        return fail(ENLSIPB200_EINVAL, "this family is compiled without bound rows");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ENLSIPB200_ENOGPU, "no CUDA device: the ENLSIP engine has no CPU fallback");
    if (device < 0) CU(cudaGetDevice(&device));
    if (device >= ndev) return fail(ENLSIPB200_EINVAL, "device index out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(ENLSIPB200_ENOGPU, "sm_100a (B200) device required");
    auto* h = new enlsipb200_handle_s();
    h->family = family;
    h->device = device;
    h->fi = fi;
    h->num_sms = prop.multiProcessorCount;
    memset(&h->bnd, 0, sizeof(Bounds));
    for (int j = 0; j < fi.n; ++j)
        if (x_low && isfinite(x_low[j])) { h->bnd.lo_idx[h->bnd.nlo] = j; h->bnd.lo_val[h->bnd.nlo] = x_low[j]; h->bnd.nlo++; }
    for (int j = 0; j < fi.n; ++j)
        if (x_upp && isfinite(x_upp[j])) { h->bnd.up_idx[h->bnd.nup] = j; h->bnd.up_val[h->bnd.nup] = x_upp[j]; h->bnd.nup++; }
    h->l = fi.q + fi.ni + h->bnd.nlo + h->bnd.nup;
    if (h->l == 0) { delete h; return fail(ENLSIPB200_EINVAL, "There must be at least one constraint"); }  // cnls_model.jl:367
    int rc = 0;
    if (cudaMalloc(&h->counter, sizeof(unsigned long long)) != cudaSuccess) { delete h; return fail(ENLSIPB200_ENOMEM, "cudaMalloc"); }
    if (cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) { delete h; return fail(ENLSIPB200_ECUDA, "cudaEventCreate"); }
    if (family == ENLSIPB200_FAMILY_GAUSS_PEAKS) {   // the family's shared abscissa table (enl_families.h)
        double tt[128];
        for (int i = 0; i < 128; ++i) tt[i] = FamGaussPeaks::abscissa(i);
        if (cudaMemcpyToSymbol(g_gp_t, tt, sizeof(tt)) != cudaSuccess) { enlsipb200_destroy(h); return fail(ENLSIPB200_ECUDA, "cudaMemcpyToSymbol"); }
    }
    rc = with_family(family, gp_nt(), [&](auto fam, auto g_, auto nt_) {
        return configure<decltype(fam), decltype(g_)::value, decltype(nt_)::value>(h);
    });
    if (rc != 0) { enlsipb200_destroy(h); return rc; }
    *out = h;
    return 0;
}

int enlsipb200_destroy(enlsipb200_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->counter) cudaFree(h->counter);
    for (int i = 0; i < 3; ++i)
        if (h->owned[i]) cudaFree(h->owned[i]);
    if (h->stage) cudaFree(h->stage);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
    return 0;
}

int enlsipb200_dims(enlsipb200_handle h, int* n, int* m, int* nb_eq, int* nb_constraints, int* lmax) {
    if (!h) return fail(ENLSIPB200_EINVAL, "null handle");
    if (n) *n = h->fi.n;
    if (m) *m = h->fi.m;
    if (nb_eq) *nb_eq = h->fi.q;
    if (nb_constraints) *nb_constraints = h->l;
    if (lmax) *lmax = h->fi.q + h->fi.ni + h->fi.maxb;
    return 0;
}

int enlsipb200_set_data(enlsipb200_handle h, int slot, const double* ptr, long long count, int on_device, void* stream) {
    if (!h || slot < 0 || slot > 2) return fail(ENLSIPB200_EINVAL, "bad handle/slot");
    CU(cudaSetDevice(h->device));
    if (on_device) { h->data[slot] = ptr; return 0; }
    if (h->owned_count[slot] < count) {
        if (h->owned[slot]) CU(cudaFree(h->owned[slot]));
        h->owned[slot] = nullptr;
        CU(cudaMalloc(&h->owned[slot], (size_t)count * sizeof(double)));
        h->owned_count[slot] = count;
    }
    CU(cudaMemcpyAsync(h->owned[slot], ptr, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    h->data[slot] = h->owned[slot];
    return 0;
}

int enlsipb200_solve_batch(enlsipb200_handle h, long long B, const double* x0, const enlsipb200_options* opt, double* x,
                           double* f, int* exit_code, int* status, int* iters, int* nact, int* active, int* counters,
                           double* trace, int trace_cap, int on_device, void* stream) {
    if (!h) return fail(ENLSIPB200_EINVAL, "null handle");
    if (B < 0 || !x0 || !x || !f || !exit_code || !status || !iters || !nact)
        return fail(ENLSIPB200_EINVAL, "x0, x, f, exit_code, status, iters, nact are required");
    if ((h->family == ENLSIPB200_FAMILY_GAUSS_PEAKS || h->family == ENLSIPB200_FAMILY_OSBORNE2) && (!h->data[0] || !h->data[1]))
        return fail(ENLSIPB200_EINVAL, "this family needs data slots 0 and 1 (GAUSS_PEAKS: y, S; OSBORNE2: t, y)");
    if (B == 0) return 0;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = h->fi.n, lmax = h->fi.q + h->fi.ni + h->fi.maxb;
    const int row_w = TRACE_HDR + n;
    KernelArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B;
    a.counter = h->counter;
    a.fd = FamilyData{h->data[0], h->data[1], h->data[2]};
    a.opt = make_options(opt, h->fi.n, h->fi.m);
    a.bnd = h->bnd;
    a.out.trace_cap = trace ? trace_cap : 0;
    if (on_device) {
        a.x0 = x0;
        a.out.x = x; a.out.f = f; a.out.exit_code = exit_code; a.out.status = status; a.out.iters = iters;
        a.out.nact = nact; a.out.active = active; a.out.counters = counters; a.out.trace = trace;
    } else {
        // one staging allocation: [x0 | x | f | trace | ints...]
        size_t nd = (size_t)B * n * 2 + (size_t)B + (trace ? (size_t)B * trace_cap * row_w : 0);
        size_t ni = (size_t)B * (4 + lmax + 2);
        size_t bytes = nd * 8 + ni * 4;
        if (h->stage_bytes < bytes) {
            if (h->stage) CU(cudaFree(h->stage));
            h->stage = nullptr;
            h->stage_bytes = 0;
            if (cudaMalloc(&h->stage, bytes) != cudaSuccess) return fail(ENLSIPB200_ENOMEM, "cudaMalloc(staging)");
            h->stage_bytes = bytes;
        }
        double* dd = (double*)h->stage;
        double* d_x0 = dd; dd += (size_t)B * n;
        a.out.x = dd; dd += (size_t)B * n;
        a.out.f = dd; dd += (size_t)B;
        if (trace) { a.out.trace = dd; dd += (size_t)B * trace_cap * row_w; }
        int* di = (int*)dd;
        a.out.exit_code = di; di += B;
        a.out.status = di; di += B;
        a.out.iters = di; di += B;
        a.out.nact = di; di += B;
        a.out.active = di; di += (size_t)B * lmax;
        a.out.counters = di;
        CU(cudaMemcpyAsync(d_x0, x0, (size_t)B * n * 8, cudaMemcpyHostToDevice, st));
        if (trace) CU(cudaMemsetAsync(a.out.trace, 0, (size_t)B * trace_cap * row_w * 8, st));
        a.x0 = d_x0;
    }
    int rc = with_family(h->family, h->nt, [&](auto fam, auto g_, auto nt_) {
        return launch<decltype(fam), decltype(g_)::value, decltype(nt_)::value>(h, a, st);
    });
    if (rc != 0) return rc;
    if (!on_device) {
        CU(cudaMemcpyAsync(x, a.out.x, (size_t)B * n * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(f, a.out.f, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(exit_code, a.out.exit_code, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(status, a.out.status, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(iters, a.out.iters, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(nact, a.out.nact, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
        if (active) CU(cudaMemcpyAsync(active, a.out.active, (size_t)B * lmax * 4, cudaMemcpyDeviceToHost, st));
        if (counters) CU(cudaMemcpyAsync(counters, a.out.counters, (size_t)B * 2 * 4, cudaMemcpyDeviceToHost, st));
        if (trace) CU(cudaMemcpyAsync(trace, a.out.trace, (size_t)B * trace_cap * row_w * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    } else if (!stream) {
        CU(cudaStreamSynchronize(st));
    }
    return 0;
}

int enlsipb200_last_kernel_ms(enlsipb200_handle h, float* ms) {
    if (!h || !ms) return fail(ENLSIPB200_EINVAL, "null argument");
    if (!h->timed) return fail(ENLSIPB200_EINVAL, "no solve has been launched");
    CU(cudaSetDevice(h->device));
    CU(cudaEventSynchronize(h->ev1));
    CU(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return 0;
}

int enlsipb200_kernel_info(enlsipb200_handle h, int* regs, int* smem, int* nt, int* ctas_per_sm, int* grid, int* G) {
    if (!h) return fail(ENLSIPB200_EINVAL, "null handle");
    if (regs) *regs = h->regs;
    if (smem) *smem = h->smem;
    if (nt) *nt = h->nt;
    if (ctas_per_sm) *ctas_per_sm = h->ctas_per_sm;
    if (grid) *grid = h->grid;
    if (G) *G = h->G;
    return 0;
}

long long enlsipb200_launch_count(enlsipb200_handle h) { return h ? h->launches : 0; }

int enlsipb200_det_exp(const double* x, double* y, long long n, int on_device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(ENLSIPB200_ENOGPU, "no CUDA device: the ENLSIP engine has no CPU fallback");
    if (n <= 0) return 0;
    const double* dx = x;
    double* dy = y;
    double *tx = nullptr, *ty = nullptr;
    if (!on_device) {
        CU(cudaMalloc(&tx, n * 8));
        CU(cudaMalloc(&ty, n * 8));
        CU(cudaMemcpy(tx, x, n * 8, cudaMemcpyHostToDevice));
        dx = tx; dy = ty;
    }
    det_exp_kernel<<<(unsigned)((n + 255) / 256), 256>>>(dx, dy, n);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    if (!on_device) {
        CU(cudaMemcpy(y, ty, n * 8, cudaMemcpyDeviceToHost));
        cudaFree(tx); cudaFree(ty);
    }
    return 0;
}

}  // extern "C"

// enl_linalg.h -- Householder QR with column pivoting (LAPACK dgeqp3/dlaqp2 semantics, SURVEY.md
// section 10), application of the orthogonal factors (dorm2r) and triangular solves, in two
// flavours: "small" (replicated, every lane of the group runs it on the group's shared state)
// and "dist" (rows distributed over the lanes of the group, reductions by warp shuffles).
//
// Replaces, for the batched regime, the LAPACK calls reached by the reference through
// `qr(., ColumnNorm())` (EF:223, 700, 722-788), `F.Q' * v`, `F.Q * v`, `J * F.Q`
// (EF:135-151, 219, 478, 526) and the triangular `\` solves (EF:133-147, 480-501, 529).
#pragma once
#include "enl_base.h"

namespace enl {

#if defined(ENL_HOST_BUILD)
// CPU reference arm only (oracle/hostport, bench.py --impl reference): when a LAPACK dgeqp3 has been bound (OpenBLAS,
// the library Julia's `qr(., ColumnNorm())` calls), the two pivoted QR routines below hand the matrix to it instead of
// running the engine's own restatement -- the CPU baseline then executes the reference's algorithm, not the engine's.
// Never compiled into the CUDA library.
typedef void (*host_dgeqp3_fn)(const int*, const int*, double*, const int*, int*, double*, double*, const int*, int*);
inline host_dgeqp3_fn& host_dgeqp3() { static host_dgeqp3_fn fn = nullptr; return fn; }
inline void host_lapack_qrcp(double* a, int ld, int rows, int cols, double* tau, int* jpvt) {
    static thread_local double work[8192];
    for (int j = 0; j < cols; ++j) jpvt[j] = 0;
    int info = 0;
    const int lwork = 8192;
    host_dgeqp3()(&rows, &cols, a, &ld, jpvt, tau, work, &lwork, &info);
    for (int j = 0; j < cols; ++j) jpvt[j] -= 1;
}
#endif

// ------------------------------------------------------------------------------------------
// small (replicated) routines.  Matrices are column major: a(r,c) = a[c*ld + r].
// ------------------------------------------------------------------------------------------
template <class V>
ENL_NOINL double nrm2_small(V a, int n) {
    double s = 0.0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) s += a[i] * a[i];
    return sqrt(s);
}

// dlaqp2 on a rows x cols matrix; returns nothing, fills tau[min(rows,cols)], jpvt[cols] (0-based)
template <class V, class VI>
ENL_NOINL void qrcp_small(V a, int ld, int rows, int cols, V tau, VI jpvt, V vn1, V vn2) {
#if defined(ENL_HOST_BUILD)
    if (host_dgeqp3() && rows > 0 && cols > 0) { host_lapack_qrcp(&a[0], ld, rows, cols, &tau[0], &jpvt[0]); return; }
#endif
#pragma unroll 1
    for (int j = 0; j < cols; ++j) {
        jpvt[j] = j;
        double nj = nrm2_small(a.off(j * ld), rows);
        vn1[j] = nj;
        vn2[j] = nj;
    }
    int k = imin(rows, cols);
#pragma unroll 1
    for (int i = 0; i < k; ++i) {
        // pivot: first index of the largest partial norm (idamax)
        int pvt = i;
        double best = vn1[i];
#pragma unroll 1
        for (int j = i + 1; j < cols; ++j)
            if (vn1[j] > best) { best = vn1[j]; pvt = j; }
        if (pvt != i) {
#pragma unroll 1
            for (int r = 0; r < rows; ++r) {
                double tmp = a[pvt * ld + r];
                a[pvt * ld + r] = a[i * ld + r];
                a[i * ld + r] = tmp;
            }
            int it = jpvt[pvt]; jpvt[pvt] = jpvt[i]; jpvt[i] = it;
            vn1[pvt] = vn1[i];
            vn2[pvt] = vn2[i];
        }
        // dlarfg on a(i:rows-1, i)
        double tau_i = 0.0;
        if (i < rows - 1) {
            double alpha = a[i * ld + i];
            double xn = nrm2_small(a.off(i * ld + i + 1), rows - i - 1);
            if (xn != 0.0) {
                double beta = -sign_of(lapy2(alpha, xn), alpha);
                tau_i = (beta - alpha) / beta;
                double sc = 1.0 / (alpha - beta);
#pragma unroll 1
                for (int r = i + 1; r < rows; ++r) a[i * ld + r] *= sc;
                a[i * ld + i] = beta;
            }
        }
        tau[i] = tau_i;
        // apply H(i)' to a(i:rows-1, i+1:cols-1)
        if (i < cols - 1 && tau_i != 0.0) {
#pragma unroll 1
            for (int c = i + 1; c < cols; ++c) {
                double wv = a[c * ld + i];
#pragma unroll 1
                for (int r = i + 1; r < rows; ++r) wv += a[i * ld + r] * a[c * ld + r];
                wv *= tau_i;
                a[c * ld + i] -= wv;
#pragma unroll 1
                for (int r = i + 1; r < rows; ++r) a[c * ld + r] -= wv * a[i * ld + r];
            }
        }
        // partial column norm update
#pragma unroll 1
        for (int j = i + 1; j < cols; ++j) {
            double v1 = vn1[j];
            if (v1 != 0.0) {
                double tq = fabs(a[j * ld + i]) / v1;
                double temp = fmax(1.0 - tq * tq, 0.0);
                double rq = v1 / vn2[j];
                double temp2 = temp * (rq * rq);
                if (temp2 <= TOL3Z) {
                    if (i < rows - 1) {
                        double nj = nrm2_small(a.off(j * ld + i + 1), rows - i - 1);
                        vn1[j] = nj;
                        vn2[j] = nj;
                    } else {
                        vn1[j] = 0.0;
                        vn2[j] = 0.0;
                    }
                } else {
                    vn1[j] = v1 * sqrt(temp);
                }
            }
        }
    }
}

// v <- Q' v  (apply H(0), H(1), ..., H(k-1) in that order); factors a (rows x k), v length rows
template <class V>
ENL_NOINL void apply_qt_small(V a, int ld, int rows, int k, V tau, V v) {
#pragma unroll 1
    for (int i = 0; i < k; ++i) {
        double ti = tau[i];
        if (ti == 0.0) continue;
        double wv = v[i];
#pragma unroll 1
        for (int r = i + 1; r < rows; ++r) wv += a[i * ld + r] * v[r];
        wv *= ti;
        v[i] -= wv;
#pragma unroll 1
        for (int r = i + 1; r < rows; ++r) v[r] -= wv * a[i * ld + r];
    }
}

// v <- Q v  (apply H(k-1), ..., H(0))
template <class V>
ENL_NOINL void apply_q_small(V a, int ld, int rows, int k, V tau, V v) {
#pragma unroll 1
    for (int i = k - 1; i >= 0; --i) {
        double ti = tau[i];
        if (ti == 0.0) continue;
        double wv = v[i];
#pragma unroll 1
        for (int r = i + 1; r < rows; ++r) wv += a[i * ld + r] * v[r];
        wv *= ti;
        v[i] -= wv;
#pragma unroll 1
        for (int r = i + 1; r < rows; ++r) v[r] -= wv * a[i * ld + r];
    }
}

// solve R[0:k,0:k] x = b (upper, back substitution).  returns false on an exactly zero diagonal
template <class V, class V2>
ENL_NOINL bool solve_upper_small(V R, int ld, int k, V2 x) {
#pragma unroll 1
    for (int i = k - 1; i >= 0; --i) {
        double s = x[i];
#pragma unroll 1
        for (int j = i + 1; j < k; ++j) s -= R[j * ld + i] * x[j];
        double d = R[i * ld + i];
        if (d == 0.0) return false;
        x[i] = s / d;
    }
    return true;
}

// solve (R[0:k,0:k])' x = b (lower triangular = transpose of the stored upper factor)
template <class V, class V2>
ENL_NOINL bool solve_upperT_small(V R, int ld, int k, V2 x) {
#pragma unroll 1
    for (int i = 0; i < k; ++i) {
        double s = x[i];
#pragma unroll 1
        for (int j = 0; j < i; ++j) s -= R[i * ld + j] * x[j];
        double d = R[i * ld + i];
        if (d == 0.0) return false;
        x[i] = s / d;
    }
    return true;
}

// EF:17-31 on diag(R) of a column-major factor with `len` diagonal entries
template <class V>
ENL_NOINL int pseudo_rank(V R, int ld, int len, double eps_rank) {
    if (len <= 0 || fabs(R[0]) < eps_rank) return 0;
    double tol = fabs(R[0]) * sqrt((double)len) * eps_rank;
    int r = 1;
#pragma unroll 1
    while (r < len && fabs(R[(r - 1) * ld + (r - 1)]) > tol) ++r;
    return r - ((r == len && fabs(R[(r - 1) * ld + (r - 1)]) > tol) ? 0 : 1);
}

// ------------------------------------------------------------------------------------------
// distributed routines: rows of an M x ncols matrix spread over the G lanes of the group
// ------------------------------------------------------------------------------------------
template <class Grp, int G, int MS, int NT>
struct Dist {
    Grp g;   // lane / mask are recomputed from threadIdx: nothing per-lane is stored across calls

    // sum over rows >= r0 of a(:,c)^2
    ENL_NOINL double colsq(DM<G, MS, NT> a, int c, int r0) const {
        double s = 0.0;
#pragma unroll
        for (int sl = 0; sl < MS; ++sl) {
            int row = sl * G + g.lane;
            double v = a.at(sl, c);
            if (row >= r0) s += v * v;
        }
        return g.sum(s);
    }

    // J <- J * Q  where Q = H(0)...H(k-1) from small factors fa (n x k): per row, local
    template <class V>
    ENL_NOINL void mul_q_right(DM<G, MS, NT> J, int n, V fa, int ld, int k, V tau) const {
#pragma unroll 1
        for (int i = 0; i < k; ++i) {
            double ti = tau[i];
            if (ti == 0.0) continue;
#pragma unroll
            for (int sl = 0; sl < MS; ++sl) {
                double wv = J.at(sl, i);
#pragma unroll 1
                for (int c = i + 1; c < n; ++c) wv += J.at(sl, c) * fa[i * ld + c];
                wv *= ti;
                J.at(sl, i) -= wv;
#pragma unroll 1
                for (int c = i + 1; c < n; ++c) J.at(sl, c) -= wv * fa[i * ld + c];
            }
        }
    }

    // dlaqp2 on the M x ncols matrix `a` (M = number of real rows; padded rows are zero).
    // Writes tau2[min(M,ncols)], jpvt[ncols]; copies the leading kk x ncols upper trapezoid to
    // Rout (column major, ld = ldr) where kk = min(M, ncols).
    template <class V, class VI>
    ENL_NOINL void qrcp(DM<G, MS, NT> a, int M, int ncols, V tau2, VI jpvt, V vn1, V vn2, V Rout, int ldr) const {
#if defined(ENL_HOST_BUILD)
        if (host_dgeqp3() && G == 1 && NT == 1 && M > 0 && ncols > 0) {   // host layout: column major, ld = MS
            host_lapack_qrcp(&a.at(0, 0), MS, M, ncols, &tau2[0], &jpvt[0]);
            const int kk = imin(M, ncols);
            for (int c = 0; c < ncols; ++c)
                for (int r = 0; r < kk; ++r) Rout[c * ldr + r] = (r <= c) ? a.row(r, c) : 0.0;
            return;
        }
#endif
#pragma unroll 1
        for (int j = 0; j < ncols; ++j) {
            jpvt[j] = j;
            double nj = sqrt(colsq(a, j, 0));
            vn1[j] = nj;
            vn2[j] = nj;
        }
        int k = imin(M, ncols);
#pragma unroll 1
        for (int i = 0; i < k; ++i) {
            int pvt = i;
            double best = vn1[i];
#pragma unroll 1
            for (int j = i + 1; j < ncols; ++j)
                if (vn1[j] > best) { best = vn1[j]; pvt = j; }
            if (pvt != i) {
#pragma unroll
                for (int sl = 0; sl < MS; ++sl) {
                    double tmp = a.at(sl, pvt);
                    a.at(sl, pvt) = a.at(sl, i);
                    a.at(sl, i) = tmp;
                }
                int it = jpvt[pvt]; jpvt[pvt] = jpvt[i]; jpvt[i] = it;
                vn1[pvt] = vn1[i];
                vn2[pvt] = vn2[i];
            }
            g.sync();
            double tau_i = 0.0;
            if (i < M - 1) {
                double alpha = a.row(i, i);
                double xn = sqrt(colsq(a, i, i + 1));
                if (xn != 0.0) {
                    double beta = -sign_of(lapy2(alpha, xn), alpha);
                    tau_i = (beta - alpha) / beta;
                    double sc = 1.0 / (alpha - beta);
                    g.sync();   // everyone has read alpha before the owner overwrites it
#pragma unroll
                    for (int sl = 0; sl < MS; ++sl) {
                        int row = sl * G + g.lane;
                        if (row > i) a.at(sl, i) *= sc;
                        else if (row == i) a.at(sl, i) = beta;
                    }
                }
            }
            tau2[i] = tau_i;
            if (i < ncols - 1 && tau_i != 0.0) {
#pragma unroll 1
                for (int c = i + 1; c < ncols; ++c) {
                    double part = 0.0;
#pragma unroll
                    for (int sl = 0; sl < MS; ++sl) {
                        int row = sl * G + g.lane;
                        double vv = (row > i) ? a.at(sl, i) : ((row == i) ? 1.0 : 0.0);
                        part += vv * a.at(sl, c);
                    }
                    double wv = g.sum(part) * tau_i;
#pragma unroll
                    for (int sl = 0; sl < MS; ++sl) {
                        int row = sl * G + g.lane;
                        if (row > i) a.at(sl, c) -= wv * a.at(sl, i);
                        else if (row == i) a.at(sl, c) -= wv;
                    }
                }
            }
            g.sync();
#pragma unroll 1
            for (int j = i + 1; j < ncols; ++j) {
                double v1 = vn1[j];
                if (v1 != 0.0) {
                    double tq = fabs(a.row(i, j)) / v1;
                    double temp = fmax(1.0 - tq * tq, 0.0);
                    double rq = v1 / vn2[j];
                    double temp2 = temp * (rq * rq);
                    if (temp2 <= TOL3Z) {
                        if (i < M - 1) {
                            double nj = sqrt(colsq(a, j, i + 1));
                            vn1[j] = nj;
                            vn2[j] = nj;
                        } else {
                            vn1[j] = 0.0;
                            vn2[j] = 0.0;
                        }
                    } else {
                        vn1[j] = v1 * sqrt(temp);
                    }
                }
            }
        }
        g.sync();
#pragma unroll 1
        for (int c = 0; c < ncols; ++c)
#pragma unroll 1
            for (int r = 0; r < k; ++r) Rout[c * ldr + r] = (r <= c) ? a.row(r, c) : 0.0;
    }

    // d <- Q' d for the distributed factor `a` with k reflectors; d is a distributed vector
    // (DM with one column)
    template <class V>
    ENL_NOINL void apply_qt(DM<G, MS, NT> a, int k, V tau2, DM<G, MS, NT> d) const {
#pragma unroll 1
        for (int i = 0; i < k; ++i) {
            double ti = tau2[i];
            if (ti == 0.0) continue;
            double part = 0.0;
#pragma unroll
            for (int sl = 0; sl < MS; ++sl) {
                int row = sl * G + g.lane;
                double vv = (row > i) ? a.at(sl, i) : ((row == i) ? 1.0 : 0.0);
                part += vv * d.at(sl, 0);
            }
            double wv = g.sum(part) * ti;
#pragma unroll
            for (int sl = 0; sl < MS; ++sl) {
                int row = sl * G + g.lane;
                if (row > i) d.at(sl, 0) -= wv * a.at(sl, i);
                else if (row == i) d.at(sl, 0) -= wv;
            }
        }
        g.sync();
    }

    // v <- Q v for a register-resident distributed vector (v[s] = row s*G+lane): H(k-1) ... H(0)
    template <class V>
    ENL_NOINL void apply_q_regs(DM<G, MS, NT> a, int k, V tau2, double* v) const {
#pragma unroll 1
        for (int i = k - 1; i >= 0; --i) {
            double ti = tau2[i];
            if (ti == 0.0) continue;
            double part = 0.0;
#pragma unroll
            for (int sl = 0; sl < MS; ++sl) {
                int row = sl * G + g.lane;
                double vv = (row > i) ? a.at(sl, i) : ((row == i) ? 1.0 : 0.0);
                part += vv * v[sl];
            }
            double wv = g.sum(part) * ti;
#pragma unroll
            for (int sl = 0; sl < MS; ++sl) {
                int row = sl * G + g.lane;
                double vv = (row > i) ? a.at(sl, i) : ((row == i) ? 1.0 : 0.0);
                v[sl] -= wv * vv;
            }
        }
    }

    // sum_{row < len} d(row)^2
    ENL_NOINL double prefix_sq(DM<G, MS, NT> d, int len) const {
        double s = 0.0;
#pragma unroll
        for (int sl = 0; sl < MS; ++sl) {
            int row = sl * G + g.lane;
            double v = d.at(sl, 0);
            if (row < len) s += v * v;
        }
        return g.sum(s);
    }
};

}  // namespace enl

"""CPU oracle for the ENLSIP hot path -- TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (numpy + SciPy's LAPACK, i.e. the same reference-LAPACK
``dgeqp3`` / ``dormqr`` / ``dtrtrs`` / ``dpotrf`` algorithms that Julia's LinearAlgebra calls)
of the algorithm in the reference ``src/enlsip_functions.jl`` (abbreviated EF below),
``src/structures.jl``, ``src/cnls_model.jl`` and the option plumbing of ``src/solver.jl``.

It is the *checker* for the CUDA engine.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product path
(``enlsip.jl_b200``) never does.

PARITY UNPINNED: the reference is pure Julia, no ``julia`` binary exists in this image (nor on
the GPU box), and the reference's own tests pin no iterates / iteration counts / exit codes
(SURVEY.md section 8c).  The oracle is pinned only by (i) the integer working-set known-answer
test of ``test/internal/working_set.jl``, (ii) the HS65 optimum/objective published in
``docs/src/tutorial.md:126-128`` and (iii) the ``time_limit=-1`` status of
``test/problems/chained_rosenbrock.jl:71-73``.  Fidelity to Enlsip.jl is otherwise by
construction: every function cites the reference lines it follows, and the quirks listed in
SURVEY.md section 9 (aliasing of ``Iteration.rx/cx``, always-reverted first-order deletions,
``min_norm_w!`` restarting from ``K[4]`` ...) are reproduced, not fixed.

Indices: constraint ids and list positions are kept 1-based exactly as in the reference
(``active``/``inactive`` arrays hold 1-based ids, 0 = empty slot) so traces can be compared
with a Julia run verbatim.  Python containers are of course 0-based; helpers convert.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np
from scipy.linalg import lapack

EPS = float(np.finfo(np.float64).eps)
SQRT_EPS = math.sqrt(EPS)
F = np.float64


class ReferenceWouldThrow(Exception):
    """Raised where the Julia code would raise (BoundsError, destructuring error, ...)."""


class ReferenceWouldHang(Exception):
    """Raised where the Julia code provably never terminates (see evaluate_violated_constraints)."""


EXIT_WOULD_THROW = -99   # engine/oracle convention: the reference raises an exception here
EXIT_WOULD_HANG = -98    # engine/oracle convention: the reference loops forever here


# --------------------------------------------------------------------------------------
# Pivoted QR objects (Julia ``qr(M, ColumnNorm())`` -> LAPACK geqp3)
# --------------------------------------------------------------------------------------
class QRPivoted:
    """``qr(M, ColumnNorm())`` of the reference (EF:223, 700, 722, 724, ...).

    ``R = triu(factors[1:min(m,n), :])``, ``p`` = jpvt (kept 0-based here), ``Q`` is the full
    square orthogonal factor applied with ``dormqr``.
    """

    def __init__(self, M):
        M = np.array(M, dtype=F, order="F", copy=True)
        if M.ndim != 2:
            raise ValueError("matrix expected")
        self.m, self.n = M.shape
        self.k = min(self.m, self.n)
        if self.k == 0:
            self.factors = M
            self.tau = np.zeros(0)
            self.p = np.arange(self.n)
        else:
            qr, jpvt, tau, _work, info = lapack.dgeqp3(M)
            if info != 0:
                raise RuntimeError("dgeqp3 info=%d" % info)
            self.factors = np.asfortranarray(qr)
            self.tau = tau
            self.p = np.asarray(jpvt, dtype=np.int64) - 1
        self.R = np.triu(self.factors[: self.k, :])

    def diagR(self):
        return np.array([self.R[i, i] for i in range(self.k)], dtype=F)

    def invperm(self):
        ip = np.empty_like(self.p)
        ip[self.p] = np.arange(self.p.size)
        return ip

    def Pmat(self):
        # F.P : A*P = A[:, p]  ->  P[p[j], j] = 1
        P = np.zeros((self.n, self.n))
        for j in range(self.n):
            P[self.p[j], j] = 1.0
        return P

    def _ormqr(self, side, trans, C):
        if self.k == 0:
            return np.array(C, dtype=F, copy=True)
        C = np.array(C, dtype=F, order="F", copy=True)
        a = np.asfortranarray(self.factors[:, : self.k])
        lwork = max(1, 64 * max(C.shape))
        cq, _w, info = lapack.dormqr(side, trans, a, self.tau, C, lwork)
        if info != 0:
            raise RuntimeError("dormqr info=%d" % info)
        return cq

    def Qt_mul(self, v):  # F.Q' * v
        v = np.asarray(v, dtype=F)
        if v.size == 0:      # empty working set: Julia multiplies 0-length vectors without complaint
            return v.copy()
        out = self._ormqr("L", "T", v.reshape(self.m, -1))
        return out.reshape(v.shape)

    def Q_mul(self, v):  # F.Q * v
        v = np.asarray(v, dtype=F)
        if v.size == 0:
            return v.copy()
        out = self._ormqr("L", "N", v.reshape(self.m, -1))
        return out.reshape(v.shape)

    def mul_Q(self, M):  # M * F.Q
        return self._ormqr("R", "N", np.asarray(M, dtype=F))


def _solve_upper(R, b):
    """``UpperTriangular(R) \\ b`` (LAPACK trtrs)."""
    k = R.shape[0]
    if k == 0:
        return np.zeros(0)
    x, info = lapack.dtrtrs(np.asfortranarray(R), np.array(b, dtype=F), lower=0, trans=0)
    if info != 0:
        # Julia throws SingularException for an exactly zero diagonal entry
        raise ReferenceWouldThrow("SingularException in upper triangular solve")
    return x


def _solve_lower(L, b):
    """``LowerTriangular(L) \\ b``."""
    k = L.shape[0]
    if k == 0:
        return np.zeros(0)
    x, info = lapack.dtrtrs(np.asfortranarray(L), np.array(b, dtype=F), lower=1, trans=0)
    if info != 0:
        raise ReferenceWouldThrow("SingularException in lower triangular solve")
    return x


def _norm(v):
    v = np.asarray(v, dtype=F)
    if v.size == 0:
        return 0.0
    return float(np.sqrt(np.dot(v, v))) if np.all(np.isfinite(v)) else float(np.linalg.norm(v))


def _dot(a, b):
    return float(np.dot(np.asarray(a, dtype=F), np.asarray(b, dtype=F)))


def _jl_range(v, k):
    """Julia ``v[1:k]`` : empty if k <= 0, BoundsError if k > length(v)."""
    if k <= 0:
        return v[:0]
    if k > len(v):
        raise ReferenceWouldThrow("BoundsError: [1:%d] of length-%d vector" % (k, len(v)))
    return v[:k]


# --------------------------------------------------------------------------------------
# Problem description: the four callbacks of cnls_model.jl:11-62
# --------------------------------------------------------------------------------------
@dataclass
class Problem:
    """The evaluation surface the solver sees: ``r``, ``J``, ``c``, ``A`` (cnls_model.jl:40-62).

    ``cons`` returns the stacked vector ``[eq; ineq; x-x_low (finite); x_upp-x (finite)]`` and
    ``jac_cons`` the matching rows (cnls_model.jl:402-403, 416-420).
    """

    n: int
    m: int
    q: int
    l: int
    res: Callable
    jac_res: Callable
    cons: Callable
    jac_cons: Callable
    x0: np.ndarray = None
    name: str = ""
    nb_reseval: int = 0
    nb_jacres: int = 0
    nb_conseval: int = 0
    nb_jaccons: int = 0

    def reset_counters(self):
        self.nb_reseval = self.nb_jacres = self.nb_conseval = self.nb_jaccons = 0


def jac_forward_diff(h, x):
    """cnls_model.jl:65-82 : ``delta_j = max(abs(x_j),1)*sqrt(eps)``, forward difference."""
    x = np.asarray(x, dtype=F)
    delta = SQRT_EPS
    hx = np.asarray(h(x), dtype=F)
    n = x.size
    Jh = np.zeros((hx.size, n))
    for j in range(n):
        dj = max(abs(x[j]), 1.0) * delta
        e = np.zeros(n)
        e[j] = 1.0
        xf = x + dj * e
        Jh[:, j] = (np.asarray(h(xf), dtype=F) - hx) / dj
    return Jh


def make_problem(n, m, res, jac_res=None, eq=None, jac_eq=None, nb_eq=0, ineq=None, jac_ineq=None,
                 nb_ineq=0, x_low=None, x_upp=None, x0=None, name="", fd=False):
    """Stack constraints exactly like instantiate_constraints_{w,wo}_bounds (cnls_model.jl:410-496).

    Missing Jacobians: the reference uses ForwardDiff AD (= analytic to rounding).  Here a
    missing Jacobian must either be supplied analytically or ``fd=True`` selects
    ``jac_forward_diff`` (the north-star FD mode).
    """
    x_low = np.full(n, -np.inf) if x_low is None else np.asarray(x_low, dtype=F)
    x_upp = np.full(n, np.inf) if x_upp is None else np.asarray(x_upp, dtype=F)
    lo_idx = [i for i in range(n) if np.isfinite(x_low[i])]
    up_idx = [i for i in range(n) if np.isfinite(x_upp[i])]
    nb_lo, nb_up = len(lo_idx), len(up_idx)
    l = nb_eq + nb_ineq + nb_lo + nb_up
    if l == 0:
        raise AssertionError("There must be at least one constraint")  # cnls_model.jl:367

    def need(fun, jac, what):
        if jac is not None:
            return jac
        if fd:
            return lambda x: jac_forward_diff(fun, x)
        raise ValueError("analytic Jacobian of %s required (no AD in the oracle)" % what)

    jr = need(res, jac_res, "residuals")
    jeq = need(eq, jac_eq, "equalities") if eq is not None else None
    jin = need(ineq, jac_ineq, "inequalities") if ineq is not None else None
    eye = np.eye(n)

    def cons(x):
        x = np.asarray(x, dtype=F)
        parts = []
        if eq is not None:
            parts.append(np.asarray(eq(x), dtype=F).reshape(-1))
        if ineq is not None:
            parts.append(np.asarray(ineq(x), dtype=F).reshape(-1))
        if nb_lo:
            parts.append((x - x_low)[lo_idx])
        if nb_up:
            parts.append((x_upp - x)[up_idx])
        return np.concatenate(parts)

    def jac_cons(x):
        x = np.asarray(x, dtype=F)
        parts = []
        if eq is not None:
            parts.append(np.asarray(jeq(x), dtype=F).reshape(nb_eq, n))
        if ineq is not None:
            parts.append(np.asarray(jin(x), dtype=F).reshape(nb_ineq, n))
        if nb_lo:
            parts.append(eye[lo_idx, :])
        if nb_up:
            parts.append(-eye[up_idx, :])
        return np.vstack(parts)

    return Problem(n=n, m=m, q=nb_eq, l=l, res=lambda x: np.asarray(res(np.asarray(x, dtype=F)), dtype=F),
                   jac_res=lambda x: np.asarray(jr(np.asarray(x, dtype=F)), dtype=F).reshape(m, n),
                   cons=cons, jac_cons=jac_cons,
                   x0=None if x0 is None else np.array(x0, dtype=F), name=name)


# --------------------------------------------------------------------------------------
# structures.jl
# --------------------------------------------------------------------------------------
@dataclass
class Iteration:
    """structures.jl:63-98.  ``rx``/``cx`` may alias the live buffers."""

    x: np.ndarray
    p: np.ndarray
    rx: np.ndarray
    cx: np.ndarray
    t: int
    alpha: float
    index_alpha_upp: int
    lam: np.ndarray
    w: np.ndarray
    rankA: int
    rankJ2: int
    dimA: int
    dimJ2: int
    b_gn: np.ndarray
    d_gn: np.ndarray
    predicted_reduction: float
    progress: float
    grad_res: float
    speed: float
    beta: float
    restart: bool
    first: bool
    add: bool
    dele: bool
    index_del: int
    code: int
    nb_newton_steps: int

    def copy(self):  # structures.jl:93-98 : deep copies of the vectors
        return Iteration(self.x.copy(), self.p.copy(), self.rx.copy(), self.cx.copy(), self.t, self.alpha,
                         self.index_alpha_upp, self.lam.copy(), self.w.copy(), self.rankA, self.rankJ2,
                         self.dimA, self.dimJ2, self.b_gn.copy(), self.d_gn.copy(), self.predicted_reduction,
                         self.progress, self.grad_res, self.speed, self.beta, self.restart, self.first,
                         self.add, self.dele, self.index_del, self.code, self.nb_newton_steps)


@dataclass
class Constraint:
    """structures.jl:145-150 : active block."""

    cx: np.ndarray
    A: np.ndarray
    scaling: bool
    diag_scale: np.ndarray


def evaluate_scaling(C: Constraint):
    """structures.jl:160-178."""
    t = C.A.shape[0]
    C.diag_scale = np.zeros(t)
    for i in range(t):
        row_i = _norm(C.A[i, :])
        C.diag_scale[i] = row_i
        if C.scaling:
            if abs(row_i) < EPS:
                row_i = 1.0
            C.A[i, :] = C.A[i, :] / row_i
            C.cx[i] = C.cx[i] / row_i
            C.diag_scale[i] = 1.0 / row_i


class WorkingSet:
    """structures.jl:209-267.  ``active``/``inactive`` hold 1-based constraint ids, 0 = empty."""

    def __init__(self, q, l, t=None, active=None, inactive=None):
        self.q, self.l = q, l
        if active is None:  # structures.jl:223-229
            self.t = q
            self.active = np.zeros(l, dtype=np.int64)
            self.inactive = np.zeros(l - q, dtype=np.int64)
            self.active[:q] = np.arange(1, q + 1)
            self.inactive[:] = np.arange(q + 1, l + 1)
        else:
            self.t = t
            self.active = np.array(active, dtype=np.int64)
            self.inactive = np.array(inactive, dtype=np.int64)

    def remove_constraint(self, s):
        """structures.jl:234-249 : ``s`` is a 1-based POSITION in ``active``."""
        l, t = self.l, self.t
        self.inactive[l - t] = self.active[s - 1]
        self.inactive[: l - t + 1] = np.sort(self.inactive[: l - t + 1])
        for i in range(s, t):
            self.active[i - 1] = self.active[i]
        self.active[t - 1] = 0
        self.t -= 1

    def add_constraint(self, s):
        """structures.jl:254-267 : ``s`` is a 1-based POSITION in ``inactive``."""
        l, t = self.l, self.t
        self.active[t] = self.inactive[s - 1]
        self.active[: t + 1] = np.sort(self.active[: t + 1])
        for i in range(s, l - t):
            self.inactive[i - 1] = self.inactive[i]
        self.inactive[l - t - 1] = 0
        self.t += 1

    def act(self):  # 0-based indices of the active constraints, in list order
        return self.active[: self.t] - 1

    def inact(self):
        return self.inactive[: self.l - self.t] - 1


# --------------------------------------------------------------------------------------
# EF:17-31
# --------------------------------------------------------------------------------------
def pseudo_rank(diag_T, eps_rank):
    diag_T = np.asarray(diag_T, dtype=F)
    if diag_T.size == 0 or abs(diag_T[0]) < eps_rank:
        return 0
    l_diag = diag_T.size
    tol = abs(diag_T[0]) * math.sqrt(float(l_diag)) * eps_rank
    r = 1
    while r < l_diag and abs(diag_T[r - 1]) > tol:
        r += 1
    return r - (0 if (r == l_diag and abs(diag_T[r - 1]) > tol) else 1)


class Evaluator:
    """res_eval!/jacres_eval!/cons_eval!/jaccons_eval! (cnls_model.jl:40-62): in-place + counters."""

    def __init__(self, prob: Problem):
        self.prob = prob

    def res(self, x, out):
        out[:] = self.prob.res(x)
        self.prob.nb_reseval += 1

    def jacres(self, x, out):
        out[:, :] = self.prob.jac_res(x)
        self.prob.nb_jacres += 1

    def cons(self, x, out):
        out[:] = self.prob.cons(x)
        self.prob.nb_conseval += 1

    def jaccons(self, x, out):
        out[:, :] = self.prob.jac_cons(x)
        self.prob.nb_jaccons += 1


def new_point(ev: Evaluator, x, rx, cx, J, A):  # EF:34-52
    ev.res(x, rx)
    ev.jacres(x, J)
    ev.cons(x, cx)
    ev.jaccons(x, A)


# --------------------------------------------------------------------------------------
# EF:116-234 search directions
# --------------------------------------------------------------------------------------
def sub_search_direction(J1, rx, cx, F_A, F_L11, F_J2, n, t, rankA, dimA, dimJ2, code):
    """EF:116-153."""
    if code == 1:
        b = -cx[F_A.p]
        p1 = _solve_lower(F_A.R.T[:, :], b) if b.size else np.zeros(0)
        d_temp = -(J1 @ p1) - rx
        d = F_J2.Qt_mul(d_temp)
        dp2 = _solve_upper(F_J2.R[:dimJ2, :dimJ2], _jl_range(d, dimJ2))
        if n - t - dimJ2 < 0:
            raise ReferenceWouldThrow("negative zeros() length")
        p2 = np.concatenate([dp2, np.zeros(n - t - dimJ2)])[F_J2.invperm()]
    elif code == -1:
        b_buff = -cx[F_A.p]
        b = F_L11.Qt_mul(b_buff)
        if dimA > t or dimA > min(F_L11.R.shape):
            raise ReferenceWouldThrow("BoundsError: R11[1:dimA,1:dimA]")
        dp1 = _solve_upper(F_L11.R[:dimA, :dimA], _jl_range(b, dimA))
        p1 = np.concatenate([dp1, np.zeros(t - dimA)])[F_L11.invperm()][:rankA]
        d_temp = -(J1 @ p1) - rx
        d = F_J2.Qt_mul(d_temp)
        if dimA > t or dimA > F_L11.R.shape[0] or dimJ2 > F_J2.R.shape[0] or dimJ2 > F_J2.R.shape[1]:
            raise ReferenceWouldThrow("BoundsError: R[1:dim,1:dim]")
        dp2 = _solve_upper(F_J2.R[:dimJ2, :dimJ2], _jl_range(d, dimJ2))
        if n - rankA - dimJ2 < 0:
            raise ReferenceWouldThrow("negative zeros() length")
        p2 = np.concatenate([dp2, np.zeros(n - rankA - dimJ2)])[F_J2.invperm()]
    else:
        raise ReferenceWouldThrow("sub_search_direction: code %r" % code)
    y = np.concatenate([p1, p2])
    if y.size != n:
        raise ReferenceWouldThrow("DimensionMismatch in Q1*[p1;p2]")
    p = F_A.Q_mul(y)
    return p, b, d


def gn_search_direction(J, rx, cx, F_A, F_L11, rankA, t, eps_rank, it: Iteration):
    """EF:206-234."""
    code = 1 if rankA == t else -1
    n = J.shape[1]
    JQ1 = F_A.mul_Q(J)
    J1, J2 = JQ1[:, :rankA], JQ1[:, rankA:]
    F_J2 = QRPivoted(J2)
    rankJ2 = pseudo_rank(F_J2.diagR(), eps_rank)
    p_gn, b_gn, d_gn = sub_search_direction(J1, rx, cx, F_A, F_L11, F_J2, n, t, rankA, rankA, rankJ2, code)
    it.rankA = rankA
    it.rankJ2 = rankJ2
    it.dimA = rankA
    it.dimJ2 = rankJ2
    it.b_gn = b_gn
    it.d_gn = d_gn
    return p_gn, F_J2


# --------------------------------------------------------------------------------------
# EF:243-423 Newton direction
# --------------------------------------------------------------------------------------
def hessian_res(ev: Evaluator, x, rx, n, m, B):  # EF:243-278
    e1 = EPS ** (1.0 / 3.0)
    f1, f2, f3, f4 = np.zeros(m), np.zeros(m), np.zeros(m), np.zeros(m)
    for k in range(n):
        for j in range(k + 1):
            ek = max(abs(x[k]), 1.0) * e1
            ej = max(abs(x[j]), 1.0) * e1
            xw = x.copy(); xw[j] += ej; xw[k] += ek; ev.res(xw, f1)
            xw = x.copy(); xw[j] -= ej; xw[k] += ek; ev.res(xw, f2)
            xw = x.copy(); xw[j] += ej; xw[k] -= ek; ev.res(xw, f3)
            xw = x.copy(); xw[j] -= ej; xw[k] -= ek; ev.res(xw, f4)
            s = 0.0
            for i in range(m):
                s += (f1[i] - f2[i] - f3[i] + f4[i]) * rx[i]
            s /= (4 * ej * ek)
            B[k, j] = s
            if j != k:
                B[j, k] = s


def hessian_cons(ev: Evaluator, x, lam, active, n, l, t, B):  # EF:288-328
    e1 = EPS ** (1.0 / 3.0)
    f1, f2, f3, f4 = np.zeros(l), np.zeros(l), np.zeros(l), np.zeros(l)
    for k in range(n):
        for j in range(k + 1):
            ek = max(abs(x[k]), 1.0) * e1
            ej = max(abs(x[j]), 1.0) * e1
            xw = x.copy(); xw[j] += ej; xw[k] += ek; ev.cons(xw, f1)
            xw = x.copy(); xw[j] -= ej; xw[k] += ek; ev.cons(xw, f2)
            xw = x.copy(); xw[j] += ej; xw[k] -= ek; ev.cons(xw, f3)
            xw = x.copy(); xw[j] -= ej; xw[k] -= ek; ev.cons(xw, f4)
            s = 0.0
            for i in range(t):
                ii = active[i] - 1
                s += (f1[ii] - f2[ii] - f3[ii] + f4[ii]) * lam[i]
            s /= (4.0 * ek * ej)
            B[k, j] = s
            if k != j:
                B[j, k] = s


def newton_search_direction(x, ev, active_cx, W: WorkingSet, lam, rx, J, F_A, F_L11, rankA):
    """EF:348-423.  Returns (p, error)."""
    m, n = J.shape
    active = W.active
    t, l = W.t, W.l
    if t == rankA:
        b = -active_cx[F_A.p]
        p1 = _solve_lower(F_A.R.T[:, :], b) if b.size else np.zeros(0)
    else:  # t > rankA
        b = F_L11.Qt_mul(-active_cx[F_A.p])
        dp1 = _solve_upper(F_L11.R[:rankA, :rankA], b[:rankA])
        p1 = F_L11.Pmat()[:rankA, :rankA] @ dp1
    if rankA == n:
        # EF:379-381 returns a bare vector where the caller destructures (p, error): with n >= 2
        # Julia would bind p=p1[1], error=p1[2] and fail later on a non-Bool; n==1 throws.
        raise ReferenceWouldThrow("newton_search_direction with rankA == n")
    JQ1 = F_A.mul_Q(J)
    J1, J2 = JQ1[:, :rankA], JQ1[:, rankA:]
    r_mat, c_mat = np.zeros((n, n)), np.zeros((n, n))
    hessian_res(ev, x, rx, n, m, r_mat)
    hessian_cons(ev, x, lam, active, n, l, t, c_mat)
    G = r_mat - c_mat
    E = F_A.Qt_mul(F_A.mul_Q(G)) if True else None  # Q' * G * Q  (left-to-right: (Q'G)Q)
    # Julia evaluates F_A.Q' * G * F_A.Q as (Q' * G) * Q
    E = F_A.mul_Q(F_A.Qt_mul(G))
    if t > rankA:
        P2 = F_L11.p
        if P2.size != n:
            raise ReferenceWouldThrow("E[vect_P2, vect_P2] with length(P2) != n")
        E = E[np.ix_(P2, P2)]
    E21 = E[rankA:n, :rankA]
    E22 = E[rankA:n, rankA:n]
    W22 = E22 + J2.T @ J2
    W21 = E21 + J2.T @ J1
    d = -(W21 @ p1) - J2.T @ rx
    sW22 = (W22 + W22.T) * 0.5
    c, info = lapack.dpotrf(np.asfortranarray(sW22), lower=0)
    if info == 0:
        U = np.triu(c)
        y = _solve_lower(U.T.copy(), d)
        p2 = _solve_upper(U, y)
        p = F_A.Q_mul(np.concatenate([p1, p2]))
        return p, False
    return np.zeros(n), True


# --------------------------------------------------------------------------------------
# EF:461-603 multipliers
# --------------------------------------------------------------------------------------
def first_lagrange_mult_estimate(A, lam, gradf, cx, scaling_done, diag_scale, Fq: QRPivoted, it: Iteration, eps_rank):
    """EF:461-508 (in place on ``lam``)."""
    t, n = A.shape
    v = np.zeros(t)
    inv_p = Fq.invperm()
    prankA = pseudo_rank(Fq.diagR(), eps_rank)
    b = Fq.Qt_mul(gradf)
    R = Fq.R
    v[:prankA] = _solve_upper(R[:prankA, :prankA], b[:prankA])
    lam_ls = v[inv_p]
    it.grad_res = _norm(b[prankA:n]) if n > prankA else 0.0
    b = -cx[Fq.p]
    y = np.zeros(t)
    y[:prankA] = _solve_lower(R.T[:prankA, :prankA], b[:prankA])
    u = np.zeros(t)
    u[:prankA] = _solve_upper(R[:prankA, :prankA], y[:prankA])
    lam[:] = lam_ls + u[inv_p]
    if scaling_done:
        lam[:] = lam * diag_scale


def second_lagrange_mult_estimate(J, F_A, lam, rx, p_gn, t, scaling, diag_scale, eps_rank=SQRT_EPS):
    """EF:514-537."""
    prankA = pseudo_rank(F_A.diagR(), eps_rank)
    J1 = F_A.mul_Q(J)[:, :t]
    b = J1.T @ (rx + J @ p_gn)
    v = np.zeros(t)
    v[:prankA] = _solve_upper(F_A.R[:prankA, :prankA], b[:prankA])
    lam[:] = v[F_A.invperm()]
    if scaling:
        lam[:] = lam * diag_scale


def minmax_lagrangian_mult(lam, W: WorkingSet, C: Constraint):
    """EF:540-564."""
    q, t = W.q, W.t
    sq_rel = SQRT_EPS
    lam_abs_max = 0.0
    sigmin = math.inf
    if t > q:
        lam_abs_max = float(np.max(np.abs(lam)))
        rows = (1.0 / C.diag_scale) if C.scaling else C.diag_scale
        for i in range(q, t):
            li = lam[i]
            if li * rows[i] <= -sq_rel and li < sigmin:
                sigmin = float(li)
    return sigmin, lam_abs_max


def check_constraint_deletion(q, A, lam, scaling, diag_scale, grad_res):
    """EF:574-603.  Returns a 1-based position (0 = none)."""
    t = A.shape[0]
    delta = 10.0
    lam_max = 1.0 if lam.size == 0 else float(np.max(np.abs(lam)))
    sq_rel = SQRT_EPS * lam_max
    s = 0
    if t > q:
        e = sq_rel
        for i in range(q + 1, t + 1):
            row_i = (1.0 / diag_scale[i - 1]) if scaling else diag_scale[i - 1]
            if row_i * lam[i - 1] <= sq_rel and row_i * lam[i - 1] <= e:
                e = row_i * lam[i - 1]
                s = i
        if grad_res > -e * delta:
            s = 0
    return s


def evaluate_violated_constraints(cx, W: WorkingSet, index_alpha_upp, n):
    """EF:608-650."""
    eps_ = SQRT_EPS
    delta = 0.1
    bnd = min(W.l, n)
    added = False
    swaps = 0
    if W.l > W.t:
        i = 1
        while i <= W.l - W.t:
            k = int(W.inactive[i - 1])
            if cx[k - 1] < eps_ or (k == index_alpha_upp and cx[k - 1] < delta):
                if W.t >= bnd:
                    worst_k = 0
                    worst_val = -math.inf
                    for j in range(W.q + 1, W.t + 1):
                        jj = int(W.active[j - 1])
                        if cx[jj - 1] > worst_val:
                            worst_val = cx[jj - 1]
                            worst_k = j
                    if worst_k > 0 and worst_val > cx[k - 1]:
                        # After the swap-out ``inactive`` is re-sorted, so position ``i`` may no longer
                        # hold ``k``: the reference can re-add the constraint it just removed and then
                        # repeats the same swap forever (observed on ~3% of the C2 HS65 instances).
                        # The loop is deterministic, so exceeding this cap means it never ends.
                        swaps += 1
                        if swaps > 4 * W.l + 16:
                            raise ReferenceWouldHang("evaluate_violated_constraints swap cycle")
                        W.remove_constraint(worst_k)
                    else:
                        i += 1
                        continue
                W.add_constraint(i)
                added = True
            else:
                i += 1
    return added


# --------------------------------------------------------------------------------------
# EF:686-795
# --------------------------------------------------------------------------------------
def _delete_row(M, s):  # 1-based s
    return np.delete(M, s - 1, axis=0)


def update_working_set(W: WorkingSet, rx, A, C: Constraint, gradf, J, p_gn, it: Iteration, eps_rank):
    lam = np.zeros(W.t)
    F_A = QRPivoted(C.A.T)
    first_lagrange_mult_estimate(C.A, lam, gradf, C.cx, C.scaling, C.diag_scale, F_A, it, eps_rank)
    s = check_constraint_deletion(W.q, C.A, lam, C.scaling, C.diag_scale, it.grad_res)
    m, n = J.shape

    def second_order_branch(lam, F_A, F_L11, F_J2, rankA):
        if not (W.t != rankA or it.rankJ2 != min(m, n - rankA)):
            second_lagrange_mult_estimate(J, F_A, lam, rx, p_gn, W.t, C.scaling, C.diag_scale)
            s2 = check_constraint_deletion(W.q, C.A, lam, C.scaling, C.diag_scale, 0.0)
            if s2 != 0:
                index_s2 = int(W.active[s2 - 1])
                lam = np.delete(lam, s2 - 1)
                C.diag_scale = np.delete(C.diag_scale, s2 - 1)
                C.cx = np.delete(C.cx, s2 - 1)
                W.remove_constraint(s2)
                it.dele = True
                it.index_del = index_s2
                C.A = _delete_row(C.A, s2)
                F_A = QRPivoted(C.A.T)
                rankA = pseudo_rank(F_A.diagR(), eps_rank)
                F_L11 = QRPivoted(F_A.R.T)
                p_gn[:], F_J2 = gn_search_direction(J, rx, C.cx, F_A, F_L11, rankA, W.t, eps_rank, it)
        return lam, F_A, F_L11, F_J2

    if s != 0:  # EF:706-765 (always reverted; restated literally)
        cx_s = C.cx[s - 1]
        A_s = C.A[s - 1, :].copy()
        lam_s = lam[s - 1]
        ds_s = C.diag_scale[s - 1]
        index_s = int(W.active[s - 1])
        lam = np.delete(lam, s - 1)
        C.cx = np.delete(C.cx, s - 1)
        C.diag_scale = np.delete(C.diag_scale, s - 1)
        W.remove_constraint(s)
        it.dele = True
        it.index_del = index_s
        C.A = _delete_row(C.A, s)
        F_A = QRPivoted(C.A.T)
        rankA = pseudo_rank(F_A.diagR(), eps_rank)
        F_L11 = QRPivoted(F_A.R.T)
        p_gn[:], F_J2 = gn_search_direction(J, rx, C.cx, F_A, F_L11, rankA, W.t, eps_rank, it)
        As_p = 0.0 if rankA <= W.t else _dot(A_s, p_gn)
        feasible = (As_p >= -cx_s) and (As_p > 0)
        if not feasible:
            C.cx = np.insert(C.cx, s - 1, cx_s)
            lam = np.insert(lam, s - 1, lam_s)
            C.diag_scale = np.insert(C.diag_scale, s - 1, ds_s)
            s_inact = int(np.nonzero(W.inactive == index_s)[0][0]) + 1
            W.add_constraint(s_inact)
            it.index_del = 0
            it.dele = False
            rows = A[W.act(), :]
            C.A = (rows * C.diag_scale[:, None]) if C.scaling else rows.copy()
            F_A = QRPivoted(C.A.T)
            rankA = pseudo_rank(F_A.diagR(), eps_rank)
            F_L11 = QRPivoted(F_A.R.T)
            p_gn[:], F_J2 = gn_search_direction(J, rx, C.cx, F_A, F_L11, rankA, W.t, eps_rank, it)
            lam, F_A, F_L11, F_J2 = second_order_branch(lam, F_A, F_L11, F_J2, rankA)
    else:  # EF:767-791
        rankA = pseudo_rank(F_A.diagR(), eps_rank)
        F_L11 = QRPivoted(F_A.R.T)
        p_gn[:], F_J2 = gn_search_direction(J, rx, C.cx, F_A, F_L11, rankA, W.t, eps_rank, it)
        lam, F_A, F_L11, F_J2 = second_order_branch(lam, F_A, F_L11, F_J2, rankA)
    it.lam = lam
    return F_A, F_L11, F_J2


def init_working_set(cx, K, step: Iteration, q, l):
    """EF:826-859."""
    delta, eps_ = 0.1, 0.01
    for i in range(len(K)):
        K[i] = delta * np.ones(l)
    for i in range(l):
        step.w[i] = min(abs(cx[i]) + eps_, delta)
    active = np.zeros(l, dtype=np.int64)
    inactive = np.zeros(l - q, dtype=np.int64)
    t = q
    lmt = 0
    active[:q] = np.arange(1, q + 1)
    for i in range(q + 1, l + 1):
        if cx[i - 1] <= 0.0:
            t += 1
            active[t - 1] = i
        else:
            lmt += 1
            inactive[lmt - 1] = i
    step.t = t
    return WorkingSet(q, l, t, active, inactive)


# --------------------------------------------------------------------------------------
# EF:864-1176 subspace dimension heuristics
# --------------------------------------------------------------------------------------
def subspace_min_previous_step(tau, rho, rho_prk, c1, pseudo_rk, previous_dimR, progress,
                               predicted_linear_progress, prelin_previous_dim, previous_alpha):
    """EF:864-904 (1-based dims; tau/rho are Python arrays)."""
    stepb, pgb1, pgb2, predb, rlenb, c2 = 2e-1, 3e-1, 1e-1, 7e-1, 2.0, 1e2

    def g(v, i):  # Julia v[i], 1-based with BoundsError
        if i < 1 or i > len(v):
            raise ReferenceWouldThrow("BoundsError in subspace_min_previous_step")
        return v[i - 1]

    if (previous_alpha < stepb and progress <= pgb1 * predicted_linear_progress ** 2
            and progress <= pgb2 * prelin_previous_dim ** 2):
        dim = max(1, previous_dimR - 1)
        if previous_dimR > 1 and g(rho, dim) > c1 * rho_prk:
            return dim
    dim = previous_dimR
    if previous_dimR < len(tau) and ((g(rho, dim) > predb * rho_prk and rlenb * g(tau, dim) < g(tau, dim + 1))
                                     or (c2 * g(tau, dim) < g(tau, dim + 1))):
        suggested = dim
    else:
        i1 = previous_dimR - 1
        if i1 <= 0:
            suggested = pseudo_rk
        else:
            buff = [i for i in range(i1, previous_dimR + 1) if g(rho, i) > predb * rho_prk]
            suggested = pseudo_rk if not buff else min(buff)
    return suggested


def gn_previous_step(tau, tau_prk, mindim, rho, rho_prk, prank):
    """EF:909-932."""
    tau_max, rho_min = 2e-1, 5e-1
    pm1 = prank - 1
    if mindim > pm1:
        return mindim
    k = pm1
    while (tau[k - 1] >= tau_max * tau_prk or rho[k - 1] <= rho_min * rho_prk) and k > mindim:
        k -= 1
    return k if k > mindim else max(mindim, pm1)


def check_gn_direction(b1nrm, d1nrm, d1nrm_as_km1, dnrm, active_c_sum, iter_number, rankA, n, m, restart,
                       constraint_added, constraint_deleted, W: WorkingSet, cx, lam, it_km1: Iteration,
                       scaling, diag_scale):
    """EF:943-1030."""
    delta = 1e-1
    c1, c2, c3, c4, c5 = 0.5, 0.1, 4.0, 10.0, 0.05
    beta_k = math.sqrt(d1nrm ** 2 + b1nrm ** 2)
    method_code = 1
    newton_or_restart = it_km1.code == 2 or restart
    first_iter = iter_number == 0
    submin_prev = it_km1.code == -1
    add_or_del = constraint_added or constraint_deleted
    conv_lower = beta_k < c1 * it_km1.beta
    progress_not_close = (it_km1.progress > c2 * it_km1.predicted_reduction) and (dnrm <= c3 * beta_k)
    if newton_or_restart or (not first_iter and (submin_prev or not (add_or_del or conv_lower or progress_not_close))):
        method_code = -1
        nonlin_k = math.sqrt(d1nrm * d1nrm + active_c_sum)
        nonlin_km1 = math.sqrt(d1nrm_as_km1 * d1nrm_as_km1 + active_c_sum)
        to_reduce = False
        if W.q < W.t:
            rows = np.array([(1.0 / diag_scale[i]) if scaling else diag_scale[i] for i in range(W.q, W.t)])
            lam_in = lam[W.q:W.t]
            cond = bool(np.any(lam_in * rows >= -SQRT_EPS)) and bool(np.any(lam_in < 0))
            to_reduce = to_reduce or cond
        if W.l - W.t > 0:
            inact_c = cx[W.inact()]
            to_reduce = to_reduce or bool(np.any(inact_c < delta))
        newton_previously = it_km1.code == 2 and not constraint_deleted
        cond4 = active_c_sum > c2
        cond5 = constraint_deleted or constraint_added or to_reduce or (W.t == n and W.t == rankA)
        eps_ = max(1e-2, 10.0 * EPS)
        cond6 = (not ((W.l == W.q) or (rankA <= W.t))) and (not ((beta_k < eps_ * dnrm) or (b1nrm < eps_ and m == n - W.t)))
        if newton_previously or not (cond4 or cond5 or cond6):
            cond7 = (it_km1.alpha < c5 and nonlin_km1 < c2 * nonlin_k) or m == n - W.t
            cond8 = not (dnrm <= c4 * beta_k)
            if newton_previously or cond7 or cond8:
                method_code = 2
    return method_code, beta_k


def determine_solving_dim(previous_dimR, rankR, predicted_linear_progress, obj_progress, prelin_previous_dim,
                          R, y, previous_alpha, restart):
    """EF:1041-1113."""
    c1 = 0.1
    newdim = rankR
    eta = 1.0
    mindim = 1
    if rankR > 0:
        if rankR > len(y):
            raise ReferenceWouldThrow("BoundsError in determine_solving_dim")
        sd, rh = np.zeros(rankR), np.zeros(rankR)
        with np.errstate(all="ignore"):
            sd[0] = abs(y[0])
            rh[0] = abs(F(y[0]) / F(R[0, 0]))
            for i in range(1, rankR):
                sd[i] = y[i]
                rh[i] = F(y[i]) / F(R[i, i])
                rh[i] = _norm(rh[i - 1:i + 1])
                sd[i] = _norm(sd[i - 1:i + 1])
        nrm_sd = sd[rankR - 1]
        nrm_rh = rh[rankR - 1]
        dsum = 0.0
        psimax = 0.0
        for i in range(rankR):
            dsum += sd[i] ** 2
            psi_ = math.sqrt(dsum) * abs(R[i, i])
            if psi_ > psimax:
                psimax = psi_
                mindim = i + 1
        if not restart:
            if previous_dimR == rankR or previous_dimR <= 0:
                suggested = gn_previous_step(sd, nrm_sd, mindim, rh, nrm_rh, rankR)
            else:
                suggested = subspace_min_previous_step(sd, rh, nrm_rh, c1, rankR, previous_dimR, obj_progress,
                                                       predicted_linear_progress, prelin_previous_dim, previous_alpha)
            newdim = max(mindim, suggested)
        else:
            newdim = max(0, min(rankR, previous_dimR))
            if newdim != 0:
                k = max(previous_dimR - 1, 1)
                if sd[newdim - 1] != 0:
                    eta = sd[k - 1] / sd[newdim - 1]
    return newdim, eta


def choose_subspace_dimensions(rx_sum, rx, active_cx_sum, J1, t, rankJ2, rankA, b, F_L11, F_J2,
                               prev: Iteration, restart):
    """EF:1118-1176."""
    alpha_low = 0.2
    previous_alpha = prev.alpha
    if rankA <= 0:
        dimA = 0
        previous_dimA = 0
        d = -rx
    else:
        previous_dimA = abs(prev.dimA) + t - prev.t
        nrm_b_asprev = _norm(_jl_range(b, previous_dimA))
        nrm_b = _norm(b)
        constraint_progress = _dot(prev.cx, prev.cx) - active_cx_sum
        dimA, _eta = determine_solving_dim(previous_dimA, rankA, nrm_b, constraint_progress, nrm_b_asprev,
                                           F_L11.R, b, previous_alpha, restart)
        dp1 = _solve_upper(F_L11.R[:dimA, :dimA], _jl_range(b, dimA))
        if rankA - dimA < 0:
            raise ReferenceWouldThrow("negative zeros() length")
        Pm = F_L11.Pmat()
        if rankA > Pm.shape[0]:
            raise ReferenceWouldThrow("BoundsError in F_L11.P[1:rankA,1:rankA]")
        p1 = Pm[:rankA, :rankA] @ np.concatenate([dp1, np.zeros(rankA - dimA)])
        d = -(rx + J1 @ p1)
    if rankJ2 > 0:
        d = F_J2.Qt_mul(d)
    previous_dimJ2 = abs(prev.dimJ2) + prev.t - t
    nrm_d_asprev = _norm(_jl_range(d, previous_dimJ2))
    nrm_d = _norm(d)
    residual_progress = _dot(prev.rx, prev.rx) - rx_sum
    dimJ2, _eta = determine_solving_dim(previous_dimJ2, rankJ2, nrm_d, residual_progress, nrm_d_asprev,
                                        F_J2.R, d, previous_alpha, restart)
    if (not restart) and previous_alpha >= alpha_low:
        dimA = max(dimA, previous_dimA)
        dimJ2 = max(dimJ2, previous_dimJ2)
    return dimA, dimJ2


def search_direction_analys(prev: Iteration, cur: Iteration, iter_number, x, ev, rx, cx, C: Constraint,
                            active_cx_sum, p_gn, J, W: WorkingSet, F_A, F_L11, F_J2, second_derivatives):
    """EF:1191-1291."""
    m, n = J.shape
    rx_sum = _dot(rx, rx)
    active_cx = C.cx
    lam = cur.lam
    b_gn = cur.b_gn
    nrm_b1_gn = _norm(_jl_range(b_gn, cur.dimA))
    rankA = cur.rankA
    d_gn = cur.d_gn
    nrm_d_gn = _norm(d_gn)
    nrm_d1_gn = _norm(_jl_range(d_gn, cur.dimJ2))
    rankJ2 = cur.rankJ2
    prev_dimJ2m1 = prev.dimJ2 + prev.t - W.t - 1
    nrm_d1_asprev = _norm(_jl_range(d_gn, prev_dimJ2m1))
    restart = cur.restart
    error_code = 0
    method_code, beta = check_gn_direction(nrm_b1_gn, nrm_d1_gn, nrm_d1_asprev, nrm_d_gn, active_cx_sum, iter_number,
                                           rankA, n, m, restart, cur.add, cur.dele, W, cx, lam, prev,
                                           C.scaling, C.diag_scale)
    if method_code == 1:
        dimA, dimJ2 = rankA, rankJ2
        p, b, d = p_gn, b_gn, d_gn
    elif method_code == -1:
        JQ1 = F_A.mul_Q(J)
        J1 = JQ1[:, :rankA]
        b = F_L11.Qt_mul(-active_cx[F_A.p])
        dimA, dimJ2 = choose_subspace_dimensions(rx_sum, rx, active_cx_sum, J1, W.t, rankJ2, rankA, b, F_L11, F_J2,
                                                 prev, restart)
        p, b, d = sub_search_direction(J1, rx, active_cx, F_A, F_L11, F_J2, n, W.t, rankA, dimA, dimJ2, method_code)
        if dimA == rankA and dimJ2 == rankJ2:
            method_code = 1
    else:  # method_code == 2
        if second_derivatives:
            p, newton_error = newton_search_direction(x, ev, active_cx, W, lam, rx, J, F_A, F_L11, rankA)
            b, d = b_gn, d_gn
            dimA = -W.t
            dimJ2 = W.t - n
            cur.nb_newton_steps += 1
            if newton_error:
                error_code = -3
        else:
            p, b, d = p_gn, b_gn, d_gn
            dimA, dimJ2 = rankA, rankJ2
            error_code = -4
    cur.b_gn = b
    cur.d_gn = d
    cur.dimA = dimA
    cur.dimJ2 = dimJ2
    cur.code = method_code
    with np.errstate(all="ignore"):
        cur.speed = float(F(beta) / F(prev.beta))
    cur.beta = beta
    cur.p = p
    return error_code


# --------------------------------------------------------------------------------------
# EF:1307-1629 merit function and penalty weights
# --------------------------------------------------------------------------------------
def psi(x, alpha, p, ev: Evaluator, w, m, l, t, active, inactive, rx_buf=None, cx_buf=None):
    """EF:1307-1340."""
    rx_buf = np.zeros(m) if rx_buf is None else rx_buf
    cx_buf = np.zeros(l) if cx_buf is None else cx_buf
    pen = 0.0
    x_new = x + alpha * p
    ev.res(x_new, rx_buf)
    ev.cons(x_new, cx_buf)
    for i in range(t):
        j = active[i] - 1
        pen += w[j] * cx_buf[j] ** 2
    for i in range(l - t):
        j = inactive[i] - 1
        if cx_buf[j] < 0.0:
            pen += w[j] * cx_buf[j] ** 2
    return 0.5 * (_dot(rx_buf, rx_buf) + pen)


def assort(K, w, t, active):
    """EF:1344-1360."""
    for i in range(t):
        for ii in range(4):
            k = active[i] - 1
            if w[k] > K[ii][k]:
                for j in range(3, ii, -1):
                    K[j][k] = K[j - 1][k]
                K[ii][k] = w[k]


def min_norm_w(ctrl, w, w_old, y, tau, pos_index, nb_pos):
    """EF:1374-1423.  ``y /= y_norm`` rebinds a local copy; ``pos_index`` mutated in place (T7)."""
    w[:] = w_old
    if nb_pos > 0:
        y_sum = _dot(y, y)
        y_norm = _norm(y)
        if y_norm != 0.0:
            y = y / y_norm
        else:
            y = y.copy()
        tau_new = tau
        s = 0.0
        n_runch = nb_pos
        terminated = False
        while not terminated:
            tau_new -= s
            with np.errstate(all="ignore"):
                c = 1.0 if float(np.max(np.abs(y))) <= EPS else float(F(tau_new) / F(y_sum))
            y_sum, s = 0.0, 0.0
            i_stop = n_runch
            k = 1
            while k <= n_runch:
                i = pos_index[k - 1] - 1
                buff = c * y[k - 1] * y_norm
                if buff >= w_old[i]:
                    w[i] = buff
                    y_sum += y[k - 1] ** 2
                    k += 1
                else:
                    s += w_old[i] * y[k - 1] * y_norm
                    n_runch -= 1
                    for j in range(k, n_runch + 1):
                        pos_index[j - 1] = pos_index[j]
                        y[j - 1] = y[j]
            y_sum *= y_norm * y_norm
            terminated = (n_runch <= 0) or (ctrl == 2) or (i_stop == n_runch)


def euclidean_norm_weight_update(vA, cx, active, t, mu, dimA, previous_w, K):
    """EF:1429-1497."""
    w = previous_w.copy()
    if t != 0:
        z = vA ** 2
        w_old = K[3]
        act0 = active[:t] - 1
        ztw = _dot(z, w_old[act0])
        pos_index = np.zeros(t, dtype=np.int64)
        if ztw >= mu and dimA < t:
            y = np.zeros(t)
            ctrl, nb_pos, gamma = 2, 0, 0.0
            for i in range(t):
                k = active[i]
                y_elem = vA[i] * (vA[i] + cx[k - 1])
                if y_elem > 0:
                    nb_pos += 1
                    pos_index[nb_pos - 1] = k
                    y[nb_pos - 1] = y_elem
                else:
                    gamma -= y_elem * w_old[k - 1]
            min_norm_w(ctrl, w, w_old, y, gamma, pos_index, nb_pos)
        elif ztw < mu and dimA < t:
            e = np.zeros(t)
            ctrl, nb_pos, tau = 2, 0, mu
            for i in range(t):
                k = active[i]
                e_elem = -vA[i] * cx[k - 1]
                if e_elem > 0:
                    nb_pos += 1
                    pos_index[nb_pos - 1] = k
                    e[nb_pos - 1] = e_elem
                else:
                    tau -= e_elem * w_old[k - 1]
            min_norm_w(ctrl, w, w_old, e, tau, pos_index, nb_pos)
        elif ztw < mu and dimA == t:
            ctrl = 1
            pos_index[:] = active[:t]
            min_norm_w(ctrl, w, w_old, z, mu, pos_index, t)
        assort(K, w, t, active)
    return w


def penalty_weight_update(w_old, Jp, Ap, K, rx, cx, W: WorkingSet, dimA, norm_code):
    """EF:1545-1629 (norm_code == 2 only: ``solve!`` never passes ``weight_code``)."""
    delta = 0.25
    active = W.active
    t = W.t
    if dimA < 0 or dimA > W.l:
        raise ReferenceWouldThrow("active[1:dimA] out of range")
    nrm_Ap = math.sqrt(_dot(Ap, Ap))
    sel = cx[active[:dimA] - 1]
    if np.any(active[:dimA] == 0):
        raise ReferenceWouldThrow("cx[0]")
    nrm_cx = 0.0 if sel.size == 0 else max(0.0, float(np.max(np.abs(sel))))
    nrm_Jp = math.sqrt(_dot(Jp, Jp))
    nrm_rx = math.sqrt(_dot(rx, rx))
    if nrm_Jp != 0:
        Jp = Jp / nrm_Jp
    if nrm_Ap != 0:
        Ap = Ap / nrm_Ap
    if nrm_rx != 0:
        rx = rx / nrm_rx
    if nrm_cx != 0:
        cx = cx / nrm_cx
    Jp_rx = _dot(Jp, rx) * nrm_Jp * nrm_rx
    AtwA = 0.0
    BtwA = 0.0
    for i in range(dimA):
        k = active[i] - 1
        AtwA += w_old[k] * Ap[i] ** 2
        BtwA += w_old[k] * Ap[i] * cx[k]
    AtwA *= nrm_Ap ** 2
    BtwA *= nrm_Ap * nrm_cx
    alpha_w = 1.0
    if abs(AtwA + nrm_Jp ** 2) > EPS:
        alpha_w = (-BtwA - Jp_rx) / (AtwA + nrm_Jp ** 2)
    rmy = (abs(Jp_rx + nrm_Jp ** 2) / delta) - nrm_Jp ** 2
    if norm_code != 2:
        raise NotImplementedError("max-norm weights are unreachable from solve! (SURVEY.md #14)")
    w = euclidean_norm_weight_update(Ap * nrm_Ap, cx * nrm_cx, active, t, rmy, dimA, w_old, K)
    BtwA = 0.0
    AtwA = 0.0
    for i in range(t):
        k = active[i] - 1
        AtwA += w[k] * Ap[i] ** 2
        BtwA += w[k] * Ap[i] * cx[k]
    BtwA *= nrm_Ap * nrm_cx
    AtwA *= nrm_Ap ** 2
    dpsi0 = BtwA + Jp_rx
    return w, dpsi0


# --------------------------------------------------------------------------------------
# EF:1635-2143 linesearch
# --------------------------------------------------------------------------------------
def concatenate(v, rx, cx, w, m, t, l, active, inactive):
    """EF:1635-1659."""
    v[:m] = rx
    for i in range(t):
        k = active[i] - 1
        v[m + k] = math.sqrt(w[k]) * cx[k]
    if l != 0:
        for j in range(l - t):
            k = inactive[j] - 1
            v[m + k] = 0.0 if cx[k] > 0 else math.sqrt(w[k]) * cx[k]


def coefficients_linesearch(v0, v1, v2, alpha_k, rx, cx, rx_new, cx_new, w, m, t, l, active, inactive):
    """EF:1665-1689."""
    concatenate(v0, rx, cx, w, m, t, l, active, inactive)
    v_buff = np.zeros(m + l)
    concatenate(v_buff, rx_new, cx_new, w, m, t, l, active, inactive)
    with np.errstate(all="ignore"):
        v2[:] = ((v_buff - v0) / F(alpha_k) - v1) / F(alpha_k)


def minimize_quadratic(x1, y1, x2, y2, x3, y3):
    """EF:1694-1702."""
    d1, d2 = y2 - y1, y3 - y1
    s = (x3 - x1) ** 2 * d1 - (x2 - x1) ** 2 * d2
    q = 2 * ((x2 - x1) * d2 - (x3 - x1) * d1)
    with np.errstate(all="ignore"):
        return float(F(x1) - F(s) / F(q))


def _clamp(u, lo, hi):  # Julia clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x))
    if u > hi:
        return hi
    if u < lo:
        return lo
    return u


def minrn(x1, y1, x2, y2, x3, y3, alpha_min, alpha_max, p_max):
    """EF:1708-1735."""
    with np.errstate(all="ignore"):
        eps_ = float(F(SQRT_EPS) / F(p_max))
    if abs(x1 - x2) < eps_ or abs(x3 - x1) < eps_ or abs(x3 - x2) < eps_:
        return 0.0, 0.0
    u = minimize_quadratic(x1, y1, x2, y2, x3, y3)
    a = _clamp(u, alpha_min, alpha_max)
    t1 = (a - x1) * (a - x2) * y3 / ((x3 - x1) * (x3 - x2))
    t2 = (a - x3) * (a - x2) * y1 / ((x1 - x3) * (x1 - x2))
    t3 = (a - x3) * (a - x2) * y2 / ((x2 - x1) * (x2 - x3))
    return a, t1 + t2 + t3


class Poly:
    """Polynomials.jl ``Polynomial``: trailing zeros chopped at construction, Horner evaluation."""

    def __init__(self, coeffs):
        c = [float(v) for v in coeffs]
        while len(c) > 1 and c[-1] == 0.0:
            c.pop()
        if len(c) == 1 and c[0] == 0.0:
            c = [0.0]
        self.c = c

    def __call__(self, x):
        with np.errstate(all="ignore"):
            acc = F(self.c[-1])
            xx = F(x)
            for v in reversed(self.c[:-1]):
                acc = acc * xx + F(v)   # Julia: muladd (FMA); difference is one rounding
            return float(acc)

    def derivative(self):
        if len(self.c) <= 1:
            return Poly([0.0])
        return Poly([i * self.c[i] for i in range(1, len(self.c))])


def newton_raphson(x_min, Dm, ds: Poly, dds: Poly):
    """EF:1791-1811."""
    alpha, it = x_min, 0
    eps_, error = 1e-4, 1.0
    with np.errstate(all="ignore"):
        while (error > eps_ or it < 3) and it < 50:
            c = dds(alpha)
            if abs(c) < EPS:
                break
            h = float(-F(ds(alpha)) / F(c))
            alpha += h
            error = float((2 * F(Dm) * F(h) ** 2) / abs(F(c)))
            it += 1
    return alpha


def one_root(c, d, a):
    """EF:1815-1818."""
    arg1, arg2 = -c / 2 + math.sqrt(d), -c / 2 - math.sqrt(d)
    return float(np.cbrt(arg1) + np.cbrt(arg2) - a / 3)


def two_roots(b, c, d, a, x_min):
    """EF:1821-1837."""
    with np.errstate(all="ignore"):
        phi = float(np.arccos(F(abs(c / 2)) / F(-b / 3) ** F(1.5)))
    if math.isnan(phi):
        raise ReferenceWouldThrow("DomainError in acos")
    t = 2 * math.sqrt(-b / 3) if c <= 0 else -2 * math.sqrt(-b / 3)
    b1 = t * math.cos(phi / 3) - a / 3
    b2 = t * math.cos((phi + 2 * math.pi) / 3) - a / 3
    b3 = t * math.cos((phi + 4 * math.pi) / 3) - a / 3
    b1, b2, b3 = sorted([b1, b2, b3])
    return (b1, b3) if x_min <= b2 else (b3, b1)


def parameters_rm(v0, v1, v2, x_min, ds: Poly, dds: Poly):
    """EF:1739-1783."""
    with np.errstate(all="ignore"):
        dds_best = dds(x_min)
        eta, d = 0.1, 1.0
        normv2 = _dot(v2, v2)
        h0 = float(abs(F(ds(x_min)) / F(dds_best)))
        Dm = float(abs(F(6) * F(_dot(v1, v2)) + F(12) * F(x_min) * F(normv2)) + F(24) * F(h0) * F(normv2))
        hm = max(h0, 1.0)
        beta_hat = None
        if dds_best * eta < 2 * Dm * hm:
            cf = ds.c
            if len(cf) != 4:
                # (a3, a2, a1) = coeffs(ds)/(2*normv2) needs >= 3 entries; with a chopped
                # polynomial the reference either throws or mis-assigns
                if len(cf) < 3:
                    raise ReferenceWouldThrow("BoundsError destructuring coeffs(ds)")
            a3, a2, a1 = [float(F(v) / F(2 * normv2)) for v in cf[:3]]
            b = a2 - (a1 ** 2) / 3
            c = a3 - a1 * a2 / 3 + 2 * (a1 / 3) ** 3
            d = (c / 2) ** 2 + (b / 3) ** 3
            if d < 0:
                alpha_hat, beta_hat = two_roots(b, c, d, a1, x_min)
            else:
                if math.isnan(d):
                    raise ReferenceWouldThrow("NaN discriminant")
                alpha_hat = one_root(c, d, a1)
        else:
            alpha_hat = newton_raphson(x_min, Dm, ds, dds)
        if d >= 0:
            beta_hat = alpha_hat
        if beta_hat is None:
            raise ReferenceWouldThrow("UndefVarError beta_hat")
    return alpha_hat, beta_hat


def bounds(alpha_min, alpha_max, alpha, s: Poly):
    """EF:1785-1789 (Julia min/max propagate NaN)."""
    if math.isnan(alpha):
        return alpha, s(alpha)
    alpha = min(alpha, alpha_max)
    alpha = max(alpha, alpha_min)
    return alpha, s(alpha)


def minrm(v0, v1, v2, x_min, alpha_min, alpha_max):
    """EF:1841-1862."""
    s = Poly([0.5 * _dot(v0, v0), _dot(v0, v1), _dot(v0, v2) + 0.5 * _dot(v1, v1), _dot(v1, v2), 0.5 * _dot(v2, v2)])
    ds = s.derivative()
    dds = ds.derivative()
    alpha_hat, beta_hat = parameters_rm(v0, v1, v2, x_min, ds, dds)
    s_a, s_b = s(alpha_hat), s(beta_hat)
    alpha_old = alpha_hat
    alpha_hat, s_a = bounds(alpha_min, alpha_max, alpha_hat, s)
    if alpha_old == beta_hat:
        beta_hat, s_b = alpha_hat, s(alpha_hat)
    else:
        beta_hat, s_b = bounds(alpha_min, alpha_max, beta_hat, s)
    return alpha_hat, s_a, beta_hat, s_b


def check_reduction(psi_alpha, psi_k, approx_k, eta, diff_psi):
    """EF:1870-1886."""
    delta = 0.2
    if psi_alpha - approx_k >= eta * diff_psi:
        return not ((psi_alpha - psi_k < eta * diff_psi) and (psi_k > delta * psi_alpha))
    return False


def goldstein_armijo_step(psi0, dpsi0, alpha_min, tau, p_max, x, alpha0, p, ev, w, m, l, t, active, inactive,
                          rx_buf, cx_buf):
    """EF:1893-1923."""
    u = alpha0
    exit_ = (p_max * u < SQRT_EPS) or (u <= alpha_min)
    psi_u = psi(x, u, p, ev, w, m, l, t, active, inactive, rx_buf, cx_buf)
    while (not exit_) and (psi_u > psi0 + tau * u * dpsi0):
        u *= 0.5
        psi_u = psi(x, u, p, ev, w, m, l, t, active, inactive, rx_buf, cx_buf)
        exit_ = (p_max * u < SQRT_EPS) or (u <= alpha_min)
    return u, exit_


def linesearch_constrained(x, alpha0, p, ev: Evaluator, rx, cx, JpAp, w, W: WorkingSet, psi0, dpsi0,
                           alpha_low, alpha_upp, log=None):
    """EF:1940-2143."""
    m = rx.size
    l, t = W.l, W.t
    active, inactive = W.active, W.inactive
    psi_rx, psi_cx = np.zeros(m), np.zeros(l)
    rx_new, cx_new = np.zeros(m), np.zeros(l)
    v0, v2 = np.zeros(m + l), np.zeros(m + l)
    eta, tau, gamma = 0.3, 0.25, 0.4
    alpha_min, alpha_max = alpha_low, alpha_upp
    alpha_k = min(alpha0, alpha_max)
    alpha_km1 = 0.0
    psi_km1 = psi0
    p_max = float(np.max(np.abs(p))) if p.size else 0.0
    gac_error = False
    v1 = JpAp  # mutated in place (T10)
    for i in range(t):
        k = active[i] - 1
        v1[m + k] = math.sqrt(w[k]) * v1[m + k]
    for j in range(l - t):
        k = inactive[j] - 1
        v1[m + k] = 0.0 if cx[k] > 0 else math.sqrt(w[k]) * v1[m + k]

    def PSI(a):
        val = psi(x, a, p, ev, w, m, l, t, active, inactive, psi_rx, psi_cx)
        if log is not None:
            log.append((a, val))
        return val

    psi_k = PSI(alpha_k)
    diff_psi = psi0 - psi_k
    x_new = x + alpha_k * p
    ev.res(x_new, rx_new)
    ev.cons(x_new, cx_new)
    v0[:] = 0.0
    v2[:] = 0.0
    coefficients_linesearch(v0, v1, v2, alpha_k, rx, cx, rx_new, cx_new, w, m, t, l, active, inactive)
    x_min = alpha_k if diff_psi >= 0 else 0.0
    alpha_kp1, pk, beta, pbeta = minrm(v0, v1, v2, x_min, alpha_min, alpha_max)
    if alpha_kp1 != beta and pbeta < pk and beta <= alpha_k:
        alpha_kp1 = beta
        pk = pbeta
    alpha_km2 = alpha_km1
    psi_km2 = psi_km1
    alpha_km1 = alpha_k
    psi_km1 = psi_k
    alpha_k = alpha_kp1
    psi_k = PSI(alpha_k)
    if (-diff_psi <= tau * dpsi0 * alpha_km1) or (psi_km1 < gamma * psi0):
        diff_psi = psi0 - psi_k
        reduction_likely = check_reduction(psi_km1, psi_k, pk, eta, diff_psi)
        while reduction_likely:
            alpha_kp1, pk = minrn(alpha_k, psi_k, alpha_km1, psi_km1, alpha_km2, psi_km2, alpha_min, alpha_max, p_max)
            alpha_km2 = alpha_km1
            psi_km2 = psi_km1
            alpha_km1 = alpha_k
            psi_km1 = psi_k
            alpha_k = alpha_kp1
            psi_k = PSI(alpha_k)
            diff_psi = psi0 - psi_k
            reduction_likely = check_reduction(psi_km1, psi_k, pk, eta, diff_psi)
        if (psi_km1 - pk >= eta * diff_psi) and (psi_k < psi_km1):
            alpha_km1 = alpha_k
            psi_km1 = psi_k
    else:
        diff_psi = psi0 - psi_k
        if (-diff_psi <= tau * dpsi0 * alpha_k) or (psi_k < gamma * psi0):
            if psi0 <= psi_km1:
                x_min = alpha_k
                x_new = x + alpha_k * p
                ev.res(x_new, rx_new)
                ev.cons(x_new, cx_new)
                v0[:] = 0.0
                v2[:] = 0.0
                coefficients_linesearch(v0, v1, v2, alpha_k, rx, cx, rx_new, cx_new, w, m, t, l, active, inactive)
                alpha_kp1, pk, beta, pbeta = minrm(v0, v1, v2, x_min, alpha_min, alpha_max)
                if alpha_kp1 != beta and pbeta < pk and beta <= alpha_k:
                    alpha_kp1 = beta
                    pk = pbeta
                alpha_km1 = 0.0
                psi_km1 = psi0
            else:
                alpha_kp1, pk = minrn(alpha_k, psi_k, alpha_km1, psi_km1, alpha_km2, psi_km2, alpha_min, alpha_max, p_max)
            alpha_km2 = alpha_km1
            psi_km2 = psi_km1
            alpha_km1 = alpha_k
            psi_km1 = psi_k
            alpha_k = alpha_kp1
            psi_k = PSI(alpha_k)
            reduction_likely = check_reduction(psi_km1, psi_k, pk, eta, diff_psi)
            while reduction_likely:
                alpha_kp1, pk = minrn(alpha_k, psi_k, alpha_km1, psi_km1, alpha_km2, psi_km2, alpha_min, alpha_max, p_max)
                alpha_km2 = alpha_km1
                psi_km2 = psi_km1
                alpha_km1 = alpha_k
                psi_km1 = psi_k
                alpha_k = alpha_kp1
                psi_k = PSI(alpha_k)
                reduction_likely = check_reduction(psi_km1, psi_k, pk, eta, diff_psi)
            if (psi_km1 - pk >= eta * diff_psi) and (psi_k < psi_km1):
                alpha_km1 = alpha_k
                psi_km1 = psi_k
        else:
            alpha_km1, gac_error = goldstein_armijo_step(psi0, dpsi0, alpha_min, tau, p_max, x, alpha_k, p, ev, w, m, l, t,
                                                         active, inactive, psi_rx, psi_cx)
    return alpha_km1, gac_error


def upper_bound_steplength(A, cx, p, W: WorkingSet, index_del):
    """EF:2149-2178."""
    inactive = W.inactive
    t, l = W.t, W.l
    alpha_upper = math.inf
    index_alpha_upp = 0
    if inactive.size and int(np.max(np.abs(inactive))) > 0:
        for i in range(l - t):
            j = int(inactive[i])
            if j != index_del:
                g = _dot(A[j - 1, :], p)
                with np.errstate(all="ignore"):
                    a_j = float(-F(cx[j - 1]) / F(g))
                if cx[j - 1] > 0 and g < 0 and a_j < alpha_upper:
                    alpha_upper = a_j
                    index_alpha_upp = j
    return min(3.0, alpha_upper), index_alpha_upp


def check_derivatives(dpsi0, psi0, psi_k, x_old, alpha, p, ev, w, W: WorkingSet, m):
    """EF:2295-2322."""
    l, t = W.l, W.t
    psi_ma = psi(x_old, -alpha, p, ev, w, m, l, t, W.active, W.inactive)
    with np.errstate(all="ignore"):
        f = float((F(psi_k) - F(psi0)) / F(alpha))
        b = float((F(psi0) - F(psi_ma)) / F(alpha))
        c = float((F(psi_k) - F(psi_ma)) / F(2 * alpha))
    max_diff = max(abs(f - c), abs(f - b), abs(b - c))
    inconsistency = abs(f - dpsi0) > max_diff and abs(c - dpsi0) > max_diff
    return -1 if inconsistency else 0


def compute_steplength(it: Iteration, prev: Iteration, x, ev: Evaluator, rx, J, cx, A, C: Constraint,
                       W: WorkingSet, K, weight_code, ls_log=None):
    """EF:2197-2293."""
    m = J.shape[0]
    p = it.p
    dimA = it.dimA
    rankJ2 = it.rankJ2
    method_code = it.code
    ind_del = it.index_del
    previous_alpha = prev.alpha
    prev_rankJ2 = prev.rankJ2
    w_old = prev.w
    Jp = J @ p
    Ap = A @ p
    JpAp = np.concatenate([Jp, Ap])
    active_Ap = C.A @ p
    active_index = W.act()
    if C.scaling:
        active_Ap = active_Ap / C.diag_scale
    Psi_error = 0
    if method_code != 2:
        w, dpsi0 = penalty_weight_update(w_old, Jp, active_Ap, K, rx, cx, W, dimA, weight_code)
        psi0 = 0.5 * (_dot(rx, rx) + _dot(w[active_index], cx[active_index] ** 2))
        if dpsi0 >= 0:
            alpha = 1.0
            Psi_error = -1
            it.index_alpha_upp = 0
        else:
            alpha_upp, index_alpha_upp = upper_bound_steplength(A, cx, p, W, ind_del)
            alpha_low = alpha_upp / 3000.0
            magfy = 6.0 if rankJ2 < prev_rankJ2 else 3.0
            alpha0 = min(1.0, magfy * previous_alpha, alpha_upp)
            alpha, gac_error = linesearch_constrained(x, alpha0, p, ev, rx, cx, JpAp, w, W, psi0, dpsi0,
                                                      alpha_low, alpha_upp, ls_log)
            if gac_error:
                psi_k = psi(x, alpha, p, ev, w, m, W.l, W.t, W.active, W.inactive)
                Psi_error = check_derivatives(dpsi0, psi0, psi_k, x, alpha, p, ev, w, W, m)
            uppbound = min(1.0, alpha_upp)
            atwa = _dot(w[active_index], active_Ap ** 2)
            it.predicted_reduction = uppbound * (-2.0 * _dot(Jp, rx) - uppbound * _dot(Jp, Jp) + (2.0 - uppbound ** 2) * atwa)
            rx_new = np.zeros(m)
            cx_new = np.zeros(W.l)
            x_new = x + alpha * p
            ev.res(x_new, rx_new)
            ev.cons(x_new, cx_new)
            whsum = _dot(w[active_index], cx_new[active_index] ** 2)
            it.progress = 2 * psi0 - _dot(rx_new, rx_new) - whsum
            it.index_alpha_upp = 0 if (index_alpha_upp != 0 and abs(alpha - alpha_upp) > 0.1) else index_alpha_upp
    else:
        w = w_old
        it.index_alpha_upp = 0
        alpha = 1.0
    return alpha, w, Psi_error


# --------------------------------------------------------------------------------------
# EF:2399-2517 termination
# --------------------------------------------------------------------------------------
def check_termination_criteria(it: Iteration, prev: Iteration, W: WorkingSet, C: Constraint, x, cx, rx_sum, gradf,
                               max_iter, nb_iter, eps_abs, eps_rel, eps_x, eps_c, error_code, delta_time,
                               sigma_min, lam_abs_max, Psi_error):
    exit_code = 0
    rel_tol = EPS
    alfnoi = rel_tol / (_norm(it.p) + rel_tol)
    preliminary = not (it.restart or (it.code == -1 and alfnoi <= 0.25))
    if preliminary:
        necessary = (not it.dele) and (_norm(C.cx) < eps_c) and (it.grad_res < math.sqrt(eps_rel) * (1 + _norm(gradf)))
        if W.l - W.t > 0:
            necessary = necessary and bool(np.all(cx[W.inact()] > 0))
        if W.t > W.q:
            factor = (1 + rx_sum) if W.t == 1 else lam_abs_max
            necessary = necessary and (sigma_min >= eps_rel * factor)
        if necessary:
            d1 = _jl_range(it.d_gn, it.dimJ2)
            x_diff = _norm(prev.x - x)
            if _dot(d1, d1) <= rx_sum * eps_rel ** 2:
                exit_code += 10000
            if rx_sum <= eps_abs ** 2:
                exit_code += 2000
            if x_diff < eps_x * _norm(x):
                exit_code += 300
            if alfnoi > 0.25:
                exit_code += 40
            if exit_code > 0 and W.l - W.t > 0:
                feas = 1
                for ii in range(W.l - W.t):
                    jj = int(W.inactive[ii])
                    if cx[jj - 1] <= 0.0:
                        feas = -1
                        break
                exit_code *= feas
    if exit_code == 0:
        x_diff = _norm(prev.x - x)
        Atcx_nrm = _norm(C.A.T @ C.cx) if C.A.shape[0] else 0.0
        ai = W.act()
        active_penalty_sum = 0.0 if W.t == 0 else _dot(it.w[ai], it.w[ai])
        if nb_iter >= max_iter:
            exit_code = -2
        elif -5 <= error_code <= -3:
            exit_code = error_code
        elif it.nb_newton_steps > 5:
            exit_code = -9
        elif Psi_error == -1:
            exit_code = -6
        elif x_diff <= 10.0 * eps_x and Atcx_nrm <= 10.0 * eps_c and active_penalty_sum >= 1.0:
            exit_code = -10
        elif delta_time > 0:
            exit_code = -11
    return exit_code


def convert_exit_code(code):
    """cnls_model.jl:166-178."""
    if code > 0:
        return 1
    if code == -2 or code == -11:
        return code
    return -1


STATUS = {0: "unsolved", 1: "found_first_order_stationary_point", -1: "failed",
          -2: "maximum_iterations_exceeded", -11: "time_limit_exceeded"}


# --------------------------------------------------------------------------------------
# EF:2638-2880 driver
# --------------------------------------------------------------------------------------
@dataclass
class IterTrace:
    """One record per executed iteration (recorded or terminating)."""

    k: int
    x_new: np.ndarray
    f_new: float
    t: int
    active: List[int]
    rankA: int
    rankJ2: int
    dimA: int
    dimJ2: int
    code: int
    alpha: float
    p_norm: float
    index_del: int
    exit_code: int
    active_cx_sum: float = 0.0
    progress: float = 0.0


@dataclass
class Result:
    exit_code: int
    status: int
    x: np.ndarray
    f: float
    iterations: int                     # length(iterations_detail)
    nb_function_evaluations: int
    nb_jacobian_evaluations: int
    active: List[int]                   # working set at exit (1-based ids)
    trace: List[IterTrace] = field(default_factory=list)
    details: List[tuple] = field(default_factory=list)   # DisplayedInfo 5-tuples
    threw: Optional[str] = None


def enlsip(x0, prob: Problem, scaling=False, second_derivatives=True, weight_code=2, MAX_ITER=100,
           TIME_LIMIT=1000.0, eps_abs=1e-10, eps_rel=1e-5, eps_x=1e-3, eps_c=1e-4, eps_rank=1e-10,
           wallclock=True, ls_log=None) -> Result:
    """EF:2638-2880.  ``wallclock=False`` makes ``time()-start_time`` identically 0 (deterministic tests)."""
    n, m, q, l = prob.n, prob.m, prob.q, prob.l
    ev = Evaluator(prob)
    prob.reset_counters()
    x0 = np.array(x0, dtype=F)
    second_derivatives = second_derivatives and (n + m < 1000)
    nb_iteration = 0
    K = [np.zeros(l) for _ in range(4)]
    rx, cx = np.zeros(m), np.zeros(l)
    J, A = np.zeros((m, n)), np.zeros((l, n))
    new_point(ev, x0, rx, cx, J, A)
    x_opt = x0
    x = x0
    f_opt = _dot(rx, rx)
    first = Iteration(x0, np.zeros(n), rx, cx, l, 1.0, 0, np.zeros(l), np.zeros(l), 0, 0, 0, 0, np.zeros(n), np.zeros(n),
                      0.0, 0.0, 0.0, 0.0, 0.0, False, True, False, False, 0, 1, 0)
    start_time = time.time() if wallclock else 0.0
    now = (lambda: time.time()) if wallclock else (lambda: 0.0)
    W = init_working_set(cx, K, first, q, l)
    first.t = W.t
    C = Constraint(cx[W.act()].copy(), A[W.act(), :].copy(), scaling, np.zeros(W.t))
    gradf = J.T @ rx
    p_gn = np.zeros(n)
    trace: List[IterTrace] = []
    details = []
    try:
        evaluate_scaling(C)
        F_A, F_L11, F_J2 = update_working_set(W, rx, A, C, gradf, J, p_gn, first, eps_rank)
        rx_sum = _dot(rx, rx)
        active_cx_sum = _dot(cx[W.act()], cx[W.act()])
        first.t = W.t
        prev = first.copy()
        error_code = search_direction_analys(prev, first, nb_iteration, x0, ev, rx, cx, C, active_cx_sum, p_gn, J, W,
                                             F_A, F_L11, F_J2, second_derivatives)
        alpha, w, Psi_error = compute_steplength(first, prev, x0, ev, rx, J, cx, A, C, W, K, weight_code, ls_log)
        first.alpha = alpha
        first.w = w
        x = x0 + alpha * first.p
        act_at_step = [int(v) for v in W.active[:W.t]]
        new_point(ev, x, rx, cx, J, A)
        gradf = J.T @ rx
        rx_sum = _dot(rx, rx)
        first.restart = error_code < 0
        sigma_min, lam_abs_max = minmax_lagrangian_mult(first.lam, W, C)
        delta_time = (now() - start_time) - TIME_LIMIT
        exit_code = check_termination_criteria(first, prev, W, C, x, cx, rx_sum, gradf, MAX_ITER, nb_iteration, eps_abs,
                                               eps_rel, eps_x, eps_c, error_code, delta_time, sigma_min, lam_abs_max,
                                               Psi_error)
        details.append((f_opt, active_cx_sum, _norm(first.p), first.alpha, first.progress))
        trace.append(IterTrace(0, x.copy(), rx_sum, W.t, act_at_step, first.rankA, first.rankJ2, first.dimA,
                               first.dimJ2, first.code, alpha, _norm(first.p), first.index_del, exit_code,
                               active_cx_sum, first.progress))
        first.add = evaluate_violated_constraints(cx, W, first.index_alpha_upp, n)
        C.cx = cx[W.act()].copy()
        C.A = A[W.act(), :].copy()
        prev = first.copy()
        first.x = x
        first.rx = rx
        first.cx = cx
        f_opt = _dot(rx, rx)
        nb_iteration += 1
        it = first.copy()
        it.first = False
        it.add = False
        it.dele = False
        nfe = njac = 0
        took_loop = False
        while exit_code == 0:
            took_loop = True
            p_gn[:] = 0.0
            evaluate_scaling(C)
            F_A, F_L11, F_J2 = update_working_set(W, rx, A, C, gradf, J, p_gn, it, eps_rank)
            active_cx_sum = _dot(cx[W.act()], cx[W.act()])
            it.t = W.t
            error_code = search_direction_analys(prev, it, nb_iteration, x, ev, rx, cx, C, active_cx_sum, p_gn, J, W,
                                                 F_A, F_L11, F_J2, second_derivatives)
            alpha, w, Psi_error = compute_steplength(it, prev, x, ev, rx, J, cx, A, C, W, K, weight_code, ls_log)
            it.alpha = alpha
            it.w = w
            x = x + alpha * it.p
            act_at_step = [int(v) for v in W.active[:W.t]]
            new_point(ev, x, rx, cx, J, A)
            rx_sum = _dot(rx, rx)
            gradf = J.T @ rx
            it.restart = error_code < 0
            sigma_min, lam_abs_max = minmax_lagrangian_mult(it.lam, W, C)
            delta_time = (now() - start_time) - TIME_LIMIT
            exit_code = check_termination_criteria(it, prev, W, C, x, cx, rx_sum, gradf, MAX_ITER, nb_iteration, eps_abs,
                                                   eps_rel, eps_x, eps_c, error_code, delta_time, sigma_min,
                                                   lam_abs_max, Psi_error)
            trace.append(IterTrace(nb_iteration, x.copy(), rx_sum, W.t, act_at_step, it.rankA, it.rankJ2, it.dimA,
                                   it.dimJ2, it.code, alpha, _norm(it.p), it.index_del, exit_code, active_cx_sum,
                                   it.progress))
            if exit_code == 0:
                f_opt = _dot(rx, rx)
                details.append((f_opt, active_cx_sum, _norm(it.p), it.alpha, it.progress))
                it.add = evaluate_violated_constraints(cx, W, it.index_alpha_upp, n)
                C.cx = cx[W.act()].copy()
                C.A = A[W.act(), :].copy()
                nb_iteration += 1
                prev = it.copy()
                it.x = x
                it.rx = rx
                it.cx = cx
                it.dele = False
                it.add = False
            else:
                x_opt = x
                f_opt = _dot(rx, rx)
                nfe = prob.nb_reseval + prob.nb_conseval
                njac = prob.nb_jacres + prob.nb_jaccons
        if not took_loop:
            # EF:2660 + T5 : ExecutionInfo() default (one dummy record, zero counters); x_opt = x0
            details = [(0.0, 0.0, 0.0, 0.0, 0.0)]
        return Result(exit_code, convert_exit_code(exit_code), np.array(x_opt, dtype=F), float(f_opt), len(details),
                      nfe, njac, [int(v) for v in W.active[:W.t]], trace, details)
    except (ReferenceWouldThrow, ReferenceWouldHang) as e:
        code = EXIT_WOULD_HANG if isinstance(e, ReferenceWouldHang) else EXIT_WOULD_THROW
        # convention shared with the engine: report the last evaluated point and its objective
        return Result(code, -1, np.array(x, dtype=F), _dot(rx, rx), len(details), 0, 0,
                      [int(v) for v in W.active[:W.t]], trace, details, threw=str(e))


def solve(prob: Problem, x0=None, max_iter=100, scaling=False, time_limit=1e3, abs_tol=EPS, rel_tol=None,
          c_tol=None, x_tol=None, wallclock=True, ls_log=None) -> Result:
    """solver.jl:62-91 option plumbing: abs_tol is NOT forwarded (T4); eps_rank = sqrt(eps)."""
    rel_tol = math.sqrt(abs_tol) if rel_tol is None else rel_tol
    c_tol = rel_tol if c_tol is None else c_tol
    x_tol = rel_tol if x_tol is None else x_tol
    x0 = prob.x0 if x0 is None else x0
    return enlsip(x0, prob, scaling=scaling, MAX_ITER=max_iter, TIME_LIMIT=time_limit, eps_rel=rel_tol, eps_x=x_tol,
                  eps_c=c_tol, eps_rank=SQRT_EPS, wallclock=wallclock, ls_log=ls_log)

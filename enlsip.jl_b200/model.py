"""Host-side mirror of the reference's model API for the batched hot path.

reference                                   here
---------                                   ----
CnlsModel(residuals, n, m; ...)             CnlsModel(family, starting_point[B,n], data=..., x_low, x_upp, jacobian=...)
   (cnls_model.jl:345-378)                     one model object = a batch of B independent problems of one family
solve!(model; silent, max_iter, scaling,    solve(model, max_iter=..., scaling=..., time_limit=..., abs_tol=...,
       time_limit, abs_tol, rel_tol,               rel_tol=..., c_tol=..., x_tol=...)        (alias solve_b)
       c_tol, x_tol)   (solver.jl:62-91)
status(model)          (cnls_model.jl:206)  status(model)            -> list of B symbols (strings)
solution(model)        (cnls_model.jl:213)  solution(model)          -> [B, n]
sum_sq_residuals(model)(cnls_model.jl:221)  sum_sq_residuals(model)  -> [B]
constraints_values     (cnls_model.jl:293)  constraints_values(model) -> [B, q+l+2n] (bounds part only here)
total_nb_constraints   (cnls_model.jl:238)  total_nb_constraints(model)

The reference's closures cannot cross to the GPU: `family` names a device functor compiled into the
library (see include/enlsip_b200.h).  Arrays may be numpy arrays (host buffers: the library copies
host<->device inside the call) or torch CUDA tensors (device buffers, zero copy, enqueued on the
current torch stream).  There is no CPU execution path.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import capi
from .model_large import LargeCnlsModel, solve_large

dict_status_codes = {0: "unsolved", 1: "found_first_order_stationary_point", -1: "failed",
                     -2: "maximum_iterations_exceeded", -11: "time_limit_exceeded"}   # cnls_model.jl:180-186

_FAMILIES = {"hs65": capi.FAMILY_HS65, "gauss_peaks": capi.FAMILY_GAUSS_PEAKS, "osborne2": capi.FAMILY_OSBORNE2,
             "chained_rosenbrock10": capi.FAMILY_CHAINED_ROSENBROCK10, "chained_wood20": capi.FAMILY_CHAINED_WOOD20}
_FAMILY_DATA = {"hs65": (), "gauss_peaks": ("y", "S"), "osborne2": ("t", "y"), "chained_rosenbrock10": (),
                "chained_wood20": ()}
_JAC = {"analytic": capi.JAC_ANALYTIC, "forward_diff": capi.JAC_FORWARD_DIFF}


def _is_torch(a):
    return type(a).__module__.startswith("torch")


class UserFamily:
    """A problem family given as CUDA C++ source: the device-side counterpart of the closures
    ``residuals``, ``eq_constraints``, ``ineq_constraints`` (+ ``jacobian_*``) of the reference constructor
    (cnls_model.jl:345-359).  See include/enlsip_b200.h (enlsipb200_compile_family) for the functions the source
    defines.  ``data`` names up to three data slots: slots 0 / 1 hold ``stride0`` / ``stride1`` doubles per problem,
    slot 2 is shared by the whole batch."""

    def __init__(self, source, n, m, nb_eqcons=0, nb_ineqcons=0, data=(), stride0=0, stride1=0, has_jacobians=False,
                 name="user"):
        self.source, self.n, self.m, self.nb_eqcons, self.nb_ineqcons = source, int(n), int(m), int(nb_eqcons), int(nb_ineqcons)
        self.data, self.stride0, self.stride1 = tuple(data), int(stride0), int(stride1)
        self.has_jacobians, self.name = bool(has_jacobians), name
        self._lib = None

    def library(self):
        if self._lib is None:
            self._lib = capi.compile_family(self.source, self.n, self.m, self.nb_eqcons, self.nb_ineqcons, self.stride0,
                                            self.stride1, self.has_jacobians, name=self.name)
        return self._lib


class CnlsModel:
    """A batch of B constrained nonlinear least squares problems of one device family."""

    def __init__(self, family, starting_point, data=None, x_low=None, x_upp=None, jacobian="analytic", device=-1):
        if isinstance(family, UserFamily):
            self._lib, fam_id, data_keys = family.library(), capi.FAMILY_USER, family.data
            if jacobian == "analytic" and not family.has_jacobians:
                jacobian = "forward_diff"      # no jacobian_* given: the reference differentiates for the user as well
            family = family.name
        elif family in _FAMILIES:
            self._lib, fam_id, data_keys = capi.lib(), _FAMILIES[family], _FAMILY_DATA[family]
        else:
            raise AssertionError("A device problem family must be provided: %s or a UserFamily" % sorted(_FAMILIES))
        if jacobian not in _JAC:
            raise AssertionError("jacobian must be 'analytic' or 'forward_diff'")
        self.family = family
        self.jacobian = jacobian
        self.starting_point = starting_point
        self.on_device = _is_torch(starting_point) and starting_point.is_cuda
        if _is_torch(starting_point) and not self.on_device:
            raise ValueError("torch tensors must live on a CUDA device (no CPU path); pass numpy arrays for host buffers")
        if starting_point.ndim != 2:
            raise ValueError("starting_point must be [B, n]")
        self.B, n = int(starting_point.shape[0]), int(starting_point.shape[1])
        self.x_low = np.full(n, -np.inf) if x_low is None else np.ascontiguousarray(x_low, dtype=np.float64)
        self.x_upp = np.full(n, np.inf) if x_upp is None else np.ascontiguousarray(x_upp, dtype=np.float64)
        if self.x_low.shape != (n,) or self.x_upp.shape != (n,):
            raise ValueError("x_low / x_upp must have n = %d entries" % n)       # cnls_model.jl:363-365
        if _is_torch(starting_point):
            import torch
            if starting_point.dtype != torch.float64:
                raise ValueError("starting_point must be float64")
        h = ctypes.c_void_p()
        capi.check(self._lib.enlsipb200_create(fam_id, self.x_low.ctypes.data, self.x_upp.ctypes.data,
                                               device, ctypes.byref(h)), self._lib)
        self._h = h
        dims = [ctypes.c_int() for _ in range(5)]
        capi.check(self._lib.enlsipb200_dims(h, *[ctypes.byref(d) for d in dims]), self._lib)
        self.nb_parameters, self.nb_residuals, self.nb_eqcons, self.nb_constraints, self.lmax = [d.value for d in dims]
        if n != self.nb_parameters:
            raise ValueError("family %s has n=%d, starting_point has %d columns" % (family, self.nb_parameters, n))
        self._data_keep = {}
        data = data or {}
        for slot, key in enumerate(data_keys):
            if key not in data:
                raise AssertionError("family %s needs data[%r]" % (family, key))
            self.set_data(slot, data[key])
        # result fields (cnls_model.jl:159-163)
        self.status_code = None
        self.sol = starting_point
        self.obj_value = None
        self.exit_code = None
        self.iterations = None
        self.nb_active = None
        self.active = None
        self.counters = None
        self.trace = None
        self.kernel_ms = None

    # -- plumbing ---------------------------------------------------------------------------
    def _stream(self):
        if self.on_device:
            import torch
            return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        return None

    def _ptr(self, a):
        """Raw pointer of a zero-copy buffer.  Torch tensors are checked here (CUDA, contiguous, float64 or int32):
        the library reads and writes them by address, a float32 tensor would be a silent out-of-bounds access."""
        if a is None:
            return None
        if _is_torch(a):
            import torch
            if not a.is_cuda or not a.is_contiguous() or a.dtype not in (torch.float64, torch.int32):
                raise ValueError("device buffers must be contiguous CUDA tensors of dtype float64 (int32 for the integer outputs)")
            return ctypes.c_void_p(a.data_ptr())
        if a.dtype not in (np.float64, np.int32) or not a.flags["C_CONTIGUOUS"]:
            raise ValueError("host buffers must be C-contiguous float64 (int32 for the integer outputs) arrays")
        return ctypes.c_void_p(a.ctypes.data)

    def set_data(self, slot, arr):
        dev = _is_torch(arr) and arr.is_cuda
        if _is_torch(arr):
            if not arr.is_contiguous():
                arr = arr.contiguous()
            count = arr.numel()
        else:
            arr = np.ascontiguousarray(arr, dtype=np.float64)
            count = arr.size
        self._data_keep[slot] = arr        # host arrays are uploaded by the next solve: keep them alive
        capi.check(self._lib.enlsipb200_set_data(self._h, slot, self._ptr(arr), count, 1 if dev else 0, self._stream()),
                   self._lib)

    def kernel_info(self):
        v = [ctypes.c_int() for _ in range(6)]
        capi.check(self._lib.enlsipb200_kernel_info(self._h, *[ctypes.byref(d) for d in v]), self._lib)
        keys = ("regs_per_thread", "smem_bytes_per_cta", "threads_per_cta", "ctas_per_sm", "grid", "lanes_per_problem")
        return dict(zip(keys, [d.value for d in v]))

    def launch_count(self):
        return int(self._lib.enlsipb200_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.enlsipb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def solve(model: CnlsModel, silent=True, max_iter=100, scaling=False, time_limit=1e3, abs_tol=None, rel_tol=None,
          c_tol=None, x_tol=None, trace_cap=0, want_active=True, want_counters=True, out=None):
    """solve!(model; ...) (solver.jl:62-91).  Results are written into the model, returns None.

    ``out`` (optional) supplies preallocated result arrays (dict with x, f, exit_code, status, iters, nact and
    optionally active, counters) so that repeated solves allocate nothing.
    """
    if isinstance(model, LargeCnlsModel):   # one large problem: TSQR engine (model_large.py)
        if out is not None:
            raise ValueError("`out` is a batched-regime argument")
        return solve_large(model, silent=silent, max_iter=max_iter, scaling=scaling, time_limit=time_limit,
                           abs_tol=abs_tol, rel_tol=rel_tol, c_tol=c_tol, x_tol=x_tol, trace_cap=trace_cap)
    B, n = model.B, model.nb_parameters
    if not silent and trace_cap == 0 and B == 1:
        trace_cap = int(max_iter) + 1          # the iteration table of print_diagnosis needs the log
    o = capi.default_options()
    o.max_iter = int(max_iter)
    o.scaling = 1 if scaling else 0
    o.jac_mode = _JAC[model.jacobian]
    o.time_limit = float(time_limit)
    nan = float("nan")
    o.abs_tol = nan if abs_tol is None else float(abs_tol)
    o.rel_tol = nan if rel_tol is None else float(rel_tol)
    o.c_tol = nan if c_tol is None else float(c_tol)
    o.x_tol = nan if x_tol is None else float(x_tol)
    row_w = capi.TRACE_HDR + n
    if model.on_device:
        import torch
        dev = model.starting_point.device
        x0 = model.starting_point.contiguous()
        mk_d = lambda *s: torch.empty(*s, dtype=torch.float64, device=dev)
        mk_i = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
        mk_z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=dev)
    else:
        x0 = np.ascontiguousarray(model.starting_point, dtype=np.float64)
        mk_d = lambda *s: np.empty(s, dtype=np.float64)
        mk_i = lambda *s: np.empty(s, dtype=np.int32)
        mk_z = lambda *s: np.zeros(s, dtype=np.float64)
    out = out or {}
    x = out.get("x") if out.get("x") is not None else mk_d(B, n)
    f = out.get("f") if out.get("f") is not None else mk_d(B)
    ec = out.get("exit_code") if out.get("exit_code") is not None else mk_i(B)
    st = out.get("status") if out.get("status") is not None else mk_i(B)
    it = out.get("iters") if out.get("iters") is not None else mk_i(B)
    na = out.get("nact") if out.get("nact") is not None else mk_i(B)
    act = out.get("active") if out.get("active") is not None else (mk_i(B, model.lmax) if want_active else None)
    cnt = out.get("counters") if out.get("counters") is not None else (mk_i(B, 2) if want_counters else None)
    tr = mk_z(B, trace_cap, row_w) if trace_cap > 0 else None
    p = model._ptr
    capi.check(model._lib.enlsipb200_solve_batch(model._h, B, p(x0), ctypes.byref(o), p(x), p(f), p(ec), p(st), p(it),
                                                 p(na), p(act), p(cnt), p(tr), int(trace_cap),
                                                 1 if model.on_device else 0, model._stream()), model._lib)
    # solver.jl:84-87
    model.status_code = st
    model.exit_code = ec
    model.sol = x
    model.obj_value = f
    model.iterations = it
    model.nb_active = na
    model.active = act
    model.counters = cnt
    model.trace = tr
    if not silent:
        print(_diagnosis(model))
    return None


solve_b = solve   # `solve!`


def evaluate(model: CnlsModel, x=None, want=("r", "J", "c", "A")):
    """new_point! (enlsip_functions.jl:34-52) for the whole batch: residuals r [B, m], Jacobians J [B, n, m] (per problem
    m x n column major; forward differences of cnls_model.jl:65-82 when the model was built with
    jacobian="forward_diff"), constraint values c [B, lmax] and constraint Jacobians A [B, lmax, n] at `x` (default:
    the model's current solution).  Device tensors in, device tensors out."""
    import torch
    x = model.sol if x is None else x
    if not (_is_torch(x) and x.is_cuda):
        x = torch.as_tensor(np.ascontiguousarray(_to_numpy(x), dtype=np.float64)).cuda()
    x = x.contiguous()
    assert x.dtype == torch.float64 and x.shape == (model.B, model.nb_parameters)
    B, n, m, lmax = model.B, model.nb_parameters, model.nb_residuals, model.lmax
    mk = lambda *s: torch.empty(*s, dtype=torch.float64, device=x.device)
    out = {"r": mk(B, m) if "r" in want else None, "J": mk(B, n, m) if "J" in want else None,
           "c": mk(B, lmax) if "c" in want else None, "A": mk(B, lmax, n) if "A" in want else None}
    o = capi.default_options()
    o.jac_mode = _JAC[model.jacobian]
    p = model._ptr
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.check(model._lib.enlsipb200_eval_batch(model._h, B, p(x), ctypes.byref(o), p(out["r"]), p(out["J"]), p(out["c"]),
                                                p(out["A"]), 1, st), model._lib)
    return out


def gn_step(model: CnlsModel, x, ev, scaling=False):
    """One Gauss-Newton step of every problem from MATERIALISED r, J, c, A (the dict `evaluate` returns): the batched step
    kernel (enlsipb200_step_batch; update_working_set, enlsip_functions.jl:686-795).  Returns dict(p [B, n], lam [B, T],
    active [B, lmax], info [B, 5] = t, rankA, rankJ2, index_del, error)."""
    import torch
    B, n, lmax = model.B, model.nb_parameters, model.lmax
    T = min(lmax, n)
    dev = ev["r"].device
    out = {"p": torch.empty(B, n, dtype=torch.float64, device=dev), "lam": torch.empty(B, T, dtype=torch.float64, device=dev),
           "active": torch.empty(B, lmax, dtype=torch.int32, device=dev), "info": torch.empty(B, 5, dtype=torch.int32, device=dev)}
    o = capi.default_options()
    o.scaling = 1 if scaling else 0
    p = model._ptr
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.check(model._lib.enlsipb200_step_batch(model._h, B, p(x.contiguous()), p(ev["r"]), p(ev["J"]), p(ev["c"]), p(ev["A"]),
                                                ctypes.byref(o), p(out["p"]), p(out["lam"]), p(out["active"]), p(out["info"]), 1, st),
               model._lib)
    return out


def last_kernel_ms(model: CnlsModel) -> float:
    ms = ctypes.c_float()
    capi.check(model._lib.enlsipb200_last_kernel_ms(model._h, ctypes.byref(ms)), model._lib)
    return float(ms.value)


def _to_numpy(a):
    return a.detach().cpu().numpy() if _is_torch(a) else np.asarray(a)


def status(model: CnlsModel):
    """cnls_model.jl:206 : one symbol per problem (``:unsolved`` before solve!)."""
    if model.status_code is None:
        return [dict_status_codes[0]] * model.B
    return [dict_status_codes[int(c)] for c in _to_numpy(model.status_code)]


def solution(model: CnlsModel):
    return model.sol


def sum_sq_residuals(model: CnlsModel):
    return model.obj_value


def total_nb_constraints(model: CnlsModel):
    return model.nb_constraints


def constraints_values(model: CnlsModel):
    """Bounds part of cnls_model.jl:293-309: ``[x - x_low; x_upp - x]`` (length 2n, +-Inf kept)."""
    sol = _to_numpy(model.sol)
    return np.concatenate([sol - model.x_low[None, :], model.x_upp[None, :] - sol], axis=1)


def iteration_table(model, b=0):
    """The per-iteration table of the reference's `print_diagnosis` (EF:2551-2554, 2571-2580) for problem `b`, rebuilt
    from the iteration log the engine returns (needs a solve with trace_cap > 0): iteration, objective,
    ||active constraints||^2, ||p||, steplength, reduction of the objective."""
    if model.trace is None:
        return "(no iteration log: solve with trace_cap > 0 or silent=False)"
    tr = _to_numpy(model.trace)[b]
    k = min(int(_to_numpy(model.iterations)[b]), tr.shape[0])
    lines = ["iter    objective   ||active_constraints||^2  ||p||       alpha     reduction"]
    for i in range(k):
        lines.append("%4d  %.7e       %.2e         %.2e  %.2e  %.3e" % (i + 1, tr[i, 0], tr[i, 11], tr[i, 8], tr[i, 7], tr[i, 12]))
    return "\n".join(lines)


def _diagnosis(model: CnlsModel):
    if model.B == 1 and model.trace is not None:
        cnt = _to_numpy(model.counters)[0] if model.counters is not None else (0, 0)
        return "\n".join(["%s problem (n=%d, m=%d, constraints=%d)" % (model.family, model.nb_parameters, model.nb_residuals,
                                                                       model.nb_constraints),
                          iteration_table(model, 0),
                          "Number of iterations...................: %4d" % int(_to_numpy(model.iterations)[0]),
                          "Square sum of residuals................: %.7e" % float(_to_numpy(model.obj_value)[0]),
                          "Number of function evaluations.........: %4d" % int(cnt[0]),
                          "Number of Jacobian matrix evaluations..: %4d" % int(cnt[1]),
                          "Termination status.....................: %s" % status(model)[0]])
    st = status(model)
    it = _to_numpy(model.iterations)
    f = _to_numpy(model.obj_value)
    lines = ["batch of %d %s problems (n=%d, m=%d, constraints=%d)" % (model.B, model.family, model.nb_parameters,
                                                                          model.nb_residuals, model.nb_constraints)]
    for s in sorted(set(st)):
        lines.append("  %-40s %d" % (s, st.count(s)))
    lines.append("  mean iterations %.2f, mean objective %.6e" % (float(it.mean()), float(f.mean())))
    return "\n".join(lines)

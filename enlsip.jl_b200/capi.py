"""ctypes binding of include/enlsip_b200.h (the same symbols the Julia host file binds with ccall)."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libenlsip_b200.so")
SRC = os.path.join(_HERE, "csrc", "enlsip_b200.cu")
HEADERS = [os.path.join(_HERE, "csrc", f) for f in ("enl_base.h", "enl_linalg.h", "enl_families.h", "enl_solver.h")] + \
          [os.path.join(os.path.dirname(_HERE), "include", "enlsip_b200.h")]

FAMILY_HS65 = 0
FAMILY_GAUSS_PEAKS = 1
FAMILY_OSBORNE2 = 2
FAMILY_CHAINED_ROSENBROCK10 = 3
FAMILY_CHAINED_WOOD20 = 4
JAC_ANALYTIC = 0
JAC_FORWARD_DIFF = 1
TRACE_HDR = 16
EXIT_WOULD_THROW, EXIT_WOULD_HANG, EXIT_CAPACITY = -99, -98, -97

EXPORTS = ["enlsipb200_version", "enlsipb200_last_error", "enlsipb200_default_options", "enlsipb200_create",
           "enlsipb200_destroy", "enlsipb200_dims", "enlsipb200_set_data", "enlsipb200_solve_batch",
           "enlsipb200_last_kernel_ms", "enlsipb200_kernel_info", "enlsipb200_launch_count", "enlsipb200_det_exp"]


class Options(ctypes.Structure):
    _fields_ = [("max_iter", ctypes.c_int), ("scaling", ctypes.c_int), ("jac_mode", ctypes.c_int),
                ("reserved", ctypes.c_int), ("time_limit", ctypes.c_double), ("abs_tol", ctypes.c_double),
                ("rel_tol", ctypes.c_double), ("c_tol", ctypes.c_double), ("x_tol", ctypes.c_double)]


NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def build(force=False, verbose=False):
    """Compile csrc/enlsip_b200.cu for sm_100a into lib/libenlsip_b200.so (in-tree)."""
    newest = max(os.path.getmtime(p) for p in [SRC] + HEADERS)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [SRC, "-o", LIB_PATH]
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    """Load the CUDA library.  Raises if it has not been built: no fallback of any kind."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libenlsip_b200.so is missing (%s): run __graft_entry__.build(); "
                               "the engine has no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        vp, ip, dp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
        L.enlsipb200_version.restype = ctypes.c_int
        L.enlsipb200_last_error.restype = ctypes.c_char_p
        L.enlsipb200_default_options.argtypes = [ctypes.POINTER(Options)]
        L.enlsipb200_default_options.restype = None
        L.enlsipb200_create.argtypes = [ctypes.c_int, vp, vp, ctypes.c_int, ctypes.POINTER(vp)]
        L.enlsipb200_destroy.argtypes = [vp]
        L.enlsipb200_dims.argtypes = [vp, ip, ip, ip, ip, ip]
        L.enlsipb200_set_data.argtypes = [vp, ctypes.c_int, vp, ctypes.c_longlong, ctypes.c_int, vp]
        L.enlsipb200_solve_batch.argtypes = [vp, ctypes.c_longlong, vp, ctypes.POINTER(Options)] + [vp] * 9 + \
                                            [ctypes.c_int, ctypes.c_int, vp]
        L.enlsipb200_last_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
        L.enlsipb200_kernel_info.argtypes = [vp, ip, ip, ip, ip, ip, ip]
        L.enlsipb200_launch_count.argtypes = [vp]
        L.enlsipb200_launch_count.restype = ctypes.c_longlong
        L.enlsipb200_det_exp.argtypes = [vp, vp, ctypes.c_longlong, ctypes.c_int]
        _lib = L
    return _lib


class EngineError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise EngineError("enlsip_b200 error %d: %s" % (rc, lib().enlsipb200_last_error().decode()))


def default_options() -> Options:
    o = Options()
    lib().enlsipb200_default_options(ctypes.byref(o))
    return o

"""ctypes binding of the C ABI in include/ocs_b200.h (libocs_b200.so, built in-tree
by optimal-control-solvers_b200/build.py).  There is no fallback: if the library is missing the import
of anything that computes fails with an instruction to build it, and on a machine
without a CUDA device every compute call raises OcsError."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# OCS_B200_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("OCS_B200_LIB") or os.path.join(_HERE, "libocs_b200.so")

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int32)
c_lp = ctypes.POINTER(ctypes.c_long)
c_h = ctypes.c_void_p

OK, ERR_ARG, ERR_CUDA, ERR_COMPILE, ERR_STATE, ERR_NOMEM = range(6)
MATH_FAST, MATH_STRICT = 0, 1
PROBLEM_TEST = 0
REDUCE_SUM, REDUCE_MIN, REDUCE_MAX = 0, 1, 2

# every symbol include/ocs_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "ocs_b200_create": (ctypes.c_int, [ctypes.c_int, c_dp, ctypes.c_int, c_dp, ctypes.c_int, ctypes.c_int,
                                       ctypes.POINTER(c_h)]),
    "ocs_b200_create_multi": (ctypes.c_int, [ctypes.c_int, c_dp, ctypes.c_int, c_dp, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                             ctypes.c_int, ctypes.POINTER(c_h)]),
    "ocs_b200_get_devices": (ctypes.c_int, [c_h, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), c_lp, ctypes.c_int]),
    "ocs_b200_get_dims": (ctypes.c_int, [c_h, c_ip]),
    "ocs_b200_get_batch": (ctypes.c_long, [c_h]),
    "ocs_b200_evaluate_resident": (ctypes.c_int, [c_h, c_dp, ctypes.c_long, ctypes.c_int]),
    "ocs_b200_gather_results_dev": (ctypes.c_int, [c_h, ctypes.c_void_p, ctypes.c_void_p]),
    "ocs_b200_get_results": (ctypes.c_int, [c_h, c_dp, c_dp, c_ip, c_dp, c_dp]),
    "ocs_b200_kernel_stats_ex": (ctypes.c_int, [c_h, ctypes.c_int, c_dp, ctypes.c_int, c_lp]),
    "ocs_b200_destroy": (ctypes.c_int, [c_h]),
    "ocs_b200_last_error": (ctypes.c_char_p, [c_h]),
    "ocs_b200_abi_version": (ctypes.c_int, []),
    "ocs_b200_get_grid": (ctypes.c_int, [c_h, ctypes.c_int, c_dp, c_dp]),
    "ocs_b200_set_problem_builtin": (ctypes.c_int, [c_h, ctypes.c_int]),
    "ocs_b200_set_problem_source": (ctypes.c_int, [c_h, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_char_p]),
    "ocs_b200_set_params": (ctypes.c_int, [c_h, c_dp, ctypes.c_int, ctypes.c_long]),
    "ocs_b200_set_tail_control": (ctypes.c_int, [c_h, c_dp, ctypes.c_int]),
    "ocs_b200_set_control_basis": (ctypes.c_int, [c_h, ctypes.c_int, c_dp]),
    "ocs_b200_get_control_index": (ctypes.c_int, [c_h, c_ip, c_ip]),
    "ocs_b200_set_control_v": (ctypes.c_int, [c_h, c_dp, ctypes.c_long]),
    "ocs_b200_set_control_u": (ctypes.c_int, [c_h, c_dp, ctypes.c_long]),
    "ocs_b200_get_control_u": (ctypes.c_int, [c_h, ctypes.c_long, ctypes.c_long, c_dp]),
    "ocs_b200_compute_states": (ctypes.c_int, [c_h, c_dp, ctypes.c_long, c_dp, c_dp, c_ip]),
    "ocs_b200_compute_adjoints": (ctypes.c_int, [c_h, c_dp, ctypes.c_long, c_dp, c_dp]),
    "ocs_b200_fetch_states": (ctypes.c_int, [c_h, ctypes.c_int, ctypes.c_long, ctypes.c_long, c_dp]),
    "ocs_b200_fetch_adjoints": (ctypes.c_int, [c_h, ctypes.c_int, ctypes.c_long, ctypes.c_long, c_dp, c_dp]),
    "ocs_b200_fb_sweep": (ctypes.c_int, [c_h, c_dp, ctypes.c_long, c_dp, ctypes.c_long, c_dp, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_int, ctypes.c_double, c_dp, c_ip, c_ip, c_dp]),
    "ocs_b200_single_shooting": (ctypes.c_int, [c_h, c_dp, ctypes.c_long, c_dp, ctypes.c_long, c_dp, c_dp, ctypes.c_int,
                                                c_ip, c_dp, c_dp, ctypes.c_int, ctypes.c_double, c_dp, c_dp, c_ip, c_ip]),
    "ocs_b200_set_solve_control_bounds": (ctypes.c_int, [c_h, c_dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int]),
    "ocs_b200_get_solve_violation": (ctypes.c_int, [c_h, c_dp]),
    "ocs_b200_reduce_J": (ctypes.c_int, [c_h, ctypes.c_int, c_dp, c_lp]),
    "ocs_b200_set_stream": (ctypes.c_int, [c_h, ctypes.c_void_p]),
    "ocs_b200_set_control_v_dev": (ctypes.c_int, [c_h, ctypes.c_void_p, ctypes.c_long]),
    "ocs_b200_compute_states_dev": (ctypes.c_int, [c_h, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_void_p]),
    "ocs_b200_compute_adjoints_dev": (ctypes.c_int, [c_h, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
                                                     ctypes.c_void_p]),
    "ocs_b200_evaluate_dev": (ctypes.c_int, [c_h, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    "ocs_b200_synchronize": (ctypes.c_int, [c_h]),
    "ocs_b200_evaluate": (ctypes.c_int, [c_h, c_dp, ctypes.c_long, c_dp, ctypes.c_long, c_dp, c_dp, c_ip, c_dp, c_dp]),
    "ocs_b200_eval_solution": (ctypes.c_int, [c_h, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int, c_dp, c_dp, c_dp, c_dp]),
    "ocs_b200_set_control_ufunc": (ctypes.c_int, [c_h, ctypes.c_int, ctypes.c_int, c_dp]),
    "ocs_b200_compute_equilibrium": (ctypes.c_int, [c_h, c_dp, ctypes.c_int, ctypes.c_long, ctypes.c_long, c_dp, ctypes.c_long, c_dp, c_dp,
                                                    c_dp, ctypes.c_long, ctypes.c_int, ctypes.c_double, c_dp, c_dp, c_dp,
                                                    c_ip, c_ip]),
    "ocs_b200_set_fb_costate": (ctypes.c_int, [c_h, ctypes.c_int]),
    "ocs_b200_kernel_stats": (ctypes.c_int, [c_h, ctypes.c_int, c_dp, c_dp, c_lp]),
    "ocs_b200_set_profiling": (ctypes.c_int, [c_h, ctypes.c_int]),
    "ocs_b200_measure_fp64_peak": (ctypes.c_int, [ctypes.c_int, c_dp]),
    "ocs_b200_measure_fp64_peak_ex": (ctypes.c_int, [ctypes.c_int, c_dp, c_dp]),
    "ocs_b200_measure_hbm": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t, c_dp, c_dp]),
}


class OcsError(RuntimeError):
    """A non-zero status from the C ABI; .code holds it."""

    def __init__(self, code, msg):
        super().__init__(f"ocs_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Loads libocs_b200.so and types every entry point.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python optimal-control-solvers_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the RK4 sweeps.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dp(a):
    """double* of a float64 array (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64
    return a.ctypes.data_as(c_dp)


def ip(a):
    if a is None:
        return None
    assert a.dtype == np.int32
    return a.ctypes.data_as(c_ip)


def f64(a, order="F"):
    """float64 array in MATLAB (column-major) memory order."""
    return np.require(np.asarray(a, dtype=np.float64), requirements=["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS", "ALIGNED"])


def check(rc, handle=None):
    if rc != 0:
        msg = load().ocs_b200_last_error(handle)
        raise OcsError(rc, msg.decode("utf-8", "replace") if msg else "")

"""ctypes binding of libpogs_b200.so (declarations: include/pogs_b200.h).

The library is the product: hand-written sm_100a CUDA behind the reference's C
ABI.  There is no CPU fallback -- if the shared object is missing or cannot be
loaded this module raises ImportError, and every entry point returns POGS_ERROR
when no CUDA device is usable.
"""
import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libpogs_b200.so")

c_f, c_d, c_i, c_u, c_sz = ctypes.c_float, ctypes.c_double, ctypes.c_int, ctypes.c_uint, ctypes.c_size_t
P = ctypes.POINTER


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python build_native.py, or __graft_entry__.build()); pogs_b200 has no CPU fallback")
    try:
        return ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise ImportError(f"cannot load {LIB_PATH}: {e}") from e


lib = _load()


def _sig(name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


def _desc(ct):
    return [P(ct)] * 5 + [P(c_i)]


for _sfx, _ct in (("D", c_d), ("S", c_f)):
    _tail = _desc(_ct) + _desc(_ct) + [_ct, _ct, _ct, c_u, c_u, c_i, c_i, P(_ct), P(_ct), P(_ct), P(_ct), P(c_u)]
    _sig("Pogs" + _sfx, c_i, [c_i, c_sz, c_sz, P(_ct)] + _tail)
    _sig("PogsSparse" + _sfx, c_i, [c_i, c_sz, c_sz, c_sz, P(_ct), P(c_i), P(c_i)] + _tail)

for _sfx, _ct in (("d", c_d), ("s", c_f)):
    _sig("pogs_b200_create_dense_" + _sfx, ctypes.c_void_p, [c_i, c_sz, c_sz, ctypes.c_void_p, c_i])
    _sig("pogs_b200_create_dense_indirect_" + _sfx, ctypes.c_void_p, [c_i, c_sz, c_sz, ctypes.c_void_p, c_i])
    _sig("pogs_b200_create_sparse_" + _sfx, ctypes.c_void_p, [c_i, c_sz, c_sz, c_sz, P(_ct), P(c_i), P(c_i)])
    _sig("pogs_b200_set_init_" + _sfx, c_i, [ctypes.c_void_p, P(_ct), P(_ct)])
    _sig("pogs_b200_solve_" + _sfx, c_i, [ctypes.c_void_p] + _desc(_ct) + _desc(_ct))
    _sig("pogs_b200_get_solution_" + _sfx, c_i,
         [ctypes.c_void_p, P(_ct), P(_ct), P(_ct), P(_ct), P(_ct), P(c_u), P(_ct)])
    _sig("pogs_b200_prox_eval_" + _sfx, c_i, [c_sz, P(c_i)] + [P(_ct)] * 5 + [_ct, P(_ct), P(_ct)])
    _sig("pogs_b200_func_eval_" + _sfx, c_i, [c_sz, P(c_i)] + [P(_ct)] * 5 + [P(_ct), P(c_d)])
    _sig("pogs_b200_gemv_" + _sfx, c_i, [c_i, c_sz, c_sz, P(_ct), c_i, c_i, P(_ct), P(_ct)])
    _sig("pogs_b200_get_equil_" + _sfx, c_i, [ctypes.c_void_p, P(_ct), P(_ct), P(_ct)])
    _sig("pogs_b200_project_" + _sfx, c_i, [ctypes.c_void_p, P(_ct), P(_ct), P(_ct), P(_ct)])
    _sig("pogs_b200_create_dense_rowblock_" + _sfx, ctypes.c_void_p,
         [c_sz, c_sz, c_sz, ctypes.c_void_p, c_i, ctypes.c_void_p])
    _sig("pogs_b200_comm_allreduce_" + _sfx, c_i, [ctypes.c_void_p, ctypes.c_void_p, c_sz])
_sig("pogs_b200_comm_create", ctypes.c_void_p, [c_i, c_i, c_sz])
_sig("pogs_b200_comm_handle", c_i, [ctypes.c_void_p, ctypes.c_void_p])
_sig("pogs_b200_comm_open", c_i, [ctypes.c_void_p, ctypes.c_void_p])
_sig("pogs_b200_comm_destroy", None, [ctypes.c_void_p])
_sig("pogs_b200_destroy", None, [ctypes.c_void_p])
_sig("pogs_b200_set_params", c_i, [ctypes.c_void_p, c_d, c_d, c_d, c_u, c_u, c_i, c_i])
_sig("pogs_b200_set_rho", c_i, [ctypes.c_void_p, c_d])
_sig("pogs_b200_set_profile", c_i, [ctypes.c_void_p, c_i])
_sig("pogs_b200_get_timing", c_i, [ctypes.c_void_p, P(c_d)])
_sig("pogs_b200_get_stats", c_i, [ctypes.c_void_p, P(c_d)])
_sig("pogs_b200_get_pass_phases", c_i, [ctypes.c_void_p, P(c_d)])
_sig("pogs_b200_gram_s", c_i, [c_sz, c_sz, P(c_f), P(c_f), c_i])
_sig("pogs_b200_gram_debug_s", c_i, [c_sz, c_sz, P(c_f), P(c_f), P(c_f), P(c_f)])
_sig("pogs_b200_plan_sparse_tiles", c_i, [c_sz, c_sz, c_sz, ctypes.c_uint, c_sz, P(ctypes.c_ulonglong)])
_sig("pogs_b200_trim_memory", None, [])
_sig("pogs_b200_last_error", ctypes.c_char_p, [])
_sig("pogs_b200_launch_count", ctypes.c_ulonglong, [])


def ctype_of(dtype):
    return c_d if np.dtype(dtype) == np.float64 else c_f


def suffix(dtype, upper=False):
    s = "d" if np.dtype(dtype) == np.float64 else "s"
    return s.upper() if upper else s


def ptr(a, ct):
    return a.ctypes.data_as(P(ct))


def last_error():
    return lib.pogs_b200_last_error().decode(errors="replace")


def trim_memory():
    """Give the blocks cached by the library's device memory pool back to the driver."""
    lib.pogs_b200_trim_memory()


def launch_count():
    return int(lib.pogs_b200_launch_count())

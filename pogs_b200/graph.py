"""Graph-form solver API -- same surface as the reference's python/pogs/graph.py
(Function, FunctionObj, _solve_graph_form and the seven solve_* wrappers with
the same argument names, defaults and result keys), bound to libpogs_b200.so.

    minimize    sum_i f_i(y_i) + sum_j g_j(x_j)     subject to  y = A x
    f_i(v) = c_i h_i(a_i v - b_i) + d_i v + (e_i/2) v^2

Differences, all additive:
  * f and g may be given as `FunctionVector`s (struct-of-arrays descriptors built
    with numpy) instead of Python lists of `FunctionObj`; the wrappers build those
    directly, so no per-row Python objects are created (the reference builds m of
    them per call, graph.py:428);
  * `dtype=np.float32` selects the single-precision entry points (PogsS /
    PogsSparseS).  The default keeps the reference behaviour: everything is
    converted to float64 and PogsD / PogsSparseD are called (graph.py:288,318,352).
"""
import ctypes
from enum import IntEnum

import numpy as np

from . import _lib

try:
    import scipy.sparse as sp

    HAS_SCIPY = True
except ImportError:  # pragma: no cover
    HAS_SCIPY = False


class Ordering(IntEnum):
    """Matrix ordering (reference graph.py:107-111, pogs_c.h:51)."""

    COL_MAJ = 0
    ROW_MAJ = 1


class Function(IntEnum):
    """h_i / h_j tags (reference graph.py:114-132, prox_lib.h:23-38)."""

    kAbs = 0
    kExp = 1
    kHuber = 2
    kIdentity = 3
    kIndBox01 = 4
    kIndEq0 = 5
    kIndGe0 = 6
    kIndLe0 = 7
    kLogistic = 8
    kMaxNeg0 = 9
    kMaxPos0 = 10
    kNegEntr = 11
    kNegLog = 12
    kRecipr = 13
    kSquare = 14
    kZero = 15


class FunctionObj:
    """c * h(a*x - b) + d*x + e*x^2 for one coordinate (reference graph.py:135-165)."""

    def __init__(self, h=Function.kZero, a=1.0, b=0.0, c=1.0, d=0.0, e=0.0):
        self.h = h
        self.a = float(a)
        self.b = float(b)
        self.c = float(c)
        self.d = float(d)
        self.e = float(e)


class FunctionVector:
    """Struct-of-arrays descriptor of a separable function of `size` coordinates.
    Every field accepts a scalar or an array of length `size`."""

    def __init__(self, size, h=Function.kZero, a=1.0, b=0.0, c=1.0, d=0.0, e=0.0):
        self.size = int(size)
        self.h = np.ascontiguousarray(np.broadcast_to(np.asarray(h, dtype=np.int32), (self.size,)))
        self.a, self.b, self.c, self.d, self.e = (
            np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.size,)))
            for v in (a, b, c, d, e))

    def __len__(self):
        return self.size

    @staticmethod
    def from_any(f, size=None):
        if isinstance(f, FunctionVector):
            return f
        fv = FunctionVector(len(f))
        if len(f):
            fv.h = np.fromiter((int(fi.h) for fi in f), dtype=np.int32, count=len(f))
            fv.a = np.fromiter((fi.a for fi in f), dtype=np.float64, count=len(f))
            fv.b = np.fromiter((fi.b for fi in f), dtype=np.float64, count=len(f))
            fv.c = np.fromiter((fi.c for fi in f), dtype=np.float64, count=len(f))
            fv.d = np.fromiter((fi.d for fi in f), dtype=np.float64, count=len(f))
            fv.e = np.fromiter((fi.e for fi in f), dtype=np.float64, count=len(f))
        return fv

    def arrays(self, dtype):
        """(a, b, c, d, e, h) contiguous in `dtype` / int32 -- the C ABI order."""
        cast = lambda v: np.ascontiguousarray(v, dtype=dtype)
        return cast(self.a), cast(self.b), cast(self.c), cast(self.d), cast(self.e), np.ascontiguousarray(self.h, dtype=np.int32)


def _desc_ptrs(arrs, ct):
    a, b, c, d, e, h = arrs
    return [_lib.ptr(a, ct), _lib.ptr(b, ct), _lib.ptr(c, ct), _lib.ptr(d, ct), _lib.ptr(e, ct),
            _lib.ptr(h, ctypes.c_int)]


def _solve_graph_form(A, f, g, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500, verbose=0, rho=1.0,
                      adaptive_rho=True, gap_stop=True, dtype=np.float64):
    """One-shot solve through the reference C entry points (graph.py:236-390).

    Returns dict with 'x', 'y', 'l', 'optval', 'iterations' (zero-based index of
    the last iteration, as the reference reports it) and 'status'."""
    dt = np.dtype(dtype)
    ct = _lib.ctype_of(dt)
    is_sparse = HAS_SCIPY and sp.issparse(A)
    if is_sparse:
        A_csr = sp.csr_matrix(A, dtype=dt)
        m, n = A_csr.shape
        data = np.ascontiguousarray(A_csr.data, dtype=dt)
        indptr = np.ascontiguousarray(A_csr.indptr, dtype=np.int32)
        indices = np.ascontiguousarray(A_csr.indices, dtype=np.int32)
    else:
        A = np.asarray(A, dtype=dt, order="C")
        m, n = A.shape
    f = FunctionVector.from_any(f)
    g = FunctionVector.from_any(g)
    assert len(f) == m, f"f should have length {m}, got {len(f)}"
    assert len(g) == n, f"g should have length {n}, got {len(g)}"
    fa, ga = f.arrays(dt), g.arrays(dt)
    x = np.zeros(n, dt)
    y = np.zeros(m, dt)
    dual = np.zeros(m, dt)
    optval = ct()
    final_iter = ctypes.c_uint()
    tail = _desc_ptrs(fa, ct) + _desc_ptrs(ga, ct) + [
        ct(rho), ct(abs_tol), ct(rel_tol), int(max_iter), int(verbose), int(adaptive_rho), int(gap_stop),
        _lib.ptr(x, ct), _lib.ptr(y, ct), _lib.ptr(dual, ct), ctypes.byref(optval), ctypes.byref(final_iter)]
    if is_sparse:
        fn = getattr(_lib.lib, "PogsSparse" + _lib.suffix(dt, upper=True))
        status = fn(int(Ordering.ROW_MAJ), m, n, A_csr.nnz, _lib.ptr(data, ct), _lib.ptr(indptr, ctypes.c_int),
                    _lib.ptr(indices, ctypes.c_int), *tail)
    else:
        fn = getattr(_lib.lib, "Pogs" + _lib.suffix(dt, upper=True))
        status = fn(int(Ordering.ROW_MAJ), m, n, _lib.ptr(A, ct), *tail)
    if int(status) == 6:   # POGS_ERROR: device / library failure, never a solver outcome
        raise RuntimeError("pogs_b200: " + _lib.last_error())
    return {"x": x, "y": y, "l": dual, "optval": float(optval.value), "iterations": int(final_iter.value),
            "status": int(status)}


def _shape(A):
    if HAS_SCIPY and sp.issparse(A):
        return A.shape
    return np.shape(A)


# --- canonical encodings (reference graph.py:393-705) ---------------------------------
def lasso_functions(m, n, b, lambd):
    return (FunctionVector(m, Function.kSquare, 1.0, b, 1.0), FunctionVector(n, Function.kAbs, 1.0, 0.0, lambd))


def ridge_functions(m, n, b, lambd):
    return (FunctionVector(m, Function.kSquare, 1.0, b, 1.0), FunctionVector(n, Function.kSquare, 1.0, 0.0, lambd))


def elastic_net_functions(m, n, b, lambda1, lambda2):
    return (FunctionVector(m, Function.kSquare, 1.0, b, 1.0),
            FunctionVector(n, Function.kAbs, 1.0, 0.0, lambda1, 0.0, lambda2 / 2))


def logistic_functions(m, n, b, lambd):
    g = FunctionVector(n, Function.kAbs, 1.0, 0.0, lambd) if lambd > 0 else FunctionVector(n, Function.kZero)
    return FunctionVector(m, Function.kLogistic, -np.asarray(b, dtype=np.float64), 0.0, 1.0), g


def huber_functions(m, n, b, delta, lambd):
    g = FunctionVector(n, Function.kAbs, 1.0, 0.0, lambd) if lambd > 0 else FunctionVector(n, Function.kZero)
    return FunctionVector(m, Function.kHuber, 1.0 / delta, np.asarray(b, dtype=np.float64) / delta, delta * delta), g


def svm_functions(m, n, b, lambd):
    return (FunctionVector(m, Function.kMaxPos0, -np.asarray(b, dtype=np.float64), -1.0, 1.0),
            FunctionVector(n, Function.kSquare, 1.0, 0.0, lambd))


def nonneg_ls_functions(m, n, b):
    return FunctionVector(m, Function.kSquare, 1.0, b, 1.0), FunctionVector(n, Function.kIndGe0)


def solve_lasso(A, b, lambd, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500, verbose=0, rho=1.0, dtype=np.float64):
    """minimize 0.5*||A x - b||^2 + lambd*||x||_1   (reference graph.py:393-433)."""
    b = np.asarray(b, dtype=np.float64).flatten()
    m, n = _shape(A)
    f, g = lasso_functions(m, n, b, lambd)
    return _solve_graph_form(A, f, g, abs_tol, rel_tol, max_iter, verbose, rho, dtype=dtype)


def solve_ridge(A, b, lambd, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500, verbose=0, rho=1.0, dtype=np.float64):
    """minimize 0.5*||A x - b||^2 + 0.5*lambd*||x||^2   (reference graph.py:436-476)."""
    b = np.asarray(b, dtype=np.float64).flatten()
    m, n = _shape(A)
    f, g = ridge_functions(m, n, b, lambd)
    return _solve_graph_form(A, f, g, abs_tol, rel_tol, max_iter, verbose, rho, dtype=dtype)


def solve_elastic_net(A, b, lambda1, lambda2, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500, verbose=0, rho=1.0,
                      dtype=np.float64):
    """minimize 0.5*||A x - b||^2 + lambda1*||x||_1 + 0.5*lambda2*||x||^2   (reference graph.py:479-524)."""
    b = np.asarray(b, dtype=np.float64).flatten()
    m, n = _shape(A)
    f, g = elastic_net_functions(m, n, b, lambda1, lambda2)
    return _solve_graph_form(A, f, g, abs_tol, rel_tol, max_iter, verbose, rho, dtype=dtype)


def solve_logistic(A, b, lambd=0.0, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500, verbose=0, rho=1.0,
                   dtype=np.float64):
    """minimize sum_i log(1+exp(-b_i a_i'x)) + lambd*||x||_1, b in {-1,+1}   (reference graph.py:527-570)."""
    b = np.asarray(b, dtype=np.float64).flatten()
    m, n = _shape(A)
    f, g = logistic_functions(m, n, b, lambd)
    return _solve_graph_form(A, f, g, abs_tol, rel_tol, max_iter, verbose, rho, dtype=dtype)


def solve_huber(A, b, delta=1.0, lambd=0.0, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500, verbose=0, rho=1.0,
                dtype=np.float64):
    """minimize sum_i huber(a_i'x - b_i, delta) + lambd*||x||_1   (reference graph.py:573-622)."""
    b = np.asarray(b, dtype=np.float64).flatten()
    m, n = _shape(A)
    f, g = huber_functions(m, n, b, delta, lambd)
    return _solve_graph_form(A, f, g, abs_tol, rel_tol, max_iter, verbose, rho, dtype=dtype)


def solve_svm(A, b, lambd=1.0, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500, verbose=0, rho=1.0, dtype=np.float64):
    """minimize sum_i max(0, 1 - b_i a_i'x) + 0.5*lambd*||x||^2   (reference graph.py:625-665)."""
    b = np.asarray(b, dtype=np.float64).flatten()
    m, n = _shape(A)
    f, g = svm_functions(m, n, b, lambd)
    return _solve_graph_form(A, f, g, abs_tol, rel_tol, max_iter, verbose, rho, dtype=dtype)


def solve_nonneg_ls(A, b, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500, verbose=0, rho=1.0, dtype=np.float64):
    """minimize 0.5*||A x - b||^2 subject to x >= 0   (reference graph.py:668-707)."""
    b = np.asarray(b, dtype=np.float64).flatten()
    m, n = _shape(A)
    f, g = nonneg_ls_functions(m, n, b)
    return _solve_graph_form(A, f, g, abs_tol, rel_tol, max_iter, verbose, rho, dtype=dtype)

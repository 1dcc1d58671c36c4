"""Persistent solver object: Python view of the reference's C++ pogs::PogsDirect /
pogs::PogsIndirect (src/include/pogs.h:55-131,155-158), which the reference never
exposed to Python.  The matrix is uploaded and set up once; z, z~ and rho persist
between solves, so a sequence of solves with changing f / g is warm-started like
the reference's examples/cpp/lasso_path.cpp:75-107.
"""
import ctypes

import numpy as np

from . import _lib
from .graph import FunctionVector, Ordering, HAS_SCIPY

if HAS_SCIPY:
    import scipy.sparse as sp


class Solver:
    """Solver(A, dtype=np.float32|np.float64, order='r'|'c', projector='direct'|'indirect').

    A: numpy array (host), scipy.sparse matrix (-> CGLS projector) or a CUDA torch
    tensor (dense, row-major, device-resident: no host round trip).
    projector='indirect' on a dense A selects the CGLS projector (the reference's
    PogsIndirect<T, MatrixDense<T>>); sparse matrices always use it.
    Method names follow the C++ class: Solve, SetRho, SetAbsTol, ..., GetX, GetY,
    GetLambda, GetMu, GetOptval, GetFinalIter, GetRho."""

    def __init__(self, A, dtype=None, order="r", projector="direct"):
        self._h = None
        if projector not in ("direct", "indirect"):
            raise ValueError("projector must be 'direct' or 'indirect'")
        dense_create = "pogs_b200_create_dense_indirect_" if projector == "indirect" else "pogs_b200_create_dense_"
        is_sparse = HAS_SCIPY and sp.issparse(A)
        is_torch = (not is_sparse) and hasattr(A, "is_cuda")
        if dtype is None:
            if is_torch:
                dtype = np.float64 if "float64" in str(A.dtype) else np.float32
            else:
                dtype = np.float32 if getattr(A, "dtype", None) == np.float32 else np.float64
        self.dtype = np.dtype(dtype)
        self._ct = _lib.ctype_of(self.dtype)
        self._sfx = _lib.suffix(self.dtype)
        self.m, self.n = A.shape
        rowmaj = order in ("r", "R")
        ordv = int(Ordering.ROW_MAJ if rowmaj else Ordering.COL_MAJ)
        if is_sparse:
            M = sp.csr_matrix(A, dtype=self.dtype) if rowmaj else sp.csc_matrix(A, dtype=self.dtype)
            data = np.ascontiguousarray(M.data, dtype=self.dtype)
            indptr = np.ascontiguousarray(M.indptr, dtype=np.int32)
            indices = np.ascontiguousarray(M.indices, dtype=np.int32)
            h = getattr(_lib.lib, "pogs_b200_create_sparse_" + self._sfx)(
                ordv, self.m, self.n, M.nnz, _lib.ptr(data, self._ct), _lib.ptr(indptr, ctypes.c_int),
                _lib.ptr(indices, ctypes.c_int))
        elif is_torch:
            import torch

            want = torch.float64 if self.dtype == np.float64 else torch.float32
            if not A.is_cuda:
                raise ValueError("torch input must be a CUDA tensor (use a numpy array for host data)")
            At = A.to(want).contiguous() if rowmaj else A.to(want).t().contiguous()
            torch.cuda.current_stream().synchronize()
            h = getattr(_lib.lib, dense_create + self._sfx)(
                ordv, self.m, self.n, ctypes.c_void_p(At.data_ptr()), 1)
        else:
            Ah = np.ascontiguousarray(A, dtype=self.dtype) if rowmaj else np.asfortranarray(A, dtype=self.dtype)
            h = getattr(_lib.lib, dense_create + self._sfx)(
                ordv, self.m, self.n, ctypes.c_void_p(Ah.ctypes.data), 0)
        if not h:
            raise RuntimeError("pogs_b200: " + _lib.last_error())
        self._h = ctypes.c_void_p(h)
        # defaults of the reference's Python wrappers (graph.py:236-247)
        self._p = dict(rho=1.0, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500, verbose=0, adaptive_rho=True,
                       gap_stop=True)
        self._rho_dirty = True
        self.status = None

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if self._h is not None:
            _lib.lib.pogs_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- setters (pogs.h:103-118) -------------------------------------------------------------
    def SetRho(self, rho):
        self._p["rho"] = float(rho)
        self._rho_dirty = True

    def SetAbsTol(self, v):
        self._p["abs_tol"] = float(v)

    def SetRelTol(self, v):
        self._p["rel_tol"] = float(v)

    def SetMaxIter(self, v):
        self._p["max_iter"] = int(v)

    def SetVerbose(self, v):
        self._p["verbose"] = int(v)

    def SetAdaptiveRho(self, v):
        self._p["adaptive_rho"] = bool(v)

    def SetGapStop(self, v):
        self._p["gap_stop"] = bool(v)

    def SetInitX(self, x):
        self._init_x = np.ascontiguousarray(x, dtype=self.dtype)

    def SetInitLambda(self, l):
        self._init_l = np.ascontiguousarray(l, dtype=self.dtype)

    def SetProfile(self, on):
        _lib.lib.pogs_b200_set_profile(self._h, int(bool(on)))

    _init_x = None
    _init_l = None

    # -- solve ----------------------------------------------------------------------------------
    def Solve(self, f, g):
        """== PogsSeparable::Solve(f, g).  Returns the status code (0 solved, 3 max-iter, 6 error)."""
        f = FunctionVector.from_any(f)
        g = FunctionVector.from_any(g)
        if len(f) != self.m or len(g) != self.n:
            raise ValueError(f"f/g must have lengths {self.m}/{self.n}")
        p = self._p
        rho = p["rho"] if self._rho_dirty else self.GetRho()   # rho persists across solves unless reset
        _lib.lib.pogs_b200_set_params(self._h, rho, p["abs_tol"], p["rel_tol"], p["max_iter"], p["verbose"],
                                      int(p["adaptive_rho"]), int(p["gap_stop"]))
        self._rho_dirty = False
        if self._init_x is not None or self._init_l is not None:
            ct = self._ct
            getattr(_lib.lib, "pogs_b200_set_init_" + self._sfx)(
                self._h, _lib.ptr(self._init_x, ct) if self._init_x is not None else None,
                _lib.ptr(self._init_l, ct) if self._init_l is not None else None)
            self._init_x = self._init_l = None
        fa, ga = f.arrays(self.dtype), g.arrays(self.dtype)
        ct = self._ct

        def ptrs(arrs):
            a, b, c, d, e, h = arrs
            return [_lib.ptr(v, ct) for v in (a, b, c, d, e)] + [_lib.ptr(h, ctypes.c_int)]

        self.status = int(getattr(_lib.lib, "pogs_b200_solve_" + self._sfx)(self._h, *ptrs(fa), *ptrs(ga)))
        if self.status == 6:
            raise RuntimeError("pogs_b200: " + _lib.last_error())
        return self.status

    # -- getters (pogs.h:87-101) -------------------------------------------------------------------
    def _get(self):
        ct = self._ct
        x = np.empty(self.n, self.dtype); mu = np.empty(self.n, self.dtype)
        y = np.empty(self.m, self.dtype); l = np.empty(self.m, self.dtype)
        optval = ct(); it = ctypes.c_uint(); rho = ct()
        getattr(_lib.lib, "pogs_b200_get_solution_" + self._sfx)(
            self._h, _lib.ptr(x, ct), _lib.ptr(y, ct), _lib.ptr(l, ct), _lib.ptr(mu, ct), ctypes.byref(optval),
            ctypes.byref(it), ctypes.byref(rho))
        return x, y, l, mu, float(optval.value), int(it.value), float(rho.value)

    def GetX(self):
        return self._get()[0]

    def GetY(self):
        return self._get()[1]

    def GetLambda(self):
        return self._get()[2]

    def GetMu(self):
        return self._get()[3]

    def GetOptval(self):
        return self._get()[4]

    def GetFinalIter(self):
        return self._get()[5]

    def GetRho(self):
        return self._get()[6]

    def result(self):
        """Same dict as the solve_* wrappers (+ 'mu', 'rho')."""
        x, y, l, mu, optval, it, rho = self._get()
        return {"x": x, "y": y, "l": l, "mu": mu, "optval": optval, "iterations": it, "status": self.status,
                "rho": rho}

    def timing(self):
        out = (ctypes.c_double * 16)()
        _lib.lib.pogs_b200_get_timing(self._h, out)
        keys = ["h2d_ms", "setup_ms", "loop_ms", "total_ms", "iterations", "exact_iterations", "prox_ms",
                "gemvt_ms", "solve_ms", "gemv_ms", "ctrl_ms", "profiled_iterations", "cgls_iterations", "equil_ms",
                "normest_ms", "gram_ms"]
        t = {k: float(out[i]) for i, k in enumerate(keys)}
        t["factor_ms"] = max(t["setup_ms"] - t["equil_ms"] - t["normest_ms"] - t["gram_ms"], 0.0)
        st = (ctypes.c_double * 8)()
        _lib.lib.pogs_b200_get_stats(self._h, st)
        t["single_pass_iterations"] = float(st[0])
        t["normest_iterations"] = float(st[1])
        t["rare_paths"] = float(st[3])
        t["one_launch"] = float(st[4])
        t["predicted_rho_hits"] = float(st[5])
        ph = (ctypes.c_double * 16)()
        _lib.lib.pogs_b200_get_pass_phases(self._h, ph)
        t["pass_phase_us"] = [float(v) for v in ph][:9]
        return t

    # -- test hooks -----------------------------------------------------------------------------------
    def equilibration(self):
        ct = self._ct
        d = np.empty(self.m, self.dtype); e = np.empty(self.n, self.dtype); nrm = ct()
        rc = getattr(_lib.lib, "pogs_b200_get_equil_" + self._sfx)(self._h, _lib.ptr(d, ct), _lib.ptr(e, ct),
                                                                   ctypes.byref(nrm))
        if rc:
            raise RuntimeError("pogs_b200: " + _lib.last_error())
        return d, e, float(nrm.value)

    def project(self, x0, y0):
        ct = self._ct
        x0 = np.ascontiguousarray(x0, self.dtype); y0 = np.ascontiguousarray(y0, self.dtype)
        x = np.empty(self.n, self.dtype); y = np.empty(self.m, self.dtype)
        rc = getattr(_lib.lib, "pogs_b200_project_" + self._sfx)(self._h, _lib.ptr(x0, ct), _lib.ptr(y0, ct),
                                                                 _lib.ptr(x, ct), _lib.ptr(y, ct))
        if rc:
            raise RuntimeError("pogs_b200: " + _lib.last_error())
        return x, y


def lasso_path(A, b, lambdas, dtype=np.float32, lambda2=0.0, **params):
    """Warm-started regularisation path (protocol of examples/cpp/lasso_path.cpp:75-107):
    one persistent solver, g.c rewritten per lambda, state carried over.
    Returns (list of result dicts, Solver timing of the last solve)."""
    from .graph import Function

    A_shape = A.shape
    m, n = A_shape
    out = []
    with Solver(A, dtype=dtype) as s:
        for k, v in params.items():
            getattr(s, "Set" + "".join(w.capitalize() for w in k.split("_")))(v)
        f = FunctionVector(m, Function.kSquare, 1.0, np.asarray(b, dtype=np.float64), 1.0)
        for lam in lambdas:
            g = FunctionVector(n, Function.kAbs, 1.0, 0.0, float(lam), 0.0, lambda2 / 2)
            s.Solve(f, g)
            r = s.result()
            r["timing"] = s.timing()
            out.append(r)
    return out

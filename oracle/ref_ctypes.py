"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the compiled reference
(`oracle/_ref/libpogs_ref*.so`, built by oracle/build_ref.sh from the unmodified
/root/reference sources).

Mirrors the calling convention of the reference's python/pogs/graph.py:167-233,
318-390 (PogsD / PogsSparseD, always ROW_MAJ) and adds the fp32 entry points
PogsS / PogsSparseS declared in src/interface_c/pogs_c.h:84-119.  Takes the
descriptor arrays in SoA form (f_h, f_a..f_e) so that callers do not have to
build Python object lists.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def ref_path(openmp=False):
    return os.path.join(_HERE, "_ref", "libpogs_ref_omp.so" if openmp else "libpogs_ref.so")


def available(openmp=False):
    return os.path.exists(ref_path(openmp))


def _lib(openmp=False):
    key = bool(openmp)
    if key not in _LIBS:
        _LIBS[key] = ctypes.CDLL(ref_path(openmp))
    return _LIBS[key]


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _desc(d, size, dt):
    """(h,a,b,c,d,e) dict/tuple of scalars-or-arrays -> SoA arrays."""
    h, a, b, c, dd, e = d
    out = [np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=dt), (size,))) for v in (a, b, c, dd, e)]
    hh = np.ascontiguousarray(np.broadcast_to(np.asarray(h, dtype=np.int32), (size,)))
    return hh, out


def solve(A, f, g, *, dtype=np.float64, rho=1.0, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500,
          verbose=0, adaptive_rho=True, gap_stop=True, openmp=False):
    """f, g = (h, a, b, c, d, e) with scalar or per-element entries.
    A: dense ndarray (row-major used) or scipy.sparse matrix (CSR used)."""
    lib = _lib(openmp)
    dt = np.dtype(dtype)
    ct = ctypes.c_double if dt == np.float64 else ctypes.c_float
    sparse = hasattr(A, "tocsr")
    m, n = A.shape
    fh, (fa, fb, fc, fd, fe) = _desc(f, m, dt)
    gh, (ga, gb, gc, gd, ge) = _desc(g, n, dt)
    x = np.zeros(n, dt); y = np.zeros(m, dt); l = np.zeros(m, dt)
    optval = ct(); final_iter = ctypes.c_uint()
    tail = [_p(fa, ct), _p(fb, ct), _p(fc, ct), _p(fd, ct), _p(fe, ct), _p(fh, ctypes.c_int),
            _p(ga, ct), _p(gb, ct), _p(gc, ct), _p(gd, ct), _p(ge, ct), _p(gh, ctypes.c_int),
            ct(rho), ct(abs_tol), ct(rel_tol), ctypes.c_uint(max_iter), ctypes.c_uint(verbose),
            ctypes.c_int(int(adaptive_rho)), ctypes.c_int(int(gap_stop)),
            _p(x, ct), _p(y, ct), _p(l, ct), ctypes.byref(optval), ctypes.byref(final_iter)]
    if sparse:
        csr = A.tocsr()
        data = np.ascontiguousarray(csr.data, dtype=dt)
        ptr = np.ascontiguousarray(csr.indptr, dtype=np.int32)
        ind = np.ascontiguousarray(csr.indices, dtype=np.int32)
        fn = lib.PogsSparseD if dt == np.float64 else lib.PogsSparseS
        fn.restype = ctypes.c_int
        status = fn(ctypes.c_int(1), ctypes.c_size_t(m), ctypes.c_size_t(n), ctypes.c_size_t(csr.nnz),
                    _p(data, ct), _p(ptr, ctypes.c_int), _p(ind, ctypes.c_int), *tail)
    else:
        Ad = np.ascontiguousarray(A, dtype=dt)
        fn = lib.PogsD if dt == np.float64 else lib.PogsS
        fn.restype = ctypes.c_int
        status = fn(ctypes.c_int(1), ctypes.c_size_t(m), ctypes.c_size_t(n), _p(Ad, ct), *tail)
    return {"x": x, "y": y, "l": l, "optval": float(optval.value),
            "iterations": int(final_iter.value), "status": int(status)}


def persistent_available(openmp=False):
    if not available(openmp):
        return False
    return hasattr(_lib(openmp), "refp_create_dense_d")


class PersistentDense:
    """The reference's own pogs::PogsDirect<T, MatrixDense<T>> object kept alive across
    solves (oracle/ref_persistent.cpp): lazy init, cached factor, implicit warm start."""

    def __init__(self, A, dtype=np.float64, openmp=False):
        self.lib = _lib(openmp)
        self.dt = np.dtype(dtype)
        self.ct = ctypes.c_double if self.dt == np.float64 else ctypes.c_float
        self.sfx = "d" if self.dt == np.float64 else "s"
        self.A = np.ascontiguousarray(A, dtype=self.dt)   # must outlive the first solve
        self.m, self.n = self.A.shape
        fn = getattr(self.lib, "refp_create_dense_" + self.sfx)
        fn.restype = ctypes.c_void_p
        self.h = ctypes.c_void_p(fn(ctypes.c_int(1), ctypes.c_size_t(self.m), ctypes.c_size_t(self.n),
                                    _p(self.A, self.ct)))
        self.first = True

    def close(self):
        if self.h is not None:
            getattr(self.lib, "refp_destroy_" + self.sfx)(self.h)
            self.h = None

    def set_init(self, x=None, lam=None):
        xa = np.ascontiguousarray(x, self.dt) if x is not None else None
        la = np.ascontiguousarray(lam, self.dt) if lam is not None else None
        getattr(self.lib, "refp_set_init_" + self.sfx)(self.h, _p(xa, self.ct) if xa is not None else None,
                                                       _p(la, self.ct) if la is not None else None)

    def solve(self, f, g, rho=None, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500, adaptive_rho=True, gap_stop=True):
        ct = self.ct
        fh, (fa, fb, fc, fd, fe) = _desc(f, self.m, self.dt)
        gh, (ga, gb, gc, gd, ge) = _desc(g, self.n, self.dt)
        x = np.zeros(self.n, self.dt); mu = np.zeros(self.n, self.dt)
        y = np.zeros(self.m, self.dt); l = np.zeros(self.m, self.dt)
        optval = ct(); it = ctypes.c_uint(); rho_o = ct()
        set_rho = 1 if (rho is not None or self.first) else 0
        rho_v = 1.0 if rho is None else rho
        fn = getattr(self.lib, "refp_solve_" + self.sfx)
        fn.restype = ctypes.c_int
        st = fn(self.h, _p(fh, ctypes.c_int), _p(fa, ct), _p(fb, ct), _p(fc, ct), _p(fd, ct), _p(fe, ct),
                _p(gh, ctypes.c_int), _p(ga, ct), _p(gb, ct), _p(gc, ct), _p(gd, ct), _p(ge, ct),
                ct(rho_v), ctypes.c_int(set_rho), ct(abs_tol), ct(rel_tol), ctypes.c_uint(max_iter),
                ctypes.c_int(int(adaptive_rho)), ctypes.c_int(int(gap_stop)),
                _p(x, ct), _p(y, ct), _p(l, ct), _p(mu, ct), ctypes.byref(optval), ctypes.byref(it),
                ctypes.byref(rho_o))
        self.first = False
        return {"x": x, "y": y, "l": l, "mu": mu, "optval": float(optval.value), "iterations": int(it.value),
                "status": int(st), "rho": float(rho_o.value)}


class _ConeC(ctypes.Structure):
    _fields_ = [("cone", ctypes.c_int), ("indices", ctypes.POINTER(ctypes.c_uint)), ("size", ctypes.c_uint)]


def cone_solve(A, b, c, cones_x, cones_y, *, direct=True, rho=1.0, abs_tol=1e-4, rel_tol=1e-4, max_iter=10000,
               adaptive_rho=True, gap_stop=False):
    """The compiled reference's PogsConeDirectD / PogsConeD (src/interface_c/pogs_c.h:167-186), fp64, ROW_MAJ.
    cones_*: list of (cone code, indices)."""
    lib = _lib(False)
    A = np.ascontiguousarray(A, np.float64); b = np.ascontiguousarray(b, np.float64); c = np.ascontiguousarray(c, np.float64)
    m, n = A.shape
    keep = []

    def make(cones):
        out = []
        for code, idx in cones or []:
            arr = (ctypes.c_uint * len(idx))(*[int(i) for i in idx]); keep.append(arr)
            out.append(_ConeC(int(code), ctypes.cast(arr, ctypes.POINTER(ctypes.c_uint)), len(idx)))
        return ((_ConeC * len(out))(*out) if out else None), len(out)

    kx, nkx = make(cones_x); ky, nky = make(cones_y)
    x = np.zeros(n); y = np.zeros(m); l = np.zeros(m)
    optval = ctypes.c_double(); it = ctypes.c_uint()
    fn = lib.PogsConeDirectD if direct else lib.PogsConeD
    fn.restype = ctypes.c_int
    cd = ctypes.c_double
    st = fn(ctypes.c_int(1), ctypes.c_size_t(m), ctypes.c_size_t(n), _p(A, cd), _p(b, cd), _p(c, cd), kx,
            ctypes.c_size_t(nkx), ky, ctypes.c_size_t(nky), cd(rho), cd(abs_tol), cd(rel_tol), ctypes.c_uint(max_iter),
            ctypes.c_uint(0), ctypes.c_int(int(adaptive_rho)), ctypes.c_int(int(gap_stop)), _p(x, cd), _p(y, cd), _p(l, cd),
            ctypes.byref(optval), ctypes.byref(it))
    return {"x": x, "y": y, "l": l, "optval": float(optval.value), "iterations": int(it.value), "status": int(st)}

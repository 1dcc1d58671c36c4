"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/liboracle.so, the plain-C
restatement of the reference's graph-form path (oracle/pogs_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.  The product (pogs_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib_path():
    return os.path.join(_HERE, "liboracle.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(lib_path()):
            build()
        _LIB = ctypes.CDLL(lib_path())
        _LIB.oracle_lambert_w_exp.restype = ctypes.c_double
        _LIB.oracle_lambert_w_exp.argtypes = [ctypes.c_double]
        for sfx, ct in (("s", ctypes.c_float), ("d", ctypes.c_double)):
            getattr(_LIB, f"oracle_prox_base_{sfx}").restype = ct
            getattr(_LIB, f"oracle_prox_base_{sfx}").argtypes = [ctypes.c_int, ct, ct]
            getattr(_LIB, f"oracle_func_vec_{sfx}").restype = ct
            getattr(_LIB, f"oracle_create_dense_{sfx}").restype = ctypes.c_void_p
            getattr(_LIB, f"oracle_create_sparse_{sfx}").restype = ctypes.c_void_p
            getattr(_LIB, f"oracle_solve_{sfx}").restype = ctypes.c_int
    return _LIB


def _ct(dt):
    return ctypes.c_double if np.dtype(dt) == np.float64 else ctypes.c_float


def _sfx(dt):
    return "d" if np.dtype(dt) == np.float64 else "s"


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _desc(d, size, dt):
    h, a, b, c, dd, e = d
    out = [np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=dt), (size,))) for v in (a, b, c, dd, e)]
    hh = np.ascontiguousarray(np.broadcast_to(np.asarray(h, dtype=np.int32), (size,)))
    return hh, out


def prox_base(h, v, rho, dtype=np.float64):
    L = lib(); ct = _ct(dtype)
    return float(getattr(L, f"oracle_prox_base_{_sfx(dtype)}")(int(h), ct(v), ct(rho)))


def prox_vec(desc, rho, v, dtype=np.float64):
    """ProxEval over a vector.  desc = (h, a, b, c, d, e)."""
    L = lib(); dt = np.dtype(dtype); ct = _ct(dt)
    v = np.ascontiguousarray(v, dtype=dt); n = v.size
    h, (a, b, c, d, e) = _desc(desc, n, dt)
    out = np.empty(n, dt)
    getattr(L, f"oracle_prox_vec_{_sfx(dt)}")(ctypes.c_size_t(n), _p(h, ctypes.c_int), _p(a, ct), _p(b, ct),
                                              _p(c, ct), _p(d, ct), _p(e, ct), ct(rho), _p(v, ct), _p(out, ct))
    return out


def func_vec(desc, v, dtype=np.float64):
    L = lib(); dt = np.dtype(dtype); ct = _ct(dt)
    v = np.ascontiguousarray(v, dtype=dt); n = v.size
    h, (a, b, c, d, e) = _desc(desc, n, dt)
    return float(getattr(L, f"oracle_func_vec_{_sfx(dt)}")(ctypes.c_size_t(n), _p(h, ctypes.c_int), _p(a, ct),
                                                          _p(b, ct), _p(c, ct), _p(d, ct), _p(e, ct), _p(v, ct)))


def rand(n, dtype=np.float64):
    L = lib(); dt = np.dtype(dtype); ct = _ct(dt)
    x = np.empty(n, dt)
    getattr(L, f"oracle_rand_{_sfx(dt)}")(_p(x, ct), ctypes.c_size_t(n))
    return x


class Solver:
    """Persistent oracle solver == reference PogsDirect / PogsIndirect object
    (src/include/pogs.h:122-131,155-158): lazy init on first solve, cached
    equilibration / factor, persistent (z, zt, rho)."""

    def __init__(self, A, dtype=np.float64, order="r", direct=None):
        self.L = lib(); self.dt = np.dtype(dtype); self.ct = _ct(self.dt); self.sfx = _sfx(self.dt)
        self.m, self.n = A.shape
        rowmaj = 1 if order in ("r", "R") else 0
        if hasattr(A, "tocsr"):
            M = A.tocsr() if rowmaj else A.tocsc()
            data = np.ascontiguousarray(M.data, dtype=self.dt)
            ptr = np.ascontiguousarray(M.indptr, dtype=np.int32)
            ind = np.ascontiguousarray(M.indices, dtype=np.int32)
            self.h = getattr(self.L, f"oracle_create_sparse_{self.sfx}")(
                ctypes.c_int(rowmaj), ctypes.c_size_t(self.m), ctypes.c_size_t(self.n), ctypes.c_size_t(M.nnz),
                _p(data, self.ct), _p(ptr, ctypes.c_int), _p(ind, ctypes.c_int))
            self.sparse = True
        else:
            Ad = np.ascontiguousarray(A, dtype=self.dt) if rowmaj else np.asfortranarray(A, dtype=self.dt)
            if direct is None:
                direct = True
            self.h = getattr(self.L, f"oracle_create_dense_{self.sfx}")(
                ctypes.c_int(rowmaj), ctypes.c_size_t(self.m), ctypes.c_size_t(self.n),
                Ad.ctypes.data_as(ctypes.POINTER(self.ct)), ctypes.c_int(int(direct)))
            self.sparse = False
        self.h = ctypes.c_void_p(self.h)

    def close(self):
        if self.h is not None:
            getattr(self.L, f"oracle_destroy_{self.sfx}")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setup(self, want_A=False):
        ct = self.ct
        d = np.empty(self.m, self.dt); e = np.empty(self.n, self.dt); nrm = ct()
        Aeq = np.empty((self.m, self.n), self.dt) if (want_A and not self.sparse) else None
        getattr(self.L, f"oracle_setup_{self.sfx}")(self.h, _p(d, ct), _p(e, ct), ctypes.byref(nrm),
                                                   _p(Aeq, ct) if Aeq is not None else None)
        return d, e, float(nrm.value), Aeq

    def project(self, x0, y0, xinit=None, tol=1e-8):
        ct = self.ct
        x0 = np.ascontiguousarray(x0, self.dt); y0 = np.ascontiguousarray(y0, self.dt)
        x = np.zeros(self.n, self.dt) if xinit is None else np.array(xinit, self.dt)
        y = np.zeros(self.m, self.dt)
        getattr(self.L, f"oracle_project_{self.sfx}")(self.h, _p(x0, ct), _p(y0, ct), _p(x, ct), _p(y, ct), ct(tol))
        return x, y

    def solve(self, f, g, rho=None, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500, adaptive_rho=True,
              gap_stop=True, init_x=None, init_lambda=None):
        ct = self.ct
        fh, (fa, fb, fc, fd, fe) = _desc(f, self.m, self.dt)
        gh, (ga, gb, gc, gd, ge) = _desc(g, self.n, self.dt)
        if rho is None:
            r = ct(); getattr(self.L, f"oracle_get_{self.sfx}")(self.h, None, None, None, None, None, None,
                                                                 ctypes.byref(r), None)
            rho = r.value
        getattr(self.L, f"oracle_set_params_{self.sfx}")(self.h, ct(rho), ct(abs_tol), ct(rel_tol),
                                                        ctypes.c_uint(max_iter), ctypes.c_int(int(adaptive_rho)),
                                                        ctypes.c_int(int(gap_stop)))
        if init_x is not None or init_lambda is not None:
            ix = np.ascontiguousarray(init_x, self.dt) if init_x is not None else None
            il = np.ascontiguousarray(init_lambda, self.dt) if init_lambda is not None else None
            getattr(self.L, f"oracle_set_init_{self.sfx}")(self.h, _p(ix, ct) if ix is not None else None,
                                                          _p(il, ct) if il is not None else None)
        status = getattr(self.L, f"oracle_solve_{self.sfx}")(
            self.h, _p(fh, ctypes.c_int), _p(fa, ct), _p(fb, ct), _p(fc, ct), _p(fd, ct), _p(fe, ct),
            _p(gh, ctypes.c_int), _p(ga, ct), _p(gb, ct), _p(gc, ct), _p(gd, ct), _p(ge, ct))
        x = np.empty(self.n, self.dt); mu = np.empty(self.n, self.dt)
        y = np.empty(self.m, self.dt); l = np.empty(self.m, self.dt)
        optval = ct(); it = ctypes.c_uint(); rho_o = ct(); inner = ctypes.c_long()
        getattr(self.L, f"oracle_get_{self.sfx}")(self.h, _p(x, ct), _p(y, ct), _p(l, ct), _p(mu, ct),
                                                 ctypes.byref(optval), ctypes.byref(it), ctypes.byref(rho_o),
                                                 ctypes.byref(inner))
        return {"x": x, "y": y, "l": l, "mu": mu, "optval": float(optval.value), "iterations": int(it.value),
                "status": int(status), "rho": float(rho_o.value), "cgls_iters": int(inner.value)}


def solve(A, f, g, *, dtype=np.float64, order="r", direct=None, **kw):
    """One-shot solve == reference C ABI PogsD/PogsS/PogsSparseD/PogsSparseS
    (src/interface_c/pogs_c.cpp:8-108): dense -> direct, sparse -> CGLS."""
    kw.setdefault("rho", 1.0)
    s = Solver(A, dtype=dtype, order=order, direct=direct)
    try:
        return s.solve(f, g, **kw)
    finally:
        s.close()

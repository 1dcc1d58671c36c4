"""Cone-form front end, first slice: mirror of the reference's python/pogs_cone.py (solve_cone, Cone,
same arguments and result keys) on top of the PogsCone* / PogsConeDirect* entry points of libpogs_b200.so.

    minimize c^T x   subject to   b - A x in K_y,  x in K_x

Implemented cones: ZERO, NON_NEG, NON_POS (linear programs) -- for these the reference's cone objective
(src/cpu/pogs.cpp:642-790) is a separable graph-form objective and runs on the device path unchanged.
SOC / SDP / exponential cones and a quadratic objective (P) are not implemented: status 6 and a message.
"""
import ctypes
from enum import IntEnum

import numpy as np

from . import _lib
from .graph import Ordering


class Cone(IntEnum):
    ZERO = 0
    NON_NEG = 1
    NON_POS = 2
    SOC = 3
    SDP = 4
    EXP_PRIMAL = 5
    EXP_DUAL = 6


class ConeConstraintC(ctypes.Structure):
    _fields_ = [("cone", ctypes.c_int), ("indices", ctypes.POINTER(ctypes.c_uint)), ("size", ctypes.c_uint)]


def _sig():
    P = ctypes.POINTER
    for sfx, ct in (("D", ctypes.c_double), ("S", ctypes.c_float)):
        for name in ("PogsCone", "PogsConeDirect"):
            fn = getattr(_lib.lib, name + sfx)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, P(ct), P(ct), P(ct), P(ConeConstraintC),
                           ctypes.c_size_t, P(ConeConstraintC), ctypes.c_size_t, ct, ct, ct, ctypes.c_uint, ctypes.c_uint,
                           ctypes.c_int, ctypes.c_int, P(ct), P(ct), P(ct), P(ct), P(ctypes.c_uint)]


_sig()


def _make(cones):
    keep, out = [], []
    for cone_type, indices in cones or []:
        idx = (ctypes.c_uint * len(indices))(*[int(i) for i in indices])
        keep.append(idx)
        out.append(ConeConstraintC(cone=int(cone_type), indices=ctypes.cast(idx, ctypes.POINTER(ctypes.c_uint)),
                                   size=len(indices)))
    arr = (ConeConstraintC * len(out))(*out) if out else None
    return arr, len(out), keep


def solve_cone(A, b, c, cones_x, cones_y, rho=1.0, abs_tol=1e-4, rel_tol=1e-3, max_iter=10000, verbose=0,
               adaptive_rho=True, gap_stop=True, use_direct=False, P=None, dtype=np.float64):
    """Same call as the reference's pogs_cone.solve_cone (python/pogs_cone.py:183-358); dtype selects PogsCone*D / *S."""
    if P is not None:
        raise NotImplementedError("cone form with a quadratic objective is not implemented in pogs_b200")
    dt = np.dtype(dtype)
    ct = _lib.ctype_of(dt)
    A = np.ascontiguousarray(A, dtype=dt)
    b = np.ascontiguousarray(b, dtype=dt)
    c = np.ascontiguousarray(c, dtype=dt)
    m, n = A.shape
    if b.shape != (m,) or c.shape != (n,):
        raise ValueError(f"b and c must have shapes ({m},) and ({n},)")
    kx, nkx, _kx = _make(cones_x)
    ky, nky, _ky = _make(cones_y)
    x = np.zeros(n, dt); y = np.zeros(m, dt); l = np.zeros(m, dt)
    optval = ct(); it = ctypes.c_uint()
    fn = getattr(_lib.lib, ("PogsConeDirect" if use_direct else "PogsCone") + _lib.suffix(dt, upper=True))
    status = fn(int(Ordering.ROW_MAJ), m, n, _lib.ptr(A, ct), _lib.ptr(b, ct), _lib.ptr(c, ct), kx, nkx, ky, nky,
                ct(rho), ct(abs_tol), ct(rel_tol), int(max_iter), int(verbose), int(adaptive_rho), int(gap_stop),
                _lib.ptr(x, ct), _lib.ptr(y, ct), _lib.ptr(l, ct), ctypes.byref(optval), ctypes.byref(it))
    return {"x": x, "y": y, "l": l, "optval": float(optval.value), "iterations": int(it.value), "status": int(status)}

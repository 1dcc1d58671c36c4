"""CPU suite, part 3: the device prox library (pogs_b200/csrc/prox.cuh) is plain
host+device C++; compile it with g++ and check all sixteen functions, with random
(a,b,c,d,e,rho), against the oracle.  This checks the *source the kernels are built
from* without a GPU; the on-device run of the same code is checked in test_gpu_units.py."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#include "prox.cuh"
extern "C" {
void hp_prox_d(size_t n, const int* h, const double* a, const double* b, const double* c, const double* d,
               const double* e, double rho, const double* in, double* out) {
  for (size_t i = 0; i < n; ++i) out[i] = pogs_b200::prox_eval<double>(h[i], a[i], b[i], c[i], d[i], e[i], in[i], rho);
}
void hp_prox_s(size_t n, const int* h, const float* a, const float* b, const float* c, const float* d,
               const float* e, float rho, const float* in, float* out) {
  for (size_t i = 0; i < n; ++i) out[i] = pogs_b200::prox_eval<float>(h[i], a[i], b[i], c[i], d[i], e[i], in[i], rho);
}
void hp_func_d(size_t n, const int* h, const double* a, const double* b, const double* c, const double* d,
               const double* e, const double* in, double* out) {
  for (size_t i = 0; i < n; ++i) out[i] = pogs_b200::func_eval<double>(h[i], a[i], b[i], c[i], d[i], e[i], in[i]);
}
}
"""


@pytest.fixture(scope="module")
def harness():
    td = tempfile.mkdtemp()
    src = os.path.join(td, "h.cpp")
    open(src, "w").write(HARNESS)
    so = os.path.join(td, "h.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I",
                           os.path.join(ROOT, "pogs_b200", "csrc"), "-o", so, src])
    return ctypes.CDLL(so)


def random_case(rng, n, h):
    a = rng.uniform(0.5, 2.0, n) * rng.choice([-1.0, 1.0], n)
    b = rng.standard_normal(n)
    c = rng.uniform(0.1, 3.0, n)
    d = rng.standard_normal(n) * 0.5
    e = rng.uniform(0.0, 1.0, n)
    v = rng.standard_normal(n) * 3
    return np.full(n, h, np.int32), a, b, c, d, e, v


@pytest.mark.parametrize("h", range(16))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_prox_header_matches_oracle(harness, oracle, h, dtype):
    rng = np.random.default_rng(100 + h)
    n = 4000
    hh, a, b, c, d, e, v = random_case(rng, n, h)
    ct = ctypes.c_double if dtype == np.float64 else ctypes.c_float
    P = lambda arr, t=ct: arr.ctypes.data_as(ctypes.POINTER(t))
    arrs = [np.ascontiguousarray(x, dtype) for x in (a, b, c, d, e, v)]
    for rho in (0.3, 1.0, 7.0):
        out = np.empty(n, dtype)
        fn = harness.hp_prox_d if dtype == np.float64 else harness.hp_prox_s
        fn(ctypes.c_size_t(n), P(hh, ctypes.c_int), *[P(x) for x in arrs[:5]], ct(rho), P(arrs[5]), P(out))
        want = oracle.prox_vec((hh, *arrs[:5]), rho, arrs[5], dtype)
        ok = np.isfinite(want)
        # same formulas, same libm: bit-for-bit up to the compiler's evaluation order
        tol = 1e-12 if dtype == np.float64 else 2e-5
        assert np.allclose(out[ok], want[ok], rtol=tol, atol=tol), (h, rho, np.abs(out[ok] - want[ok]).max())
        assert (np.isfinite(out) == ok).all()


@pytest.mark.parametrize("h", range(16))
def test_func_header_matches_oracle(harness, oracle, h):
    rng = np.random.default_rng(200 + h)
    n = 1000
    hh, a, b, c, d, e, v = random_case(rng, n, h)
    out = np.empty(n)
    P = lambda arr, t=ctypes.c_double: arr.ctypes.data_as(ctypes.POINTER(t))
    harness.hp_func_d(ctypes.c_size_t(n), P(hh, ctypes.c_int), P(a), P(b), P(c), P(d), P(e), P(v), P(out))
    want = np.array([oracle.func_vec((hh[i:i + 1], a[i:i + 1], b[i:i + 1], c[i:i + 1], d[i:i + 1], e[i:i + 1]), v[i:i + 1])
                     for i in range(n)])
    ok = np.isfinite(want)
    assert np.allclose(out[ok], want[ok], rtol=1e-12, atol=1e-12)

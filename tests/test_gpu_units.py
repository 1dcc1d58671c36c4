"""GPU suite, unit level: each device building block against the oracle, called
through the C ABI (include/pogs_b200.h part 3)."""
import ctypes

import numpy as np
import pytest

import problems
from conftest import relerr
from test_prox_header import random_case

pytestmark = pytest.mark.gpu


def _lib():
    from pogs_b200 import _lib

    return _lib


def dev_prox(desc, rho, v, dtype):
    L = _lib(); ct = L.ctype_of(dtype); sfx = L.suffix(dtype)
    h, a, b, c, d, e = desc
    arrs = [np.ascontiguousarray(x, dtype) for x in (a, b, c, d, e, v)]
    h = np.ascontiguousarray(h, np.int32)
    out = np.empty(len(v), dtype)
    rc = getattr(L.lib, "pogs_b200_prox_eval_" + sfx)(len(v), L.ptr(h, ctypes.c_int), *[L.ptr(x, ct) for x in arrs[:5]],
                                                     ct(rho), L.ptr(arrs[5], ct), L.ptr(out, ct))
    assert rc == 0, L.last_error()
    return out


def dev_func(desc, v, dtype):
    L = _lib(); ct = L.ctype_of(dtype); sfx = L.suffix(dtype)
    h, a, b, c, d, e = desc
    arrs = [np.ascontiguousarray(x, dtype) for x in (a, b, c, d, e, v)]
    h = np.ascontiguousarray(h, np.int32)
    s = ctypes.c_double()
    rc = getattr(L.lib, "pogs_b200_func_eval_" + sfx)(len(v), L.ptr(h, ctypes.c_int), *[L.ptr(x, ct) for x in arrs[:5]],
                                                     L.ptr(arrs[5], ct), ctypes.byref(s))
    assert rc == 0, L.last_error()
    return s.value


def dev_gemv(A, trans, square, v, dtype, order="r"):
    L = _lib(); ct = L.ctype_of(dtype); sfx = L.suffix(dtype)
    m, n = A.shape
    Ad = np.ascontiguousarray(A, dtype) if order == "r" else np.asfortranarray(A, dtype)
    v = np.ascontiguousarray(v, dtype)
    out = np.empty(n if trans else m, dtype)
    rc = getattr(L.lib, "pogs_b200_gemv_" + sfx)(1 if order == "r" else 0, m, n,
                                                Ad.ctypes.data_as(ctypes.POINTER(ct)), int(trans), int(square),
                                                L.ptr(v, ct), L.ptr(out, ct))
    assert rc == 0, L.last_error()
    return out


# ---- prox / objective: all 16 functions, random (a,b,c,d,e,rho) -------------------------------
@pytest.mark.parametrize("h", range(16))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_device_prox_matches_oracle(oracle, h, dtype):
    rng = np.random.default_rng(300 + h)
    n = 5000
    hh, a, b, c, d, e, v = random_case(rng, n, h)
    for rho in (0.3, 1.0, 7.0):
        got = dev_prox((hh, a, b, c, d, e), rho, v, dtype)
        want = oracle.prox_vec((hh, a, b, c, d, e), rho, np.ascontiguousarray(v, dtype), dtype)
        ok = np.isfinite(want)
        # closed forms: few ulp (device libm vs glibc, FMA contraction);
        # iterative ones (Logistic/Exp/NegEntr/Recipr): to the bisection / Halley tolerance
        iterative = h in (problems.LOGISTIC, problems.EXP, problems.NEGENTR, problems.RECIPR)
        if dtype == np.float64:
            tol = 1e-9 if iterative else 1e-12
        else:
            tol = 5e-5 if iterative else 5e-6
        scale = np.maximum(np.abs(want[ok]), 1.0)
        assert (np.abs(got[ok] - want[ok]) / scale).max() < tol, (h, rho)
        assert (np.isfinite(got) == ok).all()


def test_device_prox_known_answers(oracle):
    """The reference's own golden values (tests/test_proximal.cpp) on the device."""
    from test_oracle import KNOWN

    for dtype in (np.float64, np.float32):
        h = np.array([k[0] for k in KNOWN], np.int32)
        v = np.array([k[1] for k in KNOWN])
        want = np.array([k[3] for k in KNOWN])
        for rho in sorted(set(k[2] for k in KNOWN)):
            sel = np.array([k[2] == rho for k in KNOWN])
            n = int(sel.sum())
            got = dev_prox((h[sel], np.ones(n), np.zeros(n), np.ones(n), np.zeros(n), np.zeros(n)), rho, v[sel], dtype)
            assert np.allclose(got, want[sel], rtol=1e-6, atol=1e-7)


def test_device_prox_mixed_tags_and_empty(oracle):
    rng = np.random.default_rng(7)
    n = 3001
    _, a, b, c, d, e, v = random_case(rng, n, 0)
    hh = rng.integers(0, 16, n).astype(np.int32)
    got = dev_prox((hh, a, b, c, d, e), 1.3, v, np.float64)
    want = oracle.prox_vec((hh, a, b, c, d, e), 1.3, v, np.float64)
    ok = np.isfinite(want)
    assert np.allclose(got[ok], want[ok], rtol=1e-9, atol=1e-9)
    assert dev_prox((hh[:0], a[:0], b[:0], c[:0], d[:0], e[:0]), 1.0, v[:0], np.float64).size == 0


@pytest.mark.parametrize("h", range(16))
def test_device_objective_matches_oracle(oracle, h):
    rng = np.random.default_rng(400 + h)
    n = 3000
    hh, a, b, c, d, e, v = random_case(rng, n, h)
    if h in (problems.NEGLOG, problems.RECIPR, problems.NEGENTR):
        v = np.abs(v) + 0.5; a = np.abs(a); b = -np.abs(b)   # keep the argument in the domain
    got = dev_func((hh, a, b, c, d, e), v, np.float64)
    want = oracle.func_vec((hh, a, b, c, d, e), v, np.float64)
    assert got == pytest.approx(want, rel=1e-10, abs=1e-9)


# ---- streaming products ------------------------------------------------------------------------
SHAPES = [(1, 1), (3, 5), (64, 64), (257, 129), (1000, 300), (300, 1000), (4099, 513), (37, 2050), (5000, 8)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("order", ["r", "c"])
def test_device_gemv(shape, dtype, order):
    """k_rowdot / k_colacc on ragged shapes (non-multiples of the vector width, fewer
    rows than warps, single element) in both storage orders, plain and squared."""
    m, n = shape
    rng = np.random.default_rng(m * 7919 + n)
    A = rng.standard_normal((m, n))
    x = rng.standard_normal(n); w = rng.standard_normal(m)
    A64 = np.asarray(np.ascontiguousarray(A, dtype), np.float64)
    tol = 1e-12 if dtype == np.float64 else 2e-5
    for trans, vec in ((0, x), (1, w)):
        v64 = np.asarray(np.ascontiguousarray(vec, dtype), np.float64)
        for sq in (0, 1):
            M = A64 * A64 if sq else A64
            want = (M.T @ v64) if trans else (M @ v64)
            got = dev_gemv(A, trans, sq, vec, dtype, order)
            scale = (np.abs(M).T @ np.abs(v64)) if trans else (np.abs(M) @ np.abs(v64))
            assert (np.abs(got - want) / np.maximum(scale, 1e-30)).max() < tol, (shape, trans, sq)


def test_device_gemv_is_deterministic():
    rng = np.random.default_rng(5)
    A = rng.standard_normal((3000, 700)); w = rng.standard_normal(3000)
    a = dev_gemv(A, 1, 0, w, np.float32)
    for _ in range(3):
        assert np.array_equal(a, dev_gemv(A, 1, 0, w, np.float32))


# ---- setup: equilibration + norm estimate + projection ------------------------------------------------
@pytest.mark.parametrize("case", ["c1_lasso_500x300", "lasso_wide_200x400", "lasso_odd_503x301"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("order", ["r", "c"])
def test_device_setup_matches_oracle(oracle, case, dtype, order):
    import pogs_b200

    p = problems.build(case)
    so = oracle.Solver(p["A"], dtype=dtype, order=order)
    d0, e0, nrm0, _ = so.setup()
    with pogs_b200.Solver(p["A"], dtype=dtype, order=order) as s:
        d1, e1, nrm1 = s.equilibration()
        tol = 1e-10 if dtype == np.float64 else 2e-5
        assert relerr(d1, d0) < tol and relerr(e1, e0) < tol
        assert nrm1 == pytest.approx(nrm0, rel=10 * tol)
        rng = np.random.default_rng(1)
        x0 = rng.standard_normal(p["A"].shape[1]); y0 = rng.standard_normal(p["A"].shape[0])
        xo, yo = so.project(x0, y0)
        xd, yd = s.project(x0, y0)
        ptol = 1e-9 if dtype == np.float64 else 5e-5
        assert relerr(xd, xo) < ptol and relerr(yd, yo) < ptol
    so.close()


def test_projection_properties():
    """Size-independent checks of the projector at a larger shape: the output lies on
    the graph (y = A^ x), the residual is orthogonal to the graph, idempotence."""
    import pogs_b200

    rng = np.random.default_rng(2)
    m, n = 6000, 900
    A = rng.standard_normal((m, n)).astype(np.float32)
    with pogs_b200.Solver(A, dtype=np.float32) as s:
        d, e, _ = s.equilibration()
        x0 = rng.standard_normal(n).astype(np.float32); y0 = rng.standard_normal(m).astype(np.float32)
        x, y = s.project(x0, y0)
        Ah = (d[:, None].astype(np.float64) * A.astype(np.float64)) * e[None, :].astype(np.float64)
        Ah /= np.linalg.norm(Ah) / np.sqrt(min(m, n)) / 1.0 if False else 1.0
        # recover A^ exactly as the device built it: A^ = D A E (d, e already carry 1/sqrt(normA) each)
        assert relerr(y, Ah @ x.astype(np.float64)) < 5e-5
        # optimality: (x - x0) + A^T (y - y0) = 0
        r = (x.astype(np.float64) - x0) + Ah.T @ (y.astype(np.float64) - y0)
        assert np.linalg.norm(r) < 5e-4 * (np.linalg.norm(x0) + np.linalg.norm(y0))
        x2, y2 = s.project(x, y)
        assert relerr(x2, x) < 5e-5 and relerr(y2, y) < 5e-5


# ---- one-time Gram matrix on the tensor cores (gram_tc.cuh) ------------------------------------------
def dev_gram(A, use_tc=1):
    L = _lib()
    A = np.ascontiguousarray(A, np.float32)
    m, n = A.shape
    G = np.zeros((n, n), np.float32)
    rc = L.lib.pogs_b200_gram_s(m, n, L.ptr(A, ctypes.c_float), L.ptr(G, ctypes.c_float), int(use_tc))
    assert rc == 0, L.last_error()
    return G


@pytest.mark.parametrize("shape", [(1000, 300), (4096, 1000), (777, 257), (300, 1030), (20000, 512)])
def test_gram_tf32x3_matches_float64(shape):
    """3xTF32 on tcgen05 must give fp32-GEMM accuracy: |G - A^T A| <= 4e-6 * (|A|^T |A|) entrywise
    (plain fp32 accumulation of m terms is ~ sqrt(m) * 6e-8; a single-TF32 product would be ~5e-4),
    G exactly symmetric, and no worse than a plain fp32 product on the CUDA cores (dense_factor.cuh)."""
    m, n = shape
    rng = np.random.default_rng(7)
    A = (rng.standard_normal((m, n)) * np.exp(rng.uniform(-3, 3, size=(1, n)))).astype(np.float32)
    ref = A.astype(np.float64).T @ A.astype(np.float64)
    scale = np.abs(A).astype(np.float64).T @ np.abs(A).astype(np.float64)
    G = dev_gram(A, use_tc=1).astype(np.float64)
    err = np.max(np.abs(G - ref) / scale)
    assert np.array_equal(G, G.T)
    assert err < 4e-6, err
    Gl = dev_gram(A, use_tc=0).astype(np.float64)   # the library's CUDA-core product (fp32 FMA accumulation)
    assert np.array_equal(Gl, Gl.T)
    err_lib = np.max(np.abs(Gl - ref) / scale)
    assert err_lib < 2e-5, err_lib
    assert err < 6 * err_lib + 1e-7, (err, err_lib)


def test_gram_tf32x3_is_deterministic():
    rng = np.random.default_rng(8)
    A = rng.standard_normal((3000, 700)).astype(np.float32)
    assert np.array_equal(dev_gram(A), dev_gram(A))


# ---- setup variants: one-pass sweeps, working precision of the factor, Gram back end ------------------
@pytest.mark.parametrize("case", ["c1_lasso_500x300", "lasso_odd_503x301", "c2s_lasso_10000x1000"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_one_pass_setup_matches_two_pass_setup(monkeypatch, case, dtype):
    """Sinkhorn-Knopp and the power iteration on the single-pass kernel (one pass over A per
    sweep) against the two-product sweeps: same d, e, norm estimate and sweep count."""
    import pogs_b200

    p = problems.build(case)
    monkeypatch.setenv("POGS_B200_FORCE_FUSE", "1")
    out = {}
    for mode in ("one_pass", "two_pass"):
        if mode == "two_pass":
            monkeypatch.setenv("POGS_B200_NO_FUSE_SETUP", "1")
        else:
            monkeypatch.delenv("POGS_B200_NO_FUSE_SETUP", raising=False)
        with pogs_b200.Solver(p["A"], dtype=dtype) as s:
            out[mode] = s.equilibration()
    (d1, e1, n1), (d2, e2, n2) = out["one_pass"], out["two_pass"]
    tol = 1e-12 if dtype == np.float64 else 5e-6
    assert relerr(d1, d2) < tol and relerr(e1, e2) < tol
    assert n1 == pytest.approx(n2, rel=20 * tol)


@pytest.mark.parametrize("case", ["c1_lasso_500x300", "c2s_lasso_10000x1000", "c4s_logistic_20000x500"])
def test_factor_variants_give_the_same_projection(monkeypatch, case):
    """fp32 data: (I + A^T A)^-1 from (a) the fp32 blocked Cholesky + triangular inverse + tensor-core
    X^T X on the tensor-core Gram matrix (default), (b) the same factorisation in fp64 on the tensor-core
    Gram matrix, (c) in fp64 on the CUDA-core Gram matrix, (d) everything in fp32 on the CUDA cores must
    project alike."""
    import pogs_b200

    p = problems.build(case)
    m, n = p["A"].shape
    rng = np.random.default_rng(5)
    x0 = rng.standard_normal(n).astype(np.float32); y0 = rng.standard_normal(m).astype(np.float32)
    res = {}
    for tag, env in (("tc", {}), ("tc_fp64", {"POGS_B200_FACTOR": "fp64"}),
                     ("lib_fp64", {"POGS_B200_FACTOR": "fp64", "POGS_B200_GRAM": "plain"}),
                     ("plain_fp32", {"POGS_B200_FACTOR": "fp32", "POGS_B200_GRAM": "plain"})):
        for k in ("POGS_B200_FACTOR", "POGS_B200_GRAM"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with pogs_b200.Solver(p["A"], dtype=np.float32) as s:
            res[tag] = s.project(x0, y0)
    for tag in ("tc", "tc_fp64", "plain_fp32"):
        assert relerr(res[tag][0], res["lib_fp64"][0]) < 2e-5, tag
        assert relerr(res[tag][1], res["lib_fp64"][1]) < 2e-5, tag


def test_memory_pool_trim_and_reuse():
    """Solver buffers come from the library's pool: building, closing and trimming repeatedly must
    neither fail nor change results."""
    import pogs_b200
    from pogs_b200 import FunctionVector, _lib

    p = problems.build("c1_lasso_500x300")
    f = FunctionVector(500, *p["f"]); g = FunctionVector(300, *p["g"])
    xs = []
    for i in range(3):
        with pogs_b200.Solver(p["A"], dtype=np.float32) as s:
            assert s.Solve(f, g) == 0
            xs.append(s.result()["x"])
        if i == 1:
            _lib.trim_memory()
    assert np.array_equal(xs[0], xs[1]) and np.array_equal(xs[0], xs[2])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(10000, 1000), (3000, 2100), (2500, 1031)])
def test_symmetric_factor_apply_matches_full_product(monkeypatch, shape, dtype):
    """x = M u from the lower triangle of M only (k_symv_*) against the full row-dot product:
    same projection to rounding, diagonal-crossing tiles and ragged edges included; deterministic."""
    import pogs_b200

    m, n = shape
    rng = np.random.default_rng(13)
    A = rng.standard_normal((m, n)).astype(dtype)
    x0 = rng.standard_normal(n).astype(dtype); y0 = rng.standard_normal(m).astype(dtype)
    res = {}
    for mode in ("symv", "full"):
        if mode == "symv":
            monkeypatch.setenv("POGS_B200_SYMV", "1")   # opt-in path
        else:
            monkeypatch.delenv("POGS_B200_SYMV", raising=False)
        with pogs_b200.Solver(A, dtype=dtype) as s:
            res[mode] = s.project(x0, y0)
            if mode == "symv":
                again = s.project(x0, y0)
                assert np.array_equal(again[0], res[mode][0])
    tol = 1e-12 if dtype == np.float64 else 2e-5
    assert relerr(res["symv"][0], res["full"][0]) < tol
    assert relerr(res["symv"][1], res["full"][1]) < tol

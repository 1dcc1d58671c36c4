"""GPU suite, sparse path: CSR/CSC operator + CGLS projector through PogsSparseD /
PogsSparseS against the oracle, the reference golden vectors and the dense path."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

import problems
from conftest import relerr
from test_gpu_solve import OPT_TOL, X_TOL, solve_dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["sparse_lasso_2000x300", "c5s_sparse_lasso_100000x10000"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sparse_matches_reference_golden(golden, name, dtype):
    p = problems.build(name)
    r = solve_dev(p, dtype)
    tag = f"{name}/{np.dtype(dtype).name}"
    assert r["status"] == int(golden[tag + "/status"]) == 0
    it_ref = int(golden[tag + "/iterations"])
    assert abs(r["iterations"] - it_ref) <= max(5, it_ref // 10)
    assert relerr(r["x"], golden[tag + "/x"]) < X_TOL
    ov = float(golden[tag + "/optval"])
    assert abs(r["optval"] - ov) <= OPT_TOL * abs(ov)
    assert abs(np.linalg.norm(r["y"].astype(np.float64)) - float(golden[tag + "/y_norm"])) <= X_TOL * float(golden[tag + "/y_norm"])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sparse_trajectory_matches_oracle(oracle, dtype):
    """tol = 0, K iterations: same iterates as the oracle's CGLS path to rounding."""
    p = problems.build("sparse_lasso_2000x300")
    for K in (1, 3, 20):
        o = oracle.solve(p["A"], p["f"], p["g"], dtype=dtype, abs_tol=0.0, rel_tol=0.0, max_iter=K)
        r = solve_dev(p, dtype, abs_tol=0.0, rel_tol=0.0, max_iter=K)
        assert r["status"] == 3 and r["iterations"] == K - 1
        tol = 1e-7 if dtype == np.float64 else 2e-3
        assert relerr(r["x"], o["x"]) < tol and relerr(r["y"], o["y"]) < tol


def test_sparse_equals_dense_solution(oracle):
    """Same matrix through the sparse (CGLS) and the dense (direct) path: same minimiser
    (the reference shows the same agreement)."""
    p = problems.build("sparse_lasso_2000x300")
    rs = solve_dev(p, np.float64)
    pd = dict(p); pd["A"] = p["A"].toarray()
    rd = solve_dev(pd, np.float64)
    assert rs["status"] == rd["status"] == 0
    assert relerr(rs["x"], rd["x"]) < 2e-3 and abs(rs["optval"] - rd["optval"]) < 1e-3 * abs(rd["optval"])


def test_sparse_csc_input_and_ragged_rows(oracle):
    """COL_MAJ (CSC) input, rows with zero entries, an empty column."""
    from pogs_b200 import FunctionVector, _lib

    rng = np.random.default_rng(3)
    m, n = 400, 90
    A = sp.random(m, n, density=0.05, format="lil", random_state=4, data_rvs=rng.standard_normal)
    A[5, :] = 0; A[17, :] = 0; A[:, 11] = 0
    A = sp.csr_matrix(A); A.eliminate_zeros()
    b = rng.standard_normal(m)
    lam = 0.1 * np.abs(A.T @ b).max()
    f = (problems.SQUARE, 1.0, b, 1.0, 0.0, 0.0); g = (problems.ABS, 1.0, 0.0, lam, 0.0, 0.0)
    o = oracle.solve(A, f, g, dtype=np.float64)
    Ac = A.tocsc()
    data = np.ascontiguousarray(Ac.data, np.float64); ptr = np.ascontiguousarray(Ac.indptr, np.int32)
    ind = np.ascontiguousarray(Ac.indices, np.int32)
    fa = FunctionVector(m, *f).arrays(np.float64); ga = FunctionVector(n, *g).arrays(np.float64)
    x = np.zeros(n); y = np.zeros(m); l = np.zeros(m); ov = ctypes.c_double(); it = ctypes.c_uint()
    ct = ctypes.c_double
    P = lambda arrs: [_lib.ptr(v, ct) for v in arrs[:5]] + [_lib.ptr(arrs[5], ctypes.c_int)]
    st = _lib.lib.PogsSparseD(0, m, n, Ac.nnz, _lib.ptr(data, ct), _lib.ptr(ptr, ctypes.c_int), _lib.ptr(ind, ctypes.c_int),
                              *P(fa), *P(ga), 1.0, 1e-4, 1e-4, 2500, 0, 1, 1, _lib.ptr(x, ct), _lib.ptr(y, ct),
                              _lib.ptr(l, ct), ctypes.byref(ov), ctypes.byref(it))
    assert st == o["status"] == 0
    assert relerr(x, o["x"]) < X_TOL and abs(ov.value - o["optval"]) <= OPT_TOL * abs(o["optval"])


def test_sparse_setup_and_projection_match_oracle(oracle):
    import pogs_b200

    p = problems.build("sparse_lasso_2000x300")
    for dtype, tol in ((np.float64, 1e-9), (np.float32, 5e-5)):
        so = oracle.Solver(p["A"], dtype=dtype)
        d0, e0, n0, _ = so.setup()
        with pogs_b200.Solver(p["A"], dtype=dtype) as s:
            d1, e1, n1 = s.equilibration()
            assert relerr(d1, d0) < tol and relerr(e1, e0) < tol and n1 == pytest.approx(n0, rel=20 * tol)
            rng = np.random.default_rng(9)
            x0 = rng.standard_normal(300); y0 = rng.standard_normal(2000)
            xo, yo = so.project(x0, y0, tol=1e-8)
            xd, yd = s.project(x0, y0)
            assert relerr(xd, xo) < max(tol, 1e-6) * 10 and relerr(yd, yo) < max(tol, 1e-6) * 10
        so.close()


def test_sparse_persistent_warm_start():
    """Second solve with a nearby lambda on the same handle needs far fewer iterations."""
    import pogs_b200
    from pogs_b200 import FunctionVector

    p = problems.build("sparse_lasso_2000x300")
    m, n = p["A"].shape
    with pogs_b200.Solver(p["A"], dtype=np.float64) as s:
        f = FunctionVector(m, *p["f"])
        s.Solve(f, FunctionVector(n, *p["g"]))
        it0 = s.GetFinalIter()
        gh, ga, gb, gc, gd, ge = p["g"]
        s.Solve(f, FunctionVector(n, gh, ga, gb, 0.9 * gc, gd, ge))
        assert s.status == 0 and s.GetFinalIter() < it0 / 3
        assert s.timing()["cgls_iterations"] > 0


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sparse_graph_loop_matches_host_driven_loop(monkeypatch, dtype):
    """The captured iteration (CGLS inner loop in a CUDA-graph WHILE node) and the host-driven loop
    run the same kernels in the same order: identical iterates, iteration and CGLS counts."""
    import pogs_b200
    from pogs_b200 import FunctionVector

    p = problems.build("sparse_lasso_2000x300")
    m, n = p["A"].shape
    f = FunctionVector(m, *p["f"]); g = FunctionVector(n, *p["g"])
    out = {}
    for mode in ("graph", "host"):
        if mode == "host":
            monkeypatch.setenv("POGS_B200_NO_GRAPH", "1")
        else:
            monkeypatch.delenv("POGS_B200_NO_GRAPH", raising=False)
        with pogs_b200.Solver(p["A"], dtype=dtype) as s:
            st = s.Solve(f, g)
            out[mode] = (st, s.result(), s.timing())
    (sg, rg, tg), (sh, rh, th) = out["graph"], out["host"]
    assert sg == sh == 0
    assert rg["iterations"] == rh["iterations"] and tg["cgls_iterations"] == th["cgls_iterations"]
    assert np.array_equal(rg["x"], rh["x"]) and np.array_equal(rg["y"], rh["y"])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sparse_tiled_layout_matches_plain_layout(monkeypatch, dtype):
    """The 2-D tiled re-layout (default for large matrices, forced here) only changes where the
    entries live and the order of the row sums: same solution as the plain CSR/CSC products, for a
    CSR input and for a CSC input with empty rows and an empty column."""
    import pogs_b200
    from pogs_b200 import FunctionVector

    rng = np.random.default_rng(11)
    A1 = problems.build("sparse_lasso_2000x300")["A"]
    A2 = sp.random(700, 260, density=0.03, format="lil", random_state=5, data_rvs=rng.standard_normal)
    A2[5, :] = 0; A2[:, 17] = 0
    A2 = A2.tocsc()
    for A in (A1, A2):
        m, n = A.shape
        b = rng.standard_normal(m)
        f = FunctionVector(m, pogs_b200.Function.kSquare, 1.0, b, 1.0)
        g = FunctionVector(n, pogs_b200.Function.kAbs, 1.0, 0.0, 0.5)
        out = {}
        for mode in ("tiled", "plain"):
            monkeypatch.setenv("POGS_B200_SPMV", mode)
            with pogs_b200.Solver(A, dtype=dtype) as s:
                st = s.Solve(f, g)
                out[mode] = (st, s.result(), s.equilibration())
        sp_, rp, ep = out["plain"]
        sb, rb, eb = out["tiled"]
        tol = 1e-9 if dtype == np.float64 else 2e-4
        assert sb == sp_ == 0
        assert relerr(eb[0], ep[0]) < tol and relerr(eb[1], ep[1]) < tol
        assert abs(rb["iterations"] - rp["iterations"]) <= max(3, rp["iterations"] // 20)
        assert relerr(rb["x"], rp["x"]) < (1e-6 if dtype == np.float64 else 1e-3)


@pytest.mark.parametrize("grid", [None, 5])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sparse_tiled_layout_with_several_column_tiles(monkeypatch, dtype, grid):
    """A matrix wide enough for several column tiles in both copies (the slice of the multiplied vector of one
    tile is bounded by shared memory, so the row sums are folded over tiles): the equilibration (50 sweeps of
    products with the squared entries over both copies), the projection (CGLS to 1e-8: products with A and
    A^T) and the solution agree with the plain CSR / CSC products.  Iteration counts of this slowly
    converging problem are not compared: they depend on the summation order (oracle 479, plain 497).
    grid = 5: the product kernel runs on 5 CTAs, each walking ~30 tiles (re-staging the slice of v, the ring
    parities carried from tile to tile)."""
    import pogs_b200
    from pogs_b200 import FunctionVector

    rng = np.random.default_rng(11)
    m, n, per_row = 120000, 50000, 8   # (scipy.sparse.random takes minutes for 6e9 cells)
    A = sp.csr_matrix((rng.standard_normal(m * per_row), (np.repeat(np.arange(m), per_row), rng.integers(0, n, m * per_row))),
                      shape=(m, n))
    b = rng.standard_normal(m)
    f = FunctionVector(m, pogs_b200.Function.kSquare, 1.0, b, 1.0)
    g = FunctionVector(n, pogs_b200.Function.kAbs, 1.0, 0.0, 0.5)
    x0 = rng.standard_normal(n); y0 = rng.standard_normal(m)
    out = {}
    if grid is not None:
        monkeypatch.setenv("POGS_B200_TL_GRID", str(grid))
    for mode in ("tiled", "plain"):
        monkeypatch.setenv("POGS_B200_SPMV", mode)
        with pogs_b200.Solver(A, dtype=dtype) as s:
            px, py = s.project(x0, y0)
            st = s.Solve(f, g)
            out[mode] = (st, s.result(), s.equilibration(), px, py)
    sp_, rp, ep, pxp, pyp = out["plain"]
    sb, rb, eb, pxb, pyb = out["tiled"]
    tol = 1e-9 if dtype == np.float64 else 2e-4
    assert relerr(eb[0], ep[0]) < tol and relerr(eb[1], ep[1]) < tol
    ptol = 1e-7 if dtype == np.float64 else 2e-4
    assert relerr(pxb, pxp) < ptol and relerr(pyb, pyp) < ptol
    assert sb == sp_ == 0
    assert abs(rb["optval"] - rp["optval"]) / abs(rp["optval"]) < 5e-4
    assert relerr(rb["x"], rp["x"]) < 2e-2

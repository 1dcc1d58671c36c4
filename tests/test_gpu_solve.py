"""GPU suite, solver level: the CUDA path through the C ABI against the oracle on
the same seeded inputs, against the committed reference golden vectors, and
through size-independent properties."""
import numpy as np
import pytest

import problems
from conftest import relerr

pytestmark = pytest.mark.gpu

DENSE_CASES = [c for c in problems.CASES if "sparse" not in c]

# Tolerances (stated): both sides stop at rel_tol = abs_tol = 1e-4 on
# slightly different trajectories (fp summation order), so x / optval agree to a small
# multiple of the solver tolerance; the strict fp64 runs below pin the fixed point itself.
X_TOL = 5e-4
OPT_TOL = 5e-4


def solve_dev(p, dtype, **kw):
    import pogs_b200
    from pogs_b200 import FunctionVector

    m, n = p["A"].shape
    f = FunctionVector(m, *p["f"]); g = FunctionVector(n, *p["g"])
    args = dict(abs_tol=1e-4, rel_tol=1e-4, max_iter=2500, verbose=0, rho=1.0, adaptive_rho=True, gap_stop=True)
    args.update(p.get("solver_kwargs", {})); args.update(kw)
    return pogs_b200._solve_graph_form(p["A"], f, g, dtype=dtype, **args)


@pytest.mark.parametrize("name", DENSE_CASES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_dense_matches_reference_golden(golden, name, dtype):
    p = problems.build(name)
    r = solve_dev(p, dtype)
    tag = f"{name}/{np.dtype(dtype).name}"
    assert r["status"] == int(golden[tag + "/status"]) == 0
    it_ref = int(golden[tag + "/iterations"])
    assert abs(r["iterations"] - it_ref) <= max(5, it_ref // 10), (r["iterations"], it_ref)
    assert relerr(r["x"], golden[tag + "/x"]) < X_TOL
    ov = float(golden[tag + "/optval"])
    assert abs(r["optval"] - ov) <= OPT_TOL * abs(ov)
    assert abs(np.linalg.norm(r["y"].astype(np.float64)) - float(golden[tag + "/y_norm"])) <= X_TOL * float(golden[tag + "/y_norm"])
    assert abs(np.linalg.norm(r["l"].astype(np.float64)) - float(golden[tag + "/l_norm"])) <= 5e-3 * float(golden[tag + "/l_norm"])


@pytest.mark.parametrize("name", ["c1_lasso_500x300", "ridge_500x300", "svm_600x200", "lasso_wide_200x400"])
def test_dense_strict_fp64(golden, name):
    """Both sides driven to 1e-7: agreement far below the default tolerance."""
    p = problems.build(name)
    r = solve_dev(p, np.float64, abs_tol=1e-7, rel_tol=1e-7, max_iter=20000)
    tag = f"{name}/strict64"
    assert r["status"] == int(golden[tag + "/status"])
    assert relerr(r["x"], golden[tag + "/x"]) < 2e-6
    assert abs(r["optval"] - float(golden[tag + "/optval"])) <= 1e-6 * abs(float(golden[tag + "/optval"]))


@pytest.mark.parametrize("name", ["c1_lasso_500x300", "c4s_logistic_20000x500", "lasso_wide_200x400"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_dense_matches_oracle_same_inputs(oracle, name, dtype):
    p = problems.build(name)
    o = oracle.solve(p["A"], p["f"], p["g"], dtype=dtype, **p["solver_kwargs"])
    r = solve_dev(p, dtype)
    assert r["status"] == o["status"]
    assert relerr(r["x"], o["x"]) < X_TOL and relerr(r["y"], o["y"]) < X_TOL
    assert abs(r["optval"] - o["optval"]) <= OPT_TOL * abs(o["optval"])


def test_trajectory_matches_oracle_iteration_by_iteration(oracle):
    """With the stopping rule disabled (tol = 0) both sides run exactly K iterations
    from the same start; the iterates must agree to rounding."""
    p = problems.build("c1_lasso_500x300")
    for K in (1, 2, 7, 60):
        o = oracle.solve(p["A"], p["f"], p["g"], dtype=np.float64, abs_tol=0.0, rel_tol=0.0, max_iter=K)
        r = solve_dev(p, np.float64, abs_tol=0.0, rel_tol=0.0, max_iter=K)
        assert r["status"] == o["status"] == 3 and r["iterations"] == o["iterations"] == K - 1
        assert relerr(r["x"], o["x"]) < 1e-8 and relerr(r["y"], o["y"]) < 1e-8 and relerr(r["l"], o["l"]) < 1e-7
        assert r["optval"] == pytest.approx(o["optval"], rel=1e-9)


def test_column_major_input(oracle):
    import ctypes

    from pogs_b200 import FunctionVector, _lib

    p = problems.build("c1_lasso_500x300")
    m, n = p["A"].shape
    o = oracle.solve(p["A"], p["f"], p["g"], dtype=np.float64)
    Af = np.asfortranarray(p["A"])
    f = FunctionVector(m, *p["f"]).arrays(np.float64); g = FunctionVector(n, *p["g"]).arrays(np.float64)
    x = np.zeros(n); y = np.zeros(m); l = np.zeros(m); ov = ctypes.c_double(); it = ctypes.c_uint()
    ct = ctypes.c_double
    P = lambda arrs: [_lib.ptr(v, ct) for v in arrs[:5]] + [_lib.ptr(arrs[5], ctypes.c_int)]
    st = _lib.lib.PogsD(0, m, n, Af.ctypes.data_as(ctypes.POINTER(ct)), *P(f), *P(g), 1.0, 1e-4, 1e-4, 2500, 0, 1, 1,
                        _lib.ptr(x, ct), _lib.ptr(y, ct), _lib.ptr(l, ct), ctypes.byref(ov), ctypes.byref(it))
    assert st == 0
    assert relerr(x, o["x"]) < X_TOL and abs(ov.value - o["optval"]) < OPT_TOL * abs(o["optval"])


def test_reference_c_interface_case():
    """tests/test_c_interface.cpp:16-73: 2x2 Lasso must return 0 with |Ax - y|_1 < 0.1."""
    import pogs_b200

    A = np.array([[1.0, 1.0], [1.0, -1.0]]); b = np.array([2.0, 0.0])
    r = pogs_b200.solve_lasso(A, b, 0.1)
    assert r["status"] == 0 and r["optval"] >= 0
    assert np.abs(A @ r["x"] - r["y"]).sum() < 0.1


def test_reference_solver_smoke_cases():
    """tests/test_solver.cpp:43-221: Lasso / Ridge / NNLS on tiny structured A."""
    import pogs_b200

    A = np.zeros((10, 5)); A[:5, :5] = np.eye(5); A[5:, :] = 0.1
    b = np.ones(10)
    for fn, args in ((pogs_b200.solve_lasso, (0.1,)), (pogs_b200.solve_ridge, (0.1,)), (pogs_b200.solve_nonneg_ls, ())):
        r = fn(A, b, *args)
        assert r["status"] == 0 and 0 < r["optval"] < 10 and np.all(np.abs(r["x"]) < 2)
    assert np.all(pogs_b200.solve_nonneg_ls(A, b)["x"] >= -1e-6)


def test_wrappers_match_oracle(oracle):
    """Every solve_* wrapper (canonical encodings of graph.py) against the oracle."""
    import pogs_b200

    rng = np.random.default_rng(21)
    m, n = 400, 120
    A = rng.standard_normal((m, n)); xs = rng.standard_normal(n) * (rng.random(n) < 0.3)
    b = A @ xs + 0.1 * rng.standard_normal(m); lab = np.sign(b); lab[lab == 0] = 1
    lam = 0.1 * np.abs(A.T @ b).max()
    S, AB, LG, HB, MP, GE, ZR = problems.SQUARE, problems.ABS, problems.LOGISTIC, problems.HUBER, problems.MAXPOS0, problems.GE0, problems.ZERO
    cases = [
        (pogs_b200.solve_lasso(A, b, lam), (S, 1, b, 1, 0, 0), (AB, 1, 0, lam, 0, 0)),
        (pogs_b200.solve_ridge(A, b, lam), (S, 1, b, 1, 0, 0), (S, 1, 0, lam, 0, 0)),
        (pogs_b200.solve_elastic_net(A, b, lam, 0.5 * lam), (S, 1, b, 1, 0, 0), (AB, 1, 0, lam, 0, 0.25 * lam)),
        (pogs_b200.solve_logistic(A, lab, 0.05 * lam), (LG, -lab, 0, 1, 0, 0), (AB, 1, 0, 0.05 * lam, 0, 0)),
        (pogs_b200.solve_logistic(A[:, :20], lab), (LG, -lab, 0, 1, 0, 0), (ZR, 1, 0, 1, 0, 0)),
        (pogs_b200.solve_huber(A, b, 1.5, lam), (HB, 1 / 1.5, b / 1.5, 2.25, 0, 0), (AB, 1, 0, lam, 0, 0)),
        (pogs_b200.solve_svm(A, lab, 2.0), (MP, -lab, -1, 1, 0, 0), (S, 1, 0, 2.0, 0, 0)),
        (pogs_b200.solve_nonneg_ls(A, b), (S, 1, b, 1, 0, 0), (GE, 1, 0, 1, 0, 0)),
    ]
    for i, (r, f, g) in enumerate(cases):
        Ai = A[:, :20] if i == 4 else A
        o = oracle.solve(Ai, f, g, dtype=np.float64)
        assert r["status"] == o["status"], i
        assert relerr(r["x"], o["x"]) < X_TOL * (5 if i == 4 else 1), i
        assert abs(r["optval"] - o["optval"]) <= OPT_TOL * abs(o["optval"]), i


def test_max_iter_status_and_final_iter():
    """A run that hits max_iter returns POGS_MAX_ITER (3); final_iter is zero-based."""
    p = problems.build("c1_lasso_500x300")
    r = solve_dev(p, np.float32, max_iter=10)
    assert r["status"] == 3 and r["iterations"] == 9
    r = solve_dev(p, np.float32, max_iter=1)
    assert r["status"] == 3 and r["iterations"] == 0


def test_solve_is_deterministic():
    p = problems.build("c2s_lasso_10000x1000")
    a = solve_dev(p, np.float32)
    b = solve_dev(p, np.float32)
    assert a["iterations"] == b["iterations"] and np.array_equal(a["x"], b["x"]) and a["optval"] == b["optval"]


def test_graph_and_direct_launch_agree(monkeypatch):
    """CUDA-graph replay and plain launches run the same kernels: identical bits."""
    p = problems.build("c1_lasso_500x300")
    a = solve_dev(p, np.float32)
    monkeypatch.setenv("POGS_B200_NO_GRAPH", "1")
    b = solve_dev(p, np.float32)
    assert a["iterations"] == b["iterations"] and np.array_equal(a["x"], b["x"])


# ---- persistent solver: warm start and lambda path ---------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_lambda_path_matches_reference_persistent_object(golden, dtype):
    """examples/cpp/lasso_path.cpp protocol against golden vectors from the reference's own
    PogsDirect object: same collapse of the iteration counts, same x per lambda."""
    import pogs_b200
    from pogs_b200 import FunctionVector

    p = problems.elastic_net(m=1000, n=200, seed=2)
    tag = f"path_enet_1000x200/{np.dtype(dtype).name}"
    lams = golden[tag + "/lambdas"]
    ref_its = golden[tag + "/iterations"]
    its = []
    with pogs_b200.Solver(p["A"], dtype=dtype) as s:
        f = FunctionVector(1000, *p["f"])
        for k, lam in enumerate(lams):
            g = FunctionVector(200, problems.ABS, 1.0, 0.0, lam, 0.0, 0.05 * p["lmax"] / 2)
            assert s.Solve(f, g) == 0
            r = s.result()
            its.append(r["iterations"])
            assert relerr(r["x"], golden[tag + "/x"][k]) < 2e-3
            assert abs(r["optval"] - golden[tag + "/optval"][k]) <= 1e-3 * abs(golden[tag + "/optval"][k])
    assert abs(its[0] - ref_its[0]) <= 10
    assert max(its[1:]) <= max(ref_its[1:]) + 5 and max(its[1:]) < its[0] / 5


def test_explicit_warm_start(oracle):
    """SetInitX + SetInitLambda (pogs.cpp:144-156) against the oracle; one-sided is an error."""
    import pogs_b200
    from pogs_b200 import FunctionVector

    p = problems.build("c1_lasso_500x300")
    cold = oracle.solve(p["A"], p["f"], p["g"], dtype=np.float64)
    so = oracle.Solver(p["A"], dtype=np.float64)
    warm_o = so.solve(p["f"], p["g"], rho=1.0, init_x=cold["x"], init_lambda=cold["l"])
    so.close()
    with pogs_b200.Solver(p["A"], dtype=np.float64) as s:
        f = FunctionVector(500, *p["f"]); g = FunctionVector(300, *p["g"])
        s.SetInitX(cold["x"]); s.SetInitLambda(cold["l"])
        assert s.Solve(f, g) == 0
        r = s.result()
        assert abs(r["iterations"] - warm_o["iterations"]) <= 5 and r["iterations"] < cold["iterations"]
        assert relerr(r["x"], warm_o["x"]) < X_TOL
        s.SetInitX(cold["x"])
        with pytest.raises(RuntimeError):
            s.Solve(f, g)


# ---- properties at a larger size (the oracle would take too long) ---------------------------------------
def test_kkt_conditions_large_lasso():
    """Lasso 40000 x 2000 fp32: primal feasibility y = A x, dual feasibility
    A^T lambda + mu = 0 with mu in lambda * d|x|, and objective consistency."""
    import pogs_b200
    from pogs_b200 import FunctionVector

    m, n = 40000, 2000
    rng = np.random.default_rng(11)
    A = rng.standard_normal((m, n), dtype=np.float32)
    xs = (rng.standard_normal(n) * (rng.random(n) < 0.2)).astype(np.float32)
    b = A @ xs + 0.1 * rng.standard_normal(m).astype(np.float32)
    lam = 0.1 * float(np.abs(A.T @ b).max())
    with pogs_b200.Solver(A, dtype=np.float32) as s:
        f = FunctionVector(m, problems.SQUARE, 1.0, b, 1.0); g = FunctionVector(n, problems.ABS, 1.0, 0.0, lam)
        assert s.Solve(f, g) == 0
        r = s.result()
    x = r["x"].astype(np.float64); y = r["y"].astype(np.float64); l = r["l"].astype(np.float64); mu = r["mu"].astype(np.float64)
    A64 = A.astype(np.float64)
    assert np.linalg.norm(A64 @ x - y) <= 2e-3 * np.linalg.norm(y)
    # lambda = -grad f(y) = -(y - b) for f = 1/2 (y-b)^2 ... sign convention: A^T lambda + mu = 0
    assert np.linalg.norm(A64.T @ l + mu) <= 5e-3 * np.linalg.norm(mu)
    assert np.all(np.abs(mu) <= lam * (1 + 5e-3))
    obj = 0.5 * np.sum((y - b) ** 2) + lam * np.abs(x).sum()
    assert r["optval"] == pytest.approx(obj, rel=1e-4)
    assert 0 < np.count_nonzero(np.abs(x) > 1e-6) < n


def test_scaling_invariance():
    """Idempotence-style property: scaling b and lambda by t scales x by t (Lasso is
    positively homogeneous); checks the whole pipeline incl. equilibration at fp32."""
    p = problems.build("c2s_lasso_10000x1000")
    r1 = solve_dev(p, np.float32, abs_tol=1e-5, rel_tol=1e-5)
    h, a, b, c, d, e = p["f"]
    p2 = dict(p); p2["f"] = (h, a, 3.0 * b, c, d, e)
    gh, ga, gb, gc, gd, ge = p["g"]; p2["g"] = (gh, ga, gb, 3.0 * gc, gd, ge)
    r2 = solve_dev(p2, np.float32, abs_tol=1e-5, rel_tol=1e-5)
    assert relerr(r2["x"], 3.0 * r1["x"]) < 2e-3


# ---- single-pass kernel (speculative next half-step) ------------------------------------------------------
@pytest.mark.parametrize("name", ["c1_lasso_500x300", "c2s_lasso_10000x1000", "c4s_logistic_20000x500", "svm_600x200",
                                  "lasso_odd_503x301", "huber_500x300", "nnls_500x300"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_single_pass_matches_two_pass(monkeypatch, name, dtype):
    """The fused pass only re-orders when things are computed: with it on and off the solver
    must reach the same point in (nearly) the same number of iterations, and a good share of
    the iterations must actually have run on one pass over A."""
    import pogs_b200
    from pogs_b200 import FunctionVector

    p = problems.build(name)
    m, n = p["A"].shape
    f = FunctionVector(m, *p["f"]); g = FunctionVector(n, *p["g"])
    out = {}
    monkeypatch.setenv("POGS_B200_FORCE_FUSE", "1")   # also on shapes where it is not the default
    for mode in ("fused", "two_pass"):
        if mode == "two_pass":
            monkeypatch.setenv("POGS_B200_NO_FUSE", "1")
        else:
            monkeypatch.delenv("POGS_B200_NO_FUSE", raising=False)
        with pogs_b200.Solver(p["A"], dtype=dtype) as s:
            assert s.Solve(f, g) == 0
            out[mode] = (s.result(), s.timing())
    rf, tf = out["fused"]; r2, t2 = out["two_pass"]
    assert t2["single_pass_iterations"] == 0
    assert tf["single_pass_iterations"] >= 0.5 * tf["iterations"]
    assert abs(rf["iterations"] - r2["iterations"]) <= max(3, r2["iterations"] // 20)
    assert relerr(rf["x"], r2["x"]) < (1e-6 if dtype == np.float64 else 5e-4)
    assert abs(rf["optval"] - r2["optval"]) <= (1e-7 if dtype == np.float64 else 2e-4) * abs(r2["optval"])


def test_single_pass_fixed_iterations_match_oracle(oracle, monkeypatch):
    """tol = 0, no adaptive rho: every iteration after the first commits the speculation; the
    iterates must still follow the oracle step for step."""
    import pogs_b200
    from pogs_b200 import FunctionVector

    monkeypatch.setenv("POGS_B200_FORCE_FUSE", "1")
    p = problems.build("c1_lasso_500x300")
    f = FunctionVector(500, *p["f"]); g = FunctionVector(300, *p["g"])
    for K in (2, 3, 8, 41):
        o = oracle.solve(p["A"], p["f"], p["g"], dtype=np.float64, abs_tol=0.0, rel_tol=0.0, max_iter=K, adaptive_rho=False)
        with pogs_b200.Solver(p["A"], dtype=np.float64) as s:
            s.SetAbsTol(0.0); s.SetRelTol(0.0); s.SetMaxIter(K); s.SetAdaptiveRho(False)
            assert s.Solve(f, g) == 3
            r = s.result(); t = s.timing()
        assert t["single_pass_iterations"] == K - 1
        assert relerr(r["x"], o["x"]) < 1e-8 and relerr(r["y"], o["y"]) < 1e-8 and relerr(r["l"], o["l"]) < 1e-7

"""CPU suite, part 1: pins the oracle (oracle/pogs_oracle.c).

 (i)   known-answer values of the base proximal maps from the reference's own unit
       tests (/root/reference/tests/test_proximal.cpp:12-371);
 (ii)  golden vectors produced by the compiled, unmodified reference
       (tests/golden/ref_golden.npz, made by tests/golden/make_golden.py);
 (iii) live comparison with oracle/_ref when it has been built on this box.
"""
import math
import os
import subprocess
import tempfile

import numpy as np
import pytest

import problems
from conftest import relerr
from problems import (ABS, BOX01, EQ0, EXP, GE0, HUBER, IDENT, LE0, LOGISTIC, MAXNEG0, MAXPOS0, NEGENTR, NEGLOG,
                      RECIPR, SQUARE, ZERO)

# (tag, v, rho, expected) -- test_proximal.cpp lines in comments
KNOWN = [
    (ZERO, 5.0, 1.0, 5.0),            # :12-19
    (IDENT, 5.0, 2.0, 4.5),           # :21-28
    (ABS, 2.0, 2.0, 1.5),             # :33-36
    (ABS, 0.3, 2.0, 0.0),             # :38-41
    (ABS, -2.0, 2.0, -1.5),           # :43-46
    (ABS, 0.5, 2.0, 0.0),             # :48-51
    (ABS, 0.0, 2.0, 0.0),             # :53-56
    (SQUARE, 6.0, 3.0, 4.5),          # :62-65
    (SQUARE, -4.0, 3.0, -3.0),        # :67-70
    (SQUARE, 0.0, 3.0, 0.0),          # :72-75
    (EQ0, 5.0, 1.0, 0.0),             # :78-85
    (GE0, 3.0, 1.0, 3.0), (GE0, -2.0, 1.0, 0.0), (GE0, 0.0, 1.0, 0.0),        # :87-104
    (LE0, -3.0, 1.0, -3.0), (LE0, 2.0, 1.0, 0.0), (LE0, 0.0, 1.0, 0.0),       # :106-123
    (BOX01, 0.5, 1.0, 0.5), (BOX01, -0.5, 1.0, 0.0), (BOX01, 1.5, 1.0, 1.0),  # :125-141
    (BOX01, 0.0, 1.0, 0.0), (BOX01, 1.0, 1.0, 1.0),                           # :143-149
    (MAXPOS0, 3.0, 2.0, 2.5), (MAXPOS0, 0.3, 2.0, 0.0), (MAXPOS0, -1.0, 2.0, -1.0),   # :152-170
    (MAXNEG0, -3.0, 2.0, -2.5), (MAXNEG0, -0.3, 2.0, 0.0), (MAXNEG0, 1.0, 2.0, 1.0),  # :173-191
    (HUBER, 0.5, 2.0, 0.5 * 2.0 / 3.0), (HUBER, 5.0, 2.0, 4.5), (HUBER, -5.0, 2.0, -4.5),
    (HUBER, 0.0, 2.0, 0.0),           # :194-219
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_known_answers(oracle, dtype):
    for h, v, rho, want in KNOWN:
        got = oracle.prox_base(h, v, rho, dtype)
        assert got == pytest.approx(want, rel=1e-6, abs=1e-7), (h, v, rho)


def test_optimality_conditions(oracle):
    # test_proximal.cpp:222-247: r + exp(r)/rho = v for kExp
    for v in (2.0, 0.0, -1.0):
        r = oracle.prox_base(EXP, v, 1.0)
        assert r < v or v <= 0
        assert r + math.exp(r) / 1.0 == pytest.approx(v, abs=1e-6)
    # :249-273: r - 1/(rho r) = v for kNegLog
    for v in (3.0, 0.5, 10.0):
        r = oracle.prox_base(NEGLOG, v, 2.0)
        assert r > 0 and r - 1.0 / (2.0 * r) == pytest.approx(v, abs=1e-6)
    # :276-300 reciprocal: positive; -1/r^2 + rho (r - v) = 0
    for v in (2.0, 10.0, 0.0):
        r = oracle.prox_base(RECIPR, v, 1.0)
        assert r >= 0 and -1.0 / r ** 2 + (r - v) == pytest.approx(0.0, abs=1e-6)
    # :303-325 neg-entropy: log r + 1 + rho (r - v) = 0
    for v in (2.0, 1.0, 0.5):
        r = oracle.prox_base(NEGENTR, v, 1.0)
        assert r > 0 and math.log(r) + 1.0 + (r - v) == pytest.approx(0.0, abs=1e-6)
    # :327-358 logistic: sigma(r) + rho (r - v) = 0
    for v in (3.0, -3.0, 0.0, 10.0):
        r = oracle.prox_base(LOGISTIC, v, 1.0)
        assert 1.0 / (1.0 + math.exp(-r)) + (r - v) == pytest.approx(0.0, abs=1e-6)
        assert r < v
    # :360-371 float vs double
    assert oracle.prox_base(ABS, 3.0, 2.0, np.float32) == pytest.approx(oracle.prox_base(ABS, 3.0, 2.0), rel=1e-5)


def test_rng_matches_libstdcxx(oracle):
    """Start vector of the norm estimate (gsl_rand.h:9-16) == libstdc++'s default engine."""
    src = r"""
#include <random>
#include <cstdio>
template <typename T> void go(int n){ std::default_random_engine g; std::uniform_real_distribution<T> d((T)0,(T)1);
  for(int i=0;i<n;++i) printf("%.17g\n",(double)d(g)); }
int main(){ go<float>(64); go<double>(64); }
"""
    with tempfile.TemporaryDirectory() as td:
        cpp = os.path.join(td, "r.cpp")
        open(cpp, "w").write(src)
        exe = os.path.join(td, "r")
        subprocess.check_call(["/usr/bin/g++", "-O1", "-o", exe, cpp])
        vals = np.array([float(v) for v in subprocess.check_output([exe]).split()])
    np.testing.assert_array_equal(oracle.rand(64, np.float32).astype(np.float64), vals[:64])
    np.testing.assert_array_equal(oracle.rand(64, np.float64), vals[64:])


FAST_CASES = [c for c in problems.CASES if c not in ("c5s_sparse_lasso_100000x10000",)]


@pytest.mark.parametrize("name", FAST_CASES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_matches_reference_golden(oracle, golden, name, dtype):
    """Same seeded instance, same parameters: the restatement must follow the reference's
    trajectory -- identical status, iteration count within rounding drift, x / y / lambda /
    optval equal far below the solver tolerance."""
    p = problems.build(name)
    r = oracle.solve(p["A"], p["f"], p["g"], dtype=dtype, **p["solver_kwargs"])
    tag = f"{name}/{np.dtype(dtype).name}"
    assert r["status"] == int(golden[tag + "/status"])
    tol = 1e-9 if dtype == np.float64 else 2e-5
    if dtype == np.float64:
        assert r["iterations"] == int(golden[tag + "/iterations"])
    else:
        assert abs(r["iterations"] - int(golden[tag + "/iterations"])) <= 3
    assert relerr(r["x"], golden[tag + "/x"]) < tol
    assert abs(r["optval"] - float(golden[tag + "/optval"])) <= tol * abs(float(golden[tag + "/optval"])) * 10
    assert abs(np.linalg.norm(r["y"].astype(np.float64)) - float(golden[tag + "/y_norm"])) <= tol * float(golden[tag + "/y_norm"]) * 10
    assert abs(np.linalg.norm(r["l"].astype(np.float64)) - float(golden[tag + "/l_norm"])) <= max(tol, 1e-6) * float(golden[tag + "/l_norm"]) * 10


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_lambda_path_matches_reference(oracle, golden, dtype):
    """Warm-started path with the reference's persistent PogsDirect object
    (examples/cpp/lasso_path.cpp protocol): iteration counts collapse after the first solve."""
    p = problems.elastic_net(m=1000, n=200, seed=2)
    tag = f"path_enet_1000x200/{np.dtype(dtype).name}"
    lams = golden[tag + "/lambdas"]
    s = oracle.Solver(p["A"], dtype=dtype)
    its = []
    for k, lam in enumerate(lams):
        g = (problems.ABS, 1.0, 0.0, lam, 0.0, 0.05 * p["lmax"] / 2)
        r = s.solve(p["f"], g, rho=1.0 if k == 0 else None)
        its.append(r["iterations"])
        assert relerr(r["x"], golden[tag + "/x"][k]) < (1e-8 if dtype == np.float64 else 5e-4)
    s.close()
    ref_its = golden[tag + "/iterations"]
    assert its[0] == ref_its[0] or dtype == np.float32
    assert max(abs(np.array(its) - ref_its)) <= (0 if dtype == np.float64 else 3)
    assert max(its[1:]) < its[0] / 5


def test_oracle_vs_live_reference(oracle):
    """When oracle/_ref is present (this container, and the GPU box via the snapshot)
    compare on a fresh instance that is not in the fixtures."""
    from oracle import ref_ctypes as R

    if not R.available():
        pytest.skip("oracle/_ref not built on this box")
    p = problems.lasso(m=700, n=150, seed=123)
    for dtype, tol in ((np.float64, 1e-10), (np.float32, 2e-5)):
        a = R.solve(p["A"], p["f"], p["g"], dtype=dtype)
        b = oracle.solve(p["A"], p["f"], p["g"], dtype=dtype)
        assert a["status"] == b["status"] == 0
        assert relerr(b["x"], a["x"]) < tol
        assert abs(a["optval"] - b["optval"]) < tol * 10 * abs(a["optval"])


def test_one_sided_warm_start_is_an_error(oracle):
    """The reference aborts (ASSERT(false), pogs.cpp:159-179); the oracle reports POGS_ERROR."""
    p = problems.lasso(m=60, n=20, seed=1)
    s = oracle.Solver(p["A"])
    r = s.solve(p["f"], p["g"], rho=1.0, init_x=np.zeros(20))
    assert r["status"] == 6
    s.close()

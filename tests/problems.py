"""Seeded synthetic instances of the BASELINE configs and their
scaled-down twins.  Everything is generated from numpy.random.default_rng(seed);
fixtures store seeds and outputs, never inputs.

A problem is a dict: A (ndarray or scipy CSR), f and g as (h, a, b, c, d, e) tuples
with scalar or per-element entries (the SoA form both oracle bindings accept).
"""
import numpy as np

# Function tags (reference prox_lib.h:23-38)
ABS, EXP, HUBER, IDENT, BOX01, EQ0, GE0, LE0, LOGISTIC, MAXNEG0, MAXPOS0, NEGENTR, NEGLOG, RECIPR, SQUARE, ZERO = range(16)


def _lasso_data(m, n, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, n))
    xs = rng.standard_normal(n) * (rng.random(n) < 0.2)
    b = A @ xs + 0.1 * rng.standard_normal(m)
    return A, xs, b, rng


def lasso(m, n, seed, lam_frac=0.1):
    """C1 / C2 recipe: A~N(0,1), 20%-dense x*, b = A x* + 0.1 N, lambda = lam_frac*|A'b|_inf."""
    A, xs, b, _ = _lasso_data(m, n, seed)
    lam = lam_frac * np.abs(A.T @ b).max()
    return dict(A=A, f=(SQUARE, 1.0, b, 1.0, 0.0, 0.0), g=(ABS, 1.0, 0.0, lam, 0.0, 0.0), b=b, lam=lam)


def ridge(m, n, seed):
    p = lasso(m, n, seed)
    p["g"] = (SQUARE, 1.0, 0.0, p["lam"], 0.0, 0.0)
    return p


def elastic_net(m, n, seed, lam_frac=0.1, lam2_frac=0.05):
    """C3 encoding: g = kAbs with c=lambda1, e=lambda2/2 (graph.py:520-522)."""
    A, xs, b, _ = _lasso_data(m, n, seed)
    lmax = np.abs(A.T @ b).max()
    return dict(A=A, f=(SQUARE, 1.0, b, 1.0, 0.0, 0.0), g=(ABS, 1.0, 0.0, lam_frac * lmax, 0.0, lam2_frac * lmax / 2),
                b=b, lmax=lmax)


def logistic(m, n, seed, lam_frac=0.01):
    """C4 recipe: labels sign(A x* + 0.1 N), f = kLogistic with a = -label, g = lambda|x|."""
    A, xs, _, rng = _lasso_data(m, n, seed)
    yl = np.sign(A @ xs + 0.1 * rng.standard_normal(m))
    yl[yl == 0] = 1.0
    lam = lam_frac * np.abs(A.T @ yl).max()
    return dict(A=A, f=(LOGISTIC, -yl, 0.0, 1.0, 0.0, 0.0), g=(ABS, 1.0, 0.0, lam, 0.0, 0.0), labels=yl, lam=lam)


def svm(m, n, seed, lam=1.0):
    A, xs, _, rng = _lasso_data(m, n, seed)
    yl = np.sign(A @ xs + 0.1 * rng.standard_normal(m))
    yl[yl == 0] = 1.0
    return dict(A=A, f=(MAXPOS0, -yl, -1.0, 1.0, 0.0, 0.0), g=(SQUARE, 1.0, 0.0, lam, 0.0, 0.0), labels=yl)


def huber(m, n, seed, delta=1.5):
    p = lasso(m, n, seed)
    b = p["b"]
    p["f"] = (HUBER, 1.0 / delta, b / delta, delta * delta, 0.0, 0.0)
    return p


def nonneg_ls(m, n, seed):
    p = lasso(m, n, seed)
    p["g"] = (GE0, 1.0, 0.0, 1.0, 0.0, 0.0)
    return p


def sparse_lasso(m, n, nnz_per_row, seed, lam_frac=0.1):
    """C5 recipe: CSR with exactly nnz_per_row entries per row at distinct uniform columns."""
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)
    # distinct columns per row: argpartition of random keys would be O(mn); sample with rejection by rows
    cols = np.empty((m, nnz_per_row), dtype=np.int32)
    for i0 in range(0, m, 4096):
        i1 = min(m, i0 + 4096)
        if n <= 4096:
            keys = rng.random((i1 - i0, n))
            cols[i0:i1] = np.argpartition(keys, nnz_per_row - 1, axis=1)[:, :nnz_per_row]
        else:
            c = rng.integers(0, n, size=(i1 - i0, nnz_per_row))
            for _ in range(8):  # re-draw duplicates
                cs = np.sort(c, axis=1)
                dup = np.zeros_like(cs, dtype=bool)
                dup[:, 1:] = cs[:, 1:] == cs[:, :-1]
                if not dup.any():
                    break
                cs[dup] = rng.integers(0, n, size=int(dup.sum()))
                c = cs
            cols[i0:i1] = c
    cols.sort(axis=1)
    vals = rng.standard_normal((m, nnz_per_row))
    indptr = np.arange(0, (m + 1) * nnz_per_row, nnz_per_row, dtype=np.int64)
    A = sp.csr_matrix((vals.ravel(), cols.ravel(), indptr), shape=(m, n))
    A.sum_duplicates()
    xs = rng.standard_normal(n) * (rng.random(n) < 0.2)
    b = A @ xs + 0.1 * rng.standard_normal(m)
    lam = lam_frac * np.abs(A.T @ b).max()
    return dict(A=A, f=(SQUARE, 1.0, b, 1.0, 0.0, 0.0), g=(ABS, 1.0, 0.0, lam, 0.0, 0.0), b=b, lam=lam)


# name -> (builder, kwargs, solver kwargs).  These are the parity cases: C1 itself and
# scaled-down twins of C2..C5 plus the other wrappers / layouts the reference supports.
CASES = {
    "c1_lasso_500x300": (lasso, dict(m=500, n=300, seed=0), {}),
    "c2s_lasso_10000x1000": (lasso, dict(m=10000, n=1000, seed=1), {}),
    "c3s_enet_5000x200": (elastic_net, dict(m=5000, n=200, seed=2), {}),
    "c4s_logistic_20000x500": (logistic, dict(m=20000, n=500, seed=3), {}),
    "c5s_sparse_lasso_100000x10000": (sparse_lasso, dict(m=100000, n=10000, nnz_per_row=10, seed=4), {}),
    "ridge_500x300": (ridge, dict(m=500, n=300, seed=5), {}),
    "svm_600x200": (svm, dict(m=600, n=200, seed=6), {}),
    "huber_500x300": (huber, dict(m=500, n=300, seed=7), {}),
    "nnls_500x300": (nonneg_ls, dict(m=500, n=300, seed=8), {}),
    "lasso_wide_200x400": (lasso, dict(m=200, n=400, seed=9), {}),
    "lasso_odd_503x301": (lasso, dict(m=503, n=301, seed=10), {}),
    "lasso_noadapt_500x300": (lasso, dict(m=500, n=300, seed=0), dict(adaptive_rho=False, gap_stop=False, max_iter=300)),
    "sparse_lasso_2000x300": (sparse_lasso, dict(m=2000, n=300, nnz_per_row=6, seed=11), {}),
}


def build(name):
    fn, kw, skw = CASES[name]
    p = fn(**kw)
    p["solver_kwargs"] = dict(skw)
    return p

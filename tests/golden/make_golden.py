"""Generates tests/golden/ref_golden.npz by running the *compiled, unmodified
reference* (oracle/_ref/libpogs_ref.so, built by oracle/build_ref.sh from
/root/reference) on the seeded cases of tests/problems.py, in fp64 and fp32,
through its own C ABI (PogsD / PogsS / PogsSparseD / PogsSparseS) with the
Python-wrapper defaults (rho=1, abs_tol=rel_tol=1e-4, max_iter=2500,
adaptive_rho=1, gap_stop=1; reference graph.py:236-247).

Also stores a strict set (abs_tol=rel_tol=1e-6... as far as 2500 iterations reach)
for the fp64 cases so that parity can be checked below the solver tolerance, and a
warm-started lambda path produced with the reference's persistent C++ object
(oracle/ref_persistent.cpp -> pogs::PogsDirect).

Run here (needs /root/reference):   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import problems  # noqa: E402
from oracle import ref_ctypes as R  # noqa: E402


def thin(v, k=64):
    """Keep fixtures small: long vectors are stored subsampled (every len/k-th entry) + norm."""
    v = np.asarray(v)
    if v.size <= 4096:
        return v
    step = v.size // 2048
    return v[::step]


def main():
    out = {}
    for name in problems.CASES:
        p = problems.build(name)
        for dt in (np.float64, np.float32):
            kw = dict(p["solver_kwargs"])
            r = R.solve(p["A"], p["f"], p["g"], dtype=dt, **kw)
            tag = f"{name}/{np.dtype(dt).name}"
            out[tag + "/x"] = r["x"]
            out[tag + "/y_thin"] = thin(r["y"])
            out[tag + "/l_thin"] = thin(r["l"])
            out[tag + "/y_norm"] = np.linalg.norm(r["y"].astype(np.float64))
            out[tag + "/l_norm"] = np.linalg.norm(r["l"].astype(np.float64))
            out[tag + "/optval"] = r["optval"]
            out[tag + "/iterations"] = r["iterations"]
            out[tag + "/status"] = r["status"]
            print(f"{tag:48s} status={r['status']} iters={r['iterations']:4d} optval={r['optval']:.10g}")
        if name in ("c1_lasso_500x300", "ridge_500x300", "svm_600x200", "c3s_enet_5000x200", "lasso_wide_200x400"):
            r = R.solve(p["A"], p["f"], p["g"], dtype=np.float64, abs_tol=1e-7, rel_tol=1e-7, max_iter=20000)
            tag = f"{name}/strict64"
            out[tag + "/x"] = r["x"]
            out[tag + "/optval"] = r["optval"]
            out[tag + "/iterations"] = r["iterations"]
            out[tag + "/status"] = r["status"]
            print(f"{tag:48s} status={r['status']} iters={r['iterations']:4d} optval={r['optval']:.12g}")

    # warm-started lambda path with the reference's persistent PogsDirect object
    if R.persistent_available():
        p = problems.elastic_net(m=1000, n=200, seed=2)
        lmax = p["lmax"]
        lams = np.exp(np.linspace(np.log(lmax), np.log(1e-2 * lmax), 10))
        for dt in (np.float64, np.float32):
            s = R.PersistentDense(p["A"], dtype=dt)
            its, opts, xs = [], [], []
            for lam in lams:
                g = (problems.ABS, 1.0, 0.0, lam, 0.0, 0.05 * lmax / 2)
                r = s.solve(p["f"], g)
                its.append(r["iterations"]); opts.append(r["optval"]); xs.append(r["x"])
            s.close()
            tag = f"path_enet_1000x200/{np.dtype(dt).name}"
            out[tag + "/lambdas"] = lams
            out[tag + "/iterations"] = np.array(its)
            out[tag + "/optval"] = np.array(opts)
            out[tag + "/x"] = np.stack(xs)
            print(f"{tag:48s} iters={its}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"), **out)
    print("wrote tests/golden/ref_golden.npz")


if __name__ == "__main__":
    main()

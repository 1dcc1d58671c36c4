"""GPU probe (dev aid): dense-indirect / sparse solves against the fp64 oracle with and without the y recurrence."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import problems
import pogs_b200
from pogs_b200 import FunctionVector
from oracle import oracle_ctypes as O

def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / np.linalg.norm(np.asarray(b, np.float64)))

for name, kw in (("svm_600x200", dict(projector="indirect")), ("c1_lasso_500x300", dict(projector="indirect")),
                 ("lasso_wide_200x400", dict(projector="indirect")), ("sparse_lasso_2000x300", {}),
                 ("c5s_sparse_lasso_100000x10000", {})):
    p = problems.build(name)
    m, n = p["A"].shape
    o = O.solve(p["A"], p["f"], p["g"], dtype=np.float64, direct=False) if kw else O.solve(p["A"], p["f"], p["g"], dtype=np.float64)
    for dtype in (np.float32, np.float64):
        for rec in ("1", "0"):
            os.environ["POGS_B200_Y_REC"] = rec
            with pogs_b200.Solver(p["A"], dtype=dtype, **kw) as s:
                st = s.Solve(FunctionVector(m, *p["f"]), FunctionVector(n, *p["g"]))
                r, t = s.result(), s.timing()
            print(name, np.dtype(dtype).name, "rec", rec, "st", st, "it", r["iterations"], "oit", o["iterations"], "cgls", int(t["cgls_iterations"]),
                  "ex %.2e" % relerr(r["x"], o["x"]), "eopt %.2e" % (abs(r["optval"] - o["optval"]) / abs(o["optval"])), flush=True)

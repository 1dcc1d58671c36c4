"""GPU probe (dev aid): C5 loop time per SpMV layout (POGS_B200_SPMV modes given on the command line)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
import pogs_b200

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c5"]
m, n = cfg["m"], cfg["n"]
A, ft, gt = bench.sparse_problem(cfg)
f, g = bench.function_vectors(ft, gt, m, n)
for mode in (sys.argv[2:] or ["tiled", "blocked"]):
    os.environ["POGS_B200_SPMV"] = mode
    t0 = time.perf_counter()
    s = pogs_b200.Solver(A, dtype=np.float32)
    s.SetAbsTol(0.0); s.SetRelTol(0.0)
    s.SetMaxIter(10); s.Solve(f, g)
    tm = s.timing()
    print(mode, os.environ.get("POGS_B200_TL"), "first solve: wall", round(time.perf_counter() - t0, 3),
          {k: round(v, 2) for k, v in tm.items() if k in ("h2d_ms", "equil_ms", "normest_ms", "loop_ms")}, flush=True)
    for K in (100, 100):
        s.SetMaxIter(K); s.Solve(f, g); tm = s.timing()
        print(mode, "K", K, "us/iter", round(tm["loop_ms"] / K * 1e3, 1), "cgls", tm["cgls_iterations"], "optval", s.result()["optval"], flush=True)
    s.close()

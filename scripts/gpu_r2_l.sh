#!/bin/bash
# round 2, call L (2+ GPUs): the one-shot C entry point on several GPUs of one process
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k one_shot 2>&1 | tail -5
timeout 900 python - <<'PY' 2>&1 | tail -12
import os, sys, time, ctypes, json
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import bench
from pogs_b200 import _lib, FunctionVector
cfg = bench.CONFIGS["c2"]; m, n = cfg["m"], cfg["n"]; K = 200
A_host = torch.empty((m, n), dtype=torch.float32, pin_memory=True)
noise = bench.fill_rows(cfg, 0, m, A_host.numpy())
rhs = bench.rhs_of(cfg, A_host.numpy(), noise, bench.x_star(cfg))
atb = np.zeros(n)
for a in range(0, m, 12500): atb += A_host.numpy()[a:a+12500].T @ rhs[a:a+12500]
ft, gt = bench.descriptor_tuples(cfg, rhs, float(np.abs(atb).max()))
f, g = bench.function_vectors(ft, gt, m, n)
fa, ga = f.arrays(np.float32), g.arrays(np.float32)
ct = ctypes.c_float
P = lambda arrs: [_lib.ptr(v, ct) for v in arrs[:5]] + [_lib.ptr(arrs[5], ctypes.c_int)]
Ap = ctypes.cast(ctypes.c_void_p(A_host.data_ptr()), ctypes.POINTER(ct))
out = {}
for G in (1, torch.cuda.device_count(), 1, torch.cuda.device_count()):
    os.environ["POGS_B200_GPUS"] = str(G)
    x = np.zeros(n, np.float32); y = np.zeros(m, np.float32); l = np.zeros(m, np.float32)
    ov = ctypes.c_float(); it = ctypes.c_uint()
    t0 = time.perf_counter()
    st = _lib.lib.PogsS(1, m, n, Ap, *P(fa), *P(ga), ct(1.0), ct(0.0), ct(0.0), K, 0, 1, 1, _lib.ptr(x, ct), _lib.ptr(y, ct), _lib.ptr(l, ct), ctypes.byref(ov), ctypes.byref(it))
    dt = time.perf_counter() - t0
    print("GPUS", G, "status", st, "iters", it.value + 1, "call_s", round(dt, 4), "it/s e2e", round(K / dt, 1), "optval", ov.value, "|x|", float(np.linalg.norm(x)), "|y|", float(np.linalg.norm(y)))
    out[f"gpus{G}"] = {"call_s": dt, "e2e_it_s": K / dt, "optval": ov.value}
json.dump(out, open("gpurun_out/r2l_oneshot_multi.json", "w"))
PY

"""GPU suite, multi-GPU: the row-block solver on 2 (or more) GPUs of one node against the
single-process oracle.  Skipped on a single-GPU box."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4])
def test_rowblock_solver_matches_oracle(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    assert p.returncode == 0 and lines, p.stdout[-3000:] + p.stderr[-3000:]
    out = json.loads(lines[-1][7:])
    print("dist world", world, json.dumps(out))
    assert out["replicas_identical"]
    for k, v in out.items():
        if k.startswith("allreduce"):
            assert v, k
        if isinstance(v, dict) and k.split("/")[0].endswith("+fuse"):
            assert v["single_pass"] >= 0.5 * (v["it"] + 1), (k, v)   # the forced single-pass path really ran
        if k == "scale_fixedK":
            assert v["st"] == 3 and v["it"] == 39 and v["single_pass"] == 39 and v["nnz"] > 0, v
            assert v["ex"] < 2e-4 and v["ey"] < 2e-4 and v["eopt"] < 5e-5, v
            continue
        if isinstance(v, dict):
            assert v["status"] == v["ostatus"] == 0, (k, v)
            assert abs(v["it"] - v["oit"]) <= max(5, v["oit"] // 10), (k, v)
            assert v["ex"] < 5e-4 and v["ey"] < 5e-4 and v["eopt"] < 5e-4, (k, v)


def test_one_shot_entry_point_drives_several_gpus(monkeypatch):
    """POGS_B200_GPUS=G: the reference-facing one-shot call (PogsS / PogsD, host pointers) splits the matrix
    into row blocks and drives G GPUs from host threads of the calling process (SURVEY 8e process model)."""
    import numpy as np

    import problems
    from conftest import relerr

    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import pogs_b200
    from oracle import oracle_ctypes as O
    from pogs_b200 import FunctionVector

    G = min(_ngpu(), 4)
    for name in ("c2s_lasso_10000x1000", "c4s_logistic_20000x500"):
        p = problems.build(name)
        m, n = p["A"].shape
        f = FunctionVector(m, *p["f"]); g = FunctionVector(n, *p["g"])
        for dtype in (np.float64, np.float32):
            monkeypatch.delenv("POGS_B200_GPUS", raising=False)
            one = pogs_b200._solve_graph_form(p["A"], f, g, dtype=dtype)
            monkeypatch.setenv("POGS_B200_GPUS", str(G))
            many = pogs_b200._solve_graph_form(p["A"], f, g, dtype=dtype)
            o = O.solve(p["A"], p["f"], p["g"], dtype=dtype)
            assert many["status"] == one["status"] == o["status"] == 0
            assert abs(many["iterations"] - o["iterations"]) <= max(5, o["iterations"] // 10)
            for k in ("x", "y", "l"):
                assert relerr(many[k], o[k]) < 5e-4, (name, dtype, k)
            assert abs(many["optval"] - o["optval"]) <= 5e-4 * abs(o["optval"])
    monkeypatch.delenv("POGS_B200_GPUS", raising=False)

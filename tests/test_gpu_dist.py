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

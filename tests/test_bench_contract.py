"""bench.py contract, CPU side: the reference arm (--impl reference) prints one JSON line with the
keys the driver reads, and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "tiny",
                        "--steps", "3", "--warmup", "1", "--cpu-rows", "2000"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ADMM iterations/sec" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 3
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    # e2e = K / (construct + first solve): includes the copy of A, equilibration, Gram + Cholesky
    e = d["e2e"]
    assert e["unit"] == "iterations/s" and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert 0 < e["value"] <= d["value"] * 1.5 and e["call_s"] > 0
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
    assert d["sanity"]["optval"] > 0


def test_reference_arm_uses_all_host_threads_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must override it."""
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "tiny",
                        "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and d["n_gpus"] == 2
    # the other ranks print nothing and exit 0
    env["RANK"] = "1"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "tiny",
                        "--steps", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_synthetic_matrix_does_not_depend_on_the_row_split():
    sys.path.insert(0, ROOT)
    import numpy as np

    import bench

    cfg = dict(bench.CONFIGS["tiny"], m=30000, n=16)
    full = np.empty((30000, 16), np.float32)
    nz = bench.fill_rows(cfg, 0, 30000, full)
    for r0, r1 in ((0, 12500), (12500, 30000), (7000, 26000)):
        part = np.empty((r1 - r0, 16), np.float32)
        nzp = bench.fill_rows(cfg, r0, r1, part)
        assert np.array_equal(part, full[r0:r1]) and np.array_equal(nzp, nz[r0:r1])


def test_product_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        return   # covered by the GPU suite / the bench itself
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "tiny", "--steps", "3"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)

#!/usr/bin/env python
"""ADMM iterations/sec of the graph-form hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c3|c4]

A "step" is one ADMM iteration (prox of f and g, projection onto y = Ax, dual update,
residuals, stopping rule / adaptive rho) on the BASELINE workload: dense Lasso
100000 x 10000 fp32 (configs[1]) with synthetic data of SURVEY 8d's recipe.  Stopping
tolerances are 0 so exactly W (warm-up) and K (timed) iterations run.

JSON keys: see the contract in the task statement.  In short
  value        K / device time of the K-iteration loop (CUDA events inside the library, A
               resident in HBM, CUDA-graph replay), whole job over all ranks;
  e2e          K / wall time of one PogsS call with HOST buffers (pinned A): H2D of A and
               the descriptors, equilibration, norm estimate, Gram + factor, K iterations,
               D2H of x, y, lambda -- the reference-facing C ABI call a user makes;
  roofline     the dominant kernel (one pass over A), algorithmic bytes m*n*4 per launch
               over its mean CUDA-event duration in a second, event-instrumented loop;
  cpu_baseline the reference CPU path (oracle/_ref, OpenBLAS on all host cores) on a
               bounded row sample of the same workload, scaled to full-size iterations/s.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    # name: (m, n, kind)
    "c2": dict(m=100000, n=10000, kind="lasso", seed=1, label="solve_lasso dense 100000x10000 fp32"),
    "c3": dict(m=50000, n=2000, kind="enet", seed=2, label="elastic-net dense 50000x2000 fp32 (one lambda)"),
    "c4": dict(m=200000, n=5000, kind="logistic", seed=3, label="solve_logistic dense 200000x5000 fp32"),
    "tiny": dict(m=4000, n=500, kind="lasso", seed=1, label="solve_lasso dense 4000x500 fp32 (debug)"),
    "c5": dict(m=1000000, n=100000, nnz_per_row=100, kind="sparse", seed=4,
               label="solve_lasso sparse CSR 1M x 100k, 100 nnz/row, fp32, CGLS projector"),
    "c5s": dict(m=100000, n=10000, nnz_per_row=10, kind="sparse", seed=4,
                label="solve_lasso sparse CSR 100k x 10k, 10 nnz/row, fp32 (scaled-down twin)"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# synthetic data
def make_problem_torch(cfg, device, rows=None, row0=0):
    """A (rows x n, fp32, on `device`), descriptors (numpy) for the local row block.
    Recipe of SURVEY 8d: A~N(0,1), x* 20% dense, b = A x* + 0.1 N, lambda from |A'b|_inf."""
    import torch

    m, n = cfg["m"], cfg["n"]
    rows = m if rows is None else rows
    g = torch.Generator(device=device)
    g.manual_seed(1000 * cfg["seed"] + row0)   # each row block has its own stream
    A = torch.randn((rows, n), generator=g, device=device, dtype=torch.float32)
    gx = torch.Generator(device=device); gx.manual_seed(cfg["seed"])
    xs = torch.randn(n, generator=gx, device=device) * (torch.rand(n, generator=gx, device=device) < 0.2)
    noise = 0.1 * torch.randn(rows, generator=g, device=device)
    Ax = A @ xs + noise
    return A, xs, Ax


def descriptors(cfg, b_or_labels, lam_scale, m_local, n):
    from pogs_b200 import Function, FunctionVector

    kind = cfg["kind"]
    if kind == "lasso":
        f = FunctionVector(m_local, Function.kSquare, 1.0, b_or_labels, 1.0)
        g = FunctionVector(n, Function.kAbs, 1.0, 0.0, 0.1 * lam_scale)
    elif kind == "enet":
        f = FunctionVector(m_local, Function.kSquare, 1.0, b_or_labels, 1.0)
        g = FunctionVector(n, Function.kAbs, 1.0, 0.0, 0.1 * lam_scale, 0.0, 0.05 * lam_scale / 2)
    elif kind == "logistic":
        f = FunctionVector(m_local, Function.kLogistic, -b_or_labels, 0.0, 1.0)
        g = FunctionVector(n, Function.kAbs, 1.0, 0.0, 0.01 * lam_scale)
    else:
        raise ValueError(kind)
    return f, g


def algorithmic_bytes(m, n, s=4):
    """dense direct, m>n: 2*m*n*s + n^2*s + 40*(m+n)*s per iteration."""
    return 2 * m * n * s + n * n * s + 40 * (m + n) * s


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import pogs_b200
    from pogs_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (pogs_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    cfg = CONFIGS[args.config]
    m, n = cfg["m"], cfg["n"]
    K, W = args.steps, max(args.warmup, 3)

    # ---- data, resident in HBM -----------------------------------------------------------------
    # N>1: row blocks (strong scaling of the ONE problem): rank g generates and keeps rows
    # [a_g, b_g) of A; x* and the lambda scale are global (one NCCL all-reduce, data plumbing).
    from pogs_b200.dist import PeerComm, RowBlockSolver, row_partition

    parts = row_partition(m, world)
    r0, r1 = parts[rank]
    A, xs, Ax = make_problem_torch(cfg, device, rows=r1 - r0, row0=r0)
    if cfg["kind"] == "logistic":
        rhs = torch.sign(Ax); rhs[rhs == 0] = 1.0
    else:
        rhs = Ax
    atb = A.t() @ rhs
    if world > 1:
        dist.all_reduce(atb)
    lam_scale = float(atb.abs().max().item())
    f, g = descriptors(cfg, rhs.double().cpu().numpy(), lam_scale, r1 - r0, n)
    torch.cuda.synchronize()

    comm = None
    if world > 1:
        comm = PeerComm(slot_bytes=max(8 * (n + 64), 1 << 22))
        solver = RowBlockSolver(A, m, comm, dtype=np.float32)
    else:
        solver = pogs_b200.Solver(A, dtype=np.float32)
    solver.SetAbsTol(0.0); solver.SetRelTol(0.0); solver.SetAdaptiveRho(True); solver.SetGapStop(True)
    # warm-up: setup (equilibrate, norm estimate, Gram, factor), graph capture, W iterations
    solver.SetMaxIter(W)
    st = solver.Solve(f, g)
    assert st == 3, st
    t_setup = solver.timing()
    setup_ms = t_setup["setup_ms"]
    setup_parts = {k: t_setup[k] for k in ("equil_ms", "normest_ms", "gram_ms", "factor_ms", "h2d_ms")}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: exactly K iterations, graph replay, device-timed ------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    solver.SetMaxIter(K)
    barrier()
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    st = solver.Solve(f, g)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = _lib.launch_count() - l0
    tm = solver.timing()
    assert st == 3 and int(tm["iterations"]) == K, (st, tm)
    loop_ms = tm["loop_ms"]
    single_pass = int(tm.get("single_pass_iterations", 0))
    clocks = sampler.stop() if rank == 0 else None

    # ---- second, event-instrumented loop for the per-kernel numbers ---------------------------------------
    solver.SetProfile(True)
    Kp = min(K, 50)
    solver.SetMaxIter(Kp)
    solver.Solve(f, g)
    pt = solver.timing()
    solver.SetProfile(False)
    npi = max(int(pt["profiled_iterations"]), 1)
    phases = {k: pt[k] / npi for k in ("prox_ms", "gemvt_ms", "solve_ms", "gemv_ms", "ctrl_ms")}
    res = solver.result()
    solver.close()
    # ---- for the record: one converged solve with the reference wrapper's defaults (abs = rel = 1e-4,
    #      adaptive rho, gap stop; SURVEY 8d asks for it next to the fixed-K loop).  Outside every timed region.
    converged = None
    if world == 1 and not args.no_converged:
        try:
            s3 = pogs_b200.Solver(A, dtype=np.float32)
            st3 = s3.Solve(f, g)
            t3, r3 = s3.timing(), s3.result()
            s3.close()
            its = int(t3["iterations"])
            converged = {"status": int(st3), "iterations": its, "exact_residual_iterations": int(t3["exact_iterations"]),
                         "single_pass_iterations": int(t3.get("single_pass_iterations", 0)), "loop_ms": t3["loop_ms"],
                         "iterations_per_s": its / (t3["loop_ms"] * 1e-3) if t3["loop_ms"] > 0 else None,
                         "setup_ms": t3["setup_ms"], "optval": r3["optval"], "nnz_x": int(np.count_nonzero(r3["x"]))}
        except Exception as e:   # never let the side record break the bench line
            converged = {"error": str(e)[:200]}
    del A
    torch.cuda.empty_cache()

    # max over ranks of the device time; whole-job value
    t = torch.tensor([loop_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    loop_ms_max = float(t.item())
    value = K / (loop_ms_max * 1e-3)      # one problem, row-sharded: strong scaling

    # ---- end-to-end with HOST buffers: H2D of A + setup + K iterations + D2H, per call --------------------
    e2e = None
    if not args.no_e2e:
        rows = r1 - r0
        A_host = torch.empty((rows, n), dtype=torch.float32, pin_memory=True)
        Ad, _, _ = make_problem_torch(cfg, device, rows=rows, row0=r0)
        A_host.copy_(Ad); del Ad
        torch.cuda.synchronize(); torch.cuda.empty_cache()
        fa, ga = f.arrays(np.float32), g.arrays(np.float32)
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            # the reference-facing C ABI call a user makes: PogsS with host pointers
            x = np.zeros(n, np.float32); y = np.zeros(m, np.float32); l = np.zeros(m, np.float32)
            ov = ctypes.c_float(); it = ctypes.c_uint()
            ct = ctypes.c_float
            P = lambda arrs: [_lib.ptr(v, ct) for v in arrs[:5]] + [_lib.ptr(arrs[5], ctypes.c_int)]
            Ap = ctypes.cast(ctypes.c_void_p(A_host.data_ptr()), ctypes.POINTER(ct))
            st = _lib.lib.PogsS(1, m, n, Ap, *P(fa), *P(ga), ct(1.0), ct(0.0), ct(0.0), K, 0, 1, 1,
                                _lib.ptr(x, ct), _lib.ptr(y, ct), _lib.ptr(l, ct), ctypes.byref(ov), ctypes.byref(it))
            assert st == 3 and it.value == K - 1, (st, it.value)
            note = "one PogsS call, pinned host A: H2D + setup + K iterations + D2H"
        else:
            # row-block handle API with a host block per rank (the C ABI one-shot is single-GPU)
            s2 = RowBlockSolver(A_host.numpy(), m, comm, dtype=np.float32)
            s2.SetAbsTol(0.0); s2.SetRelTol(0.0); s2.SetMaxIter(K)
            st = s2.Solve(f, g)
            r2 = s2.result()
            s2.close()
            assert st == 3 and r2["iterations"] == K - 1
            note = "RowBlockSolver(host block) + Solve + result per rank: H2D + setup + K iterations + D2H"
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = (m * n * 4 + 6 * 4 * (m + n * world)) / K
        d2h = (n * world + 2 * m) * 4 / K
        e2e = {"value": K / float(te.item()), "unit": "iterations/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "call_s": float(te.item()), "note": note}
        del A_host
    if comm is not None:
        comm.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------
    peak, peak_src = measured_peaks()
    m_loc = r1 - r0
    pass_bytes = m_loc * n * 4      # per GPU
    dom = "gemvt" if phases["gemvt_ms"] >= phases["gemv_ms"] else "gemv"
    dom_ms = phases[dom + "_ms"]
    achieved = pass_bytes / (dom_ms * 1e-3) / 1e9
    other = "gemv" if dom == "gemvt" else "gemvt"
    dom_kernel = "k_colacc" if dom == "gemvt" else ("k_fused_pass" if single_pass else "k_rowdot")
    traffic, traffic_src = None, None
    try:   # measured DRAM bytes per launch of the same kernel on the same workload (ncu --set full)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if world == 1 and args.config in tr and dom_kernel in tr[args.config]:
            traffic = tr[args.config][dom_kernel]["bytes"]; traffic_src = tr[args.config][dom_kernel]["source"]
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "kernel": "k_colacc (A^T t_y)" if dom == "gemvt" else
                  ("k_fused_pass (y = A x, next half-step, A^T t_y' in one pass over A)" if single_pass else "k_rowdot (A x)"),
        "bytes_per_launch": pass_bytes, "ms_per_launch": dom_ms,
        "other_pass": {"kernel": "k_rowdot (A x)" if other == "gemv" else "k_colacc (A^T t_y)",
                       "ms_per_launch": phases[other + "_ms"],
                       "achieved": pass_bytes / (phases[other + "_ms"] * 1e-3) / 1e9},
        "factor_apply": {"kernel": ("k_solve_shard (rows of M sharded over the ranks + fused all-gather)" if world > 1 else
                                    "k_symv_tiles + k_symv_fold (lower triangle of M)" if n >= 4096 else "k_rowdot (M u)"),
                         "ms_per_launch": phases["solve_ms"],
                         "achieved": n * n * 4 / (phases["solve_ms"] * 1e-3) / 1e9,
                         "note": "achieved = n*n*4 B (the full symmetric M) / time; the symmetric kernel reads half of it"},
        "iteration": {"algorithmic_bytes_per_gpu": algorithmic_bytes(m_loc, n), "ms": loop_ms_max / K,
                      "achieved": algorithmic_bytes(m_loc, n) / (loop_ms_max / K * 1e-3) / 1e9,
                      "frac": algorithmic_bytes(m_loc, n) / (loop_ms_max / K * 1e-3) / 1e9 / peak,
                      "frac_of_8TBs": algorithmic_bytes(m_loc, n) / (loop_ms_max / K * 1e-3) / 1e9 / 8000.0},
        "phases_ms": phases,
        "single_pass_iterations": single_pass,
        "note": ("iteration.* uses SURVEY 8d's two-pass algorithmic bytes; %d of %d timed iterations ran on one pass "
                 "over A (committed speculation), so iteration.frac can exceed 1") % (single_pass, K),
    }

    cpu = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_reference(cfg, steps=min(K, 20), sample_rows=args.cpu_rows)

    line = {
        "metric": "ADMM iterations/sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": loop_ms_max / K, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["label"], "m": m, "n": n, "l2_policy": "inputs larger than L2 (A = %.1f GB)" % (m * n * 4 / 1e9),
                   "parallelism": ("row-block x%d (A^T y summed over NVLink peer memory inside the A^T kernel)" % world) if world > 1 else "single", "launch": "cuda-graph replay, 2 iterations per graph",
                   "tolerances": "abs=rel=0 (exactly K iterations), adaptive_rho=1"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "setup_ms": setup_ms, "setup_parts_ms": setup_parts, "wall_ms_timed_solve": wall_ms, "converged_run": converged,
        "sanity": {"optval": res["optval"], "nnz_x": int(np.count_nonzero(res["x"]))},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def _cpu_time_per_iter(cfg, rows, steps, warmup):
    """Seconds per ADMM iteration of the reference CPU path on the first `rows` rows."""
    from oracle import ref_ctypes as R

    m, n = cfg["m"], cfg["n"]
    rng = np.random.default_rng(cfg["seed"])
    A = rng.standard_normal((rows, n), dtype=np.float32)
    xs = (rng.standard_normal(n) * (rng.random(n) < 0.2)).astype(np.float32)
    b = A @ xs + 0.1 * rng.standard_normal(rows).astype(np.float32)
    if cfg["kind"] == "logistic":
        lab = np.sign(b); lab[lab == 0] = 1
        lam = 0.01 * float(np.abs(A.T @ lab).max())
        f = (8, -lab, 0.0, 1.0, 0.0, 0.0); g = (0, 1.0, 0.0, lam, 0.0, 0.0)
    else:
        lam = 0.1 * float(np.abs(A.T @ b).max())
        f = (14, 1.0, b, 1.0, 0.0, 0.0)
        g = (0, 1.0, 0.0, lam, 0.0, 0.05 * lam / 2 if cfg["kind"] == "enet" else 0.0)
    if R.persistent_available():
        kind = "reference"
        s = R.PersistentDense(A, dtype=np.float32)
    else:
        kind = "port"
        from oracle import oracle_ctypes as O

        s = O.Solver(A, dtype=np.float32)
    t0 = time.perf_counter()
    s.solve(f, g, rho=1.0, abs_tol=0.0, rel_tol=0.0, max_iter=warmup)   # init + Cholesky + warm-up
    init_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    r = s.solve(f, g, abs_tol=0.0, rel_tol=0.0, max_iter=steps)
    dt = time.perf_counter() - t0
    s.close()
    assert r["iterations"] == steps - 1
    return dt / steps, init_s, kind


def run_sparse(args):
    """C5: sparse CSR Lasso with the CGLS projector on one GPU.  Reports ADMM iterations/s
    together with k = mean CGLS inner iterations per ADMM iteration and the implied SpMV
    bandwidth (one SpMV pass = nnz*(s+4) + 4*(rows+1) + s*(rows+cols) bytes)."""
    import torch

    import pogs_b200
    from pogs_b200 import Function, FunctionVector, _lib
    import scipy.sparse as sp

    cfg = CONFIGS[args.config]
    m, n, k = cfg["m"], cfg["n"], cfg["nnz_per_row"]
    K, W = args.steps, max(args.warmup, 3)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(cfg["seed"])
    # k distinct sorted columns per row: sorted draws from [0, n-k] plus 0..k-1
    base = torch.randint(0, n - k + 1, (m, k), generator=g, device=dev, dtype=torch.int32)
    base, _ = base.sort(dim=1)
    cols = base + torch.arange(k, device=dev, dtype=torch.int32)[None, :]
    vals = torch.randn((m, k), generator=g, device=dev, dtype=torch.float32)
    xs = torch.randn(n, generator=g, device=dev) * (torch.rand(n, generator=g, device=dev) < 0.2)
    b = (vals * xs[cols.long()]).sum(dim=1) + 0.1 * torch.randn(m, generator=g, device=dev)
    atb = torch.zeros(n, device=dev).index_add_(0, cols.reshape(-1).long(), (vals * b[:, None]).reshape(-1))
    lam = 0.1 * float(atb.abs().max().item())
    indptr = np.arange(0, (m + 1) * k, k, dtype=np.int32)
    A = sp.csr_matrix((vals.reshape(-1).cpu().numpy(), cols.reshape(-1).cpu().numpy(), indptr), shape=(m, n))
    f = FunctionVector(m, Function.kSquare, 1.0, b.double().cpu().numpy(), 1.0)
    gg = FunctionVector(n, Function.kAbs, 1.0, 0.0, lam)
    del base, cols, vals
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    s = pogs_b200.Solver(A, dtype=np.float32)
    s.SetAbsTol(0.0); s.SetRelTol(0.0)
    s.SetMaxIter(W); s.Solve(f, gg)
    setup = s.timing()
    sampler = ClockSampler(0); sampler.start()
    l0 = _lib.launch_count()
    s.SetMaxIter(K); st = s.Solve(f, gg)
    launches = _lib.launch_count() - l0
    tm = s.timing()
    clocks = sampler.stop()
    kbar = tm["cgls_iterations"] / K
    nnz = m * k
    P = nnz * 8 + 4 * (m + 1) + 4 * (m + n)
    bytes_iter = (3 + 2 * kbar) * P + kbar * (6 * n + 5 * m) * 4 + 40 * (m + n) * 4
    ms = tm["loop_ms"] / K
    peak, src = measured_peaks()
    # converged run for the record
    s2 = pogs_b200.Solver(A, dtype=np.float32)
    t1 = time.perf_counter(); s2.Solve(f, gg); conv_s = time.perf_counter() - t1
    r = s2.result(); tc = s2.timing(); s2.close(); s.close()
    line = {"metric": "ADMM iterations/sec", "value": 1e3 / ms, "unit": "iterations/s", "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": cfg["label"], "m": m, "n": n, "nnz": nnz,
                                            "l2_policy": "inputs larger than L2" if nnz * 8 > 126e6 else "inputs fit L2"},
            "clocks": clocks, "gpu_launches": int(launches), "cgls_inner_per_iteration": kbar,
            "roofline": {"bound": "hbm", "achieved": bytes_iter / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bytes_iter / (ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": src,
                         "kernel": "k_spmv_blocked / k_spmv (whole CGLS iteration)", "spmv_pass_bytes": P,
                         "algorithmic_bytes_per_iteration": bytes_iter,
                         "note": "bytes per SURVEY 8d: (3+2k)*P + k*(6n+5m)*s + 40(m+n)s with 8 B per entry; the "
                                 "implementation runs 2+2k products per iteration (start residual from y_prev) on a "
                                 "column-blocked layout with 6 B per entry"},
            "setup_ms": setup["setup_ms"], "setup_parts_ms": {q: setup[q] for q in ("equil_ms", "normest_ms", "h2d_ms")},
            "converged_run": {"status": r["status"], "iterations": r["iterations"] + 1, "wall_s": conv_s,
                              "loop_ms": tc["loop_ms"], "cgls_iterations": tc["cgls_iterations"], "optval": r["optval"]},
            "cpu_baseline": None, "e2e": None}
    print(json.dumps(line))


def cpu_reference(cfg, steps, sample_rows, warmup=2):
    """The reference's own CPU implementation (oracle/_ref: unmodified src/cpu through its
    persistent PogsDirect object, OpenBLAS on all host cores; falls back to the plain-C oracle
    port) on a bounded sample of the workload.  The per-iteration cost of this path is
    t(m) = a + b*m (a: the two n x n triangular solves and BLAS-1 on n; b: two GEMV passes and
    BLAS-1 on m), so it is measured on the first r/2 and the first r rows, with all n columns,
    and extrapolated linearly to the full m."""
    m, n = cfg["m"], cfg["n"]
    r2 = min(sample_rows, m)
    r1 = max(r2 // 2, 1)
    cores = os.cpu_count() or 1
    t1, i1, kind = _cpu_time_per_iter(cfg, r1, steps, warmup)
    t2, i2, kind = _cpu_time_per_iter(cfg, r2, steps, warmup)
    if r2 >= m:
        t_full = t2
    else:
        b = max((t2 - t1) / (r2 - r1), 0.0)
        a = max(t1 - b * r1, 0.0)
        t_full = a + b * m
    return {"value": 1.0 / t_full, "unit": "iterations/s", "cores": cores, "kind": kind,
            "sample": f"first {r1} and first {r2} of {m} rows x {n} cols, {steps} iterations each after {warmup} "
                      f"warm-up: {t1 * 1e3:.1f} and {t2 * 1e3:.1f} ms/iter; per-iteration time extrapolated "
                      f"linearly in the rows to {t_full * 1e3:.1f} ms at m={m}; init+factor {i1:.1f}+{i2:.1f} s",
            "ms_per_iter_samples": [t1 * 1e3, t2 * 1e3], "ms_per_iter_full_extrapolated": t_full * 1e3}


def run_reference(args):
    """--impl reference: the reference's own CPU path timed on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    K, W = args.steps, max(args.warmup, 1)
    cpu = cpu_reference(cfg, steps=K, sample_rows=args.cpu_rows, warmup=W)
    v = cpu["value"]
    line = {
        "impl": "reference", "metric": "ADMM iterations/sec", "value": v, "unit": "iterations/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["label"], "m": cfg["m"], "n": cfg["n"],
                   "note": "reference src/cpu (PogsDirect, fp32) on host cores; bounded row sample scaled to full size"},
        "cpu_baseline": cpu,
        "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-rows", type=int, default=20000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-converged", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif CONFIGS[args.config]["kind"] == "sparse":
        run_sparse(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""ADMM iterations/sec of the graph-form hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--config c2|c3|c3path|c4|c5|c5s|tiny]

A "step" is one ADMM iteration (prox of f and g, projection onto y = Ax, dual update,
residuals, stopping rule / adaptive rho) on the BASELINE workload: dense Lasso
100000 x 10000 fp32 (configs[1]) with synthetic data of SURVEY 8d's recipe.  Stopping
tolerances are 0 so exactly W (warm-up) and K (timed) iterations run ("fixed-K, rho frozen":
with eps = 0 rho never moves and every iteration after the first commits its speculation);
`converged` carries the in-use figure next to it: one solve with the reference wrapper's
default tolerances, rho adapting, the exact-residual branch taken when due.

Synthetic data is defined per fixed 12500-row chunk (numpy default_rng([seed, chunk])), so the
matrix does not depend on the number of GPUs, the reference arm solves the same matrix, and
`sanity.optval` must agree across N.

JSON keys: see the contract in the task statement.  In short
  value        K / device time of the K-iteration loop (CUDA events inside the library, A
               resident in HBM, CUDA-graph replay), whole job over all ranks;
  e2e          K / wall time of one PogsS call with HOST buffers (pinned A): H2D of A and
               the descriptors, equilibration, norm estimate, Gram + factor, K iterations,
               D2H of x, y, lambda -- the reference-facing C ABI call a user makes;
  roofline     the dominant kernel (one pass over A), algorithmic bytes m*n*4 per launch
               over its mean CUDA-event duration in a second, event-instrumented loop;
  cpu_baseline the reference CPU path (oracle/_ref, OpenBLAS on all host cores) on a
               bounded row sample of the same workload, scaled to full-size iterations/s;
  sanity.parity  the device path against that CPU run on the SAME sample rows, same protocol
               (W then K iterations, tolerances 0): rel. differences of x, y, optval; the
               line is marked failed if any exceeds 5e-4.
--impl reference times the reference's own CPU implementation on the FULL workload:
  value = K / time of a K-iteration solve on the already initialised persistent object,
  e2e   = K / wall of construct + first solve (copy, equilibration, norm estimate, Gram,
          Cholesky, K iterations) -- what one PogsS call costs on the CPU.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CHUNK = 12500   # rows per seeded chunk of the synthetic matrix (100000 / 8)

CONFIGS = {
    # name: (m, n, kind)
    "c2": dict(m=100000, n=10000, kind="lasso", seed=1, label="solve_lasso dense 100000x10000 fp32"),
    "c3": dict(m=50000, n=2000, kind="enet", seed=2, label="elastic-net dense 50000x2000 fp32 (one lambda)"),
    "c3path": dict(m=50000, n=2000, kind="enet", seed=2, nlambda=100,
                   label="lambda-path warm-start: 100 lambda values, elastic-net dense 50000x2000 fp32"),
    "c4": dict(m=200000, n=5000, kind="logistic", seed=3, label="solve_logistic dense 200000x5000 fp32"),
    "tiny": dict(m=4000, n=500, kind="lasso", seed=1, label="solve_lasso dense 4000x500 fp32 (debug)"),
    "c5": dict(m=1000000, n=100000, nnz_per_row=100, kind="sparse", seed=4,
               label="solve_lasso sparse CSR 1M x 100k, 100 nnz/row, fp32, CGLS projector"),
    "c5s": dict(m=100000, n=10000, nnz_per_row=10, kind="sparse", seed=4,
                label="solve_lasso sparse CSR 100k x 10k, 10 nnz/row, fp32 (scaled-down twin)"),
}
PARITY_TOL = 5e-4


def host_threads():
    return os.cpu_count() or 1


def use_all_host_threads():
    """The CPU legs must see all host cores: torchrun exports OMP_NUM_THREADS=1 to its workers,
    which made the reference arm single-threaded at N>1.  Must run before the
    reference library (and its OpenBLAS) is loaded."""
    n = str(host_threads())
    os.environ["OMP_NUM_THREADS"] = n
    os.environ["OPENBLAS_NUM_THREADS"] = n


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

    def start(self, settle=0.6):
        """settle: seconds to let nvidia-smi finish its own start-up (it takes driver locks that
        slow a host-driven loop) before the caller starts timing."""
        self._start()
        if self.proc is not None and settle > 0:
            time.sleep(settle)

    def _start(self):
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
# synthetic data (SURVEY 8d recipe), defined per chunk of CHUNK rows so that it is the same matrix
# for every N, for the device arm and for the CPU arms
def x_star(cfg):
    rng = np.random.default_rng([cfg["seed"], 1 << 20])
    n = cfg["n"]
    return (rng.standard_normal(n) * (rng.random(n) < 0.2)).astype(np.float32)


def fill_rows(cfg, r0, r1, out):
    """Rows [r0, r1) of A (fp32) into out[(r1-r0) x n]; returns the matching noise vector.
    Chunks are generated on a thread pool (numpy's generators release the GIL)."""
    n = cfg["n"]
    noise = np.empty(r1 - r0, np.float32)
    jobs = []
    c = r0 // CHUNK
    while c * CHUNK < r1:
        jobs.append(c); c += 1

    def one(c):
        a, b = c * CHUNK, min((c + 1) * CHUNK, cfg["m"])
        rng = np.random.default_rng([cfg["seed"], c])
        lo, hi = max(a, r0), min(b, r1)
        if lo == a and hi == b:
            rng.standard_normal(out=out[a - r0:b - r0], dtype=np.float32)
            nz = rng.standard_normal(b - a, dtype=np.float32)
        else:   # partial chunk: generate whole, copy the slice
            blk = rng.standard_normal((b - a, n), dtype=np.float32)
            nz = rng.standard_normal(b - a, dtype=np.float32)
            out[lo - r0:hi - r0] = blk[lo - a:hi - a]
            nz = nz[lo - a:hi - a]
        noise[lo - r0:hi - r0] = 0.1 * nz

    with ThreadPoolExecutor(max_workers=min(len(jobs), max(1, host_threads()))) as ex:
        list(ex.map(one, jobs))
    return noise


def rhs_of(cfg, A_rows, noise, xs):
    """b (regression) or labels (classification) for a block of rows."""
    v = A_rows @ xs + noise
    if cfg["kind"] == "logistic":
        v = np.sign(v); v[v == 0] = 1.0
    return v.astype(np.float32)


def descriptor_tuples(cfg, rhs, lam_scale):
    """(h, a, b, c, d, e) tuples for f (rows) and g (columns): the canonical encodings of the
    reference's Python wrappers (python/pogs/graph.py:455-562)."""
    kind = cfg["kind"]
    rhs = np.asarray(rhs, np.float64)
    if kind == "lasso":
        return (14, 1.0, rhs, 1.0, 0.0, 0.0), (0, 1.0, 0.0, 0.1 * lam_scale, 0.0, 0.0)
    if kind == "enet":
        return (14, 1.0, rhs, 1.0, 0.0, 0.0), (0, 1.0, 0.0, 0.1 * lam_scale, 0.0, 0.05 * lam_scale / 2)
    if kind == "logistic":
        return (8, -rhs, 0.0, 1.0, 0.0, 0.0), (0, 1.0, 0.0, 0.01 * lam_scale, 0.0, 0.0)
    raise ValueError(kind)


def function_vectors(ft, gt, m_local, n):
    from pogs_b200 import FunctionVector

    return FunctionVector(m_local, *ft), FunctionVector(n, *gt)


def algorithmic_bytes(m, n, s=4):
    """dense direct, m>n: 2*m*n*s + n^2*s + 40*(m+n)*s per iteration."""
    return 2 * m * n * s + n * n * s + 40 * (m + n) * s


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def config_dict(cfg, world=1):
    return {"workload": cfg["label"], "m": cfg["m"], "n": cfg["n"],
            "data": "numpy default_rng([seed, chunk]) per %d-row chunk, seed %d" % (CHUNK, cfg["seed"]),
            "tolerances": "abs=rel=0 (exactly K iterations; rho frozen), adaptive_rho=1, gap_stop=1"}


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

    # ---- data: pinned host copy (the e2e call reads it), resident device copy (the timed loop) ---------
    # N>1: row blocks (strong scaling of the ONE problem): rank g keeps rows [a_g, b_g) of A;
    # x* and the lambda scale are global (one NCCL all-reduce, data plumbing).
    from pogs_b200.dist import PeerComm, RowBlockSolver, row_partition

    parts = row_partition(m, world)
    r0, r1 = parts[rank]
    rows = r1 - r0
    A_host = torch.empty((rows, n), dtype=torch.float32, pin_memory=True)
    noise = fill_rows(cfg, r0, r1, A_host.numpy())
    xs = x_star(cfg)
    rhs = rhs_of(cfg, A_host.numpy(), noise, xs)
    A = A_host.to(device, non_blocking=False)
    atb = A.t() @ torch.from_numpy(rhs).to(device)
    if world > 1:
        dist.all_reduce(atb)
    lam_scale = float(atb.abs().max().item())
    ft, gt = descriptor_tuples(cfg, rhs, lam_scale)
    f, g = function_vectors(ft, gt, rows, n)
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
    res = solver.result()

    # ---- second, event-instrumented loop for the per-kernel numbers ---------------------------------------
    solver.SetProfile(True)
    Kp = min(K, 50)
    solver.SetMaxIter(Kp)
    solver.Solve(f, g)
    pt = solver.timing()
    solver.SetProfile(False)
    npi = max(int(pt["profiled_iterations"]), 1)
    phases = {k: pt[k] / npi for k in ("prox_ms", "gemvt_ms", "solve_ms", "gemv_ms", "ctrl_ms")}
    solver.close()

    # ---- the in-use figure: one converged solve with the reference wrapper's defaults (abs = rel = 1e-4,
    #      adaptive rho, gap stop) on a fresh solver; device-timed loop; outside the fixed-K region ----------
    converged = None
    if not args.no_converged:
        try:
            s3 = RowBlockSolver(A, m, comm, dtype=np.float32) if world > 1 else pogs_b200.Solver(A, dtype=np.float32)
            st3 = s3.Solve(f, g)
            t3, r3 = s3.timing(), s3.result()
            s3.close()
            its = int(t3["iterations"])
            lm = torch.tensor([t3["loop_ms"]], device=device, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(lm, op=dist.ReduceOp.MAX)
            lm = float(lm.item())
            converged = {"value": its / (lm * 1e-3) if lm > 0 else None, "unit": "iterations/s",
                         "status": int(st3), "iterations": its, "exact_residual_iterations": int(t3["exact_iterations"]),
                         "single_pass_iterations": int(t3.get("single_pass_iterations", 0)),
                         "predicted_rho_hits": int(t3.get("predicted_rho_hits", 0)), "loop_ms": lm,
                         "setup_ms": t3["setup_ms"], "optval": r3["optval"], "nnz_x": int(np.count_nonzero(r3["x"])),
                         "tolerances": "abs=rel=1e-4, adaptive_rho=1, gap_stop=1 (python/pogs/graph.py defaults)"}
        except Exception as e:   # never let the side record break the bench line
            converged = {"error": str(e)[:200]}
    # ---- sanity record on the reference arm's protocol (fresh solver, K then K iterations at tol 0): the
    #      reference arm prints the same three numbers for the same matrix; also independent of N ----------------
    kk = None
    try:
        s4 = RowBlockSolver(A, m, comm, dtype=np.float32) if world > 1 else pogs_b200.Solver(A, dtype=np.float32)
        s4.SetAbsTol(0.0); s4.SetRelTol(0.0); s4.SetMaxIter(K)
        s4.Solve(f, g); s4.Solve(f, g)
        r4 = s4.gather_result(parts) if world > 1 else s4.result()
        s4.close()
        kk = {"optval": r4["optval"], "x_norm": float(np.linalg.norm(r4["x"].astype(np.float64))),
              "y_norm": float(np.linalg.norm(r4["y"].astype(np.float64))), "nnz_x": int(np.count_nonzero(r4["x"]))}
    except Exception as e:
        kk = {"error": str(e)[:200]}
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
    if comm is not None:
        comm.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on a bounded sample + parity of the device path on the same sample ------------------
    cpu, parity = None, None
    if world == 1 and not args.no_cpu:
        use_all_host_threads()
        rs = min(args.cpu_rows, m)
        A_s = A_host.numpy()[:rs]
        rhs_s = rhs[:rs]
        lam_s = float(np.abs(A_s.T @ rhs_s).max())
        fts, gts = descriptor_tuples(cfg, rhs_s, lam_s)
        Kc = min(K, 20)
        cpu, ref_res = cpu_reference(cfg, A_s, fts, gts, steps=Kc, warmup=2)
        try:
            fs, gs = function_vectors(fts, gts, rs, n)
            sp = pogs_b200.Solver(A_s, dtype=np.float32)
            sp.SetAbsTol(0.0); sp.SetRelTol(0.0)
            sp.SetMaxIter(2); sp.Solve(fs, gs)
            sp.SetMaxIter(Kc); sp.Solve(fs, gs)
            rp, tp = sp.result(), sp.timing()
            sp.close()
            parity = {"rel_dx": relerr(rp["x"], ref_res["x"]), "rel_dy": relerr(rp["y"], ref_res["y"]),
                      "rel_doptval": abs(rp["optval"] - ref_res["optval"]) / abs(ref_res["optval"]),
                      "single_pass_iterations": int(tp["single_pass_iterations"]),
                      "protocol": f"first {rs} rows x {n} cols of the workload, 2 then {Kc} iterations at tol 0 on a "
                                  f"persistent solver, device fp32 vs {cpu['kind']} CPU path fp32 on the same numpy data",
                      "tolerance": PARITY_TOL}
            parity["ok"] = bool(max(parity["rel_dx"], parity["rel_dy"], parity["rel_doptval"]) <= PARITY_TOL)
        except Exception as e:
            parity = {"ok": False, "error": str(e)[:200]}
    del A_host

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------
    peak, peak_src = measured_peaks()
    m_loc = rows
    one_launch = bool(pt.get("one_launch", 0))
    pass_bytes = m_loc * n * 4      # one pass over the local rows of A
    tri_bytes = n * n * 2 // world  # packed lower triangle of the factor, dealt over the ranks
    dom = "gemvt" if phases["gemvt_ms"] >= phases["gemv_ms"] else "gemv"
    dom_ms = phases[dom + "_ms"]
    other = "gemv" if dom == "gemvt" else "gemvt"
    other_ms = phases[other + "_ms"]
    traffic, traffic_src = None, None
    if one_launch:
        # the whole committed iteration is ONE launch: algorithmic bytes = one pass over A + the packed factor
        # + the state vectors and descriptors (SURVEY 8d's 40 words per element); duration = CUDA events
        # around the launch in the event-instrumented loop (plain launches, one iteration each)
        dom_kernel = "k_admm_pass"
        launch_bytes = pass_bytes + tri_bytes + 40 * (m_loc + n) * 4
        achieved = launch_bytes / (dom_ms * 1e-3) / 1e9
        kernel_label = ("k_admm_pass: the whole committed iteration in one launch (pass over A with y = A x, the next "
                        "half-step and A^T t_y'; fold; controller; streamed packed-triangle factor apply; x half-step)")
    else:
        dom_kernel = "k_colacc" if dom == "gemvt" else ("k_fused_pass" if single_pass else "k_rowdot")
        launch_bytes = pass_bytes
        achieved = launch_bytes / (dom_ms * 1e-3) / 1e9
        kernel_label = ("k_colacc (A^T t_y)" if dom == "gemvt" else
                        ("k_fused_pass (y = A x, next half-step, A^T t_y' in one pass over A)" if single_pass else "k_rowdot (A x)"))
    try:   # measured DRAM bytes per launch of the same kernel on the same workload (ncu --set full)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if world == 1 and args.config in tr and dom_kernel in tr[args.config]:
            traffic = tr[args.config][dom_kernel]["bytes"]; traffic_src = tr[args.config][dom_kernel]["source"]
    except Exception:
        pass
    moved = m_loc * n * 4 * (1 if single_pass else 2) + tri_bytes + 40 * (m_loc + n) * 4
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "kernel": kernel_label,
        "bytes_per_launch": launch_bytes, "ms_per_launch": dom_ms,
        "pass_over_A": ({"bytes": pass_bytes, "us": (tm.get("pass_phase_us") or [0])[0],
                         "achieved": pass_bytes / ((tm.get("pass_phase_us") or [0])[0] * 1e-6) / 1e9
                         if (tm.get("pass_phase_us") or [0])[0] > 0 else None,
                         "note": "phase A of the launch alone (needs POGS_B200_PASS_TIMING=1)"} if one_launch else None),
        "iteration": {"algorithmic_bytes_per_gpu": algorithmic_bytes(m_loc, n), "ms": loop_ms_max / K,
                      "achieved": algorithmic_bytes(m_loc, n) / (loop_ms_max / K * 1e-3) / 1e9,
                      "frac": algorithmic_bytes(m_loc, n) / (loop_ms_max / K * 1e-3) / 1e9 / peak,
                      "frac_of_8TBs": algorithmic_bytes(m_loc, n) / (loop_ms_max / K * 1e-3) / 1e9 / 8000.0,
                      "bytes_moved_per_gpu": moved,
                      "frac_of_bytes_moved": moved / (loop_ms_max / K * 1e-3) / 1e9 / peak},
        "phases_ms": phases,
        "phases_note": ("event-instrumented loop, plain launches: gemv_ms = the one-launch iteration kernel, solve_ms = the "
                        "gated service kernels in front of it (they return at their gate on committed iterations)") if one_launch
                       else "event-instrumented loop: prox, A^T pass, factor apply, A pass, controller",
        "pass_phase_us": tm.get("pass_phase_us"),
        "pass_phase_note": "with POGS_B200_PASS_TIMING=1: mean us per iteration of the phases of the one-launch iteration "
                           "kernel on CTA 0 (A pass, barrier, fold B, barrier, controller, factor apply D, barrier, fold E)",
        "single_pass_iterations": single_pass,
        "note": ("iteration.* uses SURVEY 8d's two-pass algorithmic bytes (2 m n s + n^2 s + 40 (m+n) s); %d of %d timed "
                 "iterations ran on one pass over A (committed speculation) and the factor is applied from its packed "
                 "lower triangle, so iteration.frac can exceed 1; frac_of_bytes_moved counts what the implementation "
                 "really streams") % (single_pass, K),
    }

    c = config_dict(cfg, world)
    c.update({"l2_policy": "inputs larger than L2 (A = %.1f GB)" % (m * n * 4 / 1e9),
              "parallelism": ("row-block x%d (A^T y summed over NVLink peer memory inside the pass)" % world) if world > 1 else "single",
              "launch": "cuda-graph replay"})
    line = {
        "metric": "ADMM iterations/sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": loop_ms_max / K, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": c,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "converged": converged,
        "setup_ms": setup_ms, "setup_parts_ms": setup_parts, "wall_ms_timed_solve": wall_ms,
        "sanity": {"optval": res["optval"], "nnz_x": int(np.count_nonzero(res["x"])),
                   "x_norm": float(np.linalg.norm(res["x"].astype(np.float64))),
                   "note": "fixed-K state after W then K iterations; independent of the number of GPUs up to rounding",
                   "k_then_k": kk, "k_then_k_note": "fresh solver, K then K iterations at tol 0: the protocol of "
                   "--impl reference, whose sanity block must show the same optval / norms for the same matrix",
                   "parity": parity},
    }
    if parity is not None and not parity.get("ok", False):
        line["failed"] = "device path differs from the CPU reference on the sample by more than %g" % PARITY_TOL
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def _ref_solver(A, sparse=False):
    """The reference's own persistent solver object (oracle/_ref), else the plain-C port."""
    from oracle import ref_ctypes as R

    if R.persistent_available() and not sparse:
        return R.PersistentDense(A, dtype=np.float32), "reference"
    from oracle import oracle_ctypes as O

    return O.Solver(A, dtype=np.float32), "port"


def _cpu_iter_time(A, ft, gt, steps, warmup):
    """(seconds per iteration, init seconds, kind, result of the timed solve) of the CPU path."""
    s, kind = _ref_solver(A)
    t0 = time.perf_counter()
    s.solve(ft, gt, rho=1.0, abs_tol=0.0, rel_tol=0.0, max_iter=warmup)   # init + Cholesky + warm-up
    init_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    r = s.solve(ft, gt, abs_tol=0.0, rel_tol=0.0, max_iter=steps)
    dt = time.perf_counter() - t0
    s.close()
    assert r["iterations"] == steps - 1
    return dt / steps, init_s, kind, r


def cpu_reference(cfg, A_s, ft, gt, steps, warmup=2):
    """The reference's own CPU implementation (oracle/_ref: unmodified src/cpu through its
    persistent PogsDirect object, OpenBLAS on all host cores; falls back to the plain-C oracle
    port) on a bounded sample of the workload: the first r rows, all n columns.  The
    per-iteration cost of this path is t(m) = a + b*m (a: the two n x n triangular solves and
    BLAS-1 on n; b: two GEMV passes and BLAS-1 on m), so it is measured on the first r/2 and
    the first r rows and extrapolated linearly to the full m.  Returns (record, result of the
    r-row run) -- the latter is what sanity.parity compares the device path with."""
    m, n = cfg["m"], cfg["n"]
    r2 = A_s.shape[0]
    r1 = max(r2 // 2, 1)
    h, a, b, c, d, e = ft
    pick = lambda v, r: v[:r] if isinstance(v, np.ndarray) else v
    ft1 = (h, pick(a, r1), pick(b, r1), c, d, e)
    t1, i1, kind, _ = _cpu_iter_time(A_s[:r1], ft1, gt, steps, warmup)
    t2, i2, kind, res = _cpu_iter_time(A_s, ft, gt, steps, warmup)
    if r2 >= m:
        t_full = t2
    else:
        bb = max((t2 - t1) / (r2 - r1), 0.0)
        aa = max(t1 - bb * r1, 0.0)
        t_full = aa + bb * m
    rec = {"value": 1.0 / t_full, "unit": "iterations/s", "cores": host_threads(), "kind": kind,
           "sample": f"first {r1} and first {r2} of {m} rows x {n} cols, {steps} iterations each after {warmup} "
                     f"warm-up: {t1 * 1e3:.1f} and {t2 * 1e3:.1f} ms/iter; per-iteration time extrapolated "
                     f"linearly in the rows to {t_full * 1e3:.1f} ms at m={m}; init+factor {i1:.1f}+{i2:.1f} s; "
                     f"the full-size measurement is bench.py --impl reference",
           "ms_per_iter_samples": [t1 * 1e3, t2 * 1e3], "ms_per_iter_full_extrapolated": t_full * 1e3}
    return rec, res


def host_problem(cfg, rows=None):
    """Full (or first-`rows`) workload on the host: A, f, g tuples."""
    m, n = cfg["m"], cfg["n"]
    rows = m if rows is None else min(rows, m)
    A = np.empty((rows, n), np.float32)
    noise = fill_rows(cfg, 0, rows, A)
    rhs = rhs_of(cfg, A, noise, x_star(cfg))
    # |A^T rhs|_inf in chunks (keeps the temporary small)
    atb = np.zeros(n, np.float64)
    for a in range(0, rows, CHUNK):
        atb += A[a:a + CHUNK].T @ rhs[a:a + CHUNK]
    lam = float(np.abs(atb).max())
    ft, gt = descriptor_tuples(cfg, rhs, lam)
    return A, ft, gt, lam


def run_reference(args):
    """--impl reference: the reference's own CPU path on this box's host cores, full workload."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    use_all_host_threads()
    cfg = CONFIGS[args.config]
    if cfg["kind"] == "sparse":
        return run_reference_sparse(args, cfg, world)
    if args.config == "c3path":
        return run_path(args, reference_only=True)
    K, W = args.steps, max(args.warmup, 1)
    rows = cfg["m"] if args.ref_rows <= 0 else min(args.ref_rows, cfg["m"])
    t0 = time.perf_counter()
    A, ft, gt, _ = host_problem(cfg, rows)
    gen_s = time.perf_counter() - t0
    # one-shot cost: construct + first solve = copy of A, equilibration, norm estimate, Gram, Cholesky, K iterations
    t0 = time.perf_counter()
    s, kind = _ref_solver(A)
    r1 = s.solve(ft, gt, rho=1.0, abs_tol=0.0, rel_tol=0.0, max_iter=K)
    first_s = time.perf_counter() - t0
    # loop only: K more iterations on the initialised object (the first solve is this arm's warm-up)
    t0 = time.perf_counter()
    r2 = s.solve(ft, gt, abs_tol=0.0, rel_tol=0.0, max_iter=K)
    loop_s = time.perf_counter() - t0
    s.close()
    assert r1["iterations"] == K - 1 and r2["iterations"] == K - 1
    v = K / loop_s
    full = rows == cfg["m"]
    c = config_dict(cfg)
    if not full:
        c["rows_used"] = rows
    cpu = {"value": v, "unit": "iterations/s", "cores": host_threads(), "kind": kind,
           "sample": (f"{'full workload' if full else 'first %d rows' % rows}: {rows} x {cfg['n']} fp32; first solve "
                      f"(init + factor + {K} iterations) {first_s:.1f} s, then {K} timed iterations {loop_s:.2f} s; "
                      f"data generation {gen_s:.1f} s not counted")}
    line = {
        "impl": "reference", "metric": "ADMM iterations/sec", "value": v, "unit": "iterations/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * loop_s / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": c,
        "cpu_baseline": cpu,
        "e2e": {"value": K / first_s, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "call_s": first_s, "note": "construct + first Solve on the host: copy of A, equilibration, norm estimate, "
                                           "Gram matrix, Cholesky, K iterations (what one PogsS call costs)"},
        "gpu_launches": 0,
        "sanity": {"k_then_k": {"optval": r2["optval"], "x_norm": float(np.linalg.norm(r2["x"].astype(np.float64))),
                                "y_norm": float(np.linalg.norm(r2["y"].astype(np.float64))),
                                "nnz_x": int(np.count_nonzero(r2["x"]))},
                   "optval": r2["optval"],
                   "note": "state after K then K iterations: compare with sanity.k_then_k of the device arm"},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# C3: the warm-started regularisation path (examples/cpp/lasso_path.cpp:75-107 protocol)
def path_lambdas(lmax, count):
    """log-spaced from lambda_max down to 1e-2 * lambda_max (NOT the reference example's
    formula, which ends at lambda_max^0.01)."""
    return lmax * np.logspace(0.0, -2.0, count)


def run_path(args, reference_only=False):
    cfg = CONFIGS["c3path"]
    m, n, nl = cfg["m"], cfg["n"], cfg["nlambda"]
    use_all_host_threads()
    A, ft, gt, lmax = host_problem(cfg)
    lams = path_lambdas(lmax, nl)
    lam2 = 0.05 * lmax

    cfgd = dict(config_dict(cfg), tolerances="abs=rel=1e-4, adaptive_rho=1, gap_stop=1 (wrapper defaults)",
                lambdas="100 log-spaced from |A'b|_inf to 1e-2 of it; lambda2 = 0.05 |A'b|_inf",
                l2_policy="inputs larger than L2 (A = 0.4 GB)")

    def g_of(lam):
        return (0, 1.0, 0.0, float(lam), 0.0, lam2 / 2)

    def cpu_path():
        s, kind = _ref_solver(A)
        t0 = time.perf_counter()
        its, ovs = [], []
        first_s = None
        for i, lam in enumerate(lams):
            r = s.solve(ft, g_of(lam), rho=1.0 if i == 0 else None)
            its.append(r["iterations"] + 1); ovs.append(r["optval"])
            if i == 0:
                first_s = time.perf_counter() - t0
        wall = time.perf_counter() - t0
        s.close()
        return {"kind": kind, "wall_s": wall, "first_solve_s": first_s, "iterations": its, "optval": ovs,
                "x_last": r["x"]}

    if reference_only:
        cp = cpu_path()
        tot = int(sum(cp["iterations"]))
        v = tot / cp["wall_s"]
        line = {"impl": "reference", "metric": "ADMM iterations/sec", "value": v, "unit": "iterations/s", "n_gpus": 1,
                "steps": tot, "warmup": 0, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfgd,
                "cpu_baseline": {"value": v, "unit": "iterations/s", "cores": host_threads(), "kind": cp["kind"],
                                 "sample": "the full 100-lambda path on the full matrix"},
                "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "path": {"total_wall_s": cp["wall_s"], "total_iterations": tot, "first_solve_s": cp["first_solve_s"]},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch

    import pogs_b200
    from pogs_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (pogs_b200 has no CPU fallback)")
    A_host = torch.from_numpy(A).pin_memory()
    f, _ = function_vectors(ft, gt, m, n)
    from pogs_b200 import FunctionVector

    def device_path(src):
        """src: pinned host array (e2e: upload + setup inside) or device tensor (resident)."""
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        its, ovs, loop_ms, sp = [], [], 0.0, 0
        s = pogs_b200.Solver(src, dtype=np.float32)
        for lam in lams:
            st = s.Solve(f, FunctionVector(n, *g_of(lam)))
            assert st == 0, st
            t = s.timing()
            its.append(int(t["iterations"])); loop_ms += t["loop_ms"]; sp += int(t["single_pass_iterations"])
            ovs.append(s.GetOptval())
        x_last = s.GetX()
        s.close()
        torch.cuda.synchronize()
        return {"wall_s": time.perf_counter() - t0, "iterations": its, "optval": ovs, "loop_ms": loop_ms,
                "x_last": x_last, "single_pass_iterations": sp}

    A_dev = A_host.to("cuda")
    device_path(A_dev)                                   # warm-up: library init, graph capture, pool growth
    sampler = ClockSampler(0); sampler.start()
    l0 = _lib.launch_count()
    dp = device_path(A_dev)                              # timed: resident A
    launches = _lib.launch_count() - l0
    clocks = sampler.stop()
    de = device_path(A_host.numpy())                     # end to end: host A in, per-lambda x / optval out
    cp = None if args.no_cpu else cpu_path()
    tot = int(sum(dp["iterations"]))
    value = tot / (dp["loop_ms"] * 1e-3)
    peak, src = measured_peaks()
    bytes_iter = algorithmic_bytes(m, n)
    parity = None
    if cp is not None:
        ov_d, ov_c = np.array(dp["optval"]), np.array(cp["optval"])
        parity = {"max_rel_doptval_over_path": float(np.max(np.abs(ov_d - ov_c) / np.abs(ov_c))),
                  "rel_dx_last": relerr(dp["x_last"], cp["x_last"]),
                  "iterations_device": tot, "iterations_cpu": int(sum(cp["iterations"])),
                  "tolerance": "both sides stop at abs=rel=1e-4: 5e-4 on optval, 5e-3 on x"}
        parity["ok"] = bool(parity["max_rel_doptval_over_path"] < 5e-4 and parity["rel_dx_last"] < 5e-3)
    line = {"metric": "ADMM iterations/sec", "value": value, "unit": "iterations/s", "n_gpus": 1, "steps": tot, "warmup": tot,
            "ms_per_step": dp["loop_ms"] / tot, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": cfgd,
            "clocks": clocks, "gpu_launches": int(launches),
            "path": {"total_wall_s_resident": dp["wall_s"], "total_wall_s_e2e": de["wall_s"], "total_iterations": tot,
                     "device_loop_s": dp["loop_ms"] * 1e-3, "iterations_first_lambda": dp["iterations"][0],
                     "iterations_max_after_first": int(max(dp["iterations"][1:])),
                     "single_pass_iterations": dp["single_pass_iterations"],
                     "cpu_total_wall_s": None if cp is None else cp["wall_s"],
                     "cpu_total_iterations": None if cp is None else int(sum(cp["iterations"])),
                     "speedup_wall_e2e": None if cp is None else cp["wall_s"] / de["wall_s"]},
            "e2e": {"value": int(sum(de["iterations"])) / de["wall_s"], "unit": "iterations/s",
                    "h2d_bytes_per_step": (m * n * 4 + nl * 6 * 4 * (m + n)) / max(int(sum(de["iterations"])), 1),
                    "d2h_bytes_per_step": nl * (n + 2 * m) * 4 / max(int(sum(de["iterations"])), 1), "call_s": de["wall_s"],
                    "note": "Solver(host A) + 100 x Solve + results: upload, setup and every per-lambda call included"},
            "roofline": {"bound": "hbm", "achieved": bytes_iter * value / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bytes_iter * value / 1e9 / peak, "traffic": None, "peak_source": src,
                         "kernel": "whole iteration (k_rowdot + k_colacc + factor apply; rows of 8 KB use the two-pass kernels)",
                         "bytes_per_iteration": bytes_iter,
                         "note": "A (0.4 GB) exceeds L2 but short solves (5-10 iterations per lambda) are dominated by "
                                 "per-solve costs: descriptor upload, graph launch, result copies"},
            "cpu_baseline": None if cp is None else {"value": int(sum(cp["iterations"])) / cp["wall_s"], "unit": "iterations/s",
                                                     "cores": host_threads(), "kind": cp["kind"],
                                                     "sample": "the full 100-lambda path on the full matrix, persistent "
                                                               "PogsDirect object (init + factor inside the first solve)"},
            "sanity": {"optval_first": dp["optval"][0], "optval_last": dp["optval"][-1], "parity": parity}}
    if parity is not None and not parity["ok"]:
        line["failed"] = "device path differs from the CPU reference along the path"
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def sparse_problem(cfg, rows=None):
    """C5 recipe on the host: CSR with exactly k sorted distinct columns per row."""
    import scipy.sparse as sp

    m, n, k = cfg["m"], cfg["n"], cfg["nnz_per_row"]
    rows = m if rows is None else min(rows, m)
    rng = np.random.default_rng([cfg["seed"], 0])
    base = np.sort(rng.integers(0, n - k + 1, size=(rows, k), dtype=np.int32), axis=1)
    cols = base + np.arange(k, dtype=np.int32)[None, :]
    vals = rng.standard_normal((rows, k), dtype=np.float32)
    xs = x_star(cfg)
    b = (vals * xs[cols]).sum(axis=1) + 0.1 * rng.standard_normal(rows, dtype=np.float32)
    indptr = np.arange(0, (rows + 1) * k, k, dtype=np.int64)
    A = sp.csr_matrix((vals.reshape(-1), cols.reshape(-1), indptr), shape=(rows, n))
    lam = 0.1 * float(np.abs(A.T @ b).max())
    return A, (14, 1.0, b.astype(np.float64), 1.0, 0.0, 0.0), (0, 1.0, 0.0, lam, 0.0, 0.0)


def _cpu_sparse_iter_time(A, ft, gt, steps, warmup):
    from oracle import ref_ctypes as R

    # PogsSparseS is one-shot (no persistent sparse object in the reference's C ABI):
    # differential timing, SURVEY 8d: (t(K2) - t(K1)) / (K2 - K1)
    kind = "reference" if R.available() else "port"
    if kind == "reference":
        run = lambda K: R.solve(A, ft, gt, dtype=np.float32, abs_tol=0.0, rel_tol=0.0, max_iter=K)
    else:
        from oracle import oracle_ctypes as O

        run = lambda K: O.solve(A, ft, gt, dtype=np.float32, abs_tol=0.0, rel_tol=0.0, max_iter=K)
    t0 = time.perf_counter(); run(warmup); t1 = time.perf_counter()
    r = run(warmup + steps); t2 = time.perf_counter()
    return ((t2 - t1) - (t1 - t0)) / steps, t1 - t0, kind, r


def run_reference_sparse(args, cfg, world):
    K = min(args.steps, 10)
    rows = min(cfg["m"], args.cpu_rows * 10)
    A, ft, gt = sparse_problem(cfg, rows)
    t, init, kind, r = _cpu_sparse_iter_time(A, ft, gt, K, 2)
    v = 1.0 / (t * cfg["m"] / rows)
    line = {"impl": "reference", "metric": "ADMM iterations/sec", "value": v, "unit": "iterations/s", "n_gpus": world,
            "steps": K, "warmup": 2, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(cfg),
            "cpu_baseline": {"value": v, "unit": "iterations/s", "cores": host_threads(), "kind": kind,
                             "sample": f"first {rows} of {cfg['m']} rows (all columns), differential timing over {K} "
                                       f"iterations: {t * 1e3:.1f} ms/iter, scaled by the row ratio"},
            "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def measured_traffic(config, kernel):
    """dram read + write bytes per launch of `kernel` from the committed ncu --set full capture (or None)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[config][kernel]["bytes"]
    except Exception:
        return None


def run_sparse(args):
    """C5: sparse CSR Lasso with the CGLS projector on one GPU.  Reports ADMM iterations/s
    together with k = mean CGLS inner iterations per ADMM iteration and the implied SpMV
    bandwidth (one SpMV pass = nnz*(s+4) + 4*(rows+1) + s*(rows+cols) bytes)."""
    import torch

    import pogs_b200
    from pogs_b200 import FunctionVector, _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (pogs_b200 has no CPU fallback)")
    cfg = CONFIGS[args.config]
    m, n, k = cfg["m"], cfg["n"], cfg["nnz_per_row"]
    K, W = args.steps, max(args.warmup, 3)
    A, ft, gt = sparse_problem(cfg)
    f, gg = function_vectors(ft, gt, m, n)
    s = pogs_b200.Solver(A, dtype=np.float32)
    s.SetAbsTol(0.0); s.SetRelTol(0.0)
    s.SetMaxIter(W); s.Solve(f, gg)
    setup = s.timing()
    sampler = ClockSampler(0); sampler.start()
    l0 = _lib.launch_count()
    s.SetMaxIter(K); st = s.Solve(f, gg)
    launches = _lib.launch_count() - l0
    tm = s.timing()
    res = s.result()
    clocks = sampler.stop()
    kbar = tm["cgls_iterations"] / K
    nnz = m * k
    P = nnz * 8 + 4 * (m + 1) + 4 * (m + n)
    bytes_iter = (3 + 2 * kbar) * P + kbar * (6 * n + 5 * m) * 4 + 40 * (m + n) * 4
    # what the implementation moves: 1 + 1/16 + 2k products per iteration on the tiled layout, 6 B per entry
    # (start residual from y_prev; y = A x from the CGLS residual recurrence, refreshed by a product every 16th
    # iteration -- POGS_B200_Y_REC=0: 2 + 2k)
    P6 = nnz * 6 + 4 * (m + 1) + 4 * (m + n)
    y_products = 1.0 if os.environ.get("POGS_B200_Y_REC") == "0" else 1.0 / 16
    moved_iter = (1 + y_products + 2 * kbar) * P6 + kbar * (6 * n + 5 * m) * 4 + 40 * (m + n) * 4
    ms = tm["loop_ms"] / K
    peak, src = measured_peaks()
    # end to end: one PogsSparseS call with host CSR arrays
    fa, ga = f.arrays(np.float32), gg.arrays(np.float32)
    # the three CSR arrays in pinned host memory (the contract's e2e: inputs copied from pinned memory every call)
    pins = [torch.from_numpy(np.ascontiguousarray(A.data, np.float32)), torch.from_numpy(np.ascontiguousarray(A.indices, np.int32)),
            torch.from_numpy(np.ascontiguousarray(A.indptr, np.int32))]
    if os.environ.get("BENCH_SPARSE_PIN", "1") != "0":
        pins = [t.pin_memory() for t in pins]
    data, ind, ptr = (t.numpy() for t in pins)
    ct = ctypes.c_float
    P_ = lambda arrs: [_lib.ptr(v, ct) for v in arrs[:5]] + [_lib.ptr(arrs[5], ctypes.c_int)]
    x = np.zeros(n, np.float32); y = np.zeros(m, np.float32); l = np.zeros(m, np.float32)
    ov = ctypes.c_float(); it = ctypes.c_uint()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = _lib.lib.PogsSparseS(1, m, n, nnz, _lib.ptr(data, ct), _lib.ptr(ptr, ctypes.c_int), _lib.ptr(ind, ctypes.c_int),
                              *P_(fa), *P_(ga), ct(1.0), ct(0.0), ct(0.0), K, 0, 1, 1, _lib.ptr(x, ct), _lib.ptr(y, ct),
                              _lib.ptr(l, ct), ctypes.byref(ov), ctypes.byref(it))
    e2e_s = time.perf_counter() - t0
    assert st == 3 and it.value == K - 1
    # converged run for the record
    s2 = pogs_b200.Solver(A, dtype=np.float32)
    t1 = time.perf_counter(); s2.Solve(f, gg); conv_s = time.perf_counter() - t1
    r = s2.result(); tc = s2.timing(); s2.close(); s.close()
    cpu, parity = None, None
    if not args.no_cpu:
        use_all_host_threads()
        rows = min(m, args.cpu_rows * 10)
        As, fts, gts = sparse_problem(cfg, rows)
        Kc = min(K, 10)
        tci, init, kind, rc = _cpu_sparse_iter_time(As, fts, gts, Kc, 2)
        cpu = {"value": 1.0 / (tci * m / rows), "unit": "iterations/s", "cores": host_threads(), "kind": kind,
               "sample": f"first {rows} of {m} rows (all columns), PogsSparseS, differential timing over {Kc} iterations: "
                         f"{tci * 1e3:.1f} ms/iter on the sample, scaled by the row ratio (cost is linear in nnz)"}
        fs, gs = function_vectors(fts, gts, rows, n)
        sp_ = pogs_b200.Solver(As, dtype=np.float32)
        sp_.SetAbsTol(0.0); sp_.SetRelTol(0.0); sp_.SetMaxIter(2 + Kc); sp_.Solve(fs, gs)
        rp = sp_.result(); sp_.close()
        parity = {"rel_dx": relerr(rp["x"], rc["x"]), "rel_dy": relerr(rp["y"], rc["y"]),
                  "rel_doptval": abs(rp["optval"] - rc["optval"]) / abs(rc["optval"]),
                  "protocol": f"first {rows} rows, {2 + Kc} iterations at tol 0, device fp32 vs {kind} CPU path fp32",
                  "tolerance": 2e-3}
        parity["ok"] = bool(max(parity["rel_dx"], parity["rel_dy"], parity["rel_doptval"]) <= 2e-3)
    line = {"metric": "ADMM iterations/sec", "value": 1e3 / ms, "unit": "iterations/s", "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(config_dict(cfg), nnz=nnz,
                                                l2_policy="inputs larger than L2" if nnz * 8 > 126e6 else "inputs fit L2"),
            "clocks": clocks, "gpu_launches": int(launches), "cgls_inner_per_iteration": kbar,
            "roofline": {"bound": "hbm", "achieved": moved_iter / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": moved_iter / (ms * 1e-3) / 1e9 / peak,
                         "traffic": measured_traffic(args.config, "k_spmv_tiled"), "peak_source": src,
                         "kernel": "k_spmv_tiled + k_tl_fold / k_spmv (whole CGLS iteration; traffic = one product)",
                         "spmv_pass_bytes": P6,
                         "bytes_per_iteration": moved_iter,
                         "survey_bytes_per_iteration": bytes_iter,
                         "survey_achieved": bytes_iter / (ms * 1e-3) / 1e9,
                         "products_per_iteration": 1 + y_products + 2 * kbar,
                         "note": "achieved counts what the implementation moves: 1+1/16+2k products per iteration (start "
                                 "residual from y_prev, y = A x from the residual recurrence) at 6 B per entry on the 2-D tiled layout; SURVEY 8d's "
                                 "figure ((3+2k) products at 8 B per entry) is given beside it"},
            "setup_ms": setup["setup_ms"], "setup_parts_ms": {q: setup[q] for q in ("equil_ms", "normest_ms", "h2d_ms")},
            "converged": {"value": (r["iterations"] + 1) / (tc["loop_ms"] * 1e-3), "unit": "iterations/s",
                          "status": r["status"], "iterations": r["iterations"] + 1, "wall_s": conv_s,
                          "loop_ms": tc["loop_ms"], "cgls_iterations": tc["cgls_iterations"], "optval": r["optval"]},
            "cpu_baseline": cpu,
            "e2e": {"value": K / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": (nnz * 8 + 4 * (m + 1) + 24 * (m + n)) / K,
                    "d2h_bytes_per_step": (n + 2 * m) * 4 / K, "call_s": e2e_s,
                    "note": "one PogsSparseS call with pinned host CSR arrays: upload, CSR->CSC, tiled re-layout, "
                            "equilibration, norm estimate, K iterations, D2H"},
            "sanity": {"optval": res["optval"], "parity": parity}}
    if parity is not None and not parity["ok"]:
        line["failed"] = "device path differs from the CPU reference on the sample"
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-rows", type=int, default=20000, help="rows of the cpu_baseline sample (product arm)")
    ap.add_argument("--ref-rows", type=int, default=0, help="--impl reference: rows to use (0 = the full workload)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-converged", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c3path":
        if int(os.environ.get("RANK", "0")) == 0:
            run_path(args)
    elif CONFIGS[args.config]["kind"] == "sparse":
        run_sparse(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

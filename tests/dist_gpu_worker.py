"""Worker for the multi-GPU parity test: launched by torch.distributed.run, one rank per GPU.
Rank 0 prints one line 'RESULT {json}' with the comparison against the oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import problems
    from oracle import oracle_ctypes as O
    from pogs_b200 import FunctionVector
    from pogs_b200.dist import PeerComm, RowBlockSolver, row_partition, slice_function

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local % torch.cuda.device_count())
    dist.init_process_group("gloo")
    out = {"world": world}
    comm = PeerComm(slot_bytes=1 << 20)
    # 1. the exchange kernel on its own
    for dt in (torch.float32, torch.float64):
        t = torch.arange(10000, device="cuda", dtype=dt) * (rank + 1)
        comm.allreduce_(t)
        want = torch.arange(10000, device="cuda", dtype=dt) * (world * (world + 1) / 2)
        out[f"allreduce_{dt}"] = bool(torch.equal(t, want))
        big = torch.full((3_000_000,), float(rank + 1), device="cuda", dtype=dt)   # several slot-sized pieces
        comm.allreduce_(big)
        out[f"allreduce_big_{dt}"] = bool((big == world * (world + 1) / 2).all().item())
    # 2. row-block solves against the single-process oracle
    cases = ["c1_lasso_500x300", "c2s_lasso_10000x1000", "c4s_logistic_20000x500", "svm_600x200"]
    if os.environ.get("POGS_DIST_CASES"):   # subset for quick checks
        cases = os.environ["POGS_DIST_CASES"].split(",")
    # "<case>+fuse": same case with the single-pass kernel forced on (its fold phase then carries
    # the cross-GPU exchange of A^T t_y; the BASELINE shape takes that path by default)
    cases = cases + [c + "+fuse" for c in cases if c.startswith("c2s") or c.startswith("c4s")]
    for name in cases:
        if name.endswith("+fuse"):
            os.environ["POGS_B200_FORCE_FUSE"] = "1"
            name_key, name = name, name[:-5]
        else:
            os.environ.pop("POGS_B200_FORCE_FUSE", None)
            name_key = name
        p = problems.build(name)
        m, n = p["A"].shape
        parts = row_partition(m, world)
        a, b = parts[rank]
        f_full = FunctionVector(m, *p["f"]); g = FunctionVector(n, *p["g"])
        for dtype in (np.float64, np.float32):
            s = RowBlockSolver(p["A"][a:b], m, comm, dtype=dtype)
            st = s.Solve(slice_function(f_full, a, b), g)
            r = s.gather_result(parts)
            tm = s.timing()
            s.close()
            if rank == 0:
                o = O.solve(p["A"], p["f"], p["g"], dtype=dtype)
                rel = lambda u, v: float(np.linalg.norm(u.astype(np.float64) - v) / np.linalg.norm(v))
                out[f"{name_key}/{np.dtype(dtype).name}"] = dict(
                    status=st, ostatus=o["status"], it=r["iterations"], oit=o["iterations"], ex=rel(r["x"], o["x"]),
                    ey=rel(r["y"], o["y"]), el=rel(r["l"], o["l"]),
                    eopt=abs(r["optval"] - o["optval"]) / abs(o["optval"]), loop_ms=tm["loop_ms"],
                    single_pass=tm.get("single_pass_iterations", 0.0))
            # replicas must be bit-identical across ranks
            xs = [None] * world
            dist.all_gather_object(xs, r["x"])
            out.setdefault("replicas_identical", True)
            out["replicas_identical"] &= all(np.array_equal(xs[0], v) for v in xs)
    # 3. bench-scale templates (n = 10000: NV = 5, full ring, streamed packed factor, tensor-core Gram) on row
    #    blocks against the compiled reference in fp64, fixed K at tolerance 0
    if os.environ.get("POGS_DIST_SCALE", "1") != "0":
        os.environ.pop("POGS_B200_FORCE_FUSE", None)
        os.environ.setdefault("OPENBLAS_NUM_THREADS", str(os.cpu_count() or 1))
        from oracle import ref_ctypes as R

        if R.available():
            m, n, K = 12000 * world, 10000, 40
            rng = np.random.default_rng(1)
            A = rng.standard_normal((m, n), dtype=np.float32)
            xs_ = (rng.standard_normal(n) * (rng.random(n) < 0.2)).astype(np.float32)
            b = (A @ xs_ + 0.1 * rng.standard_normal(m).astype(np.float32)).astype(np.float64)
            lam = 0.1 * float(np.abs(A.T.astype(np.float64) @ b).max())
            ft, gt = (problems.SQUARE, 1.0, b, 1.0, 0.0, 0.0), (problems.ABS, 1.0, 0.0, lam, 0.0, 0.0)
            parts = row_partition(m, world)
            a, b_ = parts[rank]
            s = RowBlockSolver(A[a:b_], m, comm, dtype=np.float32)
            s.SetAbsTol(0.0); s.SetRelTol(0.0); s.SetMaxIter(K)
            st = s.Solve(slice_function(FunctionVector(m, *ft), a, b_), FunctionVector(n, *gt))
            r = s.gather_result(parts)
            tm = s.timing()
            s.close()
            if rank == 0:
                o = R.solve(A, ft, gt, dtype=np.float64, abs_tol=0.0, rel_tol=0.0, max_iter=K)
                rel = lambda u, v: float(np.linalg.norm(u.astype(np.float64) - v) / np.linalg.norm(v))
                out["scale_fixedK"] = dict(st=st, it=r["iterations"], ex=rel(r["x"], o["x"]), ey=rel(r["y"], o["y"]),
                                           eopt=abs(r["optval"] - o["optval"]) / abs(o["optval"]),
                                           single_pass=tm.get("single_pass_iterations", 0.0), nnz=int(np.count_nonzero(o["x"])))
            xs = [None] * world
            dist.all_gather_object(xs, r["x"])
            out["replicas_identical"] &= all(np.array_equal(xs[0], v) for v in xs)
    comm.close()
    dist.barrier()
    if rank == 0:
        print("RESULT " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

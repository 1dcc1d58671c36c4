"""CPU suite, part 4: the N>1 path with world_size-2 gloo.

Checks (a) the host-side sharding helpers of pogs_b200/dist.py (row partition, descriptor
slicing, result gather, handle all-gather) and (b) the row-block decomposition of the ADMM
iteration itself, by running the numpy model of tests/dist_model.py on two gloo ranks and
comparing with the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_partition():
    from pogs_b200.dist import row_partition

    for m, w in ((100000, 8), (200000, 8), (10, 4), (7, 2), (3, 8), (1, 1), (1000, 3)):
        parts = row_partition(m, w)
        assert len(parts) == w and parts[0][0] == 0 and parts[-1][1] == m
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(p[1] >= p[0] for p in parts)
        if m >= 4 * w:
            sizes = [b - a for a, b in parts]
            assert max(sizes[:-1]) - min(sizes[:-1]) == 0 and all(s % 4 == 0 for s in sizes[:-1])
            assert sizes[-1] - sizes[0] < 4 * w + 4


def test_slice_function():
    from pogs_b200 import Function, FunctionVector
    from pogs_b200.dist import slice_function

    f = FunctionVector(10, Function.kSquare, 1.0, np.arange(10.0), 1.0)
    s = slice_function(f, 3, 7)
    assert len(s) == 4 and s.b.tolist() == [3.0, 4.0, 5.0, 6.0] and s.h.tolist() == [14] * 4
    assert s.a.flags.c_contiguous


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch.distributed as dist

        import dist_model
        import problems
        from oracle import oracle_ctypes as O
        from pogs_b200 import FunctionVector
        from pogs_b200.dist import all_gather_bytes, gather_rows, row_partition, slice_function

        dist.init_process_group("gloo", rank=rank, world_size=world)
        out = {}
        blobs = all_gather_bytes(bytes([rank]) * 64)
        out["handles_ok"] = [b[0] for b in blobs] == list(range(world)) and all(len(b) == 64 for b in blobs)
        for name in ("c1_lasso_500x300", "svm_600x200"):
            p = problems.build(name)
            m, n = p["A"].shape
            a, b = row_partition(m, world)[rank]
            f_full = FunctionVector(m, *p["f"])
            fl = slice_function(f_full, a, b)
            f_loc = (fl.h, fl.a, fl.b, fl.c, fl.d, fl.e)
            r = dist_model.solve_rowblock(O, p["A"][a:b], m, f_loc, p["g"])
            y = gather_rows(r["y"], None); l = gather_rows(r["l"], None)
            out[name] = dict(x=r["x"], y=y, l=l, optval=r["optval"], iterations=r["iterations"], status=r["status"])
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, out))
    except Exception as e:   # pragma: no cover
        import traceback

        q.put((rank, "ERR " + traceback.format_exc()))


def test_rowblock_model_matches_oracle_gloo_world2(oracle):
    import torch.multiprocessing as mp

    import problems
    from conftest import relerr

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r in range(2):
        assert not isinstance(res[r], str), res[r]
        assert res[r]["handles_ok"]
    for name in ("c1_lasso_500x300", "svm_600x200"):
        p = problems.build(name)
        o = oracle.solve(p["A"], p["f"], p["g"], dtype=np.float64)
        for r in range(2):
            got = res[r][name]
            assert got["status"] == o["status"] == 0
            assert abs(got["iterations"] - o["iterations"]) <= 2
            assert relerr(got["x"], o["x"]) < 1e-6 and relerr(got["y"], o["y"]) < 1e-6
            assert relerr(got["l"], o["l"]) < 1e-5
            assert abs(got["optval"] - o["optval"]) < 1e-7 * abs(o["optval"])
        # both ranks hold identical replicas
        assert np.array_equal(res[0][name]["x"], res[1][name]["x"])

"""Row-block multi-GPU solver: one process per GPU (torchrun), A split by rows.

The reference has no multi-device path; this is the design of SURVEY.md section 8e.
Rank g holds A_g (its row block, row-major), the matching slices of f, y and lambda,
and a replica of g, x, mu, the n x n factor and every scalar.  Per iteration the ranks
exchange one n-vector (A_g^T t_y, summed) and five doubles -- inside the A^T kernel and
the controller kernel, over NVLink peer memory mapped with CUDA IPC.  torch.distributed
is used for plumbing only: the rendezvous, the all-gather of the 64-byte IPC handles and
the gather of result slices.
"""
import ctypes

import numpy as np

from . import _lib
from .graph import FunctionVector
from .solver import Solver


def row_partition(m, world):
    """Contiguous, balanced row blocks: list of (start, stop), multiples of 4 rows where
    possible (the last block takes the remainder)."""
    base = (m // world) // 4 * 4 if m >= 4 * world else m // world
    bounds = [min(i * base, m) for i in range(world)] + [m]
    if base == 0:   # fewer rows than ranks: first m ranks get one row
        bounds = [min(i, m) for i in range(world)] + [m]
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def slice_function(f, start, stop):
    """Rows [start, stop) of a FunctionVector (the f descriptors follow the rows of A)."""
    f = FunctionVector.from_any(f)
    out = FunctionVector(stop - start)
    out.h = np.ascontiguousarray(f.h[start:stop])
    for k in ("a", "b", "c", "d", "e"):
        setattr(out, k, np.ascontiguousarray(getattr(f, k)[start:stop]))
    return out


def all_gather_bytes(payload: bytes, group=None):
    """All-gather of a small byte string over torch.distributed (any backend)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, payload, group=group)
    return out


def gather_rows(local, parts, group=None):
    """Concatenate per-rank row slices (numpy) on every rank."""
    import torch.distributed as dist

    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, np.asarray(local), group=group)
    return np.concatenate(out)


class PeerComm:
    """NVLink peer-memory communicator (include/pogs_b200.h part 2b)."""

    def __init__(self, slot_bytes, group=None):
        import torch.distributed as dist

        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.group = group
        self._c = None
        c = _lib.lib.pogs_b200_comm_create(self.rank, self.world, int(slot_bytes))
        if not c:
            raise RuntimeError("pogs_b200: " + _lib.last_error())
        self._c = ctypes.c_void_p(c)
        buf = ctypes.create_string_buffer(64)
        if _lib.lib.pogs_b200_comm_handle(self._c, buf):
            raise RuntimeError("pogs_b200: " + _lib.last_error())
        handles = all_gather_bytes(buf.raw, group)
        blob = b"".join(handles)
        if _lib.lib.pogs_b200_comm_open(self._c, ctypes.c_char_p(blob)):
            raise RuntimeError("pogs_b200: " + _lib.last_error())
        dist.barrier(group)

    def allreduce_(self, t):
        """In-place sum of a CUDA torch tensor (float32 / float64, contiguous, size padded
        to a multiple of 4) over the ranks -- test hook for the exchange kernel."""
        import torch

        assert t.is_cuda and t.is_contiguous()
        sfx = "d" if t.dtype == torch.float64 else "s"
        torch.cuda.current_stream().synchronize()
        rc = getattr(_lib.lib, "pogs_b200_comm_allreduce_" + sfx)(self._c, ctypes.c_void_p(t.data_ptr()), t.numel())
        if rc:
            raise RuntimeError("pogs_b200: " + _lib.last_error())
        return t

    def close(self):
        if self._c is not None:
            _lib.lib.pogs_b200_comm_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RowBlockSolver(Solver):
    """Solver(A_local, m_global, comm): persistent solver on one row block.

    A_local: CUDA torch tensor or numpy array, (m_local x n), row-major.  f passed to
    Solve() is the LOCAL slice (m_local entries); g is the full n-vector descriptor.
    result() returns x (replicated) and the local slices of y and lambda; use
    gather_result() for the full vectors."""

    def __init__(self, A_local, m_global, comm, dtype=np.float32):
        # deliberately not calling Solver.__init__: different creation entry point
        self._h = None
        self.comm = comm
        self.dtype = np.dtype(dtype)
        self._ct = _lib.ctype_of(self.dtype)
        self._sfx = _lib.suffix(self.dtype)
        self.m, self.n = A_local.shape
        self.m_global = int(m_global)
        # every rank knows m_global and the world size: refuse up front, on all ranks alike, a split
        # that leaves a rank without rows (it would fail alone while its peers wait inside a kernel)
        if self.m_global < comm.world or self.m == 0:
            raise ValueError(f"row-block solver needs at least one row per rank (m={self.m_global}, world={comm.world})")
        if hasattr(A_local, "is_cuda"):
            import torch

            want = torch.float64 if self.dtype == np.float64 else torch.float32
            At = A_local.to(want).contiguous()
            torch.cuda.current_stream().synchronize()
            ptr, on_dev = ctypes.c_void_p(At.data_ptr()), 1
        else:
            At = np.ascontiguousarray(A_local, dtype=self.dtype)
            ptr, on_dev = ctypes.c_void_p(At.ctypes.data), 0
        h = getattr(_lib.lib, "pogs_b200_create_dense_rowblock_" + self._sfx)(self.m, self.n, self.m_global, ptr,
                                                                             on_dev, comm._c)
        if not h:
            raise RuntimeError("pogs_b200: " + _lib.last_error())
        self._h = ctypes.c_void_p(h)
        self._p = dict(rho=1.0, abs_tol=1e-4, rel_tol=1e-4, max_iter=2500, verbose=0, adaptive_rho=True,
                       gap_stop=True)
        self._rho_dirty = True
        self.status = None

    def gather_result(self, parts=None):
        r = self.result()
        r["y"] = gather_rows(r["y"], parts, self.comm.group)
        r["l"] = gather_rows(r["l"], parts, self.comm.group)
        return r

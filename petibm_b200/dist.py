"""Multi-GPU wiring: one process per GPU, z-slab partition of the DMDA pressure grid.

torch.distributed is the host transport only (what MPI is inside PetIBM): it all-gathers the 64-byte
CUDA IPC handles of the solvers' exchange arenas and broadcasts the NCCL unique id.  The data path
never goes through it: halos are peer stores / peer copies over NVLink issued by the solver's own
kernels and stream, the scalar reductions are either in-kernel mailbox all-reduces over NVLink or
ncclAllReduce on the solver's stream (reference: VecScatter + MPI_Allreduce inside PETSc,
SURVEY.md section 8e)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .mesh import slab_range

REDUCE = {"p2p": _lib.REDUCE_P2P, "nccl": _lib.REDUCE_NCCL}
HALO = {"store": _lib.HALO_STORE, "memcpy": _lib.HALO_MEMCPY}


class Comm:
    """rank/nranks/device of this process plus the host transport."""

    def __init__(self, rank: int, nranks: int, device: int, reduce: str = "p2p", halo: str = "store", group=None):
        if reduce not in REDUCE or halo not in HALO:
            raise ValueError("reduce must be p2p|nccl and halo store|memcpy")
        if reduce == "p2p" and halo == "memcpy" and nranks > 1:
            raise ValueError("memcpy halos need reduce='nccl' (the all-reduce is their ordering point)")
        self.rank, self.nranks, self.device = int(rank), int(nranks), int(device)
        self.reduce, self.halo = reduce, halo
        self.group = group

    @staticmethod
    def from_env(reduce: str = "p2p", halo: str = "store", backend: str | None = None) -> "Comm":
        """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun) and joins the process group."""
        import torch
        import torch.distributed as dist

        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", str(rank)))
        if world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
        return Comm(rank, world, local, reduce, halo)

    # ---- host transport ------------------------------------------------------------------
    def allgather_bytes(self, payload: bytes) -> list:
        if self.nranks == 1:
            return [payload]
        import torch.distributed as dist

        out = [None] * self.nranks
        dist.all_gather_object(out, payload, group=self.group)
        return out

    def broadcast_bytes(self, payload: bytes | None, src: int = 0) -> bytes:
        if self.nranks == 1:
            return payload
        import torch.distributed as dist

        box = [payload]
        dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def barrier(self):
        if self.nranks > 1:
            import torch.distributed as dist

            dist.barrier(group=self.group)

    def allreduce_max(self, value: float) -> float:
        if self.nranks == 1:
            return float(value)
        return max(float(v) for v in self.allgather_bytes(float(value)))

    # ---- solver plumbing -----------------------------------------------------------------
    def init_solver(self, solver):
        _lib.check(solver._L.b200ls_comm_init(solver._h, self.rank, self.nranks, REDUCE[self.reduce], HALO[self.halo]),
                   solver._h)

    def connect_solver(self, solver):
        """Exchange arena handles (and the NCCL id) after the operator has been set on every rank."""
        L = solver._L
        mine = C.create_string_buffer(64)
        _lib.check(L.b200ls_comm_export(solver._h, mine), solver._h)
        handles = self.allgather_bytes(mine.raw)
        blob = C.create_string_buffer(b"".join(handles), 64 * self.nranks)
        _lib.check(L.b200ls_comm_connect(solver._h, blob, self.nranks), solver._h)
        if self.reduce == "nccl":
            uid = None
            if self.rank == 0:
                buf = C.create_string_buffer(128)
                _lib.check(L.b200ls_nccl_unique_id(buf))
                uid = buf.raw
            uid = self.broadcast_bytes(uid, 0)
            _lib.check(L.b200ls_nccl_init(solver._h, C.create_string_buffer(uid, 128)), solver._h)
        self.barrier()

    # ---- vector helpers (natural ordering <-> rank-local DMDA block) ------------------------
    def local_block(self, full: np.ndarray, n) -> np.ndarray:
        """Rank-local part of a natural-ordering global vector for a slab partition along the slowest axis
        (z in 3-D, y in 2-D)."""
        if len(n) == 2:
            nx, ny, nz = n[0], 1, n[1]
        else:
            nx, ny, nz = n
        lo, hi = slab_range(nz, self.rank, self.nranks)
        return np.ascontiguousarray(full.reshape(nz, ny, nx)[lo:hi].reshape(-1))

    def gather_blocks(self, local: np.ndarray) -> np.ndarray:
        parts = self.allgather_bytes(np.ascontiguousarray(local, dtype=np.float64).tobytes())
        return np.concatenate([np.frombuffer(p, dtype=np.float64) for p in parts])

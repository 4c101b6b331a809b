"""Multi-GPU wiring: one process per GPU, slab partition of the DMDA pressure grid (z-slabs in 3-D, y-slabs in
2-D); vectors that arrive in the box partition PETSc's DMDA chose are re-partitioned on the host (Repart).

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

    def allgather_f64(self, local: np.ndarray) -> np.ndarray:
        """Concatenation (rank order) of every rank's float64 vector; the parts may differ in length.  Tensor all-gather of
        torch.distributed (over NVLink with the nccl backend: the vector takes the GPU as a staging buffer; gloo on CPU)
        instead of the pickling object path -- this is the per-solve transport of the replicated solves."""
        local = np.ascontiguousarray(local, dtype=np.float64)
        if self.nranks == 1:
            return local.copy()
        import torch
        import torch.distributed as dist

        sizes = [int(v) for v in self.allgather_bytes(int(local.size))]
        m = max(sizes)
        on_gpu = dist.get_backend(self.group) == "nccl"
        dev = torch.device("cuda", self.device) if on_gpu else torch.device("cpu")
        mine = torch.zeros(m, dtype=torch.float64, device=dev)
        mine[: local.size] = torch.from_numpy(local).to(dev)
        out = torch.empty(m * self.nranks, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, mine, group=self.group)
        out = out.cpu().numpy().reshape(self.nranks, m)
        return np.concatenate([out[r, : sizes[r]] for r in range(self.nranks)])

    def allgather_f64_device(self, local: np.ndarray):
        """allgather_f64 that leaves the concatenated vector ON THE GPU (a CUDA float64 tensor), for solvers that take device
        pointers: one H2D copy of the local part, the all-gather over NVLink, no host round trip.  nccl backend only;
        returns None otherwise."""
        import torch
        import torch.distributed as dist

        if self.nranks == 1 or dist.get_backend(self.group) != "nccl":
            return None
        local = np.ascontiguousarray(local, dtype=np.float64)
        sizes = [int(v) for v in self.allgather_bytes(int(local.size))]
        m = max(sizes)
        dev = torch.device("cuda", self.device)
        mine = torch.zeros(m, dtype=torch.float64, device=dev)
        mine[: local.size] = torch.from_numpy(local).to(dev)
        out = torch.empty(m * self.nranks, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, mine, group=self.group)
        if all(sz == m for sz in sizes):
            return out
        return torch.cat([out[r * m: r * m + sizes[r]] for r in range(self.nranks)])

    def gather_matrix(self, nrows, indptr, indices, data, has_const=False, null_vecs=None):
        """All-gathers a row-distributed CSR matrix (global column indices) and its explicit null-space vectors: every
        rank receives the whole system -- the replicated solve of general systems on several ranks (LinSolverB200).
        Returns (indptr, indices, data, row offsets of the ranks, has_const, null_vecs)."""
        parts = self.allgather_bytes((int(nrows), np.asarray(indptr), np.asarray(indices), np.asarray(data), bool(has_const),
                                      None if null_vecs is None else np.asarray(null_vecs)))
        offs = np.concatenate([[0], np.cumsum([p[0] for p in parts])]).astype(np.int64)
        ips, base = [np.zeros(1, dtype=np.int64)], 0
        for p in parts:
            ip = np.asarray(p[1], dtype=np.int64)
            ips.append(ip[1:] + base)
            base += int(ip[-1])
        nv = None
        if parts[0][5] is not None:
            nv = np.ascontiguousarray(np.concatenate([np.atleast_2d(p[5]) for p in parts], axis=1), dtype=np.float64)
        return (np.concatenate(ips), np.concatenate([np.asarray(p[2], dtype=np.int32) for p in parts]),
                np.concatenate([np.asarray(p[3], dtype=np.float64) for p in parts]), offs, parts[0][4], nv)

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


class Repart:
    """DMDA box partition <-> slab partition (b200ls_repart_* of the C ABI; reference: the DMDA ownership of
    src/mesh/cartesianmesh.cpp:500-538 and the PETSc ordering of :709-721).  Index planning happens in libb200ls.so;
    this class only carries the plan of ONE rank and moves bytes with the host transport."""

    def __init__(self, dim: int, n, procs, rank: int):
        self._L = _lib.lib()
        self._p = C.c_void_p()
        n3 = (C.c_int64 * 3)(*(list(n) + [1] * (3 - len(n))))
        p3 = (C.c_int * 3)(*(list(procs) + [1] * (3 - len(procs))))
        _lib.check(self._L.b200ls_repart_create(C.byref(self._p), int(dim), n3, p3, int(rank)))
        self.dim, self.rank = int(dim), int(rank)
        self.procs = tuple(int(v) for v in list(procs)[:dim]) + (1,) * (3 - dim)
        self.nranks = int(np.prod(self.procs))
        lo, hi = (C.c_int64 * 3)(), (C.c_int64 * 3)()
        nbox, slo, shi, nslab, ident = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        _lib.check(self._L.b200ls_repart_info(self._p, lo, hi, C.byref(nbox), C.byref(slo), C.byref(shi), C.byref(nslab),
                                              C.byref(ident)))
        self.box_lo, self.box_hi = tuple(lo), tuple(hi)
        self.nbox, self.nslab = nbox.value, nslab.value
        self.slab = (slo.value, shi.value)
        self.identity = bool(ident.value)
        arr = [np.zeros(self.nranks, dtype=np.int64) for _ in range(4)]
        _lib.check(self._L.b200ls_repart_counts(self._p, *[a.ctypes.data_as(_lib._i64p) for a in arr]))
        self.box_counts, self.box_displs, self.slab_counts, self.slab_displs = arr

    def __del__(self):
        if getattr(self, "_p", None) is not None and self._p:
            self._L.b200ls_repart_destroy(self._p)
            self._p = C.c_void_p()

    # ---- index maps
    def box_rows(self) -> np.ndarray:
        out = np.empty(self.nbox, dtype=np.int64)
        _lib.check(self._L.b200ls_repart_box_rows(self._p, out.ctypes.data_as(_lib._i64p)))
        return out

    def petsc_to_natural(self, idx) -> np.ndarray:
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        out = np.empty_like(idx)
        _lib.check(self._L.b200ls_repart_petsc_to_natural(self._p, idx.size, idx.ctypes.data_as(_lib._i32p),
                                                          out.ctypes.data_as(_lib._i32p)))
        return out

    @staticmethod
    def candidates(dim: int, n, sizes) -> list:
        """Process grids whose boxes have exactly the per-rank sizes `sizes` (the Mat does not carry its DMDA)."""
        L = _lib.lib()
        n3 = (C.c_int64 * 3)(*(list(n) + [1] * (3 - len(n))))
        sz = np.ascontiguousarray(sizes, dtype=np.int64)
        out = np.zeros(3 * 64, dtype=np.int32)
        found = C.c_int(0)
        _lib.check(L.b200ls_repart_candidates(int(dim), n3, int(sz.size), sz.ctypes.data_as(_lib._i64p),
                                              out.ctypes.data_as(_lib._ip), 64, C.byref(found)))
        return [tuple(int(v) for v in out[3 * c: 3 * c + 3]) for c in range(min(found.value, 64))]

    # ---- local halves of the exchange
    def unpack_slab(self, recvbuf: np.ndarray) -> np.ndarray:
        recvbuf = np.ascontiguousarray(recvbuf, dtype=np.float64)
        assert recvbuf.size == self.nslab
        out = np.empty(self.nslab, dtype=np.float64)
        _lib.check(self._L.b200ls_repart_unpack_slab(self._p, recvbuf.ctypes.data_as(_lib._dp), out.ctypes.data_as(_lib._dp)))
        return out

    def pack_slab(self, slab: np.ndarray) -> np.ndarray:
        slab = np.ascontiguousarray(slab, dtype=np.float64)
        assert slab.size == self.nslab
        out = np.empty(self.nslab, dtype=np.float64)
        _lib.check(self._L.b200ls_repart_pack_slab(self._p, slab.ctypes.data_as(_lib._dp), out.ctypes.data_as(_lib._dp)))
        return out

    # ---- the exchange itself: ONE all-to-all per direction on the host transport (MPI_Alltoallv in the shim)
    def _alltoall(self, send: np.ndarray, send_counts, recv_counts, group=None) -> np.ndarray:
        import torch
        import torch.distributed as dist

        inp = torch.from_numpy(np.ascontiguousarray(send, dtype=np.float64))
        # the harness transport may be NCCL (one process per GPU under torchrun), which only moves device tensors; inside
        # PetIBM this is MPI_Alltoallv(send, counts, displs, MPI_DOUBLE, recv, counts, displs, MPI_DOUBLE, PETSC_COMM_WORLD)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        out = torch.empty(int(np.sum(recv_counts)), dtype=torch.float64, device=dev)
        dist.all_to_all_single(out, inp.to(dev), [int(c) for c in recv_counts], [int(c) for c in send_counts], group=group)
        return out.cpu().numpy()

    def box_to_slab(self, box: np.ndarray, group=None) -> np.ndarray:
        """Box-ordered local vector (PETSc ordering) -> slab-ordered local vector (natural order inside the slab)."""
        if self.nranks == 1 or self.identity:
            return np.ascontiguousarray(box, dtype=np.float64)
        assert box.size == self.nbox
        return self.unpack_slab(self._alltoall(box, self.box_counts, self.slab_counts, group))

    def slab_to_box(self, slab: np.ndarray, group=None) -> np.ndarray:
        if self.nranks == 1 or self.identity:
            return np.ascontiguousarray(slab, dtype=np.float64)
        return self._alltoall(self.pack_slab(slab), self.slab_counts, self.box_counts, group)

"""Host-side mirror of PetIBM's linear-solver plugin interface for the B200 backend.

Mirrors, name for name, ``petibm::linsolver::LinSolverBase`` (include/petibm/linsolver.h:59-147) and the
factory ``createLinSolver`` (src/linsolver/linsolver.cpp:57-91); ``LinSolverB200`` is the third
backend next to ``LinSolverKSP`` (src/linsolver/linsolverksp.cpp) and ``LinSolverAmgX``
(src/linsolver/linsolveramgx.cpp).  All arithmetic happens in libb200ls.so (hand-written sm_100a CUDA
behind the C ABI of include/b200ls.h); this module only moves pointers.  The C++ shim that does the
same with real PETSc ``Mat``/``Vec`` objects is petibm_b200/csrc/petibm_shim/ (see INTEGRATION.md).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import B200Error
from .mesh import Grid, slab_range


class Mat:
    """What ``setMatrix`` receives: the rank-local rows of an assembled AIJ matrix (CSR; global column
    indices in the PETSc ordering of the DMDA, which is the natural ordering on one rank and for a 1 x 1 x P
    process grid) plus the null space the application attached with
    ``MatSetNullSpace`` (navierstokes.cpp:404-413, ibpm.cpp:251-267)."""

    def __init__(self, indptr, indices, data, ncols=None):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = np.ascontiguousarray(data, dtype=np.float64)
        self.nrows = int(self.indptr.size - 1)
        self.ncols = int(ncols) if ncols is not None else self.nrows
        self.null_has_const = False
        self.null_vecs = None

    @staticmethod
    def from_scipy(A) -> "Mat":
        A = A.tocsr()
        A.sort_indices()
        return Mat(A.indptr, A.indices, A.data, A.shape[1])

    def setNullSpace(self, has_const: bool, vecs=None) -> "Mat":
        """MatNullSpaceCreate(comm, has_cnst, n, vecs) + MatSetNullSpace."""
        self.null_has_const = bool(has_const)
        self.null_vecs = None if vecs is None else np.ascontiguousarray(np.atleast_2d(vecs), dtype=np.float64)
        return self


class LinSolverBase:
    """include/petibm/linsolver.h:59-147."""

    def __init__(self, solverName: str, file: str):
        self.name = solverName
        self.config = file
        self.type = "undefined"

    def getType(self) -> str:
        return self.type

    def printInfo(self) -> str:  # linsolver.cpp:27-46
        info = "=" * 80 + "\n"
        info += f"Linear Solver {self.name}:\n"
        info += "=" * 80 + "\n"
        info += f"\tType: {self.type}\n\n"
        info += f"\tConfig file: {self.config}\n\n"
        info += "=" * 80 + "\n"
        return info

    def destroy(self):
        self.name = self.config = self.type = ""

    # pure virtuals
    def setMatrix(self, A):
        raise NotImplementedError

    def solve(self, x, b):
        raise NotImplementedError

    def getIters(self) -> int:
        raise NotImplementedError

    def getResidual(self) -> float:
        raise NotImplementedError


def _is_torch_cuda(t) -> bool:
    return hasattr(t, "is_cuda") and bool(t.is_cuda)


class LinSolverB200(LinSolverBase):
    """The B200-native backend.

    ``type`` reads "PETSc KSP" on purpose: the applications dispatch their null-space handling on that
    string and abort on anything else (navierstokes.cpp:401-426, ibpm.cpp:248-280); reporting the KSP
    type makes them attach the constant null space to DBNG exactly as they do for LinSolverKSP, which is
    the behaviour this backend reproduces.  ``backend`` carries the real name for printInfo()."""

    backend = "B200 sm_100a (libb200ls)"

    def __init__(self, solverName: str, file: str, node=None, device: int | None = None, comm=None):
        super().__init__(solverName, file)
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self._node = node
        self._comm = comm
        self._grid = None
        self._procs = None     # DMDA process grid, if the caller knows it (setProcessGrid)
        self._repart = None    # box <-> slab exchange plan when the vectors arrive as DMDA boxes
        self._staggered = True # setMatrix may use the line-coefficient form for velocity / IBPM matrices
        self._replicated = None  # (row_lo, row_hi, nrows_global) when several ranks solve replicas of a general system
        self._mg_rep = None      # (lo, hi, n) in natural ordering: several ranks, pc_type mg -> every rank solves the whole grid
        self._handle_wired = comm is not None and comm.nranks > 1   # the C handle is wired to the communicator
        self.operator = None   # "stencil" | "hybrid" | "staggered" | "csr" after setMatrix
        if device is None:
            device = comm.device if comm is not None else 0
        _lib.check(self._L.b200ls_create(C.byref(self._h), int(device)))
        self.init()

    # ---- LinSolverKSP::init (linsolverksp.cpp:48-69)
    def init(self):
        self.type = "PETSc KSP"
        opts = _lib.Options()
        self._L.b200ls_default_options(C.byref(opts))
        if self.config and self.config != "None":
            with open(self.config, "r") as fh:
                text = fh.read()
            err = C.create_string_buffer(512)
            rc = self._L.b200ls_parse_options(text.encode(), (self.name + "_").encode(), C.byref(opts), err, 512)
            if rc != _lib.OK:
                raise B200Error(rc, err.value.decode())
        _lib.check(self._L.b200ls_set_options(self._h, C.byref(opts)), self._h)
        if self._comm is not None and self._comm.nranks > 1:
            self._comm.init_solver(self)

    def printInfo(self) -> str:
        return super().printInfo().replace(f"\tType: {self.type}\n", f"\tType: {self.type} (interface) / {self.backend}\n")

    def destroy(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.b200ls_destroy(self._h)
            self._h = C.c_void_p()
        super().destroy()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # ---- options access for tests / bench
    def options(self) -> _lib.Options:
        o = _lib.Options()
        _lib.check(self._L.b200ls_get_options(self._h, C.byref(o)), self._h)
        return o

    def setOptions(self, **kw):
        o = self.options()
        names = {"cg": _lib.KSP_CG, "bcgs": _lib.KSP_BCGS, "preonly": _lib.KSP_PREONLY, "none": _lib.PC_NONE,
                 "jacobi": _lib.PC_JACOBI, "mg": _lib.PC_MG, "lu": _lib.PC_LU,
                 "preconditioned": _lib.NORM_PRECONDITIONED, "unpreconditioned": _lib.NORM_UNPRECONDITIONED,
                 "natural": _lib.NORM_NATURAL}
        for k, v in kw.items():
            if isinstance(v, str):
                v = names[v]
            setattr(o, k, v)
        _lib.check(self._L.b200ls_set_options(self._h, C.byref(o)), self._h)

    def setTuning(self, key: str, value: int):
        _lib.check(self._L.b200ls_set_tuning(self._h, key.encode(), int(value)), self._h)

    # ---- operator
    def setGrid(self, grid: Grid):
        """Grid description used to recognise the separable operator inside setMatrix; the factory fills
        it from the YAML node (mesh, boundary conditions, dt) exactly like the C++ shim does."""
        self._grid = grid

    def _slab(self, grid: Grid):
        multi = self._comm is not None and self._comm.nranks > 1
        # slabs along the slowest axis: z in 3-D; y in 2-D when several GPUs share the grid
        nslow = grid.n[2] if grid.dim == 3 else (grid.n[1] if multi else 1)
        if multi:
            return slab_range(nslow, self._comm.rank, self._comm.nranks)
        return 0, nslow

    def _local_size(self, grid: Grid, lo: int, hi: int) -> int:
        multi = self._comm is not None and self._comm.nranks > 1
        if grid.dim == 2:
            return int(grid.n[0] * (hi - lo)) if multi else int(grid.n[0] * grid.n[1])
        return int(grid.n[0] * grid.n[1] * (hi - lo))

    def setStencil(self, grid: Grid):
        """Matrix-free operator D (dt I) G straight from the grid (what setMatrix does after recognising
        the assembled matrix); used by the benchmark where assembling a 256^3 CSR on the host is not
        part of the measured path."""
        self._grid = grid
        n = (C.c_int64 * 3)(*(list(grid.n) + [1] * (3 - grid.dim)))
        per = (C.c_int * 3)(*[int(bool(p)) for p in grid.periodic])
        w = [np.ascontiguousarray(a, dtype=np.float64) for a in grid.widths]
        dz = w[2].ctypes.data_as(_lib._dp) if grid.dim == 3 else None
        lo, hi = self._slab(grid)
        multi = self._comm is not None and self._comm.nranks > 1
        self._mg_rep = None
        if multi and self.options().pc_type == _lib.PC_MG:
            # The multigrid preconditioner runs on one GPU.  On several ranks every rank sets up the
            # WHOLE grid on its own GPU and solves a bit-identical replica: solve() all-gathers b and keeps its own slab
            # of x.  Functional, not scalable -- but one multigrid solve on one GPU (13.8 ms at 256^3) is shorter than the
            # 500 plain CG iterations it replaces on eight.
            if self._handle_wired:
                self._new_handle(with_comm=False)
            nslow = grid.n[2] if grid.dim == 3 else grid.n[1]
            _lib.check(self._L.b200ls_set_poisson_stencil(
                self._h, grid.dim, n, per, w[0].ctypes.data_as(_lib._dp), w[1].ctypes.data_as(_lib._dp), dz,
                float(grid.dt), 0, nslow if grid.dim == 3 else 1), self._h)
            plane = int(grid.n[0] * grid.n[1]) if grid.dim == 3 else int(grid.n[0])
            self._mg_rep = (lo * plane, hi * plane, int(grid.size))
            self.operator = "stencil"
            self.nlocal = self._local_size(grid, lo, hi)
            return
        if multi and not self._handle_wired:
            self._new_handle(with_comm=True)     # the previous operator was solved as replicas: distributed handle again
        if multi:
            # replacing the operator frees the exchange arena: every rank first drops its mappings of the peers
            _lib.check(self._L.b200ls_comm_disconnect(self._h), self._h)
            self._comm.barrier()
        _lib.check(self._L.b200ls_set_poisson_stencil(
            self._h, grid.dim, n, per, w[0].ctypes.data_as(_lib._dp), w[1].ctypes.data_as(_lib._dp), dz,
            float(grid.dt), lo, hi), self._h)
        self.operator = "stencil"
        self.nlocal = self._local_size(grid, lo, hi)
        if self._comm is not None and self._comm.nranks > 1:
            self._comm.connect_solver(self)

    def setProcessGrid(self, procs):
        """Process grid (m, n[, p]) of the DMDA the vectors and matrix rows arrive in (PetIBM: mesh->nProc,
        cartesianmesh.cpp:547-553).  Optional: without it setMatrix tries every grid that reproduces the ranks'
        local sizes against the matrix entries."""
        self._procs = None if procs is None else tuple(int(v) for v in procs)

    def _verify_boxes(self, A: Mat):
        """Several ranks: the rows of A are this rank's DMDA box, its columns PETSc global indices (for a
        1 x 1 x P process grid that IS the natural ordering).  Finds the process grid under which A is the
        separable stencil of the mesh; returns the exchange plan or None.  Collective."""
        from .dist import Repart

        comm, grid = self._comm, self._grid
        sizes = [int(v) for v in comm.allgather_bytes(int(A.nrows))]
        cands = Repart.candidates(grid.dim, grid.n, sizes)
        if getattr(self, "_procs", None) is not None:
            want = self._procs + (1,) * (3 - len(self._procs))
            cands = [c for c in cands if c == want]
        if not cands:
            return None
        self.setStencil(grid)   # slab partition; allocates and connects the exchange arenas
        for c in cands:
            plan = Repart(grid.dim, grid.n, c[: grid.dim], comm.rank)
            rows = plan.box_rows()
            cols = plan.petsc_to_natural(A.indices)
            diff = C.c_double(0.0)
            # diag_ulps = 4: partition-dependent accumulation order of PETSc's parallel MatMatMult (b200ls.h)
            rc = self._L.b200ls_verify_csr_rows(self._h, A.nrows, rows.ctypes.data_as(_lib._i64p),
                                                A.indptr.ctypes.data_as(_lib._i64p), cols.ctypes.data_as(_lib._i32p),
                                                A.data.ctypes.data_as(_lib._dp), 4, C.byref(diff))
            if rc not in (_lib.OK, _lib.ERR_MISMATCH):
                _lib.check(rc, self._h)
            if all(comm.allgather_bytes(rc == _lib.OK)):   # every rank must agree on the mapping
                return plan
        return None

    def setStaggered(self, enable: bool):
        """Whether setMatrix may keep a staggered-grid matrix that is not the pressure stencil in line-coefficient
        form (b200ls_set_staggered) instead of plain CSR.  On by default; the structure is verified against the
        matrix either way."""
        self._staggered = bool(enable)

    def _staggered_layouts(self, nrows: int):
        """Field layouts the matrix may have, from the mesh: the packed velocity vector [u | v | w]
        (cartesianmesh.cpp:251-273: one point fewer than cells along the field's own direction unless periodic), or
        the pressure block followed by IBPM's Lagrangian force rows (ibpm.cpp:164-194)."""
        g = self._grid
        n = list(g.n) + [1] * (3 - g.dim)
        per = [int(bool(p)) for p in g.periodic][:3]
        vel = [[n[d] - (1 if (d == f and not per[d]) else 0) for d in range(3)] for f in range(g.dim)]
        out = []
        if nrows == sum(int(np.prod(v)) for v in vel) and all(min(v) >= 1 for v in vel):
            out.append(vel)
        if nrows > g.size:
            out.append([n])
        return out, per

    def _try_hybrid(self, A: Mat) -> bool:
        """IBPM's modified Poisson system: the pressure operator of the mesh (stretched grid: coefficients with face areas)
        followed by the Lagrangian coupling (b200ls_set_poisson_hybrid)."""
        g = self._grid
        if A.nrows <= g.size:
            return False
        n = (C.c_int64 * 3)(*(list(g.n) + [1] * (3 - g.dim)))
        per = (C.c_int * 3)(*[int(bool(p)) for p in g.periodic][:3])
        w = [np.ascontiguousarray(a, dtype=np.float64) for a in g.widths]
        dz = w[2].ctypes.data_as(_lib._dp) if g.dim == 3 else None
        rc = self._L.b200ls_set_poisson_hybrid(self._h, g.dim, n, per, w[0].ctypes.data_as(_lib._dp), w[1].ctypes.data_as(_lib._dp),
                                               dz, float(g.dt), A.nrows, A.indptr.ctypes.data_as(_lib._i64p),
                                               A.indices.ctypes.data_as(_lib._i32p), A.data.ctypes.data_as(_lib._dp))
        if rc == _lib.OK:
            self.operator = "hybrid"
            self.nlocal = A.nrows
            return True
        if rc != _lib.ERR_MISMATCH:
            _lib.check(rc, self._h)
        return False

    def _try_staggered(self, A: Mat) -> bool:
        """Same order as LinSolverB200::setMatrix in the PetIBM shim (linsolverb200.cpp): the packed velocity layout,
        then the hybrid form (pressure operator of the mesh + coupling rows), then the single-field line-coefficient
        form with a remainder -- a matrix that fits more than one form is the same operator kind in both front ends."""
        layouts, per = self._staggered_layouts(A.nrows)

        def staggered(dims):
            d = np.ascontiguousarray(dims, dtype=np.int64).reshape(-1)
            rc = self._L.b200ls_set_staggered(self._h, len(dims), d.ctypes.data_as(_lib._i64p), (C.c_int * 3)(*per), A.nrows,
                                              A.indptr.ctypes.data_as(_lib._i64p), A.indices.ctypes.data_as(_lib._i32p),
                                              A.data.ctypes.data_as(_lib._dp))
            if rc == _lib.OK:
                self.operator = "staggered"
                self.nlocal = A.nrows
                return True
            if rc != _lib.ERR_MISMATCH:
                _lib.check(rc, self._h)
            return False

        for dims in layouts:
            if len(dims) > 1 and staggered(dims):       # packed [u | v | w]
                return True
        if self._try_hybrid(A):
            return True
        for dims in layouts:
            if len(dims) == 1 and staggered(dims):      # pressure block + remainder
                return True
        return False

    def _replicate(self, A: Mat) -> Mat:
        """Collective: all-gathers the rows (PETSc global column indices) and the explicit null-space vectors over the
        host transport, turns this solver into a single-GPU replica (fresh handle, same options) and returns the
        global matrix.  solve() all-gathers b and hands back the caller's rows of x."""
        comm = self._comm
        indptr, indices, data, offs, has_const, nv = comm.gather_matrix(A.nrows, A.indptr, A.indices, A.data, A.null_has_const,
                                                                        A.null_vecs)
        G = Mat(indptr, indices, data, int(offs[-1]))
        G.setNullSpace(has_const, nv)
        self._new_handle(with_comm=False)
        self._replicated = (int(offs[comm.rank]), int(offs[comm.rank + 1]), int(offs[-1]))
        return G

    def _new_handle(self, with_comm: bool):
        """A fresh C handle with the options of the old one, single-rank (replica) or wired to the communicator again.
        Collective: every rank drops its peer mappings first."""
        opts = self.options()
        _lib.check(self._L.b200ls_comm_disconnect(self._h), self._h)
        self._comm.barrier()
        self._L.b200ls_destroy(self._h)
        self._h = C.c_void_p()
        _lib.check(self._L.b200ls_create(C.byref(self._h), int(self._comm.device)))
        _lib.check(self._L.b200ls_set_options(self._h, C.byref(opts)), self._h)
        if with_comm:
            self._comm.init_solver(self)
        self._handle_wired = bool(with_comm)

    def setMatrix(self, A: Mat):
        """LinSolverKSP::setMatrix (linsolverksp.cpp:72-82).  The matrix is copied/recognised here, the
        caller keeps ownership (as with AmgXSolver::setA, linsolveramgx.cpp:84)."""
        if not isinstance(A, Mat):
            A = Mat.from_scipy(A)
        recognised = False
        self._repart = None
        multi = self._comm is not None and self._comm.nranks > 1
        if multi and getattr(self, "_replicated", None) is not None:
            self._new_handle(with_comm=True)     # the previous matrix was solved as replicas: distributed handle again
        self._replicated = None
        if self._grid is not None and multi:
            self._repart = self._verify_boxes(A)
            recognised = self._repart is not None
            if recognised:
                self.nlocal = self._repart.nbox    # what the caller's vectors hold: its DMDA box
        elif self._grid is not None:
            lo, hi = self._slab(self._grid)
            if A.nrows == self._local_size(self._grid, lo, hi):
                self.setStencil(self._grid)
                diff = C.c_double(0.0)
                rc = self._L.b200ls_verify_csr(self._h, A.nrows, A.indptr.ctypes.data_as(_lib._i64p),
                                               A.indices.ctypes.data_as(_lib._i32p), A.data.ctypes.data_as(_lib._dp),
                                               C.byref(diff))
                if rc == _lib.OK:
                    recognised = True
                elif rc != _lib.ERR_MISMATCH:
                    _lib.check(rc, self._h)
        if not recognised and multi:
            # Any other system on several ranks (velocity system, IBPM's modified Poisson system, forces system, BN > 1;
            # LinSolverKSP solves them on any rank count, linsolverksp.cpp:72-105): every rank gathers the whole matrix
            # and solves the same system on its own GPU -- bit-identical replicas, each keeps its rows of x.
            A = self._replicate(A)
            multi = False
        if not recognised and not multi and self._grid is not None and self._staggered and self._replicated is None:
            recognised = self._try_staggered(A)
        if not recognised:
            # general assembled operator, still on the GPU (IBPM modified Poisson, velocity system, BN > 1)
            _lib.check(self._L.b200ls_set_csr(self._h, A.nrows, A.indptr.ctypes.data_as(_lib._i64p),
                                              A.indices.ctypes.data_as(_lib._i32p), A.data.ctypes.data_as(_lib._dp)),
                       self._h)
            self.operator = "csr"
            self.nlocal = A.nrows
        if self._replicated is not None:
            self.nlocal = self._replicated[1] - self._replicated[0]   # the caller's vectors hold its own rows
        nv = 0 if A.null_vecs is None else int(A.null_vecs.shape[0])
        pv = A.null_vecs.ctypes.data_as(_lib._dp) if nv else None
        _lib.check(self._L.b200ls_set_nullspace(self._h, int(A.null_has_const), nv, pv), self._h)

    def setNullSpace(self, has_const: bool, vecs=None):
        nv = 0 if vecs is None else int(np.atleast_2d(vecs).shape[0])
        arr = None if vecs is None else np.ascontiguousarray(np.atleast_2d(vecs), dtype=np.float64)
        _lib.check(self._L.b200ls_set_nullspace(self._h, int(bool(has_const)), nv,
                                                arr.ctypes.data_as(_lib._dp) if nv else None), self._h)

    def _solve_gathered(self, b_local: np.ndarray, lo: int, hi: int, ntot: int):
        """Replicated solves: all-gather b, solve the whole system on this GPU, return (rc, own part of x).  With the nccl
        backend the gathered vector and the solution stay on the device (H2D of the local part, all-gather over NVLink,
        b200ls_solve_device, D2H of the own part only); otherwise through host buffers."""
        bd = self._comm.allgather_f64_device(b_local)
        if bd is not None:
            import torch

            xd = torch.empty(ntot, dtype=torch.float64, device=bd.device)
            torch.cuda.current_stream(bd.device).synchronize()
            rc = self._L.b200ls_solve_device(self._h, C.c_void_p(bd.data_ptr()), C.c_void_p(xd.data_ptr()))
            self._sync_stream(bd.device)
            return rc, xd[lo:hi].cpu().numpy()
        bf = self._comm.allgather_f64(b_local)
        xf = np.empty(ntot, dtype=np.float64)
        rc = self._L.b200ls_solve(self._h, C.c_void_p(bf.ctypes.data), C.c_void_p(xf.ctypes.data))
        return rc, np.ascontiguousarray(xf[lo:hi])

    # ---- LinSolverKSP::solve (linsolverksp.cpp:85-105): zero initial guess, error if reason < 0
    def solve(self, x, b):
        rep = getattr(self, "_replicated", None)
        if rep is not None:
            lo, hi, ntot = rep
            if _is_torch_cuda(b) or _is_torch_cuda(x) or not isinstance(x, np.ndarray):
                raise ValueError("vectors of a replicated solve are host numpy arrays")
            if np.size(b) != hi - lo or x.size != hi - lo:
                raise ValueError("vector length does not match the operator")
            rc, xs = self._solve_gathered(np.ascontiguousarray(b, dtype=np.float64), lo, hi, ntot)
            if rc in (_lib.OK, _lib.ERR_DIVERGED):
                x[...] = xs.reshape(x.shape)
            _lib.check(rc, self._h)
            return x
        mgr = getattr(self, "_mg_rep", None)
        if mgr is not None:
            # pc_type mg on several ranks: boxes -> slabs (if the vectors arrive as DMDA boxes), all-gather the slabs (rank
            # order = natural ordering), solve the whole grid on this GPU, keep the own slab
            lo, hi, ntot = mgr
            rp = getattr(self, "_repart", None)
            boxes = rp is not None and not rp.identity
            if _is_torch_cuda(b) or _is_torch_cuda(x) or not isinstance(x, np.ndarray):
                raise ValueError("vectors of a replicated solve are host numpy arrays")
            bs = rp.box_to_slab(np.asarray(b, dtype=np.float64), self._comm.group) if boxes else np.ascontiguousarray(b, dtype=np.float64)
            if bs.size != hi - lo:
                raise ValueError("vector length does not match the operator")
            rc, xs = self._solve_gathered(bs, lo, hi, ntot)
            if rc in (_lib.OK, _lib.ERR_DIVERGED):
                x[...] = (rp.slab_to_box(xs, self._comm.group) if boxes else xs).reshape(x.shape)
            _lib.check(rc, self._h)
            return x
        rp = getattr(self, "_repart", None)
        if rp is not None and not rp.identity:
            # the caller's vectors are DMDA boxes: one all-to-all into the solver's slabs and one back (what
            # VecScatter does inside PETSc; MPI_Alltoallv in the C++ shim)
            if _is_torch_cuda(b) or _is_torch_cuda(x) or not isinstance(x, np.ndarray):
                raise ValueError("box-partitioned vectors are host numpy arrays")
            if np.size(b) != rp.nbox or x.size != rp.nbox:
                raise ValueError("vector length does not match the operator")
            bs = rp.box_to_slab(np.asarray(b, dtype=np.float64), self._comm.group)
            xs = np.empty(rp.nslab, dtype=np.float64)
            rc = self._L.b200ls_solve(self._h, C.c_void_p(bs.ctypes.data), C.c_void_p(xs.ctypes.data))
            if rc in (_lib.OK, _lib.ERR_DIVERGED):
                x[...] = rp.slab_to_box(xs, self._comm.group).reshape(x.shape)
            _lib.check(rc, self._h)
            return x
        if _is_torch_cuda(b) or _is_torch_cuda(x):
            if not (_is_torch_cuda(b) and _is_torch_cuda(x)):
                raise ValueError("x and b must live on the same side")
            import torch

            if b.dtype != torch.float64 or x.dtype != torch.float64 or not b.is_contiguous() or not x.is_contiguous():
                raise ValueError("device vectors must be contiguous float64")
            if b.numel() != self.nlocal or x.numel() != self.nlocal:
                raise ValueError("vector length does not match the operator")
            torch.cuda.current_stream(b.device).synchronize()
            rc = self._L.b200ls_solve_device(self._h, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()))
        else:
            if hasattr(b, "data_ptr"):   # CPU torch tensors (pinned or not)
                bp, xp, nb, nx_ = b.data_ptr(), x.data_ptr(), b.numel(), x.numel()
            else:
                if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous):
                    raise ValueError("x must be a contiguous float64 numpy array (it is written in place)")
                b = np.ascontiguousarray(b, dtype=np.float64)
                bp, xp, nb, nx_ = b.ctypes.data, x.ctypes.data, b.size, x.size
            if nb != self.nlocal or nx_ != self.nlocal:
                raise ValueError("vector length does not match the operator")
            rc = self._L.b200ls_solve(self._h, C.c_void_p(bp), C.c_void_p(xp))
        _lib.check(rc, self._h)   # reason < 0 raises, like SETERRQ2(PETSC_ERR_CONV_FAILED)
        return x

    def apply(self, x):
        """y = A x through the device operator (tests)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        mgr = getattr(self, "_mg_rep", None)
        if mgr is not None:                      # replicated whole-grid operator: slab in, slab out
            lo, hi, ntot = mgr
            xf = self._comm.allgather_f64(x)
            yf = np.empty(ntot, dtype=np.float64)
            _lib.check(self._L.b200ls_apply(self._h, C.c_void_p(xf.ctypes.data), C.c_void_p(yf.ctypes.data)), self._h)
            return np.ascontiguousarray(yf[lo:hi])
        y = np.empty_like(x)
        _lib.check(self._L.b200ls_apply(self._h, C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data)), self._h)
        return y

    # ---- the operators on either side of the pressure solve (extension beyond LinSolverBase):
    # matrix-free D, G and BN G of the mesh given to setStencil / recognised by setMatrix, bit-identical to MatMult on the
    # assembled matrices (navierstokes.cpp:442, 540-551, 583-615).  numpy arrays or CUDA float64 torch tensors.
    def velocitySize(self):
        nv, npr = C.c_int64(0), C.c_int64(0)
        _lib.check(self._L.b200ls_velocity_size(self._h, C.byref(nv), C.byref(npr)), self._h)
        return nv.value, npr.value

    def _sync_stream(self, device):
        """The *_device entry points are asynchronous on the solver's own stream: wait for it before torch reads the result
        on its stream (an application that chains b200ls calls only would not need this)."""
        import torch

        torch.cuda.ExternalStream(int(self._L.b200ls_stream(self._h)), device=device).synchronize()

    @staticmethod
    def _dev(t, n):
        import torch

        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() == n):
            raise ValueError("device vectors must be contiguous CUDA float64 tensors of the operator's size")
        return C.c_void_p(t.data_ptr())

    def divergence(self, u, out=None):
        """out = D u   (MatMult(D, UGlobal, rhs2), navierstokes.cpp:548)."""
        nv, npr = self.velocitySize()
        if _is_torch_cuda(u):
            import torch

            out = torch.empty(npr, dtype=torch.float64, device=u.device) if out is None else out
            torch.cuda.current_stream(u.device).synchronize()
            _lib.check(self._L.b200ls_divergence_device(self._h, self._dev(u, nv), self._dev(out, npr)), self._h)
            self._sync_stream(u.device)
            return out
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.empty(npr)
        assert u.size == nv
        _lib.check(self._L.b200ls_divergence(self._h, C.c_void_p(u.ctypes.data), C.c_void_p(out.ctypes.data)), self._h)
        return out

    def gradient(self, p, with_bn=False, out=None):
        """out = G p, or (BN G) p with BN = dt I   (navierstokes.cpp:442, 595)."""
        nv, npr = self.velocitySize()
        if _is_torch_cuda(p):
            import torch

            out = torch.empty(nv, dtype=torch.float64, device=p.device) if out is None else out
            torch.cuda.current_stream(p.device).synchronize()
            _lib.check(self._L.b200ls_gradient_device(self._h, self._dev(p, npr), self._dev(out, nv), int(bool(with_bn))), self._h)
            self._sync_stream(p.device)
            return out
        p = np.ascontiguousarray(p, dtype=np.float64)
        out = np.empty(nv)
        assert p.size == npr
        _lib.check(self._L.b200ls_gradient(self._h, C.c_void_p(p.ctypes.data), C.c_void_p(out.ctypes.data), int(bool(with_bn))),
                   self._h)
        return out

    def ghostedSizes(self):
        sz = (C.c_int64 * 3)()
        _lib.check(self._L.b200ls_ghosted_sizes(self._h, sz), self._h)
        return [int(v) for v in sz]

    def convection(self, qlocal, out=None):
        """out = N(q): the convection MatShell of createconvection.cpp on the ghosted local arrays qlocal[f] (one ghost
        layer on every side, i fastest; what Boundary::copyValues2LocalVecs leaves in ctx->qLocal); packed [u|v|w] out."""
        nv, _ = self.velocitySize()
        sizes = self.ghostedSizes()
        dim = 3 if sizes[2] else 2
        if _is_torch_cuda(qlocal[0]):
            import torch

            dev = qlocal[0].device
            out = torch.empty(nv, dtype=torch.float64, device=dev) if out is None else out
            torch.cuda.current_stream(dev).synchronize()
            ptrs = [self._dev(qlocal[f], sizes[f]) for f in range(dim)] + ([C.c_void_p(0)] if dim == 2 else [])
            _lib.check(self._L.b200ls_convection_device(self._h, ptrs[0], ptrs[1], ptrs[2], self._dev(out, nv)), self._h)
            self._sync_stream(dev)
            return out
        q = [np.ascontiguousarray(qlocal[f], dtype=np.float64) for f in range(dim)]
        for f in range(dim):
            assert q[f].size == sizes[f], (q[f].size, sizes[f])
        out = np.empty(nv)
        _lib.check(self._L.b200ls_convection(self._h, C.c_void_p(q[0].ctypes.data), C.c_void_p(q[1].ctypes.data),
                                             C.c_void_p(q[2].ctypes.data if dim == 3 else 0), C.c_void_p(out.ctypes.data)), self._h)
        return out

    def ghostedFromPacked(self, packed, qlocal):
        """Interior and periodic wrap layers of the ghosted device arrays from a packed device vector (DMGlobalToLocal)."""
        import torch

        nv, _ = self.velocitySize()
        sizes = self.ghostedSizes()
        dim = 3 if sizes[2] else 2
        torch.cuda.current_stream(packed.device).synchronize()
        ptrs = [self._dev(qlocal[f], sizes[f]) for f in range(dim)] + ([C.c_void_p(0)] if dim == 2 else [])
        _lib.check(self._L.b200ls_ghosted_from_packed_device(self._h, self._dev(packed, nv), ptrs[0], ptrs[1], ptrs[2]), self._h)
        self._sync_stream(packed.device)
        return qlocal

    def project(self, u, p, dp):
        """u <- u - (BN G) dp ; p <- p + dp, in place   (navierstokes.cpp:583-615)."""
        nv, npr = self.velocitySize()
        if _is_torch_cuda(u):
            import torch

            torch.cuda.current_stream(u.device).synchronize()
            _lib.check(self._L.b200ls_project_device(self._h, self._dev(u, nv), self._dev(p, npr), self._dev(dp, npr)), self._h)
            self._sync_stream(u.device)
            return u, p
        for a, m in ((u, nv), (p, npr)):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.size == m):
                raise ValueError("u and p are updated in place: contiguous float64 numpy arrays of the operator's sizes")
        dp = np.ascontiguousarray(dp, dtype=np.float64)
        _lib.check(self._L.b200ls_project(self._h, C.c_void_p(u.ctypes.data), C.c_void_p(p.ctypes.data),
                                          C.c_void_p(dp.ctypes.data)), self._h)
        return u, p

    def getIters(self) -> int:
        v = C.c_int(0)
        _lib.check(self._L.b200ls_get_iters(self._h, C.byref(v)), self._h)
        return v.value

    def getResidual(self) -> float:
        v = C.c_double(0.0)
        _lib.check(self._L.b200ls_get_residual(self._h, C.byref(v)), self._h)
        return v.value

    def getReason(self) -> int:
        v = C.c_int(0)
        _lib.check(self._L.b200ls_get_reason(self._h, C.byref(v)), self._h)
        return v.value

    def getHistory(self) -> np.ndarray:
        n = C.c_int(0)
        _lib.check(self._L.b200ls_get_history(self._h, None, 0, C.byref(n)), self._h)
        buf = np.empty(max(n.value, 1), dtype=np.float64)
        _lib.check(self._L.b200ls_get_history(self._h, buf.ctypes.data_as(_lib._dp), n.value, C.byref(n)), self._h)
        return buf[: n.value].copy()

    # ---- measurement hooks
    def timing(self):
        s, l, k = C.c_double(0), C.c_double(0), C.c_int64(0)
        _lib.check(self._L.b200ls_get_timing(self._h, C.byref(s), C.byref(l), C.byref(k)), self._h)
        e = C.c_double(0)
        _lib.check(self._L.b200ls_get_e2e_ms(self._h, C.byref(e)), self._h)
        return {"solve_ms": s.value, "loop_ms": l.value, "launches": k.value, "e2e_ms": e.value}

    def setProfile(self, enable: bool):
        _lib.check(self._L.b200ls_set_profile(self._h, int(bool(enable))), self._h)

    def profile(self, kclass: int):
        t, c = C.c_double(0), C.c_int64(0)
        _lib.check(self._L.b200ls_get_profile(self._h, int(kclass), C.byref(t), C.byref(c)), self._h)
        return t.value, c.value

    def setTrace(self, capacity: int):
        _lib.check(self._L.b200ls_set_trace(self._h, int(capacity)), self._h)

    def getTrace(self, capacity: int = 4096) -> np.ndarray:
        buf = np.zeros((capacity, 5), dtype=np.uint64)
        n = C.c_int(0)
        _lib.check(self._L.b200ls_get_trace(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), capacity, C.byref(n)), self._h)
        return buf[: min(n.value, capacity)].copy()

    def timeKernel(self, kclass: int, reps: int = 20, flush_l2: bool = True) -> float:
        v = C.c_double(0)
        _lib.check(self._L.b200ls_time_kernel(self._h, int(kclass), int(reps), int(bool(flush_l2)), C.byref(v)), self._h)
        return v.value


def createLinSolver(solverName: str, node, device: int | None = None, comm=None) -> LinSolverBase:
    """petibm::linsolver::createLinSolver (src/linsolver/linsolver.cpp:57-91) with the third branch.

    ``parameters.<name>Solver.type``: "CPU" -> LinSolverKSP and "GPU" -> LinSolverAmgX exist only inside
    PetIBM (PETSc / AmgX are not part of this package), "B200" -> LinSolverB200."""
    key = solverName + "Solver"
    params = node.get("parameters", {}).get(key, {})
    stype = str(params.get("type", "CPU"))
    config = str(params.get("config", "None"))
    if config != "None" and not config.startswith("/"):
        config = os.path.join(str(node["directory"]), config)
    if stype == "B200":
        solver = LinSolverB200(solverName, config, node=node, device=device, comm=comm)
        if "mesh" in node:
            solver.setGrid(Grid.from_config(node))
        return solver
    if stype in ("CPU", "GPU"):
        raise ValueError(f'solver type "{stype}" of "{solverName}" is PetIBM\'s own PETSc KSP / AmgX backend; '
                         'this package provides only type "B200"')
    raise ValueError(f'Unrecognized value "{stype}" of the type of the linear solver "{solverName}"')

"""TEST SCAFFOLDING: a stand-in for the `petibm_b200` module whose LinSolverB200 runs on the CPU emulation of the kernel
sources (tests/emu) instead of libb200ls.so's device path.  Its only purpose is to execute the BODIES of the `-m gpu`
tests that were written without a GPU at hand (tests/test_zzz_gpu_*.py) on a CPU-only machine, so that a typo, a wrong
tolerance or a wrong expectation in a test does not wait for a GPU box to be found (tests/test_gpu_tests_on_emulation.py).
It mirrors the dispatch of petibm_b200.linsolver.LinSolverB200.setMatrix (separable stencil -> hybrid -> line-coefficient ->
CSR) with the same host-side analyses of libb200ls.so, and the same KSP error behaviour (reason < 0 raises).  Never
imported by the product, by bench.py or by any `-m gpu` test."""
from __future__ import annotations

import ctypes as C

import numpy as np

import petibm_b200 as _real
from oracle import oracle as orc
from petibm_b200 import _lib
from petibm_b200.staggered import analyze, analyze_hybrid
from tests import test_emulated_kernels as K
from tests import test_emulated_mg as MG

B200Error = _real.B200Error
Grid = _real.Grid
Mat = _real.Mat


class _Opts:
    def __init__(self):
        self.ksp_type, self.pc_type, self.norm_type = 0, 0, 1
        self.max_it, self.rtol, self.atol, self.divtol = 10000, 1e-5, 1e-50, 1e4
        self.mg_levels, self.mg_smooth_its, self.mg_coarse_its = 0, 2, 16


_NAMES = {"cg": 0, "bcgs": 1, "preonly": 2, "none": 0, "jacobi": 1, "mg": 2, "lu": 3, "preconditioned": 1, "unpreconditioned": 2, "natural": 3}


def make_module(emu):
    """Returns an object with the attributes of `petibm_b200` the tests use, bound to the loaded emulation library."""

    class LinSolverB200:
        def __init__(self, solverName, file, **_):
            self.name, self.config = solverName, file
            self._o = _Opts()
            self._grid = None
            self._staggered = True
            self.operator = None
            self.nlocal = 0
            self._M = None
            self._null = (False, None)
            self._res = None
            self._tuning = {}
            if file and file != "None":
                L = _lib.lib()
                o = _lib.Options()
                L.b200ls_default_options(C.byref(o))
                err = C.create_string_buffer(512)
                rc = L.b200ls_parse_options(open(file).read().encode(), (solverName + "_").encode(), C.byref(o), err, 512)
                if rc != 0:
                    raise B200Error(rc, err.value.decode())
                for k in vars(self._o):
                    setattr(self._o, k, getattr(o, k))

        # ---- options
        def options(self):
            return self._o

        def setOptions(self, **kw):
            for k, v in kw.items():
                setattr(self._o, k, _NAMES[v] if isinstance(v, str) else v)

        def setTuning(self, key, value):
            self._tuning[key] = int(value)

        def _tiles(self, bcgs=False):
            """The kernel choice of sep_tile_xr (sep_solver.inc) for the line-coefficient operator."""
            xr = self._tuning.get("sep_tile", -1)
            if xr < 0 and not bcgs:
                xr = 0
            if xr < 0:
                big = sum(d[0] * d[1] * d[2] for d in self._dims) >= (1 << 20) if self.operator == "staggered" else False
                xr = 2 if big and not any(self._grid.periodic) and any(d[2] >= 8 and d[0] >= 48 for d in self._dims) else 0
            return K._sep_tiles(emu, 2 if xr == 2 else 0, self._tuning.get("sep_zchunk", 0), stages=self._tuning.get("sep_stages", 3))

        # ---- operator
        def setGrid(self, grid):
            self._grid = grid

        def setStaggered(self, enable):
            self._staggered = bool(enable)

        def setStencil(self, grid):
            self._grid = grid
            self.operator = "stencil"
            self.nlocal = grid.size

        def setNullSpace(self, has_const, vecs=None):
            self._null = (bool(has_const), None if vecs is None else np.ascontiguousarray(np.atleast_2d(vecs))[0])

        def _velocity_dims(self):
            g = self._grid
            n = list(g.n) + [1] * (3 - g.dim)
            per = [int(bool(p)) for p in g.periodic][:3]
            return [[n[d] - (1 if (d == f and not per[d]) else 0) for d in range(3)] for f in range(g.dim)], per

        def setMatrix(self, A):
            if not isinstance(A, Mat):
                A = Mat.from_scipy(A)
            import scipy.sparse as sp

            self._M = sp.csr_matrix((A.data, A.indices, A.indptr), shape=(A.nrows, A.ncols))
            self._M.sort_indices()
            self.nlocal = A.nrows
            self.operator = "csr"
            self._dims = None
            g = self._grid
            if g is not None and A.nrows == g.size:
                ref = orc.assemble_dbng(g.widths, [int(p) for p in g.periodic][: g.dim] + [0] * (3 - g.dim), g.dt)
                rp, col, val = ref.arrays()
                if np.array_equal(rp, A.indptr) and np.array_equal(col, A.indices) and np.array_equal(val, A.data):
                    self.operator = "stencil"
            elif g is not None and self._staggered:
                per = [int(bool(p)) for p in g.periodic][:3]
                if A.nrows > g.size:
                    try:
                        analyze_hybrid(g.widths, per[: g.dim], g.dt, A.indptr, A.indices, A.data)
                        self.operator = "hybrid"
                    except B200Error:
                        pass
                if self.operator == "csr":
                    vel, per3 = self._velocity_dims()
                    layouts = []
                    if A.nrows == sum(int(np.prod(v)) for v in vel):
                        layouts.append(vel)
                    if A.nrows > g.size:
                        layouts.append([list(g.n) + [1] * (3 - g.dim)])
                    for dims in layouts:
                        try:
                            analyze(dims, per3, A.indptr, A.indices, A.data)
                            self.operator, self._dims = "staggered", dims
                            break
                        except B200Error:
                            pass
            self._null = (A.null_has_const, None if A.null_vecs is None else A.null_vecs[0])

        # ---- y = A x
        def apply(self, x):
            x = np.ascontiguousarray(x, dtype=np.float64)
            g = self._grid
            if self.operator == "stencil":
                return K._apply(emu, g.widths, [int(p) for p in g.periodic][: g.dim], x)
            if self.operator == "csr":
                return orc.Csr.from_arrays(self._M.shape[0], self._M.shape[1], self._M.indptr, self._M.indices, self._M.data).spmv(x)
            per = [int(bool(p)) for p in g.periodic][:3]
            with self._tiles():
                if self.operator == "hybrid":
                    return K._sep_solve(emu, None, per[: g.dim], self._M, x, mode="apply", hybrid_widths=g.widths, dt=g.dt)[0]
                return K._sep_solve(emu, self._dims, per, self._M, x, mode="apply")[0]

        # ---- KSPSolve
        def solve(self, x, b):
            o, g = self._o, self._grid
            pc = {0: "none", 1: "jacobi", 2: "mg", 3: "lu"}[o.pc_type]
            has_const, nv = self._null
            kw = dict(rtol=o.rtol, atol=o.atol, max_it=o.max_it)
            if (o.ksp_type == 2) != (pc == "lu"):
                raise B200Error(-3, "ksp_type preonly and pc_type lu are only implemented together")
            if o.ksp_type == 2:
                if self.operator != "csr":
                    raise B200Error(-3, "the direct solve works on a small assembled matrix kept as CSR")
                rc, X = K._dense_solve(emu, self._M, b)
                x[...] = X[0]
                self._res = (np.zeros(0), 1 if rc == 0 else 0, 4 if rc == 0 else -11)
                if rc != 0:
                    raise B200Error(-5, "diverged: KSPConvergedReason -11")
                return x
            if self.operator == "stencil":
                per = [int(p) for p in g.periodic][: g.dim]
                if o.ksp_type != 0:
                    raise B200Error(-3, "the separable stencil operator is solved with cg")
                if pc == "mg":
                    xs, hist, its, reason, _ = MG._mg(emu, g.widths, per, b, has_const=has_const, levels=o.mg_levels,
                                                      smooth=o.mg_smooth_its, coarse=o.mg_coarse_its, **kw)
                else:
                    xs, hist, its, reason, _ = K._cg(emu, g.widths, per, b, pc=pc, has_const=has_const, norm=o.norm_type, **kw)
            elif pc == "mg":
                if self.operator != "hybrid":
                    raise B200Error(-3, "pc_type mg needs the pressure operator of the mesh")
                xs, hist, its, reason = MG._hybrid_mg(emu, g.widths, self._M, b, nullvec=nv, has_const=has_const,
                                                      smooth=o.mg_smooth_its, coarse=o.mg_coarse_its, dt=g.dt, **kw)
            elif self.operator == "csr":
                if o.ksp_type == 1 and (has_const or nv is not None):
                    raise B200Error(-3, "bcgs with a null space attached is not supported")
                xs, hist, its, reason = K._csr_solve(emu, self._M, b, bcgs=o.ksp_type == 1, pc=pc, has_const=has_const, nullvec=nv, **kw)
            else:
                per = [int(bool(p)) for p in g.periodic][:3]
                mode = "bcgs" if o.ksp_type == 1 else "cg"
                with self._tiles(bcgs=(mode == "bcgs")):
                    if self.operator == "hybrid":
                        xs, hist, its, reason = K._sep_solve(emu, None, per[: g.dim], self._M, b, mode=mode, pc=pc, has_const=has_const,
                                                             nullvec=nv, hybrid_widths=g.widths, dt=g.dt, **kw)
                    else:
                        xs, hist, its, reason = K._sep_solve(emu, self._dims, per, self._M, b, mode=mode, pc=pc, has_const=has_const,
                                                             nullvec=nv, **kw)
            x[...] = xs
            self._res = (hist, its, reason)
            if reason < 0:
                raise B200Error(-5, f"diverged: KSPConvergedReason {reason}")
            return x

        def getHistory(self):
            return self._res[0].copy()

        def getIters(self):
            return self._res[1]

        def getReason(self):
            return self._res[2]

        def getResidual(self):
            return float(self._res[0][-1]) if self._res[0].size else 0.0

        # ---- operators around the solve
        def velocitySize(self):
            vel, _ = self._velocity_dims()
            return sum(int(np.prod(v)) for v in vel), self._grid.size

        def _ops(self, mode, vin, io, io2=None):
            g = self._grid
            dim, n, p, w, dz = K._grid_args(g.widths, [int(q) for q in g.periodic][: g.dim])
            emu.emu_stag_ops.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int64), K._ip, K._dp, K._dp, K._dp, C.c_double, K._dp, K._dp,
                                         K._dp]
            vin = np.ascontiguousarray(vin, dtype=np.float64)
            emu.emu_stag_ops(mode, dim, n, p, w[0].ctypes.data_as(K._dp), w[1].ctypes.data_as(K._dp), dz, float(g.dt),
                             vin.ctypes.data_as(K._dp), io.ctypes.data_as(K._dp), None if io2 is None else io2.ctypes.data_as(K._dp))
            return io

        def divergence(self, u):
            return self._ops(0, u, np.empty(self._grid.size))

        def gradient(self, p, with_bn=False):
            return self._ops(2 if with_bn else 1, p, np.empty(self.velocitySize()[0]))

        def project(self, u, p, dp):
            self._ops(3, dp, u, p)
            return u, p

        def destroy(self):
            pass

    class _Module:
        pass

    m = _Module()
    m.LinSolverB200, m.Mat, m.Grid, m.B200Error = LinSolverB200, Mat, Grid, B200Error
    return m

"""ctypes front-end of the CPU oracle (oracle/petibm_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of petibm_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (petibm_b200) never does.

Parity status: grid/operator part pinned against the reference's mesh golden
vectors; KSP part "parity unpinned" (PETSc 3.16 is not available, restated from
its published algorithm).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "petibm_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)
    return _LIB_PATH


class OrcKspOpts(C.Structure):
    _fields_ = [
        ("ksp_type", C.c_int),
        ("pc_type", C.c_int),
        ("norm_type", C.c_int),
        ("max_it", C.c_int),
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("divtol", C.c_double),
        ("has_const_nullspace", C.c_int),
        ("n_nullvecs", C.c_int),
    ]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    try:
        L = C.CDLL(_LIB_PATH)
    except OSError:
        build(force=True)
        L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    L.orc_stretch_grid.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double, _dp]
    L.orc_axis_from_subdomains.argtypes = [C.c_double, C.c_int, _dp, _ip, _dp, _dp]
    L.orc_axis_from_subdomains.restype = C.c_int
    L.orc_pressure_coord.argtypes = [C.c_int, C.c_double, _dp, _dp]
    L.orc_vertex_coord.argtypes = [C.c_int, C.c_double, _dp, _dp]
    L.orc_velocity_axis.argtypes = [C.c_int, C.c_double, C.c_double, _dp, C.c_int, C.c_int, _dp, _dp, _ip]
    L.orc_velocity_axis.restype = C.c_int
    for name in ("orc_csr_nrows", "orc_csr_ncols", "orc_csr_nnz"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int64
    L.orc_csr_free.argtypes = [vp]
    L.orc_csr_export.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int32), _dp]
    L.orc_csr_import.argtypes = [C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32), _dp]
    L.orc_csr_import.restype = vp
    L.orc_assemble_divergence.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, _dp]
    L.orc_assemble_divergence.restype = vp
    L.orc_assemble_gradient.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp]
    L.orc_assemble_gradient.restype = vp
    L.orc_bnhead_order1.argtypes = [C.c_int64, C.c_double]
    L.orc_bnhead_order1.restype = vp
    L.orc_bnhead.argtypes = [vp, C.c_double, C.c_double, C.c_int]
    L.orc_bnhead.restype = vp
    L.orc_matmatmult.argtypes = [vp, vp]
    L.orc_matmatmult.restype = vp
    L.orc_assemble_dbng_literal.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_double, _dp]
    L.orc_assemble_dbng_literal.restype = vp
    L.orc_assemble_dbng_closed.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_double]
    L.orc_assemble_dbng_closed.restype = vp
    L.orc_convection.argtypes = [C.c_int, _ip, C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp)]
    L.orc_convection.restype = None
    L.orc_set_fast.argtypes = [C.c_int, C.c_int]
    L.orc_max_threads.restype = C.c_int
    L.orc_spmv.argtypes = [vp, _dp, _dp]
    L.orc_ksp_solve.argtypes = [vp, C.POINTER(OrcKspOpts), _dp, _dp, _dp, _ip, _dp, _dp, C.c_int, _ip]
    L.orc_ksp_solve.restype = C.c_int
    _lib = L
    return L


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


# --------------------------------------------------------------------------- grid
def stretch_grid(bg, ed, n, r):
    out = np.empty(n)
    lib().orc_stretch_grid(bg, ed, n, r, out.ctypes.data_as(_dp))
    return out


def axis_from_subdomains(start, subdomains):
    """subdomains: list of dicts {end, cells, stretchRatio} (the YAML of parser.cpp:334-356)."""
    ends, pe = _d([s["end"] for s in subdomains])
    cells, pc = _i([s["cells"] for s in subdomains])
    ratios, pr = _d([s["stretchRatio"] for s in subdomains])
    out = np.empty(int(cells.sum()))
    n = lib().orc_axis_from_subdomains(float(start), len(subdomains), pe, pc, pr, out.ctypes.data_as(_dp))
    assert n == out.size
    return out


def pressure_coord(dLp, vmin):
    dLp, p = _d(dLp)
    out = np.empty(dLp.size)
    lib().orc_pressure_coord(dLp.size, vmin, p, out.ctypes.data_as(_dp))
    return out


def vertex_coord(dLp, vmin):
    dLp, p = _d(dLp)
    out = np.empty(dLp.size + 1)
    lib().orc_vertex_coord(dLp.size, vmin, p, out.ctypes.data_as(_dp))
    return out


def velocity_axis(dLp, vmin, vmax, same_dir, periodic):
    """Returns (n_valid, dLTrue, coordTrue) of cartesianmesh.cpp:212-355 for one (comp, dir)."""
    dLp, p = _d(dLp)
    dl = np.empty(dLp.size + 3)
    co = np.empty(dLp.size + 3)
    ln = C.c_int(0)
    n = lib().orc_velocity_axis(dLp.size, vmin, vmax, p, int(same_dir), int(periodic),
                                dl.ctypes.data_as(_dp), co.ctypes.data_as(_dp), C.byref(ln))
    return n, dl[: ln.value].copy(), co[: ln.value].copy()


def field_sizes(n, periodic):
    """Points of the velocity fields per direction (cartesianmesh.cpp:212-355: n - 1 along the field's own direction,
    n when that direction is periodic; the pressure count otherwise)."""
    dim = len(n)
    return [[int(n[d]) - (1 if (d == f and not periodic[d]) else 0) for d in range(dim)] for f in range(dim)]


def field_spacings(widths, periodic):
    """mesh->dL[f][d] (interior part) for every field and direction, through orc_velocity_axis."""
    dim = len(widths)
    out = []
    for f in range(dim):
        row = []
        for d in range(dim):
            w = np.asarray(widths[d], dtype=np.float64)
            nv, dl, _ = velocity_axis(w, 0.0, float(w.sum()), f == d, bool(periodic[d]))
            row.append(np.ascontiguousarray(dl[1: 1 + nv]))
        out.append(row)
    return out


def convection(widths, periodic, qlocal):
    """N(q) of createconvection.cpp on the ghosted local arrays qlocal[f] (shape (nz_f+2, ny_f+2, nx_f+2), 2-D:
    (ny_f+2, nx_f+2)); returns the list of interior fields (the blocks of the packed vector)."""
    dim = len(widths)
    n = [len(w) for w in widths]
    nf = field_sizes(n, periodic)
    dL = field_spacings(widths, periodic)
    nf_flat = np.zeros(9, dtype=np.int32)
    PD = _dp * 9
    PQ = _dp * 3
    dl_ptrs, q_ptrs, o_ptrs = PD(), PQ(), PQ()
    keep, outs = [], []
    for f in range(dim):
        for d in range(dim):
            nf_flat[3 * f + d] = nf[f][d]
            dl_ptrs[3 * f + d] = dL[f][d].ctypes.data_as(_dp)
        q = np.ascontiguousarray(qlocal[f], dtype=np.float64)
        assert q.shape == tuple(m + 2 for m in reversed(nf[f])), (q.shape, nf[f])
        keep.append(q)
        q_ptrs[f] = q.ctypes.data_as(_dp)
        o = np.empty(tuple(reversed(nf[f])))
        outs.append(o)
        o_ptrs[f] = o.ctypes.data_as(_dp)
    lib().orc_convection(dim, nf_flat.ctypes.data_as(_ip), dl_ptrs, q_ptrs, o_ptrs)
    return outs


# ---------------------------------------------------------------------------- CSR
class Csr:
    """Owning handle on an orc_csr."""

    def __init__(self, handle):
        if not handle:
            raise ValueError("oracle returned a NULL matrix")
        self.h = C.c_void_p(handle)

    def __del__(self):
        try:
            if self.h:
                lib().orc_csr_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def shape(self):
        return (lib().orc_csr_nrows(self.h), lib().orc_csr_ncols(self.h))

    @property
    def nnz(self):
        return lib().orc_csr_nnz(self.h)

    def arrays(self):
        n, nnz = self.shape[0], self.nnz
        rp = np.empty(n + 1, dtype=np.int64)
        col = np.empty(nnz, dtype=np.int32)
        val = np.empty(nnz, dtype=np.float64)
        lib().orc_csr_export(self.h, rp.ctypes.data_as(C.POINTER(C.c_int64)),
                             col.ctypes.data_as(C.POINTER(C.c_int32)), val.ctypes.data_as(_dp))
        return rp, col, val

    def to_scipy(self):
        import scipy.sparse as sp

        rp, col, val = self.arrays()
        return sp.csr_matrix((val, col, rp), shape=self.shape)

    def spmv(self, x):
        x, px = _d(x)
        y = np.empty(self.shape[0])
        lib().orc_spmv(self.h, px, y.ctypes.data_as(_dp))
        return y

    @staticmethod
    def from_arrays(nrows, ncols, rowptr, col, val):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        return Csr(lib().orc_csr_import(nrows, ncols, rowptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                        col.ctypes.data_as(C.POINTER(C.c_int32)), val.ctypes.data_as(_dp)))


def _grid_args(dim, widths, periodic):
    np_, pn = _i([len(w) for w in widths[:dim]] + [1] * (3 - dim))
    per, pp = _i(list(periodic[:dim]) + [0] * (3 - dim))
    arrs = [_d(w) for w in widths[:dim]] + [_d([1.0])] * (3 - dim)
    return (np_, pn, per, pp, arrs)


def assemble_divergence(widths, periodic, a0=None):
    dim = len(widths)
    np_, pn, per, pp, arrs = _grid_args(dim, widths, periodic)
    a0p = None
    if a0 is not None:
        a0, a0p = _d(a0)
    return Csr(lib().orc_assemble_divergence(dim, pn, pp, arrs[0][1], arrs[1][1], arrs[2][1], a0p))


def assemble_gradient(widths, periodic):
    dim = len(widths)
    np_, pn, per, pp, arrs = _grid_args(dim, widths, periodic)
    return Csr(lib().orc_assemble_gradient(dim, pn, pp, arrs[0][1], arrs[1][1], arrs[2][1]))


def bnhead_order1(n, dt):
    return Csr(lib().orc_bnhead_order1(n, dt))


def bnhead(Op: Csr, dt, coeff, N):
    """createBnHead for any order N (createbn.cpp:19-95)."""
    return Csr(lib().orc_bnhead(Op.h, float(dt), float(coeff), int(N)))


def matmatmult(A: Csr, B: Csr):
    return Csr(lib().orc_matmatmult(A.h, B.h))


def assemble_dbng(widths, periodic, dt, a0=None, literal=True):
    """The pressure-Poisson matrix D*(dt*G) of navierstokes.cpp:347-356.

    literal=True runs the reference's assembly pipeline (D, G, BN, two MatMatMult);
    literal=False uses the closed form (only valid with a0 == 0)."""
    dim = len(widths)
    np_, pn, per, pp, arrs = _grid_args(dim, widths, periodic)
    if literal:
        a0p = None
        if a0 is not None:
            a0, a0p = _d(a0)
        return Csr(lib().orc_assemble_dbng_literal(dim, pn, pp, arrs[0][1], arrs[1][1], arrs[2][1], dt, a0p))
    assert a0 is None or not np.any(np.asarray(a0))
    return Csr(lib().orc_assemble_dbng_closed(dim, pn, pp, arrs[0][1], arrs[1][1], arrs[2][1], dt))


# ---------------------------------------------------------------------------- KSP
KSP_TYPES = {"cg": 0, "bcgs": 1}
PC_TYPES = {"none": 0, "jacobi": 1}
NORM_TYPES = {"none": 0, "preconditioned": 1, "unpreconditioned": 2, "natural": 3}
REASONS = {
    2: "CONVERGED_RTOL", 3: "CONVERGED_ATOL", 4: "CONVERGED_ITS",
    -3: "DIVERGED_ITS", -4: "DIVERGED_DTOL", -5: "DIVERGED_BREAKDOWN",
    -8: "DIVERGED_INDEFINITE_PC", -9: "DIVERGED_NANORINF", -10: "DIVERGED_INDEFINITE_MAT",
}


@dataclass
class KspResult:
    x: np.ndarray
    its: int
    rnorm: float
    reason: int
    history: np.ndarray = field(repr=False)


def set_fast(fast: bool, nthreads: int = 0):
    """fast=False: strict serial summation order (parity); True: OpenMP + SIMD reductions (timing)."""
    lib().orc_set_fast(int(fast), int(nthreads))


def max_threads():
    return lib().orc_max_threads()


def ksp_solve(A: Csr, b, ksp_type="cg", pc_type="none", norm_type="preconditioned",
              rtol=1e-5, atol=1e-50, divtol=1e4, max_it=10000, const_nullspace=False,
              nullvecs=None) -> KspResult:
    """KSPSolve as LinSolverKSP::solve issues it (linsolverksp.cpp:85-105), zero initial guess."""
    b, pb = _d(b)
    n = A.shape[0]
    assert b.size == n
    nv = 0
    pv = None
    if nullvecs is not None:
        nullvecs, pv = _d(np.atleast_2d(nullvecs))
        nv = nullvecs.shape[0]
    o = OrcKspOpts(KSP_TYPES[ksp_type], PC_TYPES[pc_type], NORM_TYPES[norm_type], int(max_it),
                   float(rtol), float(atol), float(divtol), int(bool(const_nullspace)), nv)
    x = np.empty(n)
    hist = np.empty(int(max_it) + 2)
    its, hl = C.c_int(0), C.c_int(0)
    rn = C.c_double(0.0)
    reason = lib().orc_ksp_solve(A.h, C.byref(o), pv, pb, x.ctypes.data_as(_dp), C.byref(its),
                                 C.byref(rn), hist.ctypes.data_as(_dp), hist.size, C.byref(hl))
    return KspResult(x, its.value, rn.value, reason, hist[: min(hl.value, hist.size)].copy())

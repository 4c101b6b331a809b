"""GPU parity of the line-coefficient operator (b200ls_set_staggered, sep_kernels.cuh): the velocity system
A = I/dt - c nu L with BiCGStab + Jacobi and IBPM's modified Poisson system
(row a11) without storing the stencil part of the matrix.  The same kernels run on the CPU emulation in
tests/test_emulated_kernels.py; here they run on the device, through the C ABI, against the oracle and against the
plain-CSR path of the same library (identical row sums => identical histories)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import petibm_b200

    return petibm_b200


def _pair(pb, name, grid, M, tile=0, **opts):
    """Two solvers on the same matrix: line-coefficient form and plain CSR.  tile = 0: the row-per-thread kernels of
    sep_kernels.cuh, whose dot products are summed like those of the CSR kernels (bit-identical histories); 2: the
    tiled plane-marching kernels of sep_tile.cuh (same row sums bit for bit, dot products in another order)."""
    out = []
    for staggered in (True, False):
        s = pb.LinSolverB200(name, "None")
        s.setOptions(**opts)
        s.setGrid(grid)
        s.setStaggered(staggered)
        s.setTuning("sep_tile", tile)
        out.append(s)
    return out


@pytest.mark.parametrize("shape,per", [((14, 12, 10), (0, 0, 0)), ((11, 9), (0, 0)), ((9, 8, 7), (1, 0, 1)), ((70, 5, 4), (0, 1, 0))])
def test_velocity_system_spmv_bit_exact_and_bcgs(pb, shape, per):
    widths = H.make_widths(shape)
    A, _ = H.velocity_system(widths, per, dt=0.01, nu=0.01, c=0.5)
    Ao = orc.Csr.from_arrays(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)
    opts = dict(ksp_type="bcgs", pc_type="jacobi", rtol=0.0, atol=1e-9, max_it=500)
    s, c = _pair(pb, "velocity", H.grid_of(widths, per), A, **opts)
    s.setMatrix(pb.Mat.from_scipy(A))
    c.setMatrix(pb.Mat.from_scipy(A))
    assert s.operator == "staggered" and c.operator == "csr"
    rng = np.random.default_rng(1)
    x = rng.standard_normal(A.shape[0])
    y = s.apply(x)
    assert np.array_equal(y, Ao.spmv(x)) and np.array_equal(y, c.apply(x))
    b = rng.standard_normal(A.shape[0])
    xs, xc = np.empty_like(b), np.empty_like(b)
    s.solve(xs, b)
    c.solve(xc, b)
    ref = orc.ksp_solve(Ao, b, ksp_type="bcgs", pc_type="jacobi", rtol=0.0, atol=1e-9, max_it=500)
    assert s.getReason() == c.getReason() == ref.reason == 3
    assert s.getIters() == c.getIters() and np.array_equal(s.getHistory(), c.getHistory()) and np.array_equal(xs, xc)
    assert abs(s.getIters() - ref.its) <= 2
    m = min(6, ref.history.size, s.getHistory().size)
    np.testing.assert_allclose(s.getHistory()[:m], ref.history[:m], rtol=1e-9)
    np.testing.assert_allclose(xs, ref.x, rtol=0, atol=1e-6 * np.abs(ref.x).max())
    s.destroy(); c.destroy()


@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_ibpm_modified_poisson_stencil_block_plus_remainder(pb, pc):
    shape, nf = (30, 26), 14
    widths = H.make_widths(shape)
    G = orc.assemble_gradient(widths, [0, 0, 0]).to_scipy()
    rng = np.random.default_rng(9)
    rows = rng.integers(0, G.shape[0], size=(nf, 16))
    R = sp.csr_matrix((rng.uniform(0.01, 1.0, nf * 16), (rows.ravel(), np.repeat(np.arange(nf), 16))), shape=(G.shape[0], nf))
    K = sp.hstack([G, -R]).tocsr()
    M = (-(K.T @ K) * 0.01).tocsr()
    M.sort_indices()
    pN = G.shape[1]
    nv = np.zeros(M.shape[0]); nv[:pN] = 1.0 / np.sqrt(pN)
    Mo = orc.Csr.from_arrays(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)
    xs = rng.standard_normal(M.shape[0]); xs -= (xs @ nv) * nv
    b = M @ xs
    nit = 40
    s, c = _pair(pb, "poisson", pb.Grid(widths, (False, False, False), 0.01), M, pc_type=pc, rtol=0.0, atol=0.0, max_it=nit)
    s.setMatrix(pb.Mat.from_scipy(M).setNullSpace(False, nv))
    c.setMatrix(pb.Mat.from_scipy(M).setNullSpace(False, nv))
    assert s.operator == "staggered" and c.operator == "csr" and s.nlocal == pN + nf
    assert np.array_equal(s.apply(xs), Mo.spmv(xs))
    ref = orc.ksp_solve(Mo, b, pc_type=pc, rtol=0.0, atol=0.0, max_it=nit, nullvecs=nv)
    x, xc = np.empty_like(b), np.empty_like(b)
    for solver, out in ((s, x), (c, xc)):
        with pytest.raises(pb.B200Error):
            solver.solve(out, b)
    assert np.array_equal(s.getHistory(), c.getHistory()) and np.array_equal(x, xc)
    np.testing.assert_allclose(s.getHistory(), ref.history, rtol=1e-10)
    np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-9 * np.abs(ref.x).max())
    s.destroy(); c.destroy()


def test_a_matrix_without_the_structure_keeps_the_csr_operator(pb):
    shape, per = (10, 9, 8), (0, 0, 0)
    widths = H.make_widths(shape)
    A, _ = H.velocity_system(widths, per)
    B = A.copy()
    B.data[B.indptr[33] + 1] *= 1.0 + 1e-13
    s = pb.LinSolverB200("velocity", "None")
    s.setGrid(H.grid_of(widths, per))
    s.setMatrix(pb.Mat.from_scipy(B))
    assert s.operator == "csr"
    s.setMatrix(pb.Mat.from_scipy(A))          # and the operator can be replaced by the structured one afterwards
    assert s.operator == "staggered"
    s.destroy()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("dim", [2, 3])
def test_hybrid_operator_on_a_stretched_ibpm_system(pb, dim):
    """[D;E] BN [G,-H] assembled like PetIBM does (tests/helpers.ibpm_system) on a stretched grid: setMatrix recognises the
    pressure operator of the mesh plus a remainder ("hybrid"), SpMV and CG are those of the CSR kernels bit for bit
    (the multigrid block preconditioner on this operator: tests/test_zzz_gpu_4_multigrid.py)."""
    n_side, n_band = (6, 10) if dim == 3 else (12, 24)
    sub = [{"end": 0.6, "cells": n_side, "stretchRatio": 1.0 / 1.2}, {"end": 1.4, "cells": n_band, "stretchRatio": 1.0},
           {"end": 2.0, "cells": n_side, "stretchRatio": 1.2}]
    w = orc.axis_from_subdomains(0.0, sub)
    widths = [w.copy() for _ in range(dim)]
    M, pN, nv = H.ibpm_system(widths, dt=0.01, nb=14)
    Mo = orc.Csr.from_arrays(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)
    rng = np.random.default_rng(4)
    xs = rng.standard_normal(M.shape[0]); xs -= (xs @ nv) * nv
    b = Mo.spmv(xs)
    grid = pb.Grid(widths, (False, False, False), 0.01)
    nit = 20
    s, c = _pair(pb, "poisson", grid, M, pc_type="jacobi", rtol=0.0, atol=0.0, max_it=nit)
    s.setMatrix(pb.Mat.from_scipy(M).setNullSpace(False, nv))
    c.setMatrix(pb.Mat.from_scipy(M).setNullSpace(False, nv))
    assert s.operator == "hybrid" and c.operator == "csr"
    assert np.array_equal(s.apply(xs), b) and np.array_equal(c.apply(xs), b)
    x, xc = np.empty_like(b), np.empty_like(b)
    for solver, out in ((s, x), (c, xc)):
        with pytest.raises(pb.B200Error):
            solver.solve(out, b)
    assert np.array_equal(s.getHistory(), c.getHistory()) and np.array_equal(x, xc)
    ref = orc.ksp_solve(Mo, b, pc_type="jacobi", rtol=0.0, atol=0.0, max_it=nit, nullvecs=nv)
    np.testing.assert_allclose(s.getHistory()[:8], ref.history[:8], rtol=1e-10)
    s.destroy(); c.destroy()


@pytest.mark.parametrize("zchunk,stages", [(0, 3), (5, 3), (0, 4), (3, 4)])
@pytest.mark.parametrize("shape,per", [((70, 19, 12), (0, 0, 0)), ((9, 8, 7), (1, 0, 1)), ((40, 20, 16), (0, 1, 0)), ((66, 17), (0, 1))])
def test_tiled_kernels_velocity_system(pb, shape, per, zchunk, stages):
    """sep_tile.cuh on the device: several tiles in x and y, ragged edges, one and many z chunks, three and four stages in
    the per-thread cp.async queue, periodic axes (wrapped cells and nothing else go through sep_row).  y = A x is bit-identical to MatMult_SeqAIJ on the assembled matrix;
    BiCGStab + Jacobi follows the row-per-thread kernels (dot products summed in another order: 1e-9 over the first
    entries) and the oracle."""
    widths = H.make_widths(shape)
    A, _ = H.velocity_system(widths, per, dt=0.01, nu=0.01, c=0.5)
    Ao = orc.Csr.from_arrays(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)
    opts = dict(ksp_type="bcgs", pc_type="jacobi", rtol=0.0, atol=1e-9, max_it=500)
    t, _c = _pair(pb, "velocity", H.grid_of(widths, per), A, tile=2, **opts)
    s, _c2 = _pair(pb, "velocity", H.grid_of(widths, per), A, tile=0, **opts)
    _c.destroy(); _c2.destroy()
    t.setTuning("sep_zchunk", zchunk)
    t.setTuning("sep_stages", stages)
    t.setMatrix(pb.Mat.from_scipy(A))
    s.setMatrix(pb.Mat.from_scipy(A))
    assert t.operator == s.operator == "staggered"
    rng = np.random.default_rng(1)
    for _ in range(2):
        x = rng.standard_normal(A.shape[0])
        assert np.array_equal(t.apply(x), Ao.spmv(x))
    b = rng.standard_normal(A.shape[0])
    xt, xs = np.empty_like(b), np.empty_like(b)
    t.solve(xt, b)
    s.solve(xs, b)
    ref = orc.ksp_solve(Ao, b, ksp_type="bcgs", pc_type="jacobi", rtol=0.0, atol=1e-9, max_it=500)
    assert t.getReason() == s.getReason() == ref.reason == 3
    assert abs(t.getIters() - s.getIters()) <= 1 and abs(t.getIters() - ref.its) <= 2
    m = min(6, ref.history.size, t.getHistory().size, s.getHistory().size)
    np.testing.assert_allclose(t.getHistory()[:m], s.getHistory()[:m], rtol=1e-9)
    np.testing.assert_allclose(t.getHistory()[:m], ref.history[:m], rtol=1e-9)
    np.testing.assert_allclose(xt, ref.x, rtol=0, atol=1e-6 * np.abs(ref.x).max())
    t.destroy(); s.destroy()


@pytest.mark.parametrize("stages", [3, 4])
@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_tiled_kernels_ibpm_like_system_with_remainder(pb, pc, stages):
    """Stencil block + remainder rows and columns, CG with the explicit null-space vector (ibpm.cpp:251-267) on the tiled
    kernels: the remainder entries of a stencil row follow its stencil terms, the rows behind the block take sep_row."""
    shape, nf = (70, 26), 14
    widths = H.make_widths(shape)
    G = orc.assemble_gradient(widths, [0, 0, 0]).to_scipy()
    rng = np.random.default_rng(9)
    rows = rng.integers(0, G.shape[0], size=(nf, 16))
    R = sp.csr_matrix((rng.uniform(0.01, 1.0, nf * 16), (rows.ravel(), np.repeat(np.arange(nf), 16))), shape=(G.shape[0], nf))
    K = sp.hstack([G, -R]).tocsr()
    M = (-(K.T @ K) * 0.01).tocsr()
    M.sort_indices()
    pN = G.shape[1]
    nv = np.zeros(M.shape[0]); nv[:pN] = 1.0 / np.sqrt(pN)
    Mo = orc.Csr.from_arrays(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)
    xs = rng.standard_normal(M.shape[0]); xs -= (xs @ nv) * nv
    b = M @ xs
    nit = 40
    t, c = _pair(pb, "poisson", pb.Grid(widths, (False, False, False), 0.01), M, tile=2, pc_type=pc, rtol=0.0, atol=0.0, max_it=nit)
    c.destroy()
    t.setTuning("sep_stages", stages)
    t.setMatrix(pb.Mat.from_scipy(M).setNullSpace(False, nv))
    assert t.operator == "staggered"
    assert np.array_equal(t.apply(xs), Mo.spmv(xs))
    ref = orc.ksp_solve(Mo, b, pc_type=pc, rtol=0.0, atol=0.0, max_it=nit, nullvecs=nv)
    x = np.empty_like(b)
    with pytest.raises(pb.B200Error):
        t.solve(x, b)
    np.testing.assert_allclose(t.getHistory(), ref.history, rtol=1e-10)
    np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-9 * np.abs(ref.x).max())
    t.destroy()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("stages", [3, 4])
def test_tiled_kernels_hybrid_operator_on_a_stretched_ibpm_system(pb, stages):
    """The hybrid form (pressure block of a stretched grid with its face areas + remainder) on the tiled kernels."""
    sub = [{"end": 0.6, "cells": 6, "stretchRatio": 1.0 / 1.2}, {"end": 1.4, "cells": 10, "stretchRatio": 1.0},
           {"end": 2.0, "cells": 6, "stretchRatio": 1.2}]
    w = orc.axis_from_subdomains(0.0, sub)
    widths = [w.copy() for _ in range(3)]
    M, pN, nv = H.ibpm_system(widths, dt=0.01, nb=14)
    Mo = orc.Csr.from_arrays(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)
    rng = np.random.default_rng(4)
    xs = rng.standard_normal(M.shape[0]); xs -= (xs @ nv) * nv
    b = Mo.spmv(xs)
    nit = 20
    t, c = _pair(pb, "poisson", pb.Grid(widths, (False, False, False), 0.01), M, tile=2, pc_type="jacobi", rtol=0.0, atol=0.0, max_it=nit)
    c.destroy()
    t.setTuning("sep_stages", stages)
    t.setMatrix(pb.Mat.from_scipy(M).setNullSpace(False, nv))
    assert t.operator == "hybrid"
    assert np.array_equal(t.apply(xs), b)
    x = np.empty_like(b)
    with pytest.raises(pb.B200Error):
        t.solve(x, b)
    ref = orc.ksp_solve(Mo, b, pc_type="jacobi", rtol=0.0, atol=0.0, max_it=nit, nullvecs=nv)
    np.testing.assert_allclose(t.getHistory()[:8], ref.history[:8], rtol=1e-10)
    np.testing.assert_allclose(t.getHistory(), ref.history, rtol=1e-6)
    t.destroy()


def test_default_kernel_choice_of_the_line_coefficient_operator(pb):
    """Without a tuning key the library picks the kernels itself (sep_tile_xr in sep_solver.inc); whatever it picks, the
    product is the assembled MatMult bit for bit and the solve converges like the oracle's."""
    shape, per = (72, 20, 16), (0, 0, 0)
    widths = H.make_widths(shape)
    A, _ = H.velocity_system(widths, per, dt=0.01, nu=0.01, c=0.5)
    Ao = orc.Csr.from_arrays(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)
    s = pb.LinSolverB200("velocity", "None")
    s.setOptions(ksp_type="bcgs", pc_type="jacobi", rtol=0.0, atol=1e-9, max_it=500)
    s.setGrid(H.grid_of(widths, per))
    s.setMatrix(pb.Mat.from_scipy(A))
    assert s.operator == "staggered"
    rng = np.random.default_rng(2)
    x = rng.standard_normal(A.shape[0])
    assert np.array_equal(s.apply(x), Ao.spmv(x))
    b = rng.standard_normal(A.shape[0])
    xs = np.empty_like(b)
    s.solve(xs, b)
    ref = orc.ksp_solve(Ao, b, ksp_type="bcgs", pc_type="jacobi", rtol=0.0, atol=1e-9, max_it=500)
    assert s.getReason() == ref.reason == 3 and abs(s.getIters() - ref.its) <= 2
    np.testing.assert_allclose(s.getHistory()[:6], ref.history[:6], rtol=1e-9)
    s.destroy()

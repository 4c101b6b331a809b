"""GPU checks of the geometric multigrid preconditioner (-<name>_pc_type mg; mg_kernels.cuh): the device cycle against
the independent numpy/scipy restatement (tests/mg_reference.py), preconditioned-CG histories, and the point of it --
the same converged solution as plain CG on the same library in a fraction of the iterations.  The same kernel sources
run on the CPU emulation in tests/test_emulated_mg.py."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers as H
from tests import mg_reference as R

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.fixture(scope="module")
def pb():
    import petibm_b200

    return petibm_b200


def _solver(pb, widths, per, **opts):
    s = pb.LinSolverB200("poisson", "None")
    s.setOptions(pc_type="mg", **opts)
    s.setStencil(H.grid_of(widths, per))
    s.setNullSpace(True)
    return s


@pytest.mark.parametrize("shape,per", [((32, 24, 16), (0, 0, 0)), ((20, 32, 24), (1, 0, 1)), ((70, 45), (0, 0)), ((33, 40), (1, 1)),
                                       ((31, 17, 13), (0, 0, 0))])
def test_pcg_with_multigrid_matches_the_restatement(pb, shape, per):
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b, xs = H.consistent_rhs(A)
    V = R.VCycle(widths, per, 0.01)
    nit = 6
    xr, hr, _, _ = R.pcg(A.to_scipy(), b, V.apply, True, 0.0, 0.0, nit)
    s = _solver(pb, widths, per, rtol=0.0, atol=0.0, max_it=nit)
    x = np.empty_like(b)
    with pytest.raises(pb.B200Error):
        s.solve(x, b)                              # DIVERGED_ITS after the fixed number of iterations, like KSP
    assert (s.getIters(), s.getReason()) == (nit, -3)
    np.testing.assert_allclose(s.getHistory(), hr, rtol=1e-8)
    np.testing.assert_allclose(x, xr, rtol=0, atol=1e-9 * np.abs(xr).max())
    # to convergence, against plain CG of the same library
    s.setOptions(rtol=1e-10, atol=1e-50, max_it=200)
    s.solve(x, b)
    its_mg = s.getIters()
    assert s.getReason() == 2
    p = pb.LinSolverB200("poisson", "None")
    p.setOptions(rtol=1e-10, atol=1e-50, max_it=5000)
    p.setStencil(H.grid_of(widths, per))
    p.setNullSpace(True)
    xp = np.empty_like(b)
    p.solve(xp, b)
    assert p.getReason() == 2 and its_mg <= 25 and 3 * its_mg <= p.getIters(), (its_mg, p.getIters())
    np.testing.assert_allclose(x, xs, rtol=0, atol=1e-7 * np.abs(xs).max())
    np.testing.assert_allclose(x, xp, rtol=0, atol=1e-7 * np.abs(xs).max())
    hist = s.getHistory()
    assert np.all(np.diff(np.log(hist)) < 0)
    s.destroy(); p.destroy()


def test_multigrid_through_the_options_file_and_on_128_cubed(pb, tmp_path):
    cfg = tmp_path / "poisson_solver.info"
    cfg.write_text("-poisson_ksp_type cg\n-poisson_pc_type mg\n-poisson_mg_levels_ksp_max_it 2\n-poisson_ksp_rtol 1e-8\n"
                   "-poisson_ksp_atol 1e-50\n-poisson_ksp_max_it 100\n")
    s = pb.LinSolverB200("poisson", str(cfg))
    assert s.options().pc_type == 2
    grid = pb.Grid.uniform((128, 128, 128), dt=0.01)
    s.setStencil(grid)
    s.setNullSpace(True)
    rng = np.random.default_rng(3)
    xs = rng.standard_normal(grid.size)
    xs -= xs.mean()
    b = s.apply(xs)
    x = np.empty_like(b)
    s.solve(x, b)
    assert s.getReason() == 2 and s.getIters() <= 20, s.getIters()      # grid-independent: ~11 at 32^3 and 64^3 as well
    res = s.apply(x) - b
    assert np.linalg.norm(res) <= 1e-6 * np.linalg.norm(b)
    np.testing.assert_allclose(x - x.mean(), xs, rtol=0, atol=1e-6 * np.abs(xs).max())
    s.destroy()


def test_multigrid_on_the_stretched_cylinder_grid(pb):
    """The 450 x 450 grid of the 2-D cylinder case (125 stretched + 200 uniform + 125 stretched cells
    per axis, aspect ratios up to 40): width-equalising coarsening keeps the iteration count at the level of a uniform
    grid; Jacobi-preconditioned CG needs about 2 450 iterations for the same tolerance."""
    sub = [{"end": -0.75, "cells": 125, "stretchRatio": 1.0 / 1.02}, {"end": 0.75, "cells": 200, "stretchRatio": 1.0},
           {"end": 15.0, "cells": 125, "stretchRatio": 1.02}]
    w = orc.axis_from_subdomains(-15.0, sub)
    widths = [w, w.copy()]
    s = _solver(pb, widths, (0, 0), rtol=1e-8, atol=1e-50, max_it=200)
    rng = np.random.default_rng(8)
    xs = rng.standard_normal(w.size * w.size)
    xs -= xs.mean()
    b = s.apply(xs)
    x = np.empty_like(b)
    s.solve(x, b)
    assert s.getReason() == 2 and s.getIters() <= 25, s.getIters()
    np.testing.assert_allclose(x, xs, rtol=0, atol=1e-5 * np.abs(xs).max())
    s.destroy()


def test_multigrid_needs_the_separable_operator(pb):
    shape, per = (9, 8, 7), (0, 0, 0)
    A = H.oracle_matrix(H.make_widths(shape), per)
    s = pb.LinSolverB200("poisson", "None")
    s.setOptions(pc_type="mg")
    s.setMatrix(H.mat_of(A))                      # no grid description: the matrix is kept as CSR
    assert s.operator == "csr"
    with pytest.raises(pb.B200Error) as ei:
        s.solve(np.empty(A.shape[0]), np.ones(A.shape[0]))
    assert ei.value.code == -3
    s.destroy()


@pytest.mark.parametrize("dim", [2, 3])
def test_block_multigrid_on_a_stretched_ibpm_system(pb, dim):
    """[D;E] BN [G,-H] on a stretched grid in hybrid form (tests/test_zzz_gpu_2_staggered.py) with -poisson_pc_type mg:
    V-cycle on the pressure block, diagonal on the force rows, the explicit null-space vector of ibpm.cpp:251-267 --
    histories of the numpy restatement, same solution as Jacobi CG in a fraction of the iterations."""
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
    V = R.VCycle(widths, (0,) * dim, 0.01)
    dg = M.diagonal()

    def block_pc(v):
        z = np.empty_like(v)
        z[:pN] = V.apply(v[:pN])
        z[pN:] = v[pN:] / dg[pN:]
        return z

    xr, hr, _, _ = R.pcg(M, b, block_pc, False, 0.0, 0.0, 6, nullvec=nv)
    s = pb.LinSolverB200("poisson", "None")
    s.setOptions(pc_type="mg", rtol=0.0, atol=0.0, max_it=6)
    s.setGrid(pb.Grid(widths, (False, False, False), 0.01))
    s.setMatrix(pb.Mat.from_scipy(M).setNullSpace(False, nv))
    assert s.operator == "hybrid"
    x = np.empty_like(b)
    with pytest.raises(pb.B200Error):
        s.solve(x, b)
    np.testing.assert_allclose(s.getHistory(), hr, rtol=1e-8)
    s.setOptions(rtol=1e-9, atol=1e-50, max_it=500)
    s.solve(x, b)
    its_mg = s.getIters()
    s.setOptions(pc_type="jacobi", max_it=5000)
    xj = np.empty_like(b)
    s.solve(xj, b)
    assert s.getReason() == 2 and 3 * its_mg <= s.getIters(), (its_mg, s.getIters())
    np.testing.assert_allclose(x, xs, rtol=0, atol=1e-6 * np.abs(xs).max())
    s.destroy()


@pytest.mark.parametrize("shape,per", [((48, 40, 32), (0, 0, 0)), ((70, 45), (0, 0))])
def test_the_launch_switches_do_not_change_the_numbers(pb, shape, per):
    """mg_tail (coarse levels as one launch), mg_fuse (residual update and sums inside the cycle) and mg_graph (CUDA-graph
    replay) are on by default since round 2 (profiles/r02_tts_multigrid.log); every combination gives the same history."""
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b, _ = H.consistent_rhs(A)
    ref = None
    for tail, fuse, graph in ((1, 1, 1), (0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)):
        s = pb.LinSolverB200("poisson", "None")
        s.setOptions(pc_type="mg", rtol=1e-10, atol=1e-50, max_it=100)
        s.setTuning("mg_tail", tail); s.setTuning("mg_fuse", fuse); s.setTuning("mg_graph", graph)
        s.setStencil(H.grid_of(widths, per))
        s.setNullSpace(True)
        x = np.empty_like(b)
        s.solve(x, b)
        h = s.getHistory()
        if ref is None:
            ref = (h, x.copy())
        else:
            assert h.size == ref[0].size
            np.testing.assert_allclose(h, ref[0], rtol=1e-9)
            np.testing.assert_allclose(x, ref[1], rtol=0, atol=1e-10 * np.abs(ref[1]).max())
        s.destroy()


def test_graph_replay_of_the_assembled_operator_loops_does_not_change_the_numbers(pb):
    """csr_graph: CG / BiCGStab batches of the CSR and line-coefficient paths as CUDA graphs -- same history on and off."""
    widths = H.make_widths((14, 12, 10))
    A, _ = H.velocity_system(widths, (0, 0, 0), dt=0.01, nu=0.01, c=0.5)
    b = np.random.default_rng(4).standard_normal(A.shape[0])
    out = {}
    for stag in (True, False):
        for graph in (0, 1):
            s = pb.LinSolverB200("velocity", "None")
            s.setOptions(ksp_type="bcgs", pc_type="jacobi", rtol=0.0, atol=1e-10, max_it=200)
            s.setTuning("csr_graph", graph)
            s.setGrid(H.grid_of(widths, (0, 0, 0)))
            s.setStaggered(stag)
            s.setMatrix(pb.Mat.from_scipy(A))
            x = np.empty(A.shape[0])
            s.solve(x, b)
            out[(stag, graph)] = (s.getHistory(), x.copy(), s.operator)
            s.destroy()
    assert out[(True, 0)][2] == "staggered" and out[(False, 0)][2] == "csr"
    for stag in (True, False):
        assert np.array_equal(out[(stag, 0)][0], out[(stag, 1)][0]) and np.array_equal(out[(stag, 0)][1], out[(stag, 1)][1])

"""The configurations BASELINE.json lists, as parity cases (the benchmark itself is bench.py):
  C1  2-D lid-driven cavity 32 x 32 through the YAML settings + options files, time-step-like solve sequence
  C2  3-D 128^3 uniform Poisson CG, direct oracle comparison at full size
  C3  2-D IBPM-style modified Poisson at the size of the shipped cylinder case (CSR operator + explicit null vector)
  BN  a wider (25-point) operator standing in for BN order > 1: must take the verified CSR fallback."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import yaml

from oracle import oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import petibm_b200

    return petibm_b200


def _cavity32_settings():
    """The 32 x 32 lid-driven cavity case (BASELINE.json config 1): unit square, uniform cells, Dirichlet walls,
    moving lid, dt = 0.01, both solvers switched to the B200 backend.  Built here and round-tripped through
    YAML text, which is how the settings reach createLinSolver inside PetIBM."""
    axis = lambda d: {"direction": d, "start": 0.0, "subDomains": [{"end": 1.0, "cells": 32, "stretchRatio": 1.0}]}
    wall = lambda loc, lid=0.0: {"location": loc, "u": ["DIRICHLET", lid], "v": ["DIRICHLET", 0.0]}
    node = {"mesh": [axis("x"), axis("y")],
            "flow": {"nu": 0.01, "initialVelocity": [0.0, 0.0],
                     "boundaryConditions": [wall("xMinus"), wall("xPlus"), wall("yMinus"), wall("yPlus", 1.0)]},
            "parameters": {"dt": 0.01, "startStep": 0, "nt": 1000, "nsave": 1000, "nrestart": 1000,
                           "convection": "ADAMS_BASHFORTH_2", "diffusion": "CRANK_NICOLSON",
                           "velocitySolver": {"type": "B200", "config": "config/velocity_solver.info"},
                           "poissonSolver": {"type": "B200", "config": "config/poisson_solver.info"}}}
    return yaml.safe_load(yaml.safe_dump(node))


def test_c1_cavity32_through_the_config_files(pb, tmp_path):
    node = _cavity32_settings()
    node["directory"] = str(tmp_path)
    (tmp_path / "config").mkdir()
    (tmp_path / "config" / "poisson_solver.info").write_text(
        "-poisson_ksp_type cg\n-poisson_ksp_atol 1.0E-06\n-poisson_ksp_rtol 0.0\n-poisson_ksp_max_it 1000\n-poisson_pc_type jacobi\n")
    (tmp_path / "config" / "velocity_solver.info").write_text(
        "-velocity_ksp_type bcgs\n-velocity_ksp_atol 1.0E-06\n-velocity_ksp_rtol 0.0\n-velocity_ksp_max_it 1000\n-velocity_pc_type jacobi\n")
    pS = pb.createLinSolver("poisson", node)
    vS = pb.createLinSolver("velocity", node)
    assert pS.getType() == vS.getType() == "PETSc KSP"
    assert vS.options().ksp_type == 1 and vS.options().pc_type == 1
    grid = pb.Grid.from_config(node)
    assert grid.n == (32, 32) and grid.dt == 0.01
    A = H.oracle_matrix(grid.widths, (0, 0), grid.dt)
    pS.setMatrix(H.mat_of(A, const_nullspace=True))          # what NavierStokesSolver::init does (navierstokes.cpp:164)
    assert pS.operator == "stencil"
    # velocity-like system on the same grid for the second solver (A = I/dt - c nu L flavoured)
    n = A.shape[0]
    M = (sp.identity(n) / grid.dt - 0.5 * 0.01 * (A.to_scipy() / grid.dt * 1e3)).tocsr()
    M.sort_indices()
    vS.setMatrix(pb.Mat.from_scipy(M))
    assert vS.operator == "csr"
    Mo = orc.Csr.from_arrays(n, n, M.indptr, M.indices, M.data)
    rng = np.random.default_rng(4)
    x = np.empty(n)
    for step in range(5):                                     # a few "time steps": solve, read iterations/residual
        rhs = rng.standard_normal(n)
        vS.solve(x, rhs)
        ref = orc.ksp_solve(Mo, rhs, ksp_type="bcgs", pc_type="jacobi", rtol=0.0, atol=1e-6, max_it=1000)
        assert vS.getIters() == ref.its and abs(vS.getResidual() - ref.rnorm) <= 1e-6 * ref.history[0] + 1e-9 * ref.rnorm
        np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-8 * np.abs(ref.x).max())
        rhs2 = A.spmv(rng.standard_normal(n))                 # div of something: consistent
        pS.solve(x, rhs2)
        ref = orc.ksp_solve(A, rhs2, pc_type="jacobi", rtol=0.0, atol=1e-6, max_it=1000, const_nullspace=True)
        assert pS.getReason() == ref.reason == 3 and abs(pS.getIters() - ref.its) <= 1
        np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-6 * np.abs(ref.x).max())
    pS.destroy(); vS.destroy()


@pytest.mark.parametrize("pc", ["none", "jacobi"])
@pytest.mark.parametrize("n,nit", [((128, 128, 128), 60), ((256, 256, 256), 40)])
def test_c2_128cubed_and_headline_256cubed_against_the_oracle(pb, pc, n, nit):
    """C2 and the headline bench configuration itself (256^3 uniform): the residual history of the first nit
    iterations against the oracle's KSPSolve_CG (OpenMP summation mode: about 1 s at 256^3), rtol 1e-10."""
    grid = pb.Grid.uniform(n, dt=0.01)
    A = orc.assemble_dbng(grid.widths, (0, 0, 0), 0.01, literal=False)
    b, xs = H.consistent_rhs(A)
    orc.set_fast(True, 0)
    ref = orc.ksp_solve(A, b, pc_type=pc, rtol=0.0, atol=0.0, max_it=nit, const_nullspace=True)
    orc.set_fast(False, 0)
    s = pb.LinSolverB200("poisson", "None")
    s.setOptions(pc_type=pc, rtol=0.0, atol=0.0, max_it=nit)
    s.setStencil(grid)
    s.setNullSpace(True)
    assert np.array_equal(s.apply(xs), b)                     # bit-exact SpMV at full size
    x = np.empty_like(b)
    with pytest.raises(pb.B200Error):
        s.solve(x, b)
    np.testing.assert_allclose(s.getHistory(), ref.history, rtol=1e-10)
    np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-9 * np.abs(ref.x).max())
    s.destroy()


def test_large_ragged_stretched_grid_on_the_default_kernel(pb):
    """Out-of-L2 grid whose extents are not multiples of the tile (300 x 290 x 130, stretched): the launch-shape model
    picks the TMA kernel with several z chunks (ragged last tiles in x and y, ragged last chunk) -- bit-exact SpMV and the
    first 25 residual norms against the oracle."""
    n = (300, 290, 130)
    widths = H.make_widths(n)
    A = orc.assemble_dbng(widths, (0, 0, 0), 0.01, literal=False)
    b, xs = H.consistent_rhs(A)
    nit = 25
    orc.set_fast(True, 0)
    ref = orc.ksp_solve(A, b, rtol=0.0, atol=0.0, max_it=nit, const_nullspace=True)
    orc.set_fast(False, 0)
    s = pb.LinSolverB200("poisson", "None")
    s.setOptions(rtol=0.0, atol=0.0, max_it=nit)
    s.setStencil(H.grid_of(widths, (0, 0, 0)))
    s.setNullSpace(True)
    assert np.array_equal(s.apply(xs), b)
    x = np.empty_like(b)
    with pytest.raises(pb.B200Error):
        s.solve(x, b)
    np.testing.assert_allclose(s.getHistory(), ref.history, rtol=1e-10)
    np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-9 * np.abs(ref.x).max())
    s.destroy()


def test_c3_ibpm_sized_modified_poisson(pb):
    """450 x 450 stretched grid + 158 Lagrangian points x 2 force components."""
    sub = [{"end": -0.75, "cells": 125, "stretchRatio": 1.0 / 1.02}, {"end": 0.75, "cells": 200, "stretchRatio": 1.0},
           {"end": 15.0, "cells": 125, "stretchRatio": 1.02}]
    w = orc.axis_from_subdomains(-15.0, sub)
    widths = [w, w.copy()]
    G = orc.assemble_gradient(widths, [0, 0, 0]).to_scipy()
    nf = 2 * 158
    rng = np.random.default_rng(9)
    # each force dof couples to a 4 x 4 patch of velocity points (delta support), like E^T
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
    s = pb.LinSolverB200("poisson", "None")
    s.setOptions(pc_type="jacobi", rtol=0.0, atol=0.0, max_it=40)
    s.setGrid(pb.Grid(widths, (False, False, False), 0.01))   # the grid is known, the matrix is NOT its stencil
    s.setStaggered(False)                                      # plain CSR here; line-coefficient form: test_zzz_gpu_2_staggered.py
    s.setMatrix(pb.Mat.from_scipy(M).setNullSpace(False, nv))
    assert s.operator == "csr" and s.nlocal == pN + nf
    ref = orc.ksp_solve(Mo, b, pc_type="jacobi", rtol=0.0, atol=0.0, max_it=40, nullvecs=nv)
    x = np.empty_like(b)
    with pytest.raises(pb.B200Error):
        s.solve(x, b)
    np.testing.assert_allclose(s.getHistory(), ref.history, rtol=1e-10)
    np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-9 * np.abs(ref.x).max())
    s.destroy()


def test_wider_stencil_takes_the_csr_fallback(pb):
    """BN order N > 1 widens DBNG (createbn.cpp:59-92): setMatrix must not mistake it for the 7-point form."""
    shape, per = (12, 11, 10), (0, 0, 0)
    widths = H.make_widths(shape)
    _, Lap = H.velocity_system(widths, per, dt=0.01, nu=0.01, c=0.5)
    Lo = orc.Csr.from_arrays(Lap.shape[0], Lap.shape[1], Lap.indptr, Lap.indices, Lap.data)
    # the reference's own product: BN = dt I + dt^2 (c nu) L (createbn.cpp:59-92), then D (BN G) (navierstokes.cpp:351-356)
    B2 = orc.bnhead(Lo, 0.01, 0.5 * 0.01, 2)
    W = orc.matmatmult(orc.assemble_divergence(widths, per), orc.matmatmult(B2, orc.assemble_gradient(widths, per))).to_scipy().tocsr()
    W.sort_indices()
    assert W.getnnz(axis=1).max() > 7                          # 13-/25-point rows
    Wo = orc.Csr.from_arrays(W.shape[0], W.shape[1], W.indptr, W.indices, W.data)
    b = W @ (lambda v: v - v.mean())(np.random.default_rng(6).standard_normal(W.shape[0]))
    s = pb.LinSolverB200("poisson", "None")
    s.setOptions(rtol=0.0, atol=0.0, max_it=30)
    s.setGrid(H.grid_of(widths, per))
    s.setMatrix(pb.Mat.from_scipy(W).setNullSpace(True))
    assert s.operator == "csr"
    ref = orc.ksp_solve(Wo, b, rtol=0.0, atol=0.0, max_it=30, const_nullspace=True)
    x = np.empty_like(b)
    with pytest.raises(pb.B200Error):
        s.solve(x, b)
    np.testing.assert_allclose(s.getHistory(), ref.history, rtol=1e-10)
    s.destroy()

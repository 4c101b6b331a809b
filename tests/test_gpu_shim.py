"""The C++ shim (petibm_b200/csrc/petibm_shim/LinSolverB200 : LinSolverBase) compiled against single-process
stand-ins for PETSc (tests/cpp/petibm_stub.h) and driven like the applications drive LinSolverKSP, checked
against the oracle.  The same two files compile against real PETSc inside PetIBM."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "petibm_b200", "csrc", "petibm_shim")


def _build(tmp):
    exe = os.path.join(tmp, "shim_driver")
    cmd = ["g++", "-std=c++14", "-O1", "-DB200_SHIM_STUB", "-I", os.path.join(ROOT, "tests", "cpp"), "-I", SHIM,
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "shim_driver.cpp"),
           os.path.join(SHIM, "linsolverb200.cpp"), "-o", exe, "-L", os.path.join(ROOT, "petibm_b200"), "-lb200ls",
           "-Wl,-rpath," + os.path.join(ROOT, "petibm_b200")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_shim_compiles_against_the_stub(tmp_path):
    """CPU: the shim is valid C++14 against the interface it claims to implement and links to the C ABI."""
    _build(str(tmp_path))


def _write_case(d, widths, per, A, b, const_ns, with_grid, nsolves=1):
    rp, col, val = A.arrays()
    dim = len(widths)
    n = [len(w) for w in widths] + [1] * (3 - dim)
    meta = np.array([dim, *n, *(list(per) + [0] * (3 - dim)), int(const_ns), int(with_grid), nsolves], dtype=np.int32)
    meta.tofile(os.path.join(d, "meta.bin"))
    np.array([0.01]).tofile(os.path.join(d, "dt.bin"))
    rp.astype(np.int32).tofile(os.path.join(d, "rowptr.bin"))
    col.astype(np.int32).tofile(os.path.join(d, "col.bin"))
    val.tofile(os.path.join(d, "val.bin"))
    for name, w in zip(("dx", "dy", "dz"), widths):
        np.asarray(w, dtype=np.float64).tofile(os.path.join(d, name + ".bin"))
    np.asarray(b, dtype=np.float64).tofile(os.path.join(d, "b.bin"))


@pytest.mark.gpu
@pytest.mark.parametrize("with_grid", [True, False])
def test_shim_solves_like_ksp(tmp_path, with_grid):
    d = str(tmp_path)
    exe = _build(d)
    shape, per = (18, 14, 10), (0, 0, 0)
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b, xs = H.consistent_rhs(A)
    _write_case(d, widths, per, A, b, True, with_grid, nsolves=2)
    cfg = os.path.join(d, "poisson_solver.info")
    open(cfg, "w").write("-poisson_ksp_type cg\n-poisson_pc_type jacobi\n-poisson_ksp_atol 1e-8\n-poisson_ksp_rtol 0\n"
                         "-poisson_ksp_max_it 2000\n")
    res = subprocess.run([exe, d, cfg], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "type=PETSc KSP" in res.stdout
    ref = orc.ksp_solve(A, b, pc_type="jacobi", rtol=0.0, atol=1e-8, max_it=2000, const_nullspace=True)
    its, resn, ierr, is_stencil = np.fromfile(os.path.join(d, "out.bin"))
    x = np.fromfile(os.path.join(d, "x.bin"))
    assert ierr == 0 and ref.reason == 3
    assert bool(is_stencil) == with_grid          # without the grid description the CSR operator is used
    assert abs(int(its) - ref.its) <= 1 and resn < 1e-8
    np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-7 * np.abs(ref.x).max())


@pytest.mark.gpu
def test_shim_reports_divergence_like_linsolverksp(tmp_path):
    d = str(tmp_path)
    exe = _build(d)
    shape, per = (12, 10), (0, 0)
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b, _ = H.consistent_rhs(A)
    _write_case(d, widths, per, A, b, True, True)
    cfg = os.path.join(d, "poisson_solver.info")
    open(cfg, "w").write("-poisson_ksp_rtol 1e-14\n-poisson_ksp_max_it 3\n")
    res = subprocess.run([exe, d, cfg], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0
    its, resn, ierr, _ = np.fromfile(os.path.join(d, "out.bin"))
    assert int(ierr) == 82 and int(its) == 3            # PETSC_ERR_CONV_FAILED, as linsolverksp.cpp:100-103
    assert "diverged with reason -3" in res.stderr
    # an unsupported preconditioner in the options file is a construction-time error
    open(cfg, "w").write("-poisson_pc_type gamg\n")
    res = subprocess.run([exe, d, cfg], capture_output=True, text=True, timeout=300)
    assert "not implemented by the B200 backend" in res.stderr


def _write_matrix_case(d, widths, per, M, b, null_mode, nullvec=None, with_grid=True):
    """Any assembled matrix (scipy CSR) with the grid description of the mesh: null_mode 0 none, 1 constant, 2 one vector."""
    M = M.tocsr(); M.sort_indices()
    dim = len(widths)
    n = [len(w) for w in widths] + [1] * (3 - dim)
    np.array([dim, *n, *(list(per) + [0] * (3 - dim)), int(null_mode), int(with_grid), 1], dtype=np.int32).tofile(os.path.join(d, "meta.bin"))
    np.array([0.01]).tofile(os.path.join(d, "dt.bin"))
    M.indptr.astype(np.int32).tofile(os.path.join(d, "rowptr.bin"))
    M.indices.astype(np.int32).tofile(os.path.join(d, "col.bin"))
    M.data.astype(np.float64).tofile(os.path.join(d, "val.bin"))
    for name, w in zip(("dx", "dy", "dz"), widths):
        np.asarray(w, dtype=np.float64).tofile(os.path.join(d, name + ".bin"))
    np.asarray(b, dtype=np.float64).tofile(os.path.join(d, "b.bin"))
    if nullvec is not None:
        np.asarray(nullvec, dtype=np.float64).tofile(os.path.join(d, "nullvec.bin"))


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_shim_picks_the_structured_operators_for_the_other_systems(tmp_path):
    """The three other matrices the applications hand to setMatrix, through the C++ shim: the velocity system (bcgs + jacobi,
    shipped velocity_solver.info) -> line-coefficient operator; IBPM's modified Poisson system on a stretched grid with
    pc_type mg -> hybrid operator with the block multigrid; the decoupled-IBPM forces system (shipped forces_solver.info)
    -> direct solve."""
    d = str(tmp_path)
    exe = _build(d)
    rng = np.random.default_rng(8)
    # velocity system
    shape, per = (14, 12, 10), (0, 0, 0)
    widths = H.make_widths(shape)
    A, _ = H.velocity_system(widths, per, dt=0.01, nu=0.01)
    b = rng.standard_normal(A.shape[0])
    _write_matrix_case(d, widths, per, A, b, 0)
    cfg = os.path.join(d, "velocity_solver.info")
    open(cfg, "w").write("-velocity_ksp_type bcgs\n-velocity_ksp_atol 1.0E-08\n-velocity_ksp_rtol 0.0\n-velocity_ksp_max_it 1000\n"
                         "-velocity_pc_type jacobi\n-velocity_pc_jacobi_type diagonal\n")
    res = subprocess.run([exe, d, cfg, "velocity"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    its, resn, ierr, kind = np.fromfile(os.path.join(d, "out.bin"))
    Ao = orc.Csr.from_arrays(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)
    ref = orc.ksp_solve(Ao, b, ksp_type="bcgs", pc_type="jacobi", rtol=0.0, atol=1e-8, max_it=1000)
    assert ierr == 0 and kind == 3 and abs(int(its) - ref.its) <= 2
    np.testing.assert_allclose(np.fromfile(os.path.join(d, "x.bin")), ref.x, rtol=0, atol=1e-6 * np.abs(ref.x).max())
    # IBPM modified Poisson system, stretched grid, multigrid block preconditioner
    sub = [{"end": 0.6, "cells": 10, "stretchRatio": 1.0 / 1.2}, {"end": 1.4, "cells": 20, "stretchRatio": 1.0},
           {"end": 2.0, "cells": 10, "stretchRatio": 1.2}]
    w = orc.axis_from_subdomains(0.0, sub)
    widths = [w, w.copy()]
    M, pN, nv = H.ibpm_system(widths, dt=0.01, nb=14)
    xs = rng.standard_normal(M.shape[0]); xs -= (xs @ nv) * nv
    b = M @ xs
    _write_matrix_case(d, widths, (0, 0), M, b, 2, nullvec=nv)
    cfg = os.path.join(d, "poisson_solver.info")
    open(cfg, "w").write("-poisson_ksp_type cg\n-poisson_pc_type mg\n-poisson_ksp_rtol 1e-9\n-poisson_ksp_atol 1e-50\n-poisson_ksp_max_it 500\n")
    res = subprocess.run([exe, d, cfg, "poisson"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    its, resn, ierr, kind = np.fromfile(os.path.join(d, "out.bin"))
    assert ierr == 0 and kind == 2 and int(its) <= 80
    np.testing.assert_allclose(np.fromfile(os.path.join(d, "x.bin")), xs, rtol=0, atol=1e-6 * np.abs(xs).max())
    # forces system, direct solve
    F = (-M[pN:, pN:]).tocsr()
    b = rng.standard_normal(F.shape[0])
    _write_matrix_case(d, widths, (0, 0), F, b, 0, with_grid=False)
    cfg = os.path.join(d, "forces_solver.info")
    open(cfg, "w").write("-forces_ksp_type preonly\n-forces_pc_type lu\n-forces_pc_factor_mat_solver_type superlu_dist\n")
    res = subprocess.run([exe, d, cfg, "forces"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    its, resn, ierr, kind = np.fromfile(os.path.join(d, "out.bin"))
    assert ierr == 0 and kind == 0 and int(its) == 1 and resn == 0.0
    ref = np.linalg.solve(F.toarray(), b)
    np.testing.assert_allclose(np.fromfile(os.path.join(d, "x.bin")), ref, rtol=0, atol=1e-11 * np.abs(ref).max())

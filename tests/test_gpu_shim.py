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

"""GPU check of the direct solve (-forces_ksp_type preonly -forces_pc_type lu, the forces system of PetIBM's decoupled
IBPM; dense_kernels.cuh) through the C ABI: KSPSolve_PREONLY semantics (one PCApply, its = 1, KSP_CONVERGED_ITS, no
residual norm) and the solution of numpy's LAPACK solve.  The same kernels run on the CPU emulation in
tests/test_emulated_kernels.py::test_emulated_direct_solve_of_the_forces_system."""
import numpy as np
import pytest
import scipy.sparse as sp

from tests.test_emulated_kernels import _forces_system

from tests import helpers as H  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.fixture(scope="module")
def pb():
    import petibm_b200

    return petibm_b200


def test_forces_system_direct_solve(pb, tmp_path):
    cfg = tmp_path / "forces_solver.info"      # the file every decoupled-IBPM example ships
    cfg.write_text("# forces solver: prefix `-forces_`\n-forces_ksp_type preonly\n-forces_pc_type lu\n"
                   "-forces_pc_factor_mat_solver_type superlu_dist\n")
    F = _forces_system(n_side=10, n_band=20, nb=40)
    s = pb.LinSolverB200("forces", str(cfg))
    assert (s.options().ksp_type, s.options().pc_type) == (2, 3)
    s.setMatrix(pb.Mat.from_scipy(F))
    assert s.operator == "csr"
    rng = np.random.default_rng(6)
    A = F.toarray()
    x = np.empty(F.shape[0])
    for _ in range(3):                          # one factorisation, several time steps
        b = rng.standard_normal(F.shape[0])
        s.solve(x, b)
        assert (s.getIters(), s.getReason(), s.getResidual()) == (1, 4, 0.0) and s.getHistory().size == 0
        ref = np.linalg.solve(A, b)
        np.testing.assert_allclose(x, ref, rtol=0, atol=1e-11 * np.abs(ref).max())
    # a second setMatrix (moving bodies: rigidkinematics.cpp:119-140) factorises again
    s.setMatrix(pb.Mat.from_scipy((2.0 * F).tocsr()))
    s.solve(x, b)
    np.testing.assert_allclose(x, 0.5 * ref, rtol=0, atol=1e-11 * np.abs(ref).max())
    s.destroy()


@pytest.mark.parametrize("n", [33, 130, 1000])
def test_direct_solve_pivots(pb, n):
    """Partial pivoting on the device (blocked LU, dense_kernels.cuh): zero diagonal, badly scaled rows, sizes that are
    not multiples of the panel width; against LAPACK."""
    rng = np.random.default_rng(100 + n)
    A = rng.standard_normal((n, n))
    A[np.arange(n), np.arange(n)] = 0.0
    A[rng.integers(0, n, 2)] *= 1e6
    s = pb.LinSolverB200("forces", "None")
    s.setOptions(ksp_type="preonly", pc_type="lu")
    s.setMatrix(pb.Mat.from_scipy(sp.csr_matrix(A)))
    b = rng.standard_normal(n)
    x = np.empty(n)
    s.solve(x, b)
    ref = np.linalg.solve(A, b)
    np.testing.assert_allclose(x, ref, rtol=0, atol=1e-13 * np.linalg.cond(A) * np.abs(ref).max())
    assert np.abs(A @ x - b).max() <= 1e-12 * (np.abs(A).max() * np.abs(x).max() * n)
    s.destroy()


def test_direct_solve_refusals(pb):
    s = pb.LinSolverB200("forces", "None")
    with pytest.raises(pb.B200Error):
        s.setOptions(ksp_type="preonly")                      # without pc_type lu
    s.setOptions(ksp_type="preonly", pc_type="lu")
    S = sp.csr_matrix(np.array([[1.0, 2.0, 0.0], [2.0, 4.0, 0.0], [0.0, 0.0, 1.0]]))
    s.setMatrix(pb.Mat.from_scipy(S))
    with pytest.raises(pb.B200Error) as ei:
        s.solve(np.empty(3), np.ones(3))
    assert ei.value.code == -5 and s.getReason() == -11      # zero pivot: KSP_DIVERGED_PC_FAILED
    s.destroy()

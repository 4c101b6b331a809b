"""Row a10 of SURVEY section 8: the velocity system A = I/dt - c nu L (createlaplacian.cpp:108-262,
navierstokes.cpp:342-344), solved by BiCGStab + Jacobi in the shipped configs.  The operator is assembled on
the test side (tests/helpers.velocity_system) from the oracle's velocity-mesh arrays; the GPU solve goes through
the general CSR operator and is compared with the oracle's KSPSolve_BCGS restatement."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers as H


def test_velocity_operator_structure():
    widths = H.make_widths((10, 9, 8))
    A, L = H.velocity_system(widths, (0, 0, 0), dt=0.01, nu=0.01, c=0.5)
    nu_, nv_, nw_ = 9 * 9 * 8, 10 * 8 * 8, 10 * 9 * 7
    assert A.shape == (nu_ + nv_ + nw_,) * 2
    # rows that touch no wall: Laplacian row sums vanish; the fields do not couple
    rs = np.abs(L.sum(axis=1).A1)
    assert (rs < 1e-9 * np.abs(L).max()).sum() > 0.3 * L.shape[0]
    assert L[:nu_, nu_:].nnz == 0 and L[nu_:nu_ + nv_, :nu_].nnz == 0
    # stretched grid: non-symmetric (row scaling by dL of the row), strongly diagonally dominant
    assert abs(A - A.T).max() > 1e-3
    off = np.abs(A).sum(axis=1).A1 - np.abs(A.diagonal())
    assert np.all(np.abs(A.diagonal()) > off)
    # uniform grid: symmetric
    Au, _ = H.velocity_system([np.full(8, 0.125)] * 3, (0, 0, 0))
    assert abs(Au - Au.T).max() < 1e-9 * abs(Au).max()
    # periodic direction: the u field has as many points as cells along x
    Ap, _ = H.velocity_system(H.make_widths((9, 8)), (1, 0))
    assert Ap.shape[0] == 9 * 8 + 9 * 7


@pytest.mark.parametrize("shape,per", [((10, 9, 8), (0, 0, 0)), ((7, 6, 5), (1, 0, 1)), ((9, 8), (0, 1))])
def test_vectorised_assembly_is_the_same_matrix(shape, per):
    widths = H.make_widths(shape)
    A, _ = H.velocity_system(widths, per, dt=0.02, nu=0.03, c=0.5)
    B = H.velocity_system_fast(widths, per, dt=0.02, nu=0.03, c=0.5)
    assert A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    assert np.array_equal(A.data, B.data)


@pytest.mark.gpu
@pytest.mark.parametrize("dt,nu", [(0.01, 0.01), (0.5, 1.0)])
def test_velocity_system_bcgs_jacobi(tmp_path, dt, nu):
    import petibm_b200 as pb

    widths = H.make_widths((14, 12, 10))
    A, _ = H.velocity_system(widths, (0, 0, 0), dt=dt, nu=nu, c=0.5)
    Ao = orc.Csr.from_arrays(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)
    cfg = tmp_path / "velocity_solver.info"
    cfg.write_text("-velocity_ksp_type bcgs\n-velocity_ksp_atol 1.0E-08\n-velocity_ksp_rtol 0.0\n"
                   "-velocity_ksp_max_it 1000\n-velocity_pc_type jacobi\n")   # as examples/*/config/velocity_solver.info
    s = pb.LinSolverB200("velocity", str(cfg))
    s.setMatrix(pb.Mat.from_scipy(A))
    assert s.operator == "csr"
    rng = np.random.default_rng(2)
    x = np.empty(A.shape[0])
    for _ in range(3):
        b = rng.standard_normal(A.shape[0])
        kw = dict(ksp_type="bcgs", pc_type="jacobi", rtol=0.0, atol=1e-8, max_it=1000)
        ref = orc.ksp_solve(Ao, b, **kw)
        orc.set_fast(True, 4)
        alt = orc.ksp_solve(Ao, b, **kw)          # same algorithm, another summation order: the comparison floor
        orc.set_fast(False, 0)
        s.solve(x, b)
        hist = s.getHistory()
        assert s.getReason() == ref.reason == 3
        assert hist.size == s.getIters() + 1 and hist[-1] < 1e-8
        # BiCGStab amplifies round-off quickly; iteration counts of two correct implementations differ by a few
        assert abs(s.getIters() - ref.its) <= max(2, abs(alt.its - ref.its) + 2), (s.getIters(), ref.its, alt.its)
        m = min(hist.size, ref.history.size, alt.history.size)
        sens = np.maximum.accumulate(np.abs(alt.history[:m] - ref.history[:m]) / ref.history[:m])
        rel = np.abs(hist[:m] - ref.history[:m]) / ref.history[:m]
        assert np.all(rel <= np.maximum(1e-10, 1e3 * sens)), (rel.max(), sens.max())
        # the answer itself: true (Jacobi-preconditioned) residual at the requested tolerance
        res = (A @ x - b) / A.diagonal()
        assert np.linalg.norm(res) < 1e-7
        np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-6 * np.abs(ref.x).max())
    s.destroy()

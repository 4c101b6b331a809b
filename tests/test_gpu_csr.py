"""GPU parity of the general assembled-operator path (CSR): IBPM-style modified Poisson systems with an
explicit null-space vector, the velocity system with BiCGStab + Jacobi, and the verified fallback taken
when setMatrix receives a matrix the separable stencil does not reproduce."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu

HIST_RTOL = 1e-10


@pytest.fixture(scope="module")
def pb():
    import petibm_b200

    return petibm_b200


def _orc_csr(M):
    M = M.tocsr()
    M.sort_indices()
    return orc.Csr.from_arrays(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)


def _velocity_like(shape=(10, 9, 8), seed=5):
    """A = I/dt - c nu L flavoured: non-symmetric, diagonally dominant (navierstokes.cpp:342-344)."""
    rng = np.random.default_rng(seed)
    A = H.oracle_matrix(H.make_widths(shape), (0, 0, 0)).to_scipy()
    n = A.shape[0]
    M = (-A + sp.identity(n) * 50.0 + sp.diags(rng.uniform(0, 1, n - 1), 1) + sp.diags(rng.uniform(0, 0.5, n - 7), -7)).tocsr()
    M.sort_indices()
    return M


def _ibpm_like(shape=(14, 12), nf=9, seed=11):
    """-(K^T K) with K = [G | R]: the pressure block keeps the constant in its null space, the force block
    does not (ibpm.cpp:164-194, 251-267)."""
    rng = np.random.default_rng(seed)
    widths = H.make_widths(shape)
    G = orc.assemble_gradient(widths, [0, 0, 0]).to_scipy()           # UN x pN, rows sum to zero
    R = sp.random(G.shape[0], nf, density=0.04, random_state=seed, format="csr")
    K = sp.hstack([G, R]).tocsr()
    M = (-(K.T @ K) * 0.01).tocsr()
    M.sort_indices()
    pN = G.shape[1]
    nv = np.zeros(M.shape[0])
    nv[:pN] = 1.0 / np.sqrt(pN)
    assert np.abs(M @ nv).max() < 1e-13
    return M, nv


def test_csr_spmv_bit_exact(pb):
    M = _velocity_like()
    Ao = _orc_csr(M)
    s = pb.LinSolverB200("velocity", "None")
    s.setMatrix(pb.Mat.from_scipy(M))
    assert s.operator == "csr"
    x = np.random.default_rng(0).standard_normal(M.shape[0])
    assert np.array_equal(s.apply(x), Ao.spmv(x))
    s.destroy()


@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_bcgs_matches_oracle(pb, pc):
    M = _velocity_like()
    Ao = _orc_csr(M)
    rng = np.random.default_rng(1)
    xs = rng.standard_normal(M.shape[0])
    b = M @ xs
    ref = orc.ksp_solve(Ao, b, ksp_type="bcgs", pc_type=pc, rtol=1e-10, max_it=500)
    assert ref.reason == 2
    s = pb.LinSolverB200("velocity", "None")
    s.setOptions(ksp_type="bcgs", pc_type=pc, rtol=1e-10, max_it=500)
    s.setMatrix(pb.Mat.from_scipy(M))
    x = np.empty_like(b)
    s.solve(x, b)
    assert s.getReason() == 2 and abs(s.getIters() - ref.its) <= 1
    hist = s.getHistory()
    m = min(hist.size, ref.history.size, 12)
    np.testing.assert_allclose(hist[:m], ref.history[:m], rtol=1e-9)
    np.testing.assert_allclose(x, xs, rtol=0, atol=1e-7 * np.abs(xs).max())
    # fixed iteration count: DIVERGED_ITS exactly like KSP
    s.setOptions(rtol=0.0, atol=0.0, max_it=5)
    with pytest.raises(pb.B200Error):
        s.solve(x, b)
    ref5 = orc.ksp_solve(Ao, b, ksp_type="bcgs", pc_type=pc, rtol=0.0, atol=0.0, max_it=5)
    assert s.getReason() == ref5.reason == -3 and s.getIters() == ref5.its == 5
    np.testing.assert_allclose(s.getHistory(), ref5.history, rtol=1e-9)
    s.destroy()


@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_cg_with_explicit_nullspace_vector(pb, pc):
    M, nv = _ibpm_like()
    Ao = _orc_csr(M)
    rng = np.random.default_rng(2)
    xs = rng.standard_normal(M.shape[0])
    xs -= (xs @ nv) * nv
    b = M @ xs
    nit = 30
    ref = orc.ksp_solve(Ao, b, pc_type=pc, rtol=0.0, atol=0.0, max_it=nit, nullvecs=nv)
    s = pb.LinSolverB200("poisson", "None")
    s.setOptions(pc_type=pc, rtol=0.0, atol=0.0, max_it=nit)
    s.setMatrix(pb.Mat.from_scipy(M).setNullSpace(False, nv))
    assert s.operator == "csr"
    x = np.empty_like(b)
    with pytest.raises(pb.B200Error):
        s.solve(x, b)
    assert s.getIters() == nit and s.getReason() == -3
    np.testing.assert_allclose(s.getHistory(), ref.history, rtol=HIST_RTOL)
    np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-9 * np.abs(ref.x).max())
    # and to convergence
    s.setOptions(rtol=1e-9, atol=1e-50, max_it=3000)
    s.solve(x, b)
    ref2 = orc.ksp_solve(Ao, b, pc_type=pc, rtol=1e-9, max_it=3000, nullvecs=nv)
    assert s.getReason() == ref2.reason == 2 and abs(s.getIters() - ref2.its) <= 2
    s.destroy()


@pytest.mark.parametrize("nullspace", [False, True])
def test_fallback_when_the_matrix_is_not_the_stencil(pb, nullspace):
    shape, per = (9, 8, 7), (0, 0, 0)
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    rp, col, val = A.arrays()
    val = val.copy()
    # symmetric perturbation that keeps the row sums: not D*(dt*G) of this grid any more
    S = sp.csr_matrix((val, col, rp), shape=A.shape).tolil()
    i, j = 5, 6
    S[i, j] *= 1.5; S[j, i] = S[i, j]
    S[i, i] = 0; S[i, i] = -S[i].sum()
    S[j, j] = 0; S[j, j] = -S[j].sum()
    S = S.tocsr(); S.sort_indices()
    Ao = _orc_csr(S)
    b = S @ (lambda v: v - v.mean())(np.random.default_rng(3).standard_normal(S.shape[0]))
    s = pb.LinSolverB200("poisson", "None")
    s.setOptions(pc_type="jacobi", rtol=0.0, atol=0.0, max_it=25)
    s.setGrid(H.grid_of(widths, per))
    s.setMatrix(pb.Mat.from_scipy(S).setNullSpace(nullspace))
    assert s.operator == "csr"
    ref = orc.ksp_solve(Ao, b, pc_type="jacobi", rtol=0.0, atol=0.0, max_it=25, const_nullspace=nullspace)
    x = np.empty_like(b)
    with pytest.raises(pb.B200Error):
        s.solve(x, b)
    np.testing.assert_allclose(s.getHistory(), ref.history, rtol=HIST_RTOL)
    np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-9 * np.abs(ref.x).max())
    s.destroy()


def test_unsupported_combinations_fail_loudly(pb):
    M = _velocity_like()
    s = pb.LinSolverB200("velocity", "None")
    s.setOptions(ksp_type="bcgs")
    s.setMatrix(pb.Mat.from_scipy(M).setNullSpace(True))
    with pytest.raises(pb.B200Error) as ei:
        s.solve(np.empty(M.shape[0]), np.ones(M.shape[0]))
    assert ei.value.code == -3
    s.destroy()

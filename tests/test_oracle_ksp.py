"""KSP restatement (KSPSolve_CG / KSPSolve_BCGS / KSPConvergedDefault / MatNullSpaceRemove).
PETSc is not installed, so these check the restatement against independent numpy/scipy
formulations of the same recurrences and against the semantics SURVEY.md section 8c lists."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import oracle as orc


def _problem(shape=(12, 10, 8), per=(0, 0, 0), seed=20240521, stretched=True):
    rng = np.random.default_rng(seed)
    widths = [(rng.uniform(0.7, 1.3, n) if stretched else np.ones(n)) / n for n in shape]
    A = orc.assemble_dbng(widths, per, 0.01)
    xs = rng.standard_normal(A.shape[0])
    xs -= xs.mean()
    return A, A.spmv(xs), xs


def _numpy_cg(A, b, pc, nullspace, nit):
    """Textbook PCG written independently of the C code, preconditioned-norm history."""
    n = b.size
    dinv = 1.0 / A.diagonal() if pc == "jacobi" else np.ones(n)

    def B(r):
        z = r * dinv
        if nullspace:
            z = z + z.sum() / (-1.0 * n)
        return z

    x = np.zeros(n); r = b.copy(); z = B(r)
    hist = [np.sqrt(z @ z)]
    beta = z @ r; p = z.copy()
    for i in range(nit):
        if i:
            p = z + (beta / betaold) * p
        w = A @ p
        a = beta / (p @ w)
        x = x + a * p; r = r - a * w; z = B(r)
        hist.append(np.sqrt(z @ z))
        betaold = beta; beta = z @ r
    return x, np.array(hist)


@pytest.mark.parametrize("pc", ["none", "jacobi"])
@pytest.mark.parametrize("per", [(0, 0, 0), (1, 1, 1), (0, 1, 0)])
def test_cg_history_matches_independent_numpy(pc, per):
    A, b, xs = _problem(per=per)
    res = orc.ksp_solve(A, b, pc_type=pc, rtol=0, atol=0, max_it=40, const_nullspace=True)
    assert res.reason == -3 and res.its == 40 and res.history.size == 41  # DIVERGED_ITS
    x, hist = _numpy_cg(A.to_scipy(), b, pc, True, 40)
    np.testing.assert_allclose(res.history, hist, rtol=1e-10)
    np.testing.assert_allclose(res.x, x, rtol=0, atol=1e-10 * np.abs(x).max())


def test_cg_converges_on_negative_definite_operator():
    # finding 2 of SURVEY.md: DBNG is negative semi-definite; KSPCG only tests sign changes
    A, b, xs = _problem()
    res = orc.ksp_solve(A, b, rtol=1e-12, atol=1e-50, max_it=2000, const_nullspace=True)
    assert res.reason == 2  # CONVERGED_RTOL
    assert res.history[-1] <= 1e-12 * res.history[0]
    assert res.rnorm == res.history[-1]
    assert res.its == res.history.size - 1
    np.testing.assert_allclose(res.x - res.x.mean(), xs, atol=1e-7 * np.abs(xs).max())
    # scipy on -A as a loose cross-check of the solution
    xsci, info = spla.cg(-A.to_scipy(), -b, rtol=1e-12, maxiter=5000)
    assert info == 0
    np.testing.assert_allclose(res.x - res.x.mean(), xsci - xsci.mean(), atol=1e-7 * np.abs(xs).max())


def test_converged_default_semantics():
    A, b, _ = _problem(shape=(8, 8, 4))
    # atol branch: ttol = max(rtol*dp0, atol); reason ATOL iff dp < atol
    r = orc.ksp_solve(A, b, rtol=0, atol=1e-6, max_it=1000, const_nullspace=True)
    assert r.reason == 3 and r.history[-1] < 1e-6 and np.all(r.history[:-1] >= 1e-6)
    r = orc.ksp_solve(A, b, rtol=1e-3, atol=1e-50, max_it=1000, const_nullspace=True)
    assert r.reason == 2 and r.history[-1] <= 1e-3 * r.history[0]
    r = orc.ksp_solve(A, b, rtol=0, atol=0, max_it=3, const_nullspace=True)
    assert r.reason == -3 and r.its == 3
    # zero rhs: dp0 = 0 <= ttol -> converged at iteration 0; 0 < atol(1e-50) -> ATOL
    r = orc.ksp_solve(A, np.zeros_like(b), const_nullspace=True)
    assert r.its == 0 and r.reason == 3 and r.history.size == 1
    # NaN in rhs
    bn = b.copy(); bn[3] = np.nan
    r = orc.ksp_solve(A, bn, const_nullspace=True)
    assert r.reason == -9


def test_indefinite_matrix_detected():
    # p.Ap changes sign between iterations -> KSP_DIVERGED_INDEFINITE_MAT (-10)
    n = 6
    rp = np.arange(n + 1)
    A = orc.Csr.from_arrays(n, n, rp, np.arange(n), [1.0, -2.0, 3.0, -1.0, 2.0, 0.5])
    r = orc.ksp_solve(A, np.ones(n), rtol=1e-14, max_it=50)
    assert r.reason == -10


def test_norm_types():
    A, b, _ = _problem(shape=(8, 6, 4))
    rp = orc.ksp_solve(A, b, pc_type="jacobi", norm_type="preconditioned", rtol=0, atol=0, max_it=10, const_nullspace=True)
    ru = orc.ksp_solve(A, b, pc_type="jacobi", norm_type="unpreconditioned", rtol=0, atol=0, max_it=10, const_nullspace=True)
    rn = orc.ksp_solve(A, b, pc_type="jacobi", norm_type="natural", rtol=0, atol=0, max_it=10, const_nullspace=True)
    np.testing.assert_allclose(rp.x, ru.x, rtol=0, atol=1e-12 * np.abs(rp.x).max())
    np.testing.assert_allclose(rp.x, rn.x, rtol=0, atol=1e-12 * np.abs(rp.x).max())
    assert not np.allclose(rp.history, ru.history)


def test_bcgs_solves_nonsymmetric():
    rng = np.random.default_rng(5)
    A, b, _ = _problem(shape=(8, 7, 5))
    S = A.to_scipy().tolil()
    M = (-S).tocsr()
    import scipy.sparse as sp
    M = (M + sp.identity(M.shape[0]) * 50.0 + sp.diags(rng.uniform(0, 1, M.shape[0] - 1), 1)).tocsr()
    M.sort_indices()
    Ao = orc.Csr.from_arrays(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)
    xs = rng.standard_normal(M.shape[0])
    bb = M @ xs
    r = orc.ksp_solve(Ao, bb, ksp_type="bcgs", pc_type="jacobi", rtol=1e-12, max_it=500)
    assert r.reason == 2
    np.testing.assert_allclose(r.x, xs, atol=1e-8)

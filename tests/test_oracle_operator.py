"""Operator restatement: D, G, BN, D*(BN*G) (createdivergence.cpp, creategradient.cpp,
createbn.cpp, navierstokes.cpp:347-356) against an independent scipy assembly and the
properties SURVEY.md appendix A.1 states (symmetric, zero row sums, closed form)."""
import itertools

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc


def _widths(rng, n, stretched):
    if not stretched:
        return np.full(n, 1.0 / n)
    return rng.uniform(0.5, 1.5, size=n) / n


def _scipy_dbng(widths, periodic, dt):
    """Independent, index-based scipy build of D and G, then D @ (dt*G)."""
    dim = len(widths)
    npc = [len(w) for w in widths] + [1] * (3 - dim)
    w3 = list(widths) + [np.ones(1)] * (3 - dim)
    per = list(periodic) + [0] * (3 - dim)
    nv = [[npc[d] - 1 + per[d] if d == f else npc[d] for d in range(3)] for f in range(dim)]
    voff = np.cumsum([0] + [int(np.prod(nv[f])) for f in range(dim)])
    pN = int(np.prod(npc))

    def pidx(i, j, k):
        return i + npc[0] * (j + npc[1] * k)

    Dr, Dc, Dv, Gr, Gc, Gv = [], [], [], [], [], []
    for f in range(dim):
        for k, j, i in itertools.product(range(nv[f][2]), range(nv[f][1]), range(nv[f][0])):
            s = (i, j, k)[f]
            u = voff[f] + i + nv[f][0] * (j + nv[f][1] * k)
            lo = [i, j, k]
            hi = [i, j, k]
            hi[f] = (s + 1) % npc[f]
            h = 0.5 * (w3[f][hi[f]] + w3[f][s])
            area = np.prod([w3[d][(i, j, k)[d]] for d in range(3) if d != f]) if dim == 3 else w3[1 - f][(i, j, k)[1 - f]] * 1.0
            Gr += [u, u]; Gc += [pidx(*lo), pidx(*hi)]; Gv += [-1.0 / h, 1.0 / h]
            Dr += [pidx(*lo), pidx(*hi)]; Dc += [u, u]; Dv += [area, -area]
    D = sp.csr_matrix((Dv, (Dr, Dc)), shape=(pN, voff[-1]))
    G = sp.csr_matrix((Gv, (Gr, Gc)), shape=(voff[-1], pN))
    return D, G, (D @ (dt * G)).tocsr()


CASES = [
    ((6, 5), (0, 0), True), ((6, 5), (0, 1), True), ((4, 4), (1, 1), False),
    ((5, 4, 3), (0, 0, 0), True), ((4, 3, 5), (1, 0, 1), True), ((3, 3, 3), (1, 1, 1), True),
    ((2, 3), (1, 0), True), ((2, 2, 2), (1, 1, 1), True),
]


@pytest.mark.parametrize("shape,per,stretched", CASES)
def test_dbng_literal_vs_scipy_and_closed(shape, per, stretched):
    rng = np.random.default_rng(7)
    widths = [_widths(rng, n, stretched) for n in shape]
    dt = 0.013
    D = orc.assemble_divergence(widths, per).to_scipy()
    G = orc.assemble_gradient(widths, per).to_scipy()
    Ds, Gs, As = _scipy_dbng(widths, per, dt)
    assert abs(D - Ds).max() == 0.0
    assert abs(G - Gs).max() == 0.0
    A_lit = orc.assemble_dbng(widths, per, dt, literal=True)
    A = A_lit.to_scipy()
    # scipy sums products in a different order: allow a few ulp
    assert abs(A - As).max() <= 4e-16 * abs(As).max()
    # symmetric, zero row sums (constant null space), negative semi-definite diagonal
    assert abs(A - A.T).max() == 0.0
    assert np.abs(A @ np.ones(A.shape[0])).max() <= 1e-13 * abs(A).max()
    assert np.all(A.diagonal() < 0)
    # closed form is bit-identical to the literal MatMatMult pipeline
    A_cl = orc.assemble_dbng(widths, per, dt, literal=False)
    rl, cl, vl = A_lit.arrays()
    rc, cc, vc = A_cl.arrays()
    np.testing.assert_array_equal(rl, rc)
    np.testing.assert_array_equal(cl, cc)
    np.testing.assert_array_equal(vl, vc)


def test_bnhead_rowsum_kat():
    # tests/operators/createbnhead_test.cpp:17-61, N = 1 term: sum of all entries = nx*ny*dt
    nx, ny, dt = 10, 12, 2.3
    B = orc.bnhead_order1(nx * ny, dt).to_scipy()
    assert abs(B.sum() - nx * ny * dt) <= 1.0e-11


def test_neumann_wall_breaks_symmetry():
    # createdivergence.cpp:231-242 with a0 = 1 (singleboundaryneumann.cpp:27) on xMinus
    widths = [np.full(5, 0.2), np.full(4, 0.25)]
    a0 = [1.0, 0.0, 0.0, 0.0, 0.0, 0.0]
    A = orc.assemble_dbng(widths, (0, 0), 0.01, a0=a0, literal=True).to_scipy()
    assert abs(A - A.T).max() > 0


def test_spmv_matches_scipy():
    rng = np.random.default_rng(3)
    widths = [_widths(rng, n, True) for n in (7, 6, 5)]
    A = orc.assemble_dbng(widths, (0, 1, 0), 0.01)
    x = rng.standard_normal(A.shape[0])
    np.testing.assert_allclose(A.spmv(x), A.to_scipy() @ x, rtol=0, atol=1e-15 * abs(A.to_scipy()).max() * 10)


def test_bnhead_row_sum_kat_of_the_reference():
    """The reference's own known-answer test for createBnHead (tests/operators/createbnhead_test.cpp:17-61):
    Op = (2/dt) I on a 10 x 12 grid, dt = 2.3, c = 0.5; the sum of all entries of the N-th order BnHead is
    nx*ny*dt * sum_{t=1..N} (c*dt*val)^(t-1), checked to 1e-11 for N = 1..10."""
    dt, c = 2.3, 0.5
    val = 2.0 / dt
    nx, ny = 10, 12
    n = nx * ny
    Op = orc.Csr.from_arrays(n, n, np.arange(n + 1), np.arange(n), np.full(n, val))
    ans = nx * ny * dt
    for N in range(1, 11):
        B = orc.bnhead(Op, dt, c, N)
        assert B.shape == (n, n)
        if N > 1:
            ans += dt * nx * ny * (c * dt * val) ** (N - 1)
        assert abs(B.to_scipy().sum() - ans) <= 1.0e-11


def test_bnhead_order1_is_dt_identity_and_order2_widens_the_poisson_stencil():
    import scipy.sparse as sp
    from tests import helpers as H

    widths = H.make_widths((7, 6, 5))
    per = [0, 0, 0]
    A, Lap = H.velocity_system(widths, per, dt=0.01, nu=0.01, c=0.5)
    Lo = orc.Csr.from_arrays(Lap.shape[0], Lap.shape[1], Lap.indptr, Lap.indices, Lap.data)
    B1 = orc.bnhead(Lo, 0.01, 0.5 * 0.01, 1).to_scipy()
    assert abs(B1 - 0.01 * sp.identity(Lap.shape[0])).max() == 0.0
    D = orc.assemble_divergence(widths, per)
    G = orc.assemble_gradient(widths, per)
    B2 = orc.bnhead(Lo, 0.01, 0.5 * 0.01, 2)
    dbng1 = orc.assemble_dbng(widths, per, 0.01).to_scipy()
    dbng2 = orc.matmatmult(D, orc.matmatmult(B2, G)).to_scipy()
    # order 2 = order 1 + dt^2 c nu D L G: wider rows (up to 25 points), constants still in the null space
    assert dbng2.getnnz(axis=1).max() > dbng1.getnnz(axis=1).max() >= 7 - 1
    assert abs(dbng2 @ np.ones(dbng2.shape[0])).max() < 1e-12 * abs(dbng2).max()
    ref = dbng1 + (0.01 ** 2) * (0.5 * 0.01) * (D.to_scipy() @ Lap @ G.to_scipy())
    assert abs(dbng2 - ref).max() < 1e-13 * abs(dbng2).max()

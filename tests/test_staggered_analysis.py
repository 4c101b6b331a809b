"""b200ls_staggered_analyze on CPU: the line-coefficient structure read out of an assembled staggered-grid matrix is
LOSSLESS -- rebuilding the matrix from (1-D coefficient arrays, diagonal, remainder) gives back every entry bitwise --
for the velocity system A = I/dt - c nu L (navierstokes.cpp:342-344, createlaplacian.cpp:134-159; test-side assembly
in tests/helpers.velocity_system) and for an IBPM-style modified Poisson system (ibpm.cpp:100-203); matrices that do
not have that structure are refused."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc
from petibm_b200 import B200Error
from petibm_b200.staggered import analyze
from tests import helpers as H


def velocity_dims(shape, per):
    dim = len(shape)
    n = list(shape) + [1] * (3 - dim)
    p = list(per) + [0] * (3 - dim)
    return [[n[d] - (1 if (d == f and not p[d]) else 0) for d in range(3)] for f in range(dim)], p


def rebuild(dims, per, st, nrows):
    """The matrix the structure describes, as scipy CSR with sorted rows."""
    rows, cols, vals = [], [], []
    off = 0
    for f, n in enumerate(dims):
        n0, n1, n2 = n
        stride = (1, n0, n0 * n1)
        size = n0 * n1 * n2
        l = np.arange(size)
        idx = (l % n0, (l // n0) % n1, l // (n0 * n1))
        rows.append(off + l); cols.append(off + l); vals.append(st["diag"][off:off + size])
        for d in range(3):
            cm, cp = st["coef"][f][d]
            s = idx[d]
            wrap = bool(per[d]) and n[d] >= 3
            for coef, step in ((cm, -1), (cp, +1)):
                nb = s + step
                ok = (nb >= 0) & (nb < n[d])
                if wrap:
                    nb = nb % n[d]
                    ok = np.ones_like(ok)
                c = coef[s]
                keep = ok & (c != 0.0)
                rows.append(off + l[keep])
                cols.append(off + l[keep] + (nb[keep] - s[keep]) * stride[d])
                vals.append(c[keep])
        off += size
    rp, rc, rv = st["rem"]
    rr = np.repeat(np.arange(nrows), np.diff(rp))
    rows.append(rr); cols.append(rc.astype(np.int64)); vals.append(rv)
    M = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nrows, nrows))
    M.sort_indices()
    return M


def same_matrix(A, B):
    A = A.tocsr(); A.sort_indices()
    keep = A.data != 0.0                         # explicit zeros carry no information
    if not keep.all():
        A = A.copy(); A.eliminate_zeros()
    return (np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
            and np.array_equal(A.data.view(np.int64), B.data.view(np.int64)))


@pytest.mark.parametrize("shape,per", [((9, 8), (0, 0)), ((9, 8), (1, 0)), ((7, 6, 5), (0, 0, 0)), ((6, 7, 5), (0, 1, 1)),
                                       ((5, 5, 5), (1, 1, 1)), ((12, 3, 4), (0, 0, 0))])
def test_velocity_system_structure_is_lossless(shape, per):
    A, _ = H.velocity_system(H.make_widths(shape), per, dt=0.01, nu=0.02, c=0.5)
    dims, p = velocity_dims(shape, per)
    assert A.shape[0] == sum(int(np.prod(d)) for d in dims)
    st = analyze(dims, p, A.indptr, A.indices, A.data)
    assert st["nsep"] == A.shape[0] and st["rem"][1].size == 0
    assert np.array_equal(st["diag"], A.diagonal())
    assert same_matrix(A, rebuild(dims, p, st, A.shape[0]))
    # stretched grid: the x coefficients of u really vary along x and nowhere else (they are 1-D arrays)
    cm, cp = st["coef"][0][0]
    assert np.unique(cm[1:]).size > 1 and cm[0] == (0.0 if not per[0] else cm[0])
    if not per[0]:
        assert cm[0] == 0.0 and cp[-1] == 0.0        # wall-side neighbours are ghosts: no entry, folded into the diagonal


def _ibpm_like(shape, nf, seed):
    rng = np.random.default_rng(seed)
    widths = H.make_widths(shape)
    G = orc.assemble_gradient(widths, [0, 0, 0]).to_scipy()
    R = sp.random(G.shape[0], nf, density=0.05, random_state=seed, format="csr")
    K = sp.hstack([G, -R]).tocsr()
    M = (-(K.T @ K) * 0.01).tocsr()
    M.sort_indices()
    return M, G.shape[1]


@pytest.mark.parametrize("shape", [(14, 12), (8, 7, 6)])
def test_ibpm_style_system_splits_into_stencil_block_and_remainder(shape):
    M, pN = _ibpm_like(shape, 9, 11)
    dims = [list(shape) + [1] * (3 - len(shape))]
    st = analyze(dims, (0, 0, 0), M.indptr, M.indices, M.data)
    assert st["nsep"] == pN
    rp, rc, rv = st["rem"]
    assert rc.size > 0 and np.all(rc[: rp[pN]] >= pN)             # pressure rows: only force columns are left over
    assert np.array_equal(np.diff(rp)[pN:], np.diff(M.indptr)[pN:])  # force rows: kept whole
    assert same_matrix(M, rebuild(dims, (0, 0, 0), st, M.shape[0]))


def test_the_pressure_operator_itself_is_not_a_line_coefficient_stencil_on_a_stretched_grid():
    """DBNG carries face areas (dy*dz etc.), so its x coefficient varies with j and k: refused -- that operator has its
    own matrix-free form (b200ls_set_poisson_stencil)."""
    shape = (7, 6, 5)
    A = H.oracle_matrix(H.make_widths(shape), (0, 0, 0)).to_scipy().tocsr()
    A.sort_indices()
    with pytest.raises(B200Error) as ei:
        analyze([list(shape)], (0, 0, 0), A.indptr, A.indices, A.data)
    assert ei.value.code == -6 and "line coefficient" in str(ei.value)


def test_refusals():
    shape, per = (7, 6, 5), (0, 0, 0)
    A, _ = H.velocity_system(H.make_widths(shape), per)
    dims, p = velocity_dims(shape, per)
    ok = analyze(dims, p, A.indptr, A.indices, A.data)
    assert ok["nsep"] == A.shape[0]
    # one off-diagonal entry changed in its last bit
    B = A.copy()
    q = B.indptr[40] + (0 if B.indices[B.indptr[40]] != 40 else 1)
    B.data[q] = np.nextafter(B.data[q], 0.0)
    with pytest.raises(B200Error) as ei:
        analyze(dims, p, B.indptr, B.indices, B.data)
    assert ei.value.code == -6
    # an entry outside the 7-point pattern
    C2 = A.tolil(); C2[3, 200] = 0.25; C2 = C2.tocsr(); C2.sort_indices()
    with pytest.raises(B200Error):
        analyze(dims, p, C2.indptr, C2.indices, C2.data)
    # a missing entry
    D = A.tolil(); D[50, 51] = 0.0; D = D.tocsr(); D.eliminate_zeros(); D.sort_indices()
    with pytest.raises(B200Error):
        analyze(dims, p, D.indptr, D.indices, D.data)
    # wrong field layout (fields swapped) and unsorted rows
    with pytest.raises(B200Error):
        analyze(dims[::-1], p, A.indptr, A.indices, A.data)
    U = A.copy()
    a, e = U.indptr[10], U.indptr[11]
    U.indices[a:e] = U.indices[a:e][::-1].copy(); U.data[a:e] = U.data[a:e][::-1].copy()
    with pytest.raises(B200Error):
        analyze(dims, p, U.indptr, U.indices, U.data)
    # periodic axis with two cells cannot be separated
    with pytest.raises(B200Error):
        analyze([[2, 4, 1]], (1, 0, 0), *[getattr(sp.identity(8, format="csr"), k) for k in ("indptr", "indices", "data")])


def test_an_explicit_zero_is_harmless():
    shape, per = (6, 5), (0, 0)
    A, _ = H.velocity_system(H.make_widths(shape), per)
    dims, p = velocity_dims(shape, per)
    Z = A.tolil(); Z[2, 17] = 1.0; Z = Z.tocsr(); Z.sort_indices()
    Z.data[(Z.indices == 17) & (np.repeat(np.arange(Z.shape[0]), np.diff(Z.indptr)) == 2)] = 0.0   # stored, but zero
    st = analyze(dims, p, Z.indptr, Z.indices, Z.data)
    assert same_matrix(A, rebuild(dims, p, st, A.shape[0]))


def _stretched_axis(n_side, n_band, ratio):
    sub = [{"end": 0.6, "cells": n_side, "stretchRatio": 1.0 / ratio}, {"end": 1.4, "cells": n_band, "stretchRatio": 1.0},
           {"end": 2.0, "cells": n_side, "stretchRatio": ratio}]
    return orc.axis_from_subdomains(0.0, sub)


@pytest.mark.parametrize("dim", [2, 3])
def test_ibpm_system_on_a_stretched_grid_is_pressure_operator_plus_remainder(dim):
    """[D;E] BN [G,-H] assembled through the oracle's MatMatMult restatement (tests/helpers.ibpm_system): the pressure
    block is D (dt I) G of the mesh bit for bit, so b200ls_hybrid_analyze accepts it; its coefficients carry face areas,
    so the line-coefficient analysis does not (that is why the hybrid form exists)."""
    from petibm_b200.staggered import analyze_hybrid

    w = _stretched_axis(5, 8, 1.3) if dim == 3 else _stretched_axis(8, 14, 1.2)
    widths = [w.copy() for _ in range(dim)]
    M, pN, nv = H.ibpm_system(widths, dt=0.01, nb=10)
    assert abs(M @ nv).max() < 1e-15 and abs(M - M.T).max() < 1e-14 * abs(M).max()
    st = analyze_hybrid(widths, (0,) * dim, 0.01, M.indptr, M.indices, M.data)
    assert st["nsep"] == pN and np.array_equal(st["diag"], M.diagonal()[:pN])
    rp, rc, rv = st["rem"]
    assert rc.size > 0 and np.all(rc[: rp[pN]] >= pN) and np.array_equal(np.diff(rp)[pN:], np.diff(M.indptr)[pN:])
    # rebuild: closed-form stencil block (area * g) + stored diagonal + remainder == M, bitwise
    n3 = [w.size] * dim + [1] * (3 - dim)
    d3 = widths + [np.ones(1)] * (3 - dim)
    rows, cols, vals = [np.arange(pN)], [np.arange(pN)], [st["diag"]]
    l = np.arange(pN)
    idx = (l % n3[0], (l // n3[0]) % n3[1], l // (n3[0] * n3[1]))
    stride = (1, n3[0], n3[0] * n3[1])
    for d in range(3):
        a, b = (1 if d == 0 else 0), (1 if d == 2 else 2)
        area = d3[a][idx[a]] * d3[b][idx[b]]
        for gsel, step in ((0, -1), (1, +1)):
            c = area * st["g"][d][idx[d] + gsel]
            keep = c != 0.0
            rows.append(l[keep]); cols.append(l[keep] + step * stride[d]); vals.append(c[keep])
    rr = np.repeat(np.arange(M.shape[0]), np.diff(rp))
    rows.append(rr); cols.append(rc.astype(np.int64)); vals.append(rv)
    B = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=M.shape)
    B.sort_indices()
    assert same_matrix(M, B)
    with pytest.raises(B200Error):
        analyze([n3], (0, 0, 0), M.indptr, M.indices, M.data)          # not a line-coefficient stencil
    # a changed pressure entry or a wrong time step is refused
    Bad = M.copy()
    Bad.data[Bad.indptr[5]] = np.nextafter(Bad.data[Bad.indptr[5]], 0.0)
    with pytest.raises(B200Error):
        analyze_hybrid(widths, (0,) * dim, 0.01, Bad.indptr, Bad.indices, Bad.data)
    with pytest.raises(B200Error):
        analyze_hybrid(widths, (0,) * dim, 0.02, M.indptr, M.indices, M.data)

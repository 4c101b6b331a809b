"""DMDA box partition <-> slab partition (b200ls_repart_* of the C ABI) on CPU: every rank's plan is built in one
process and the all-to-all is played with numpy copies, so the index logic is checked against an independent
restatement of PETSc's DMDA ordering (rank = px + m*(py + n*pz), one box per rank numbered
i-fastest, first M mod m ranks one cell longer; src/mesh/cartesianmesh.cpp:500-538, 709-721)."""
import itertools

import numpy as np
import pytest

from petibm_b200.dist import Repart
from petibm_b200.mesh import slab_range


def _split(M, m):
    base, rem = divmod(M, m)
    return [q * base + min(q, rem) for q in range(m + 1)]


def _dmda_reference(n, procs):
    """Independent restatement: for every rank the natural indices of its box in PETSc (box-local, i-fastest) order."""
    nx, ny, nz = n
    m, nn, pp = procs
    sx, sy, sz = _split(nx, m), _split(ny, nn), _split(nz, pp)
    boxes = []
    for pz, py, px in itertools.product(range(pp), range(nn), range(m)):   # rank = px + m*(py + n*pz): px fastest
        k, j, i = np.meshgrid(np.arange(sz[pz], sz[pz + 1]), np.arange(sy[py], sy[py + 1]), np.arange(sx[px], sx[px + 1]),
                              indexing="ij")
        boxes.append((i + nx * (j + ny * k)).reshape(-1))
    return boxes


CASES3 = [((8, 6, 9), (2, 2, 2)), ((10, 7, 9), (2, 3, 1)), ((5, 4, 7), (1, 1, 3)), ((16, 8, 8), (4, 1, 2)),
          ((7, 11, 13), (1, 4, 2)), ((6, 6, 6), (3, 2, 1)), ((4, 4, 8), (1, 1, 1))]
CASES2 = [((12, 10), (2, 2)), ((9, 14), (3, 1)), ((7, 5), (1, 4)), ((33, 17), (4, 2))]


def _plans(dim, n, procs):
    nranks = int(np.prod(procs))
    return [Repart(dim, n, procs, r) for r in range(nranks)]


def _exchange(plans, vecs, src_counts, src_displs, dst_counts, dst_displs, dst_sizes):
    """What MPI_Alltoallv does: block (q -> s) of rank q's send buffer lands at rank s's receive displacement q."""
    out = [np.full(sz, np.nan) for sz in dst_sizes]
    for q, pq in enumerate(plans):
        for s, ps in enumerate(plans):
            c = int(src_counts(pq)[s])
            assert c == int(dst_counts(ps)[q]), "send and receive counts disagree"
            a, b = int(src_displs(pq)[s]), int(dst_displs(ps)[q])
            out[s][b:b + c] = vecs[q][a:a + c]
    return out


@pytest.mark.parametrize("n,procs", CASES3 + CASES2)
def test_box_to_slab_and_back_against_a_restated_dmda_ordering(n, procs):
    dim = len(n)
    n3 = tuple(n) if dim == 3 else (n[0], 1, n[1])
    p3 = tuple(procs) if dim == 3 else (procs[0], 1, procs[1])
    plans = _plans(dim, n, procs)
    ref = _dmda_reference(n3, p3)
    nranks = len(plans)
    total = int(np.prod(n3))
    # sizes, ownership, natural indices of the own rows
    assert sum(p.nbox for p in plans) == total and sum(p.nslab for p in plans) == total
    for r, p in enumerate(plans):
        assert np.array_equal(p.box_rows(), ref[r])
        assert p.slab == slab_range(n3[2], r, nranks)
        assert p.identity == (p3[0] == 1 and p3[1] == 1)
    # PETSc global index -> natural index = concatenation of the boxes
    perm = np.concatenate(ref)
    assert np.array_equal(plans[0].petsc_to_natural(np.arange(total)), perm)
    assert np.array_equal(plans[-1].petsc_to_natural(np.arange(total)[::-1]), perm[::-1])
    # a global field whose value encodes its natural index, handed over in box (PETSc) order
    field = 1000.0 + np.arange(total, dtype=np.float64)
    boxes = [field[ref[r]] for r in range(nranks)]
    recv = _exchange(plans, boxes, lambda p: p.box_counts, lambda p: p.box_displs, lambda p: p.slab_counts,
                     lambda p: p.slab_displs, [p.nslab for p in plans])
    slabs = [p.unpack_slab(recv[r]) for r, p in enumerate(plans)]
    for r, p in enumerate(plans):
        lo, hi = p.slab
        assert np.array_equal(slabs[r], field[lo * n3[0] * n3[1]: hi * n3[0] * n3[1]]), "slab is not in natural order"
    # and back
    send = [p.pack_slab(slabs[r]) for r, p in enumerate(plans)]
    back = _exchange(plans, send, lambda p: p.slab_counts, lambda p: p.slab_displs, lambda p: p.box_counts,
                     lambda p: p.box_displs, [p.nbox for p in plans])
    for r in range(nranks):
        assert np.array_equal(back[r], boxes[r])
    # the send side of box -> slab needs no packing: its blocks tile the local vector in rank order
    for p in plans:
        assert np.array_equal(p.box_displs[p.box_counts > 0],
                              (np.cumsum(p.box_counts) - p.box_counts)[p.box_counts > 0])
        assert int(p.box_counts.sum()) == p.nbox and int(p.slab_counts.sum()) == p.nslab


@pytest.mark.parametrize("n,procs", CASES3 + CASES2)
def test_process_grid_is_found_from_the_local_sizes(n, procs):
    dim = len(n)
    plans = _plans(dim, n, procs)
    sizes = [p.nbox for p in plans]
    cands = Repart.candidates(dim, n, sizes)
    want = tuple(procs) + (1,) * (3 - dim)
    assert want in cands
    # every candidate reproduces the sizes (they differ only in where the cells go, which the matrix check settles)
    for c in cands:
        assert [Repart(dim, n, c[:dim], r).nbox for r in range(len(sizes))] == sizes


def test_uneven_sizes_single_out_the_grid():
    # 10 x 7 x 9 on 2 x 3 x 1: sizes 5*3*9, 5*3*9, 5*2*9, ... are not reproduced by any other factorisation of 6
    n, procs = (10, 7, 9), (2, 3, 1)
    sizes = [p.nbox for p in _plans(3, n, procs)]
    assert Repart.candidates(3, n, sizes) == [procs]


def test_bad_arguments_are_rejected():
    from petibm_b200 import B200Error

    with pytest.raises(B200Error):
        Repart(3, (4, 4, 2), (1, 1, 4), 0)      # more ranks than planes of the slow axis
    with pytest.raises(B200Error):
        Repart(3, (4, 4, 8), (2, 2, 1), 4)      # rank out of range
    with pytest.raises(B200Error):
        Repart(3, (4, 4, 8), (8, 1, 1), 0)      # more processes than cells along x
    with pytest.raises(B200Error):
        Repart(3, (8, 8, 8), (2, 2, 2), 0).petsc_to_natural([8 ** 3])


def test_dmda_split_matches_the_mirror():
    import ctypes as C

    from petibm_b200 import _lib

    L = _lib.lib()
    for M, m in [(256, 8), (10, 3), (7, 7), (9, 4), (1, 1)]:
        out = (C.c_int64 * (m + 1))()
        assert L.b200ls_dmda_split(M, m, out) == 0
        assert list(out) == _split(M, m)
        for r in range(m):
            assert slab_range(M, r, m) == (out[r], out[r + 1])


def test_box_local_rows_of_the_oracle_matrix_map_back_to_the_natural_rows():
    """The per-rank view the multi-GPU checks hand to setMatrix (tests/helpers.box_local_system): every box row,
    with its PETSc column indices translated by the plan, is the corresponding natural row of the assembled DBNG."""
    from tests import helpers as H

    shape, per, procs = (6, 5, 7), (1, 0, 0), (2, 1, 2)
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    rp, col, val = A.arrays()
    b = np.arange(A.shape[0], dtype=np.float64)
    seen = 0
    for rank in range(4):
        M, bl, plan = H.box_local_system(A, b, 3, shape, procs, rank)
        rows = plan.box_rows()
        assert np.array_equal(bl, b[rows]) and M.nrows == rows.size
        nat_cols = plan.petsc_to_natural(M.indices)
        for q, r in enumerate(rows):
            a, e = M.indptr[q], M.indptr[q + 1]
            assert np.all(np.diff(M.indices[a:e]) > 0)          # PETSc keeps rows sorted by its own column index
            got = dict(zip(nat_cols[a:e].tolist(), M.data[a:e].tolist()))
            want = dict(zip(col[rp[r]:rp[r + 1]].tolist(), val[rp[r]:rp[r + 1]].tolist()))
            assert got == want
        seen += rows.size
    assert seen == A.shape[0]

"""Property tests (hypothesis) of the host-side index logic added to the C ABI: the DMDA box <-> slab exchange plan for
random grids and process grids, and the losslessness of the line-coefficient analysis for random stencil matrices."""
import numpy as np
import scipy.sparse as sp
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from petibm_b200.dist import Repart
from petibm_b200.staggered import analyze
from tests.test_repartition import _dmda_reference, _exchange
from tests.test_staggered_analysis import rebuild, same_matrix

_settings = settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])


@st.composite
def grids(draw):
    dim = draw(st.sampled_from([2, 3]))
    procs = [draw(st.integers(1, 3)) for _ in range(dim)]
    nranks = int(np.prod(procs))
    n = [draw(st.integers(procs[d], procs[d] + 7)) for d in range(dim)]
    n[-1] = max(n[-1], nranks)                      # every rank owns at least one plane of the slab axis
    return dim, tuple(n), tuple(procs)


@_settings
@given(grids())
def test_box_slab_round_trip_for_random_process_grids(g):
    dim, n, procs = g
    n3 = n if dim == 3 else (n[0], 1, n[1])
    p3 = procs if dim == 3 else (procs[0], 1, procs[1])
    plans = [Repart(dim, n, procs, r) for r in range(int(np.prod(procs)))]
    ref = _dmda_reference(n3, p3)
    total = int(np.prod(n3))
    field = np.arange(total, dtype=np.float64) * 0.5 - 3.0
    boxes = [field[ref[r]] for r in range(len(plans))]
    for r, p in enumerate(plans):
        assert np.array_equal(p.box_rows(), ref[r])
    recv = _exchange(plans, boxes, lambda p: p.box_counts, lambda p: p.box_displs, lambda p: p.slab_counts,
                     lambda p: p.slab_displs, [p.nslab for p in plans])
    slabs = [p.unpack_slab(recv[r]) for r, p in enumerate(plans)]
    assert np.array_equal(np.concatenate(slabs), field)             # slabs in rank order = natural order
    back = _exchange(plans, [p.pack_slab(slabs[r]) for r, p in enumerate(plans)], lambda p: p.slab_counts,
                     lambda p: p.slab_displs, lambda p: p.box_counts, lambda p: p.box_displs, [p.nbox for p in plans])
    for r in range(len(plans)):
        assert np.array_equal(back[r], boxes[r])
    assert procs + (1,) * (3 - dim) in Repart.candidates(dim, n, [p.nbox for p in plans])


@st.composite
def stencil_matrices(draw):
    """Random block-diagonal stencil matrix with line coefficients (what the velocity system looks like), optionally
    periodic, plus a random remainder behind the blocks."""
    nf = draw(st.integers(1, 3))
    per = [draw(st.booleans()) for _ in range(3)]
    dims = []
    for _ in range(nf):
        dims.append([draw(st.integers(3 if per[d] else 1, 6)) for d in range(3)])
    seed = draw(st.integers(0, 2 ** 31 - 1))
    nextra = draw(st.integers(0, 4))
    return dims, per, seed, nextra


@_settings
@given(stencil_matrices())
def test_line_coefficient_analysis_is_lossless_for_random_stencils(case):
    dims, per, seed, nextra = case
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    off = 0
    for n in dims:
        n0, n1, n2 = n
        size = n0 * n1 * n2
        stride = (1, n0, n0 * n1)
        l = np.arange(size)
        idx = (l % n0, (l // n0) % n1, l // (n0 * n1))
        rows.append(off + l); cols.append(off + l); vals.append(rng.uniform(5.0, 9.0, size))
        for d in range(3):
            if n[d] == 1:
                continue
            cm, cp = rng.uniform(-1.0, -0.1, n[d]), rng.uniform(-1.0, -0.1, n[d])
            wrap = per[d] and n[d] >= 3
            for coef, step in ((cm, -1), (cp, +1)):
                nb = idx[d] + step
                ok = (nb >= 0) & (nb < n[d])
                if wrap:
                    nb, ok = nb % n[d], np.ones_like(ok)
                rows.append(off + l[ok]); cols.append(off + l[ok] + (nb[ok] - idx[d][ok]) * stride[d]); vals.append(coef[idx[d]][ok])
        off += size
    nrows = off + nextra
    if nextra:
        k = 3 * nextra
        rr = np.concatenate([rng.integers(0, nrows, k), np.arange(off, nrows)])
        cc = np.concatenate([rng.integers(off, nrows, k), np.arange(off, nrows)])
        rows.append(rr); cols.append(cc); vals.append(rng.uniform(0.1, 1.0, rr.size))
    M = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nrows, nrows))
    M.sum_duplicates(); M.sort_indices()
    st_ = analyze(dims, per, M.indptr, M.indices, M.data)
    assert st_["nsep"] == off
    assert same_matrix(M, rebuild(dims, per, st_, nrows))

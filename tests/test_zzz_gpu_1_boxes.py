"""Operator recognition for matrix rows that arrive as DMDA boxes (b200ls_verify_csr_rows): one GPU holds the
operator, the rows of every simulated rank are checked against it in box order with translated columns.  The
multi-GPU solve through the box <-> slab exchange is in tests/mgpu_check.py (run_box_case)."""
import ctypes as C

import numpy as np
import pytest

from petibm_b200 import _lib
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _verify(s, M, rows, cols, ulps):
    diff = C.c_double(0.0)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    rc = s._L.b200ls_verify_csr_rows(s._h, M.nrows, rows.ctypes.data_as(_lib._i64p), M.indptr.ctypes.data_as(_lib._i64p),
                                     cols.ctypes.data_as(_lib._i32p), M.data.ctypes.data_as(_lib._dp), int(ulps),
                                     C.byref(diff))
    return rc, diff.value


@pytest.mark.parametrize("shape,per,procs", [((10, 9, 8), (0, 0, 0), (2, 2, 1)), ((9, 8, 7), (1, 0, 1), (2, 1, 2)),
                                             ((14, 11), (0, 1), (2, 2))])
def test_box_rows_verify_against_the_stencil(shape, per, procs):
    import petibm_b200 as pb

    dim = len(shape)
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b = np.zeros(A.shape[0])
    s = pb.LinSolverB200("poisson", "None")
    s.setStencil(H.grid_of(widths, per))
    nranks = int(np.prod(procs))
    for rank in range(nranks):
        M, _, plan = H.box_local_system(A, b, dim, shape, procs, rank)
        rows, cols = plan.box_rows(), plan.petsc_to_natural(M.indices)
        rc, diff = _verify(s, M, rows, cols, 0)
        assert rc == _lib.OK and diff == 0.0
        # the wrong process grid maps the columns elsewhere: rejected
        wrong = tuple(reversed(procs)) if tuple(reversed(procs)) != tuple(procs) else None
        if wrong is not None and all(w <= m for w, m in zip(wrong, shape)) and nranks <= shape[-1]:
            from petibm_b200.dist import Repart

            other = Repart(dim, shape, wrong, rank)
            if other.nbox == M.nrows:
                rc2, _ = _verify(s, M, other.box_rows(), other.petsc_to_natural(M.indices), 4)
                assert rc2 == _lib.ERR_MISMATCH
        # a diagonal that is one unit in the last place off (partition-dependent accumulation order of PETSc's parallel
        # MatMatMult): rejected bitwise, accepted within 4 ulp; an off-diagonal entry gets no such slack
        q = M.nrows // 2
        a, e = M.indptr[q], M.indptr[q + 1]
        d = a + int(np.where(cols[a:e] == rows[q])[0][0])
        keep = M.data[d]
        M.data[d] = np.nextafter(keep, 0.0)
        assert _verify(s, M, rows, cols, 0)[0] == _lib.ERR_MISMATCH
        assert _verify(s, M, rows, cols, 4)[0] == _lib.OK
        M.data[d] = keep * (1.0 + 1e-12)
        assert _verify(s, M, rows, cols, 4)[0] == _lib.ERR_MISMATCH
        M.data[d] = keep
        o = a if a != d else a + 1
        keep = M.data[o]
        M.data[o] = np.nextafter(keep, 0.0)
        assert _verify(s, M, rows, cols, 4)[0] == _lib.ERR_MISMATCH
        M.data[o] = keep
        assert _verify(s, M, rows, cols, 0)[0] == _lib.OK
    s.destroy()

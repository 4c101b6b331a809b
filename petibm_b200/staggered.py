"""Host-side view of b200ls_staggered_analyze (include/b200ls.h): the line-coefficient structure of an assembled
staggered-grid matrix -- the velocity system A = I/dt - c nu L (navierstokes.cpp:342-344) or IBPM's modified Poisson
system (ibpm.cpp:100-203) -- read out of the matrix and verified against it.  No device needed; used by tests and
diagnostics, the solver itself goes through b200ls_set_staggered."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def analyze(dims, periodic, indptr, indices, data):
    """dims: [[n0, n1, n2], ...] per field.  Returns dict(coef=[[(cm, cp) per axis] per field], diag, rem=(indptr,
    indices, data)); raises B200Error(ERR_MISMATCH) with the first offending entry when the matrix does not fit."""
    L = _lib.lib()
    d = np.ascontiguousarray(dims, dtype=np.int64)
    nf = d.shape[0]
    per = (C.c_int * 3)(*([int(bool(p)) for p in periodic] + [0] * 3)[:3])
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float64)
    nrows = indptr.size - 1
    ncoef = L.b200ls_staggered_coef_size(nf, d.ctypes.data_as(_lib._i64p))
    if ncoef < 0:
        raise _lib.B200Error(_lib.ERR_ARG, "bad field description")
    nsep = int(np.prod(d, axis=1).sum())
    coef = np.zeros(ncoef)
    diag = np.zeros(max(nsep, 1))
    rp = np.zeros(nrows + 1, dtype=np.int64)
    rc_, rv = np.zeros(max(data.size, 1), dtype=np.int32), np.zeros(max(data.size, 1))
    err = C.create_string_buffer(256)
    rc = L.b200ls_staggered_analyze(nf, d.ctypes.data_as(_lib._i64p), per, nrows, indptr.ctypes.data_as(_lib._i64p),
                                    indices.ctypes.data_as(_lib._i32p), data.ctypes.data_as(_lib._dp),
                                    coef.ctypes.data_as(_lib._dp), diag.ctypes.data_as(_lib._dp),
                                    rp.ctypes.data_as(_lib._i64p), rc_.ctypes.data_as(_lib._i32p),
                                    rv.ctypes.data_as(_lib._dp), err, 256)
    if rc != _lib.OK:
        raise _lib.B200Error(rc, err.value.decode() or "analysis failed")
    out, pos = [], 0
    for f in range(nf):
        axes = []
        for ax in range(3):
            n = int(d[f, ax])
            axes.append((coef[pos:pos + n].copy(), coef[pos + n:pos + 2 * n].copy()))
            pos += 2 * n
        out.append(axes)
    nrem = int(rp[-1])
    return {"coef": out, "diag": diag[:nsep].copy(), "rem": (rp, rc_[:nrem].copy(), rv[:nrem].copy()), "nsep": nsep}


def analyze_hybrid(widths, periodic, dt, indptr, indices, data):
    """b200ls_hybrid_analyze: the pressure operator D (dt I) G of the grid (checked bitwise against the closed form)
    followed by a CSR remainder -- IBPM's modified Poisson system on a stretched grid.  Returns dict(g=[gx, gy, gz],
    diag, rem=(indptr, indices, data), nsep)."""
    L = _lib.lib()
    dim = len(widths)
    w = [np.ascontiguousarray(a, dtype=np.float64) for a in widths]
    n = (C.c_int64 * 3)(*([a.size for a in w] + [1] * (3 - dim)))
    per = (C.c_int * 3)(*([int(bool(p)) for p in periodic] + [0] * 3)[:3])
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float64)
    nrows = indptr.size - 1
    n3 = [a.size for a in w] + [1] * (3 - dim)
    nsep = int(np.prod(n3))
    g = np.zeros(sum(n3) + 3)
    diag = np.zeros(max(nsep, 1))
    rp = np.zeros(nrows + 1, dtype=np.int64)
    rc_, rv = np.zeros(max(data.size, 1), dtype=np.int32), np.zeros(max(data.size, 1))
    err = C.create_string_buffer(256)
    dz = w[2].ctypes.data_as(_lib._dp) if dim == 3 else None
    rc = L.b200ls_hybrid_analyze(dim, n, per, w[0].ctypes.data_as(_lib._dp), w[1].ctypes.data_as(_lib._dp), dz, float(dt), nrows,
                                 indptr.ctypes.data_as(_lib._i64p), indices.ctypes.data_as(_lib._i32p),
                                 data.ctypes.data_as(_lib._dp), g.ctypes.data_as(_lib._dp), diag.ctypes.data_as(_lib._dp),
                                 rp.ctypes.data_as(_lib._i64p), rc_.ctypes.data_as(_lib._i32p), rv.ctypes.data_as(_lib._dp), err, 256)
    if rc != _lib.OK:
        raise _lib.B200Error(rc, err.value.decode() or "analysis failed")
    gs, pos = [], 0
    for m in n3:
        gs.append(g[pos:pos + m + 1].copy())
        pos += m + 1
    nrem = int(rp[-1])
    return {"g": gs, "diag": diag[:nsep].copy(), "rem": (rp, rc_[:nrem].copy(), rv[:nrem].copy()), "nsep": nsep}

"""Geometric multigrid preconditioner (petibm_b200/csrc/mg_kernels.cuh + the shared schedule mg_schedule.h) on the CPU
emulation of the kernel sources, against the independent numpy/scipy restatement in tests/mg_reference.py:
one V-cycle z = M^-1 r, preconditioned CG histories, and what the preconditioner is for -- the same converged
solution as plain CG in a fraction of the iterations.  The emulated plain-CG path and the oracle are the anchors."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers as H
from tests import mg_reference as R
from tests.test_emulated_kernels import emu, _grid_args, _cg, _dp, _ip  # noqa: F401  (emu is a fixture)


def _mg(L, widths, per, b, mode="pcg", has_const=True, rtol=0.0, atol=0.0, max_it=20, levels=0, smooth=2, coarse=16, tile=10,
        tail_cells=0, fuse=0):
    dim, n, p, w, dz = _grid_args(widths, per)
    L.emu_mg.argtypes = [C.c_int, C.POINTER(C.c_int64), _ip, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _ip, _ip, _ip, _ip]
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty_like(b)
    hist = np.zeros(max_it + 2)
    nh, its, reason, nl = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    rc = L.emu_mg(dim, n, p, w[0].ctypes.data_as(_dp), w[1].ctypes.data_as(_dp), dz, 0.01, {"apply": 0, "pcg": 1}[mode],
                  int(has_const), rtol, atol, max_it, levels, smooth, coarse, tile, int(tail_cells), int(fuse), b.ctypes.data_as(_dp),
                  x.ctypes.data_as(_dp),
                  hist.ctypes.data_as(_dp), hist.size, C.byref(nh), C.byref(its), C.byref(reason), C.byref(nl))
    assert rc == 0
    return x, hist[: nh.value].copy(), its.value, reason.value, nl.value


CASES = [((16, 12, 8), (0, 0, 0)), ((13, 9, 10), (0, 0, 0)), ((12, 16, 8), (1, 0, 1)), ((24, 20), (0, 0)), ((17, 12), (1, 1)),
         ((8, 8, 8), (1, 1, 1))]


@pytest.mark.parametrize("shape,per", CASES)
def test_reference_fine_operator_is_the_oracle_operator(shape, per):
    """The face-sum assembly of the checker equals the pinned oracle's literal D (dt I) G (tests/test_oracle_*.py)."""
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per).to_scipy()
    V = R.VCycle(widths, per, 0.01)
    diff = abs(V.levels[0]["A"] - A).max()
    assert diff <= 1e-14 * abs(A).max()
    # every level keeps the constant in its null space and stays symmetric negative semi-definite
    for lev in V.levels:
        Al = lev["A"]
        assert abs(Al @ np.ones(Al.shape[0])).max() <= 1e-12 * max(abs(Al).max(), 1e-300)
        assert abs(Al - Al.T).max() <= 1e-14 * max(abs(Al).max(), 1e-300)


@pytest.mark.parametrize("shape,per", CASES)
@pytest.mark.parametrize("smooth", [1, 2, 3])
def test_emulated_vcycle_matches_the_restatement(emu, shape, per, smooth):
    widths = H.make_widths(shape)
    V = R.VCycle(widths, per, 0.01, smooth_its=smooth, coarse_its=7)
    rng = np.random.default_rng(5)
    r = rng.standard_normal(int(np.prod(shape)))
    r -= r.mean()
    z, _, _, _, nl = _mg(emu, widths, per, r, mode="apply", smooth=smooth, coarse=7)
    assert nl == len(V.levels) and nl >= 2
    zr = V.apply(r)
    np.testing.assert_allclose(z, zr, rtol=0, atol=1e-11 * np.abs(zr).max())
    # the preconditioner is a symmetric operator (PCG needs that): <M u, v> = <u, M v>
    u = rng.standard_normal(r.size); u -= u.mean()
    zu, _, _, _, _ = _mg(emu, widths, per, u, mode="apply", smooth=smooth, coarse=7)
    assert abs(zu @ r - u @ z) <= 1e-10 * abs(zu @ r)


def test_emulated_vcycle_with_a_level_cap(emu):
    """-pc_mg_levels: the hierarchy stops early and the (now larger) last level gets the long Chebyshev polynomial."""
    shape, per = (16, 12, 8), (0, 0, 0)
    widths = H.make_widths(shape)
    r = np.random.default_rng(9).standard_normal(int(np.prod(shape)))
    r -= r.mean()
    for cap in (1, 2):
        V = R.VCycle(widths, per, 0.01, max_levels=cap, coarse_its=9)
        z, _, _, _, nl = _mg(emu, widths, per, r, mode="apply", levels=cap, coarse=9)
        assert nl == cap == len(V.levels)
        zr = V.apply(r)
        np.testing.assert_allclose(z, zr, rtol=0, atol=1e-11 * np.abs(zr).max())


def _petibm_like_axis(n_band, n_side, ratio):
    """Uniform fine band with geometrically stretched cells on both sides, like the axes of the shipped examples
    (examples/ibpm/cylinder2dRe100_GPU/config.yaml: stretchRatio 1.02 over 125 cells next to a 200-cell band)."""
    side = [{"end": -0.5, "cells": n_side, "stretchRatio": 1.0 / ratio}, {"end": 0.5, "cells": n_band, "stretchRatio": 1.0},
            {"end": 6.0, "cells": n_side, "stretchRatio": ratio}]
    return orc.axis_from_subdomains(-6.0, side)


def test_stretched_grid_needs_and_gets_the_width_equalising_coarsening(emu):
    """Aspect ratios like PetIBM's grids: merging every pair on every axis keeps the anisotropy on all levels and the
    point smoother stalls; coarsening the narrowest cells first does not.  Checked with the restatement (both rules) and
    with the emulated kernels (the rule that is built)."""
    w = _petibm_like_axis(20, 14, 1.25)
    widths = [w, w.copy()]
    assert w.max() / w.min() > 15
    A = H.oracle_matrix(widths, (0, 0))
    b, xs = H.consistent_rhs(A)
    As = A.to_scipy()
    V = R.VCycle(widths, (0, 0), 0.01)
    _, _, its_eq, reason = R.pcg(As, b, V.apply, True, 1e-8, 0.0, 300)
    Vp = R.VCycle(widths, (0, 0), 0.01)
    Vp.levels = R.hierarchy(widths, (0, 0), 0.01, ratio=1e30)      # every pair merges: plain 2:1 coarsening
    for lev in Vp.levels:
        dg = lev["A"].diagonal()
        lev["dinv"] = np.where(dg != 0.0, 1.0 / np.where(dg != 0.0, dg, 1.0), 0.0)
    _, _, its_pair, _ = R.pcg(As, b, Vp.apply, True, 1e-8, 0.0, 300)
    assert reason == 2 and its_eq <= 22 and its_pair >= 2 * its_eq, (its_eq, its_pair)
    x, hist, its, reason, nl = _mg(emu, widths, (0, 0), b, rtol=1e-8, max_it=60)
    assert reason == 2 and abs(its - its_eq) <= 1 and nl == len(V.levels), (its, its_eq, nl, len(V.levels))
    np.testing.assert_allclose(x, xs, rtol=0, atol=1e-6 * np.abs(xs).max())


@pytest.mark.parametrize("shape,per", [((16, 12, 8), (0, 0, 0)), ((12, 16, 8), (1, 0, 1)), ((24, 20), (0, 0)), ((13, 9, 10), (0, 0, 0))])
def test_emulated_pcg_history_and_iteration_count(emu, shape, per):
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b, xs = H.consistent_rhs(A)
    V = R.VCycle(widths, per, 0.01)
    As = A.to_scipy()
    # fixed number of iterations: histories agree with the restatement
    nit = 6
    xr, hr, _, _ = R.pcg(As, b, V.apply, True, 0.0, 0.0, nit)
    x, hist, its, reason, _ = _mg(emu, widths, per, b, max_it=nit)
    assert (its, reason) == (nit, -3) and hist.size == nit + 1
    np.testing.assert_allclose(hist, hr, rtol=1e-8)
    np.testing.assert_allclose(x, xr, rtol=0, atol=1e-9 * np.abs(xr).max())
    # to convergence: a handful of iterations instead of plain CG's dozens, same solution
    x, hist, its, reason, _ = _mg(emu, widths, per, b, rtol=1e-10, max_it=60)
    plain = orc.ksp_solve(A, b, rtol=1e-10, atol=1e-50, max_it=2000, const_nullspace=True)
    assert reason == 2 and plain.reason == 2
    assert its <= 25 and its * 3 <= plain.its, (its, plain.its)
    np.testing.assert_allclose(x, xs, rtol=0, atol=1e-7 * np.abs(xs).max())
    np.testing.assert_allclose(x, plain.x, rtol=0, atol=1e-7 * np.abs(xs).max())
    assert np.all(np.diff(np.log(hist)) < 0)      # monotone decrease of ||z||: a fixed SPD preconditioner


def _hybrid_mg(L, widths, M, b, nullvec=None, has_const=False, rtol=0.0, atol=0.0, max_it=20, smooth=2, coarse=16, dt=0.01):
    from petibm_b200.staggered import analyze_hybrid

    dim = len(widths)
    M = M.tocsr(); M.sort_indices()
    st = analyze_hybrid(widths, (0,) * dim, dt, M.indptr, M.indices, M.data)
    w3 = [np.asarray(a, dtype=np.float64) for a in widths] + [np.ones(1)] * (3 - dim)
    n3 = (C.c_int64 * 3)(*[a.size for a in w3])
    per = (C.c_int * 3)(0, 0, 0)
    wpack = np.ascontiguousarray(np.concatenate(w3)); coef = np.ascontiguousarray(np.concatenate(st["g"]))
    diag = np.ascontiguousarray(st["diag"])
    rp, rc, rv = st["rem"]
    rc = np.ascontiguousarray(rc if rc.size else np.zeros(1, dtype=np.int32)); rv = np.ascontiguousarray(rv if rv.size else np.zeros(1))
    dg = M.diagonal()
    dinv = np.where(dg != 0.0, 1.0 / np.where(dg != 0.0, dg, 1.0), 1.0)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty_like(b); hist = np.zeros(max_it + 2)
    nh, its, reason = C.c_int(0), C.c_int(0), C.c_int(0)
    nvp = None if nullvec is None else np.ascontiguousarray(nullvec, dtype=np.float64).ctypes.data_as(_dp)
    L.emu_hybrid_mg_pcg.argtypes = [C.c_int, C.POINTER(C.c_int64), _ip, _dp, C.c_double, C.c_int64, _dp, _dp, C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int32), _dp, _dp, C.c_int, _dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                    _dp, _dp, _dp, C.c_int, _ip, _ip, _ip]
    rcode = L.emu_hybrid_mg_pcg(dim, n3, per, wpack.ctypes.data_as(_dp), dt, M.shape[0], coef.ctypes.data_as(_dp),
                                diag.ctypes.data_as(_dp), rp.ctypes.data_as(C.POINTER(C.c_int64)),
                                rc.ctypes.data_as(C.POINTER(C.c_int32)), rv.ctypes.data_as(_dp), dinv.ctypes.data_as(_dp),
                                int(has_const), nvp, rtol, atol, max_it, smooth, coarse, b.ctypes.data_as(_dp), x.ctypes.data_as(_dp),
                                hist.ctypes.data_as(_dp), hist.size, C.byref(nh), C.byref(its), C.byref(reason))
    assert rcode == 0
    return x, hist[: nh.value].copy(), its.value, reason.value


@pytest.mark.parametrize("dim", [2, 3])
def test_emulated_block_preconditioner_on_the_ibpm_system(emu, dim):
    """IBPM's modified Poisson system on a stretched grid, CG with M^-1 = diag(V-cycle on the pressure block, 1/diag on the
    force rows) and the explicit null-space vector the application attaches (ibpm.cpp:251-267): histories equal the
    restatement's, and the solve needs a fraction of Jacobi-CG's iterations (what the shipped configuration runs with
    AMG instead)."""
    n_side, n_band = (5, 8) if dim == 3 else (10, 20)
    sub = [{"end": 0.6, "cells": n_side, "stretchRatio": 1.0 / 1.25}, {"end": 1.4, "cells": n_band, "stretchRatio": 1.0},
           {"end": 2.0, "cells": n_side, "stretchRatio": 1.25}]
    w = orc.axis_from_subdomains(0.0, sub)
    widths = [w.copy() for _ in range(dim)]
    M, pN, nv = H.ibpm_system(widths, dt=0.01, nb=12)
    Mo = orc.Csr.from_arrays(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)
    rng = np.random.default_rng(21)
    xs = rng.standard_normal(M.shape[0]); xs -= (xs @ nv) * nv
    b = Mo.spmv(xs)
    V = R.VCycle(widths, (0,) * dim, 0.01)
    dg = M.diagonal()

    def block_pc(v):
        z = np.empty_like(v)
        z[:pN] = V.apply(v[:pN])
        z[pN:] = v[pN:] / dg[pN:]
        return z

    nit = 6
    xr, hr, _, _ = R.pcg(M, b, block_pc, False, 0.0, 0.0, nit, nullvec=nv)
    x, hist, its, reason = _hybrid_mg(emu, widths, M, b, nullvec=nv, max_it=nit)
    assert (its, reason) == (nit, -3)
    np.testing.assert_allclose(hist, hr, rtol=1e-8)
    np.testing.assert_allclose(x, xr, rtol=0, atol=1e-9 * np.abs(xr).max())
    # the same solve with the operator product on the tiled kernels of sep_tile.cuh (tuning "sep_tile")
    emu.emu_set_sep_tile.argtypes = [C.c_int, C.c_int, C.c_int]
    for xr_tile in (2, 4):
        emu.emu_set_sep_tile(xr_tile, 0, 24)
        try:
            xt, ht, it_t, reason_t = _hybrid_mg(emu, widths, M, b, nullvec=nv, max_it=nit)
        finally:
            emu.emu_set_sep_tile(0, 0, 0)
        assert (it_t, reason_t) == (nit, -3)
        np.testing.assert_allclose(ht, hist, rtol=1e-10)
        np.testing.assert_allclose(xt, x, rtol=0, atol=1e-11 * np.abs(x).max())
    x, hist, its, reason = _hybrid_mg(emu, widths, M, b, nullvec=nv, rtol=1e-9, max_it=200)
    jac = orc.ksp_solve(Mo, b, pc_type="jacobi", rtol=1e-9, atol=1e-50, max_it=5000, nullvecs=nv)
    assert reason == 2 and jac.reason == 2 and 3 * its <= jac.its, (its, jac.its)
    np.testing.assert_allclose(x, xs, rtol=0, atol=1e-6 * np.abs(xs).max())


def test_emulated_mg_edge_cases(emu):
    """Zero and NaN right-hand sides end like KSP ends them (KSPConvergedDefault: converged at iteration 0 with a zero
    residual; KSP_DIVERGED_NANORINF), and an already tiny grid (one level) still works."""
    shape, per = (12, 8, 8), (0, 0, 0)
    widths = H.make_widths(shape)
    n = int(np.prod(shape))
    x, hist, its, reason, _ = _mg(emu, widths, per, np.zeros(n), rtol=1e-8, atol=1e-50, max_it=10)
    assert (its, reason) == (0, 3) and hist.size == 1 and hist[0] == 0.0 and not np.any(x)
    b = np.ones(n); b[5] = np.nan
    x, hist, its, reason, _ = _mg(emu, widths, per, b, rtol=1e-8, max_it=10)
    assert reason == -9 and its == 0
    # one level only: the "coarse" polynomial is the whole preconditioner
    tiny = H.make_widths((3, 3, 3))
    A = H.oracle_matrix(tiny, (0, 0, 0))
    b, xs = H.consistent_rhs(A)
    x, hist, its, reason, nl = _mg(emu, tiny, (0, 0, 0), b, rtol=1e-10, max_it=60)
    assert nl == 1 and reason == 2
    np.testing.assert_allclose(x, xs, rtol=0, atol=1e-8 * np.abs(xs).max())


@pytest.mark.parametrize("shape,per,smooth", [((16, 12, 8), (0, 0, 0), 2), ((12, 16, 8), (1, 0, 1), 1), ((24, 20), (0, 0), 3)])
def test_emulated_single_cta_tail_gives_the_same_numbers(emu, shape, per, smooth):
    """Tuning "mg_tail": the coarse levels of the cycle interpreted by ONE CTA (k_mg_tail) from the recorded schedule --
    same bodies, same order, so z = M^-1 r and the PCG history are identical bit for bit to the launch-per-step cycle."""
    widths = H.make_widths(shape)
    r = np.random.default_rng(5).standard_normal(int(np.prod(shape)))
    r -= r.mean()
    z0, _, _, _, nl = _mg(emu, widths, per, r, mode="apply", smooth=smooth, coarse=7)
    for cells in (2000, 200, 30):                      # tail starts at different levels
        z1, _, _, _, _ = _mg(emu, widths, per, r, mode="apply", smooth=smooth, coarse=7, tail_cells=cells)
        assert np.array_equal(z0, z1), cells
    A = H.oracle_matrix(widths, per)
    b, _ = H.consistent_rhs(A)
    x0, h0, i0, r0, _ = _mg(emu, widths, per, b, rtol=1e-9, max_it=40, smooth=smooth)
    x1, h1, i1, r1, _ = _mg(emu, widths, per, b, rtol=1e-9, max_it=40, smooth=smooth, tail_cells=200)
    assert (i0, r0) == (i1, r1) and np.array_equal(h0, h1) and np.array_equal(x0, x1)


@pytest.mark.parametrize("shape,per,smooth", [((16, 12, 8), (0, 0, 0), 1), ((12, 16, 8), (1, 0, 1), 2), ((24, 20), (0, 0), 3),
                                              ((3, 3, 3), (0, 0, 0), 1), ((3, 3, 3), (0, 0, 0), 2)])
def test_emulated_fused_first_and_last_steps_give_the_same_numbers(emu, shape, per, smooth):
    """Tuning "mg_fuse": r <- r - a w inside the first fine-level step and the six sums inside the last one (24 + 16 B/row
    and two launches less per iteration) -- same arithmetic, same thread-to-cell mapping, so the PCG history and the
    solution are identical bit for bit; a single-level hierarchy falls back to the separate kernels where it has to."""
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b, _ = H.consistent_rhs(A)
    x0, h0, i0, r0, _ = _mg(emu, widths, per, b, rtol=1e-9, max_it=40, smooth=smooth, coarse=smooth + 1)
    for tail in (0, 200):
        x1, h1, i1, r1, _ = _mg(emu, widths, per, b, rtol=1e-9, max_it=40, smooth=smooth, coarse=smooth + 1, fuse=1, tail_cells=tail)
        assert (i0, r0) == (i1, r1) and np.array_equal(h0, h1) and np.array_equal(x0, x1)

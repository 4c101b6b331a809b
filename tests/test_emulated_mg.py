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


def _mg(L, widths, per, b, mode="pcg", has_const=True, rtol=0.0, atol=0.0, max_it=20, levels=0, smooth=2, coarse=16, tile=10):
    dim, n, p, w, dz = _grid_args(widths, per)
    L.emu_mg.argtypes = [C.c_int, C.POINTER(C.c_int64), _ip, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _ip, _ip, _ip, _ip]
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty_like(b)
    hist = np.zeros(max_it + 2)
    nh, its, reason, nl = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    rc = L.emu_mg(dim, n, p, w[0].ctypes.data_as(_dp), w[1].ctypes.data_as(_dp), dz, 0.01, {"apply": 0, "pcg": 1}[mode],
                  int(has_const), rtol, atol, max_it, levels, smooth, coarse, tile, b.ctypes.data_as(_dp), x.ctypes.data_as(_dp),
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

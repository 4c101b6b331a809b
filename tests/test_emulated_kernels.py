"""Kernel LOGIC on a CPU-only machine: the unmodified sources of petibm_b200/csrc/{kernels,spmv2}.cuh are compiled
with g++ against tests/emu/cuda_emu.h (fiber emulation of threads, barriers, warp shuffles and late-completing
cp.async copies; -DB200_EMULATE swaps csrc/hw.cuh) and driven by tests/emu/emu_solver.cpp, then compared with the
oracle exactly like the GPU parity tests.  This is test scaffolding -- not a fallback: libb200ls.so cannot reach
it -- and it says nothing about timing or the GPU memory model; it lets the indexing / halo / ring-buffer / KSP
state-machine logic of a kernel change be checked before GPU time is spent on it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
LIB = os.path.join(EMU, "libb200emu.so")
SRC = [os.path.join(EMU, f) for f in ("emu_solver.cpp", "cuda_emu.h")] + \
      [os.path.join(ROOT, "petibm_b200", "csrc", f) for f in ("kernels.cuh", "spmv2.cuh", "spmv3.cuh", "hw.cuh", "csr_kernels.cuh",
                                                            "sep_kernels.cuh", "sep_tile.cuh", "mg_kernels.cuh", "mg_schedule.h",
                                                            "ops_kernels.cuh", "update_fly.cuh", "dense_kernels.cuh", "spmv4.cuh", "spmv5.cuh")]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in SRC):
        # hidden visibility + -Bsymbolic: the kernel functions must bind inside this library, not to the CUDA host
        # stubs of the same (mangled) names that libb200ls.so exports when it is loaded in the same process
        cmd = ["g++", "-std=c++17", "-O1", "-DB200_EMULATE", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
               "-fvisibility=hidden", "-fvisibility-inlines-hidden", "-Wl,-Bsymbolic",
               "-I", EMU, "-I", os.path.join(ROOT, "petibm_b200", "csrc"), "-o", LIB, SRC[0]]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-3000:]
    L = C.CDLL(LIB)
    L.emu_stencil_apply.argtypes = [C.c_int, C.POINTER(C.c_int64), _ip, _dp, _dp, _dp, C.c_double, C.c_int, _dp, _dp]
    L.emu_stencil_cg.argtypes = [C.c_int, C.POINTER(C.c_int64), _ip, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_int,
                                 C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp,
                                 _dp, C.c_int, _ip, _ip, _ip, _dp]
    L.emu_stencil_cg_ranks.argtypes = [C.c_int, C.POINTER(C.c_int64), _ip, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_double,
                                       C.c_double, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _ip, _ip, _ip]
    L.emu_set_schedule.argtypes = [C.c_uint64]
    L.emu_csr_solve.argtypes = [C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32), _dp, C.c_int, C.c_int, C.c_int, _dp,
                                C.c_double, C.c_double, C.c_int, _dp, _dp, _dp, C.c_int, _ip, _ip, _ip]
    L.emu_sep_solve.argtypes = [C.c_int, C.POINTER(C.c_int64), _ip, _dp, C.c_int64, _dp, _dp, C.POINTER(C.c_int64),
                                C.POINTER(C.c_int32), _dp, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_double, C.c_double, C.c_int,
                                _dp, _dp, _dp, C.c_int, _ip, _ip, _ip]
    L.emu_set_sep_tile.argtypes = [C.c_int, C.c_int, C.c_int]
    L.emu_set_sep_stages.argtypes = [C.c_int]
    return L


def _grid_args(widths, per):
    dim = len(widths)
    n = (C.c_int64 * 3)(*([len(w) for w in widths] + [1] * (3 - dim)))
    p = (C.c_int * 3)(*(list(per) + [0] * (3 - dim)))
    w = [np.ascontiguousarray(a, dtype=np.float64) for a in widths]
    dz = w[2].ctypes.data_as(_dp) if dim == 3 else None
    return dim, n, p, w, dz


def _apply(L, widths, per, x, kz=0):
    dim, n, p, w, dz = _grid_args(widths, per)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    L.emu_stencil_apply(dim, n, p, w[0].ctypes.data_as(_dp), w[1].ctypes.data_as(_dp), dz, 0.01, kz,
                        x.ctypes.data_as(_dp), y.ctypes.data_as(_dp))
    return y


def _cg(L, widths, per, b, pc="none", has_const=True, norm=1, rtol=0.0, atol=0.0, max_it=20, tile=10, kz=0, ub=3, rev=1):
    dim, n, p, w, dz = _grid_args(widths, per)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty_like(b)
    hist = np.zeros(max_it + 2)
    nh, its, reason, rn = C.c_int(0), C.c_int(0), C.c_int(0), C.c_double(0)
    rc = L.emu_stencil_cg(dim, n, p, w[0].ctypes.data_as(_dp), w[1].ctypes.data_as(_dp), dz, 0.01, int(pc == "jacobi"),
                          int(has_const), norm, rtol, atol, 1e4, max_it, tile, kz, ub, rev, b.ctypes.data_as(_dp),
                          x.ctypes.data_as(_dp), hist.ctypes.data_as(_dp), hist.size, C.byref(nh), C.byref(its),
                          C.byref(reason), C.byref(rn))
    assert rc == 0
    return x, hist[: nh.value].copy(), its.value, reason.value, rn.value


SHAPES = [((12, 10, 8), (0, 0, 0)), ((9, 7, 5), (1, 1, 1)), ((70, 13, 6), (0, 1, 0)), ((67, 9, 7), (1, 0, 1)),
          ((33, 15), (0, 0)), ((20, 12), (1, 1)), ((5, 3, 3), (0, 0, 0))]


@pytest.mark.parametrize("shape,per", SHAPES)
def test_emulated_spmv_is_bit_exact(emu, shape, per):
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    x = np.random.default_rng(3).standard_normal(A.shape[0])
    for kz in (0, 3):                      # one z-chunk and several (chunk halos)
        if len(shape) == 2 and kz:
            continue
        assert np.array_equal(_apply(emu, widths, per, x, kz), A.spmv(x))


@pytest.mark.parametrize("tile", [10, 18, 13, 15])
@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_emulated_cg_matches_oracle(emu, tile, pc):
    for shape, per, kz in (((12, 10, 8), (0, 0, 0), 3), ((70, 9, 5), (0, 1, 1), 0)):
        widths = H.make_widths(shape)
        A = H.oracle_matrix(widths, per)
        b, _ = H.consistent_rhs(A)
        nit = 15
        ref = orc.ksp_solve(A, b, pc_type=pc, rtol=0, atol=0, max_it=nit, const_nullspace=True)
        x, hist, its, reason, rn = _cg(emu, widths, per, b, pc=pc, max_it=nit, tile=tile, kz=kz)
        assert (its, reason) == (ref.its, ref.reason) == (nit, -3) and hist.size == nit + 1 and rn == hist[-1]
        np.testing.assert_allclose(hist, ref.history, rtol=1e-10)
        np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-9 * np.abs(ref.x).max())


_TMA_REFERENCE = {}


@pytest.mark.parametrize("tile", [40, 41, 42, 43, 46, 47, 50, 51, 52, 53, 60, 62])
@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_emulated_tma_kernel_matches_oracle(emu, tile, pc):
    """k_spmv4 (TMA boxes with zero fill outside the grid, mbarrier full/empty pipeline, neighbours' p rebuilt from the
    staged r / p', row sum closed one plane later): ragged tiles in x and y, several z chunks, 2-D, tiny grids."""
    for shape, per, kz in (((12, 10, 8), (0, 0, 0), 3), ((70, 9, 5), (0, 0, 0), 0), ((67, 21, 7), (0, 0, 0), 2),
                           ((33, 31), (0, 0), 0), ((5, 3, 3), (0, 0, 0), 1), ((1, 1, 7), (0, 0, 0), 0)):
        widths = H.make_widths(shape)
        key = (shape, pc)
        if key not in _TMA_REFERENCE:                # the oracle solve and the cp.async kernel's: once per case, not per tile
            A = H.oracle_matrix(widths, per)
            b, _ = H.consistent_rhs(A)
            nit = min(12, A.shape[0] - 3)            # 7 unknowns: exact convergence after 6 iterations, noise beyond
            ref = orc.ksp_solve(A, b, pc_type=pc, rtol=0, atol=0, max_it=nit, const_nullspace=True)
            _, hist2, _, _, _ = _cg(emu, widths, per, b, pc=pc, max_it=nit, tile=10, kz=kz)
            _TMA_REFERENCE[key] = (b, nit, ref, hist2)
        b, nit, ref, hist2 = _TMA_REFERENCE[key]
        x, hist, its, reason, rn = _cg(emu, widths, per, b, pc=pc, max_it=nit, tile=tile, kz=kz)
        assert (its, reason) == (ref.its, ref.reason) == (nit, -3) and hist.size == nit + 1 and rn == hist[-1]
        np.testing.assert_allclose(hist, ref.history, rtol=1e-10)
        np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-9 * np.abs(ref.x).max())
        # against the cp.async kernel: the same numbers up to the order of the block partial sums of p.w
        np.testing.assert_allclose(hist, hist2, rtol=1e-12)


def test_emulated_tma_kernel_does_not_depend_on_the_thread_schedule(emu):
    shape, per = (70, 13, 9), (0, 0, 0)
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b, _ = H.consistent_rhs(A)
    for tile in (40, 41, 46):
        x0, h0, _, _, _ = _cg(emu, widths, per, b, pc="jacobi", max_it=6, tile=tile, kz=3)
        try:
            for seed in (1, 2, 3):
                emu.emu_set_schedule(seed)
                x, h, _, _, _ = _cg(emu, widths, per, b, pc="jacobi", max_it=6, tile=tile, kz=3)
                assert np.array_equal(h, h0) and np.array_equal(x, x0)
        finally:
            emu.emu_set_schedule(0)


def test_emulated_convergence_logic_and_traversal_order(emu):
    shape, per = (10, 8, 6), (0, 0, 0)
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b, _ = H.consistent_rhs(A)
    ref = orc.ksp_solve(A, b, rtol=1e-6, atol=1e-50, max_it=500, const_nullspace=True)
    out = {}
    for rev in (0, 1):
        x, hist, its, reason, _ = _cg(emu, widths, per, b, rtol=1e-6, atol=1e-50, max_it=500, rev=rev, ub=2)
        assert reason == ref.reason == 2 and its == ref.its and hist.size == its + 1
        np.testing.assert_allclose(hist[:25], ref.history[:25], rtol=1e-10)
        np.testing.assert_allclose(hist, ref.history, rtol=1e-5)   # round-off is amplified towards convergence
        np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-7 * np.abs(ref.x).max())
        out[rev] = x
    # zero right-hand side: converged at iteration 0 with x = 0; NaN: KSP_DIVERGED_NANORINF
    x, hist, its, reason, _ = _cg(emu, widths, per, np.zeros_like(b), rtol=1e-5, atol=1e-50)
    assert (its, reason, hist.size) == (0, 3, 1) and np.all(x == 0.0)
    bn = b.copy()
    bn[7] = np.nan
    assert _cg(emu, widths, per, bn, rtol=1e-5, atol=1e-50)[3] == -9
    # without a null space, unpreconditioned norm
    refu = orc.ksp_solve(A, b, pc_type="jacobi", norm_type="unpreconditioned", rtol=0, atol=0, max_it=12, const_nullspace=False)
    _, hist, _, _, _ = _cg(emu, widths, per, b, pc="jacobi", has_const=False, norm=2, max_it=12)
    np.testing.assert_allclose(hist, refu.history, rtol=1e-10)


@pytest.mark.parametrize("tile", [30, 31, 32, 33])
@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_emulated_balanced_split_kernel(emu, tile, pc):
    """k_spmv3 (round-2 candidate): CTA ranges that start mid-tile, span two tiles, or are empty."""
    for shape, per, nctas in (((12, 10, 8), (0, 0, 0), 3), ((70, 25, 7), (0, 1, 1), 5), ((130, 9, 5), (1, 0, 0), 4),
                              ((33, 31), (0, 0), 2)):
        widths = H.make_widths(shape)
        A = H.oracle_matrix(widths, per)
        b, _ = H.consistent_rhs(A)
        nit = 12
        ref = orc.ksp_solve(A, b, pc_type=pc, rtol=0, atol=0, max_it=nit, const_nullspace=True)
        x, hist, its, reason, _ = _cg(emu, widths, per, b, pc=pc, max_it=nit, tile=tile, kz=nctas)
        assert (its, reason) == (nit, -3)
        np.testing.assert_allclose(hist, ref.history, rtol=1e-10)
        np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-9 * np.abs(ref.x).max())


@pytest.mark.parametrize("nranks", [2, 3])
@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_emulated_multi_rank_slabs(emu, nranks, pc):
    """Slab partition, ghost planes of p (recomputed) and r (pushed by the update kernel), periodic-z ring, scalars
    identical on every rank -- the multi-GPU kernel logic with an emulated NCCL-style all-reduce."""
    for shape, per, kz in (((12, 10, 11), (0, 0, 0), 0), ((70, 9, 10), (0, 1, 1), 2), ((9, 8, 7), (1, 0, 1), 0)):
        widths = H.make_widths(shape)
        A = H.oracle_matrix(widths, per)
        b, _ = H.consistent_rhs(A)
        nit = 12
        ref = orc.ksp_solve(A, b, pc_type=pc, rtol=0, atol=0, max_it=nit, const_nullspace=True)
        dim, n, p, w, dz = _grid_args(widths, per)
        x = np.empty_like(b)
        hist = np.zeros(nit + 2)
        nh, its, reason = C.c_int(0), C.c_int(0), C.c_int(0)
        rc = emu.emu_stencil_cg_ranks(nranks, n, p, w[0].ctypes.data_as(_dp), w[1].ctypes.data_as(_dp), dz, 0.01,
                                      int(pc == "jacobi"), 1, 0.0, 0.0, nit, 10, kz, b.ctypes.data_as(_dp),
                                      x.ctypes.data_as(_dp), hist.ctypes.data_as(_dp), hist.size, C.byref(nh), C.byref(its),
                                      C.byref(reason))
        assert rc == 0 and (its.value, reason.value, nh.value) == (nit, -3, nit + 1)
        np.testing.assert_allclose(hist[: nh.value], ref.history, rtol=1e-10)
        np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-9 * np.abs(ref.x).max())


@pytest.mark.parametrize("tile", [10, 18, 30, 32, 33])
def test_results_do_not_depend_on_the_thread_schedule(emu, tile):
    """Race check: shuffled fiber order and random preemption at every shared-memory access must not change a
    single bit (a missing barrier in the tile ring / stage ring / coefficient table / state staging would)."""
    # tile 30: 6 tiles x 7 planes over 5 CTAs -> ranges that end one tile and start the next (segment boundary)
    shape, per = ((70, 13, 9), (0, 1, 0)) if tile < 30 else ((70, 25, 7), (0, 1, 0))
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b, _ = H.consistent_rhs(A)
    kz = 3 if tile < 30 else 5
    emu.emu_set_schedule(0)
    x0, h0, its0, reason0, _ = _cg(emu, widths, per, b, pc="jacobi", max_it=8, tile=tile, kz=kz)
    try:
        for seed in (1, 2, 3):
            emu.emu_set_schedule(seed)
            x, h, its, reason, _ = _cg(emu, widths, per, b, pc="jacobi", max_it=8, tile=tile, kz=kz)
            assert np.array_equal(h, h0) and np.array_equal(x, x0) and (its, reason) == (its0, reason0)
    finally:
        emu.emu_set_schedule(0)


@pytest.mark.parametrize("nranks,upd_blocks", [(2, 3), (3, 3), (2, 1), (2, 8)])
def test_emulated_2d_grid_on_several_ranks(emu, nranks, upd_blocks, monkeypatch):
    """A 2-D grid on several GPUs is the 3-D code on (nx, 1, ny) with y-slabs (b200ls_set_poisson_stencil); the result
    must match the 2-D oracle operator (bit-identical arithmetic: products with 1.0, additions of exact zeros).
    upd_blocks: grid of the update kernel -- one CTA (the pusher also owns the interior), few, several pushers."""
    monkeypatch.setenv("EMU_UPD_BLOCKS", str(upd_blocks))
    for shape, per in (((33, 14), (0, 0)), ((20, 12), (1, 1))):
        widths = H.make_widths(shape)
        A = H.oracle_matrix(widths, per)                      # the 2-D reference operator
        b, _ = H.consistent_rhs(A)
        nit = 12
        ref = orc.ksp_solve(A, b, pc_type="jacobi", rtol=0, atol=0, max_it=nit, const_nullspace=True)
        n = (C.c_int64 * 3)(shape[0], 1, shape[1])
        p = (C.c_int * 3)(per[0], 0, per[1])
        wx = np.ascontiguousarray(widths[0]); wy = np.ones(1); wz = np.ascontiguousarray(widths[1])
        x = np.empty_like(b)
        hist = np.zeros(nit + 2)
        nh, its, reason = C.c_int(0), C.c_int(0), C.c_int(0)
        rc = emu.emu_stencil_cg_ranks(nranks, n, p, wx.ctypes.data_as(_dp), wy.ctypes.data_as(_dp), wz.ctypes.data_as(_dp),
                                      0.01, 1, 1, 0.0, 0.0, nit, 10, 0, b.ctypes.data_as(_dp), x.ctypes.data_as(_dp),
                                      hist.ctypes.data_as(_dp), hist.size, C.byref(nh), C.byref(its), C.byref(reason))
        assert rc == 0 and (its.value, reason.value) == (nit, -3)
        np.testing.assert_allclose(hist[: nh.value], ref.history, rtol=1e-10)
        np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-9 * np.abs(ref.x).max())


def _csr_solve(L, M, b, bcgs=False, pc="none", has_const=False, nullvec=None, rtol=0.0, atol=0.0, max_it=20):
    M = M.tocsr(); M.sort_indices()
    rp = np.ascontiguousarray(M.indptr, dtype=np.int64); col = np.ascontiguousarray(M.indices, dtype=np.int32)
    val = np.ascontiguousarray(M.data, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty_like(b); hist = np.zeros(max_it + 2)
    nh, its, reason = C.c_int(0), C.c_int(0), C.c_int(0)
    nvp = None if nullvec is None else np.ascontiguousarray(nullvec, dtype=np.float64).ctypes.data_as(_dp)
    rc = L.emu_csr_solve(M.shape[0], rp.ctypes.data_as(C.POINTER(C.c_int64)), col.ctypes.data_as(C.POINTER(C.c_int32)),
                         val.ctypes.data_as(_dp), int(bcgs), int(pc == "jacobi"), int(has_const), nvp, rtol, atol, max_it,
                         b.ctypes.data_as(_dp), x.ctypes.data_as(_dp), hist.ctypes.data_as(_dp), hist.size, C.byref(nh),
                         C.byref(its), C.byref(reason))
    assert rc == 0
    return x, hist[: nh.value].copy(), its.value, reason.value


@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_emulated_csr_paths(emu, pc):
    """The assembled-operator kernels (csr_kernels.cuh) on the emulation: BiCGStab on a velocity system, CG with
    the constant null space and with one explicit null-space vector."""
    import scipy.sparse as sp

    # velocity system, BiCGStab
    A, _ = H.velocity_system(H.make_widths((8, 7, 6)), (0, 0, 0), dt=0.5, nu=1.0)
    Ao = orc.Csr.from_arrays(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)
    b = np.random.default_rng(2).standard_normal(A.shape[0])
    ref = orc.ksp_solve(Ao, b, ksp_type="bcgs", pc_type=pc, rtol=0.0, atol=1e-8, max_it=300)
    x, hist, its, reason = _csr_solve(emu, A, b, bcgs=True, pc=pc, atol=1e-8, max_it=300)
    assert reason == ref.reason == 3 and abs(its - ref.its) <= 2
    np.testing.assert_allclose(hist[:8], ref.history[:8], rtol=1e-9)
    np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-6 * np.abs(ref.x).max())
    # pressure operator as CSR with the constant null space
    P = H.oracle_matrix(H.make_widths((9, 8, 5)), (0, 1, 0))
    bp, _ = H.consistent_rhs(P)
    refp = orc.ksp_solve(P, bp, pc_type=pc, rtol=0, atol=0, max_it=15, const_nullspace=True)
    x, hist, its, reason = _csr_solve(emu, P.to_scipy(), bp, pc=pc, has_const=True, max_it=15)
    assert (its, reason) == (15, -3)
    np.testing.assert_allclose(hist, refp.history, rtol=1e-10)
    # IBPM-style modified Poisson with an explicit null-space vector
    G = orc.assemble_gradient(H.make_widths((10, 9)), [0, 0, 0]).to_scipy()
    R = sp.random(G.shape[0], 7, density=0.05, random_state=4, format="csr")
    K = sp.hstack([G, -R]).tocsr()
    M = (-(K.T @ K) * 0.01).tocsr(); M.sort_indices()
    nv = np.zeros(M.shape[0]); nv[: G.shape[1]] = 1.0 / np.sqrt(G.shape[1])
    Mo = orc.Csr.from_arrays(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)
    xs = np.random.default_rng(5).standard_normal(M.shape[0]); xs -= (xs @ nv) * nv
    bm = M @ xs
    refm = orc.ksp_solve(Mo, bm, pc_type=pc, rtol=0, atol=0, max_it=15, nullvecs=nv)
    x, hist, its, reason = _csr_solve(emu, M, bm, pc=pc, nullvec=nv, max_it=15)
    np.testing.assert_allclose(hist, refm.history, rtol=1e-10)
    np.testing.assert_allclose(x, refm.x, rtol=0, atol=1e-9 * np.abs(refm.x).max())


def _sep_solve(L, dims, per, M, b, mode="apply", pc="none", has_const=False, nullvec=None, rtol=0.0, atol=0.0, max_it=20,
               hybrid_widths=None, dt=0.01):
    """The line-coefficient operator (sep_kernels.cuh) on the emulation; the structure is read out of M by the host
    code of libb200ls.so (b200ls_staggered_analyze), exactly as b200ls_set_staggered does before uploading it.
    hybrid_widths: the pressure operator of that grid + remainder instead (b200ls_hybrid_analyze / _set_poisson_hybrid)."""
    from petibm_b200.staggered import analyze, analyze_hybrid

    M = M.tocsr(); M.sort_indices()
    wp = None
    if hybrid_widths is not None:
        st = analyze_hybrid(hybrid_widths, per, dt, M.indptr, M.indices, M.data)
        w3 = [np.asarray(a, dtype=np.float64) for a in hybrid_widths] + [np.ones(1)] * (3 - len(hybrid_widths))
        dims = [[a.size for a in w3]]
        coef = np.ascontiguousarray(np.concatenate(st["g"]))
        wpack = np.ascontiguousarray(np.concatenate(w3))
        wp = wpack.ctypes.data_as(_dp)
    else:
        st = analyze(dims, per, M.indptr, M.indices, M.data)
        coef = np.ascontiguousarray(np.concatenate([np.concatenate(ax) for f in st["coef"] for ax in f]))
    d = np.ascontiguousarray(dims, dtype=np.int64).reshape(-1)
    p3 = (C.c_int * 3)(*([int(v) for v in per] + [0] * 3)[:3])
    diag = np.ascontiguousarray(st["diag"])
    rp, rc, rv = st["rem"]
    rc = np.ascontiguousarray(rc if rc.size else np.zeros(1, dtype=np.int32)); rv = np.ascontiguousarray(rv if rv.size else np.zeros(1))
    dg = M.diagonal()
    dinv = np.where(dg != 0.0, 1.0 / np.where(dg != 0.0, dg, 1.0), 1.0)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty_like(b); hist = np.zeros(max_it + 2)
    nh, its, reason = C.c_int(0), C.c_int(0), C.c_int(0)
    nvp = None if nullvec is None else np.ascontiguousarray(nullvec, dtype=np.float64).ctypes.data_as(_dp)
    rcode = L.emu_sep_solve(len(dims), d.ctypes.data_as(C.POINTER(C.c_int64)), p3, wp, M.shape[0], coef.ctypes.data_as(_dp),
                            diag.ctypes.data_as(_dp), rp.ctypes.data_as(C.POINTER(C.c_int64)),
                            rc.ctypes.data_as(C.POINTER(C.c_int32)), rv.ctypes.data_as(_dp), dinv.ctypes.data_as(_dp),
                            {"apply": 0, "cg": 1, "bcgs": 2}[mode], int(pc == "jacobi"), int(has_const), nvp, rtol, atol, max_it,
                            b.ctypes.data_as(_dp), x.ctypes.data_as(_dp), hist.ctypes.data_as(_dp), hist.size, C.byref(nh),
                            C.byref(its), C.byref(reason))
    assert rcode == 0
    return x, hist[: nh.value].copy(), its.value, reason.value


def _velocity_dims(shape, per):
    dim = len(shape)
    n = list(shape) + [1] * (3 - dim)
    p = list(per) + [0] * (3 - dim)
    return [[n[d] - (1 if (d == f and not p[d]) else 0) for d in range(3)] for f in range(dim)], p


@pytest.mark.parametrize("shape,per", [((9, 8), (0, 0)), ((8, 7, 6), (0, 0, 0)), ((7, 6, 5), (1, 0, 1)), ((5, 6, 7), (1, 1, 1)),
                                       ((40, 3, 3), (0, 1, 0))])
def test_emulated_line_coefficient_spmv_is_bit_identical_to_the_assembled_matmult(emu, shape, per):
    """Velocity system A = I/dt - c nu L (createlaplacian.cpp:134-159, navierstokes.cpp:342-344) in line-coefficient
    form: y = A x equals the oracle's MatMult_SeqAIJ restatement on the assembled matrix bit for bit."""
    A, _ = H.velocity_system(H.make_widths(shape), per, dt=0.01, nu=0.02)
    Ao = orc.Csr.from_arrays(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)
    dims, p = _velocity_dims(shape, per)
    rng = np.random.default_rng(3)
    for _ in range(2):
        x = rng.standard_normal(A.shape[0])
        y, _, _, _ = _sep_solve(emu, dims, p, A, x, mode="apply")
        assert np.array_equal(y, Ao.spmv(x))


@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_emulated_line_coefficient_krylov_paths(emu, pc):
    """Same systems as test_emulated_csr_paths, through sep_kernels.cuh: the residual histories must equal those of the
    CSR kernels EXACTLY (same row sums bit for bit, same reductions), and match the oracle like they do."""
    import scipy.sparse as sp

    # velocity system, BiCGStab (+ Jacobi): what vSolver->solve runs in the shipped configs
    shape, per = (8, 7, 6), (0, 0, 0)
    A, _ = H.velocity_system(H.make_widths(shape), per, dt=0.5, nu=1.0)
    Ao = orc.Csr.from_arrays(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)
    dims, p = _velocity_dims(shape, per)
    b = np.random.default_rng(2).standard_normal(A.shape[0])
    ref = orc.ksp_solve(Ao, b, ksp_type="bcgs", pc_type=pc, rtol=0.0, atol=1e-8, max_it=300)
    xc, hc, ic, rc = _csr_solve(emu, A, b, bcgs=True, pc=pc, atol=1e-8, max_it=300)
    x, hist, its, reason = _sep_solve(emu, dims, p, A, b, mode="bcgs", pc=pc, atol=1e-8, max_it=300)
    assert (its, reason) == (ic, rc) and np.array_equal(hist, hc) and np.array_equal(x, xc)
    assert reason == ref.reason == 3 and abs(its - ref.its) <= 2
    np.testing.assert_allclose(hist[:8], ref.history[:8], rtol=1e-9)
    # periodic velocity system, CG is not what PetIBM uses there but the kernel exists: compare with the CSR kernels
    shape, per = (7, 6, 5), (1, 0, 1)
    A, _ = H.velocity_system(np.array([np.full(n, 1.0 / n) for n in shape], dtype=object).tolist(), per, dt=0.5, nu=1.0)
    dims, p = _velocity_dims(shape, per)
    b = np.random.default_rng(6).standard_normal(A.shape[0])
    xc, hc, ic, rc = _csr_solve(emu, A, b, pc=pc, max_it=12)
    x, hist, its, reason = _sep_solve(emu, dims, p, A, b, mode="cg", pc=pc, max_it=12)
    assert (its, reason) == (ic, rc) and np.array_equal(hist, hc) and np.array_equal(x, xc)
    # IBPM-style modified Poisson: stencil block + remainder, CG with the explicit null-space vector (ibpm.cpp:251-267)
    gshape = (10, 9)
    G = orc.assemble_gradient(H.make_widths(gshape), [0, 0, 0]).to_scipy()
    R = sp.random(G.shape[0], 7, density=0.05, random_state=4, format="csr")
    K = sp.hstack([G, -R]).tocsr()
    M = (-(K.T @ K) * 0.01).tocsr(); M.sort_indices()
    nv = np.zeros(M.shape[0]); nv[: G.shape[1]] = 1.0 / np.sqrt(G.shape[1])
    Mo = orc.Csr.from_arrays(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)
    xs = np.random.default_rng(5).standard_normal(M.shape[0]); xs -= (xs @ nv) * nv
    bm = M @ xs
    refm = orc.ksp_solve(Mo, bm, pc_type=pc, rtol=0, atol=0, max_it=15, nullvecs=nv)
    y, _, _, _ = _sep_solve(emu, [[10, 9, 1]], (0, 0, 0), M, xs, mode="apply")
    assert np.array_equal(y, Mo.spmv(xs))
    xc, hc, ic, rc = _csr_solve(emu, M, bm, pc=pc, nullvec=nv, max_it=15)
    x, hist, its, reason = _sep_solve(emu, [[10, 9, 1]], (0, 0, 0), M, bm, mode="cg", pc=pc, nullvec=nv, max_it=15)
    assert (its, reason) == (ic, rc) and np.array_equal(hist, hc) and np.array_equal(x, xc)
    np.testing.assert_allclose(hist, refm.history, rtol=1e-10)
    np.testing.assert_allclose(x, refm.x, rtol=0, atol=1e-9 * np.abs(refm.x).max())


@pytest.mark.parametrize("dim", [2, 3])
def test_emulated_hybrid_operator_on_a_stretched_ibpm_system(emu, dim):
    """IBPM's modified Poisson system [D;E] BN [G,-H] on a stretched grid (tests/helpers.ibpm_system): the pressure block
    is rebuilt from the 1-D arrays with its face areas, the coupling comes from the remainder -- SpMV bit-identical to
    the oracle's MatMult on the assembled matrix, CG with the explicit null-space vector equal to the CSR kernels bit for
    bit (and to the oracle like they are)."""
    sub = [{"end": 0.6, "cells": 5 if dim == 3 else 8, "stretchRatio": 1.0 / 1.25}, {"end": 1.4, "cells": 8 if dim == 3 else 14, "stretchRatio": 1.0},
           {"end": 2.0, "cells": 5 if dim == 3 else 8, "stretchRatio": 1.25}]
    w = orc.axis_from_subdomains(0.0, sub)
    widths = [w.copy() for _ in range(dim)]
    M, pN, nv = H.ibpm_system(widths, dt=0.01, nb=10)
    Mo = orc.Csr.from_arrays(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)
    rng = np.random.default_rng(12)
    xs = rng.standard_normal(M.shape[0]); xs -= (xs @ nv) * nv
    y, _, _, _ = _sep_solve(emu, None, (0,) * dim, M, xs, mode="apply", hybrid_widths=widths)
    assert np.array_equal(y, Mo.spmv(xs))
    bm = Mo.spmv(xs)
    for pc in ("none", "jacobi"):
        ref = orc.ksp_solve(Mo, bm, pc_type=pc, rtol=0, atol=0, max_it=15, nullvecs=nv)
        xc, hc, ic, rc = _csr_solve(emu, M, bm, pc=pc, nullvec=nv, max_it=15)
        x, hist, its, reason = _sep_solve(emu, None, (0,) * dim, M, bm, mode="cg", pc=pc, nullvec=nv, max_it=15, hybrid_widths=widths)
        assert (its, reason) == (ic, rc) and np.array_equal(hist, hc) and np.array_equal(x, xc)
        # against the oracle: 1e-10 while the recurrence is insensitive to the summation order (the assembled matrix is
        # symmetric only up to rounding, and CG amplifies that from iteration to iteration), 1e-6 over the whole window
        np.testing.assert_allclose(hist[:8], ref.history[:8], rtol=1e-10)
        np.testing.assert_allclose(hist, ref.history, rtol=1e-6)
        np.testing.assert_allclose(x, ref.x, rtol=0, atol=1e-6 * np.abs(ref.x).max())


def _random_stencil_blocks(rng):
    import scipy.sparse as sp

    nf = int(rng.integers(1, 4))
    per = [bool(rng.integers(0, 2)) for _ in range(3)]
    dims = [[int(rng.integers(3 if per[d] else 1, 7)) for d in range(3)] for _ in range(nf)]
    nextra = int(rng.integers(0, 5))
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
            for coef, step in ((cm, -1), (cp, +1)):
                nb = idx[d] + step
                ok = (nb >= 0) & (nb < n[d])
                if per[d]:
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
    return M, dims, per


def test_emulated_line_coefficient_spmv_on_random_stencil_blocks(emu):
    """Random combinations the structured tests do not reach: one to three blocks, extents down to one cell, periodic
    axes with three cells (every neighbour wraps somewhere), a random remainder behind the blocks.  y = A x must equal the
    oracle's MatMult_SeqAIJ restatement bit for bit (same row sums in the same order)."""
    import scipy.sparse as sp

    rng = np.random.default_rng(77)
    for case in range(14):
        M, dims, per = _random_stencil_blocks(rng)
        nrows = M.shape[0]
        Mo = orc.Csr.from_arrays(nrows, nrows, M.indptr, M.indices, M.data)
        x = rng.standard_normal(nrows)
        y, _, _, _ = _sep_solve(emu, dims, per, M, x, mode="apply")
        assert np.array_equal(y, Mo.spmv(x)), (case, dims, per)


class _sep_tiles:
    """Run the line-coefficient operator through the tiled kernels of sep_tile.cuh (xr = 2 / 4: tiles of 64 / 128 cells
    in x; zchunk planes per chunk, 0 = the launch rule; target = the CTA count that rule aims for; stages = planes in the
    per-thread cp.async queue, 3 or 4)."""

    def __init__(self, L, xr, zchunk=0, target=24, stages=3):
        self.L, self.args, self.stages = L, (xr, zchunk, target), stages

    def __enter__(self):
        self.L.emu_set_sep_tile(*self.args)
        self.L.emu_set_sep_stages(self.stages)

    def __exit__(self, *exc):
        self.L.emu_set_sep_tile(0, 0, 0)
        self.L.emu_set_sep_stages(3)


TILE_VARIANTS = [(2, 0), (2, 1), (2, 3), (4, 0), (4, 2)]


@pytest.mark.parametrize("xr,zchunk", TILE_VARIANTS)
@pytest.mark.parametrize("shape,per", [((9, 8), (0, 0)), ((8, 7, 6), (0, 0, 0)), ((7, 6, 5), (1, 0, 1)), ((5, 6, 7), (1, 1, 1)),
                                       ((40, 3, 3), (0, 1, 0)), ((70, 19, 5), (0, 0, 0)), ((131, 10, 4), (1, 0, 0)),
                                       ((66, 17), (0, 1)), ((3, 3, 9), (0, 0, 1))])
def test_emulated_tiled_line_coefficient_spmv_is_bit_identical(emu, shape, per, xr, zchunk):
    """sep_tile.cuh: the plane-marching tile kernels give y = A x bit for bit like MatMult_SeqAIJ on the assembled velocity
    matrix -- several tiles in x and y, ragged edges, one or many z chunks, periodic axes (wrapped cells take sep_row)."""
    A, _ = H.velocity_system(H.make_widths(shape), per, dt=0.01, nu=0.02)
    Ao = orc.Csr.from_arrays(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)
    dims, p = _velocity_dims(shape, per)
    x = np.random.default_rng(3).standard_normal(A.shape[0])
    for stages in (3, 4):
        with _sep_tiles(emu, xr, zchunk, stages=stages):
            y, _, _, _ = _sep_solve(emu, dims, p, A, x, mode="apply")
        assert np.array_equal(y, Ao.spmv(x)), stages


def test_tiled_kernel_launch_geometry(emu):
    """sep_tile_plan (host code shared by sep_solver.inc and the emulation): the chunks of a field cover its planes exactly
    once, the automatic rule never launches more CTAs than it was given slots for unless one chunk per tile already does,
    keeps at least 4 planes per chunk, and sizes the surface / tail parts from the periodic faces and the remainder rows."""
    emu.emu_sep_tile_plan.argtypes = [C.c_int, C.POINTER(C.c_int64), _ip, C.c_int64, C.c_int, C.c_int, C.c_int, _ip]

    def plan(dims, per, extra, xr, zchunk, target):
        d = np.ascontiguousarray(dims, dtype=np.int64).reshape(-1)
        p = (C.c_int * 3)(*per)
        out = (C.c_int * (4 + 6 * len(dims)))()
        nsep = int(sum(a * b * c for a, b, c in dims))
        emu.emu_sep_tile_plan(len(dims), d.ctypes.data_as(C.POINTER(C.c_int64)), p, nsep + extra, xr, zchunk, target, out)
        o = list(out)
        return o[:4], [o[4 + 6 * f: 10 + 6 * f] for f in range(len(dims))]

    # the 128^3 velocity system on 148 SMs x 3 resident CTAs: 96 tiles per plane-set, 4 chunks of 32 planes = 384 CTAs
    (sb, surf_b, tail_b, surf_c), fields = plan([[127, 128, 128], [128, 127, 128], [128, 128, 127]], (0, 0, 0), 0, 2, 0, 444)
    assert (sb, surf_b, tail_b, surf_c) == (384, 0, 0, 0)
    assert [f[:4] for f in fields] == [[2, 32, 4, 32]] * 3 and [f[4] for f in fields] == [0, 128, 256]
    rng = np.random.default_rng(5)
    for _ in range(200):
        nf = int(rng.integers(1, 4))
        per = [int(rng.integers(0, 2)) for _ in range(3)]
        dims = [[int(rng.integers(3, 300)), int(rng.integers(3, 70)), int(rng.integers(1, 200))] for _ in range(nf)]
        xr = int(rng.choice([2, 4])); zchunk = int(rng.choice([0, 0, 1, 5, 64])); target = int(rng.choice([24, 444, 1184]))
        extra = int(rng.integers(0, 3)) * 100
        (sb, surf_b, tail_b, surf_c), fields = plan(dims, per, extra, xr, zchunk, target)
        b0 = 0; surf = 0; tiles_all = 0
        for (n0, n1, n2), (tiles_x, tiles, nchunk, zc, block0, surf0) in zip(dims, fields):
            assert tiles_x == -(-n0 // (32 * xr)) and tiles == tiles_x * -(-n1 // 8)
            assert (nchunk - 1) * zc < n2 <= nchunk * zc                       # every plane in exactly one chunk, no empty chunk
            assert block0 == b0 and surf0 == surf
            if zchunk:
                assert zc == min(zchunk, n2) or nchunk == -(-n2 // zchunk)
            else:
                assert zc >= min(4, n2)
            b0 += tiles * nchunk; tiles_all += tiles
            surf += (2 * n1 * n2 if per[0] else 0) + (2 * n0 * n2 if per[1] else 0) + (2 * n0 * n1 if (per[2] and n2 >= 3) else 0)
        assert sb == b0 and surf_c == surf
        if not zchunk:
            assert sb <= max(target, tiles_all)
        assert surf_b == min(-(-surf // 256), 128) and tail_b == min(-(-extra // 256), 64)


@pytest.mark.parametrize("xr,zchunk", [(2, 2), (4, 0)])
def test_emulated_tiled_kernels_do_not_depend_on_the_thread_schedule(emu, xr, zchunk):
    """One barrier per plane guards the double-buffered shared plane: with shuffled fiber order and random preemption at
    every shared-memory access the product stays bit-identical (a missing barrier would show as a schedule-dependent y)."""
    shape, per = (70, 19, 5), (0, 0, 0)
    A, _ = H.velocity_system(H.make_widths(shape), per, dt=0.01, nu=0.02)
    Ao = orc.Csr.from_arrays(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)
    dims, p = _velocity_dims(shape, per)
    x = np.random.default_rng(9).standard_normal(A.shape[0])
    y0 = Ao.spmv(x)
    try:
        for seed in (1, 2, 3):
            emu.emu_set_schedule(seed)
            with _sep_tiles(emu, xr, zchunk):
                y, _, _, _ = _sep_solve(emu, dims, p, A, x, mode="apply")
            assert np.array_equal(y, y0)
    finally:
        emu.emu_set_schedule(0)


@pytest.mark.parametrize("xr,zchunk", [(2, 0), (4, 3)])
def test_emulated_tiled_line_coefficient_spmv_on_random_stencil_blocks(emu, xr, zchunk):
    """The random blocks of the test above (extents down to one cell, three-cell periodic axes, random remainder)
    through the tiled kernels."""
    rng = np.random.default_rng(78)
    for case in range(10):
        M, dims, per = _random_stencil_blocks(rng)
        Mo = orc.Csr.from_arrays(M.shape[0], M.shape[0], M.indptr, M.indices, M.data)
        x = rng.standard_normal(M.shape[0])
        with _sep_tiles(emu, xr, zchunk):
            y, _, _, _ = _sep_solve(emu, dims, per, M, x, mode="apply")
        assert np.array_equal(y, Mo.spmv(x)), (case, dims, per)


@pytest.mark.parametrize("xr,zchunk", [(2, 0), (2, 2), (4, 0)])
@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_emulated_tiled_line_coefficient_krylov_paths(emu, pc, xr, zchunk):
    """BiCGStab (+ Jacobi) on the velocity system, CG on a periodic one and CG with the explicit null-space vector on an
    IBPM-style system with a remainder, through the tiled kernels: same iteration counts and reasons as the row-per-thread
    kernels, histories equal up to the summation order of the dot products (1e-12), oracle agreement as before."""
    import scipy.sparse as sp

    shape, per = (8, 7, 6), (0, 0, 0)
    A, _ = H.velocity_system(H.make_widths(shape), per, dt=0.5, nu=1.0)
    Ao = orc.Csr.from_arrays(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)
    dims, p = _velocity_dims(shape, per)
    b = np.random.default_rng(2).standard_normal(A.shape[0])
    ref = orc.ksp_solve(Ao, b, ksp_type="bcgs", pc_type=pc, rtol=0.0, atol=1e-8, max_it=300)
    x0, h0, i0, r0 = _sep_solve(emu, dims, p, A, b, mode="bcgs", pc=pc, atol=1e-8, max_it=300)
    stages = 4 if zchunk else 3
    with _sep_tiles(emu, xr, zchunk, stages=stages):
        x, hist, its, reason = _sep_solve(emu, dims, p, A, b, mode="bcgs", pc=pc, atol=1e-8, max_it=300)
    assert abs(its - i0) <= 1 and reason == r0 == ref.reason == 3
    # BiCGStab amplifies the rounding of its dot products from iteration to iteration: tight over the first entries,
    # loose over the 35-iteration window (the oracle comparison of the row-per-thread kernels has the same shape)
    np.testing.assert_allclose(hist[:10], h0[:10], rtol=1e-9)
    m = min(hist.size, h0.size)
    np.testing.assert_allclose(hist[:m], h0[:m], rtol=5e-2)
    np.testing.assert_allclose(hist[:8], ref.history[:8], rtol=1e-9)
    np.testing.assert_allclose(x, x0, rtol=0, atol=1e-6 * np.abs(x0).max())
    shape, per = (7, 6, 5), (1, 0, 1)
    A, _ = H.velocity_system(np.array([np.full(n, 1.0 / n) for n in shape], dtype=object).tolist(), per, dt=0.5, nu=1.0)
    dims, p = _velocity_dims(shape, per)
    b = np.random.default_rng(6).standard_normal(A.shape[0])
    x0, h0, i0, r0 = _sep_solve(emu, dims, p, A, b, mode="cg", pc=pc, max_it=12)
    with _sep_tiles(emu, xr, zchunk, stages=stages):
        x, hist, its, reason = _sep_solve(emu, dims, p, A, b, mode="cg", pc=pc, max_it=12)
    assert (its, reason) == (i0, r0)
    np.testing.assert_allclose(hist, h0, rtol=1e-12)
    np.testing.assert_allclose(x, x0, rtol=0, atol=1e-12 * np.abs(x0).max())
    gshape = (10, 9)
    G = orc.assemble_gradient(H.make_widths(gshape), [0, 0, 0]).to_scipy()
    R = sp.random(G.shape[0], 7, density=0.05, random_state=4, format="csr")
    K = sp.hstack([G, -R]).tocsr()
    M = (-(K.T @ K) * 0.01).tocsr(); M.sort_indices()
    nv = np.zeros(M.shape[0]); nv[: G.shape[1]] = 1.0 / np.sqrt(G.shape[1])
    Mo = orc.Csr.from_arrays(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)
    xs = np.random.default_rng(5).standard_normal(M.shape[0]); xs -= (xs @ nv) * nv
    bm = M @ xs
    refm = orc.ksp_solve(Mo, bm, pc_type=pc, rtol=0, atol=0, max_it=15, nullvecs=nv)
    with _sep_tiles(emu, xr, zchunk, stages=stages):
        y, _, _, _ = _sep_solve(emu, [[10, 9, 1]], (0, 0, 0), M, xs, mode="apply")
        x, hist, its, reason = _sep_solve(emu, [[10, 9, 1]], (0, 0, 0), M, bm, mode="cg", pc=pc, nullvec=nv, max_it=15)
    assert np.array_equal(y, Mo.spmv(xs))
    np.testing.assert_allclose(hist, refm.history, rtol=1e-10)
    np.testing.assert_allclose(x, refm.x, rtol=0, atol=1e-9 * np.abs(refm.x).max())


@pytest.mark.parametrize("xr,zchunk", [(2, 0), (4, 2)])
@pytest.mark.parametrize("dim", [2, 3])
def test_emulated_tiled_hybrid_operator_on_a_stretched_ibpm_system(emu, dim, xr, zchunk):
    """The hybrid form (pressure block with face areas + remainder, b200ls_set_poisson_hybrid) through the tiled kernels:
    SpMV bit-identical to the assembled MatMult, CG with the explicit null-space vector as the row-per-thread kernels."""
    sub = [{"end": 0.6, "cells": 5 if dim == 3 else 8, "stretchRatio": 1.0 / 1.25}, {"end": 1.4, "cells": 8 if dim == 3 else 14, "stretchRatio": 1.0},
           {"end": 2.0, "cells": 5 if dim == 3 else 8, "stretchRatio": 1.25}]
    w = orc.axis_from_subdomains(0.0, sub)
    widths = [w.copy() for _ in range(dim)]
    M, pN, nv = H.ibpm_system(widths, dt=0.01, nb=10)
    Mo = orc.Csr.from_arrays(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)
    rng = np.random.default_rng(12)
    xs = rng.standard_normal(M.shape[0]); xs -= (xs @ nv) * nv
    bm = Mo.spmv(xs)
    with _sep_tiles(emu, xr, zchunk):
        y, _, _, _ = _sep_solve(emu, None, (0,) * dim, M, xs, mode="apply", hybrid_widths=widths)
    assert np.array_equal(y, bm)
    for pc in ("none", "jacobi"):
        x0, h0, i0, r0 = _sep_solve(emu, None, (0,) * dim, M, bm, mode="cg", pc=pc, nullvec=nv, max_it=15, hybrid_widths=widths)
        with _sep_tiles(emu, xr, zchunk):
            x, hist, its, reason = _sep_solve(emu, None, (0,) * dim, M, bm, mode="cg", pc=pc, nullvec=nv, max_it=15, hybrid_widths=widths)
        assert (its, reason) == (i0, r0)
        np.testing.assert_allclose(hist[:8], h0[:8], rtol=1e-10)
        np.testing.assert_allclose(hist, h0, rtol=1e-6)


@pytest.mark.parametrize("shape,per", [((9, 7), (0, 0)), ((8, 6), (1, 0)), ((7, 6, 5), (0, 0, 0)), ((6, 5, 7), (1, 0, 1)),
                                       ((5, 5, 5), (1, 1, 1)), ((12, 3, 4), (0, 1, 0))])
def test_emulated_divergence_gradient_projection_are_the_assembled_products(emu, shape, per):
    """The operators on either side of the pressure solve (ops_kernels.cuh; navierstokes.cpp:540-551, 583-615, 442):
    rhs2 = D u, G p, (BN G) dp and the projection u -= BNG dp, p += dp are bit-identical to MatMult on the oracle's
    assembled D, G and MatMatMult(BN, G), and chaining them gives D (BN G) = the Poisson operator of the solve."""
    emu.emu_stag_ops.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int64), _ip, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp]
    widths = H.make_widths(shape)
    dim, n, p, w, dz = _grid_args(widths, per)
    per3 = list(per) + [0] * (3 - dim)
    D = orc.assemble_divergence(widths, per3)
    G = orc.assemble_gradient(widths, per3)
    BNG = orc.matmatmult(orc.bnhead_order1(G.shape[0], 0.01), G)
    UN, pN = G.shape
    rng = np.random.default_rng(11)

    def run(mode, vin, io, io2=None):
        vin = np.ascontiguousarray(vin); io = np.ascontiguousarray(io)
        emu.emu_stag_ops(mode, dim, n, p, w[0].ctypes.data_as(_dp), w[1].ctypes.data_as(_dp), dz, 0.01, vin.ctypes.data_as(_dp),
                         io.ctypes.data_as(_dp), None if io2 is None else io2.ctypes.data_as(_dp))
        return io

    u = rng.standard_normal(UN)
    assert np.array_equal(run(0, u, np.empty(pN)), D.spmv(u))
    pr = rng.standard_normal(pN)
    assert np.array_equal(run(1, pr, np.empty(UN)), G.spmv(pr))
    assert np.array_equal(run(2, pr, np.empty(UN)), BNG.spmv(pr))
    # projection: u <- u + (-1.0) * (BNG dp), p <- p + 1.0 * dp   (VecAXPY semantics)
    dp = rng.standard_normal(pN)
    u2, p2 = u.copy(), pr.copy()
    run(3, dp, u2, p2)
    assert np.array_equal(u2, u + (-1.0) * BNG.spmv(dp)) and np.array_equal(p2, pr + 1.0 * dp)
    # D (BN G) p is the operator the solver applies (its own bit-exact SpMV is checked elsewhere): same to rounding
    A = H.oracle_matrix(widths, per)
    lhs = run(0, run(2, pr, np.empty(UN)), np.empty(pN))
    np.testing.assert_allclose(lhs, A.spmv(pr), rtol=0, atol=1e-12 * np.abs(lhs).max())


@pytest.mark.parametrize("shape,per", [((9, 7), (0, 0)), ((8, 6), (1, 1)), ((7, 6, 5), (0, 0, 0)), ((6, 5, 7), (1, 0, 1)), ((70, 4, 3), (0, 1, 0))])
def test_emulated_convection_is_the_reference_stencil(emu, shape, per):
    """k_convection (ops_kernels.cuh) against the oracle's restatement of createconvection.cpp:39-332 on random ghosted
    fields: bit-identical; k_ghosted_from_packed reproduces the interior and the periodic wrap layers."""
    emu.emu_stag_ops.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int64), _ip, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp]
    widths = H.make_widths(shape)
    dim, n, p, w, dz = _grid_args(widths, per)
    q, packed, nf = H.ghosted_fields(shape, per)
    ref = np.concatenate([o.ravel() for o in orc.convection(widths, per, q)])
    qin = np.ascontiguousarray(np.concatenate([a.ravel() for a in q]))
    out = np.empty(ref.size)
    emu.emu_stag_ops(4, dim, n, p, w[0].ctypes.data_as(_dp), w[1].ctypes.data_as(_dp), dz, 0.01, qin.ctypes.data_as(_dp),
                     out.ctypes.data_as(_dp), None)
    assert np.array_equal(out, ref)
    # interior + periodic wrap layers from the packed vector; the other ghost layers keep what the caller put there
    g = np.full(qin.size, -7.0)
    emu.emu_stag_ops(5, dim, n, p, w[0].ctypes.data_as(_dp), w[1].ctypes.data_as(_dp), dz, 0.01,
                     np.ascontiguousarray(packed).ctypes.data_as(_dp), g.ctypes.data_as(_dp), None)
    off = 0
    for f in range(dim):
        a = g[off: off + q[f].size].reshape(q[f].shape)
        off += q[f].size
        filled = a != -7.0
        assert np.array_equal(a[filled], q[f][filled])
        inner = tuple([slice(1, -1)] * dim)
        assert filled[inner].all()
        for d in range(dim):                      # a face ghost layer is filled exactly when its axis is periodic
            ax = dim - 1 - d
            sl = [slice(1, -1)] * dim
            sl[ax] = 0
            assert filled[tuple(sl)].all() == bool(per[d]) and filled[tuple(sl)].any() == bool(per[d])


@pytest.mark.parametrize("shape,per", [((24, 10, 9), (0, 0, 0)), ((67, 9, 6), (1, 1, 0)), ((16, 12, 10), (0, 0, 1)), ((33, 21), (1, 0))])
def test_emulated_jacobi_diagonal_on_the_fly_is_bit_identical(emu, shape, per):
    """k_update2f (tuning upd_variant = 2, a round-2 candidate): 1/diag rebuilt from the 1-D arrays inside the update
    kernel instead of streamed from HBM.  Same accumulation order as k_jacobi_setup, so history and solution are equal to
    the stored-reciprocal path bit for bit (padded rows, periodic wrap rows, 2-D included)."""
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b, _ = H.consistent_rhs(A)
    emu.emu_set_update_variant.argtypes = [C.c_int]
    try:
        emu.emu_set_update_variant(0)
        x0, h0, i0, r0, n0 = _cg(emu, widths, per, b, pc="jacobi", max_it=12)
        emu.emu_set_update_variant(1)
        x1, h1, i1, r1, n1 = _cg(emu, widths, per, b, pc="jacobi", max_it=12)
    finally:
        emu.emu_set_update_variant(0)
    assert (i0, r0) == (i1, r1) and np.array_equal(h0, h1) and np.array_equal(x0, x1)


def _forces_system(n_side=8, n_band=14, nb=20, dt=0.01):
    """E BN H of the decoupled IBPM (decoupledibpm.cpp:149-216) for a circle of nb Lagrangian points on a stretched 2-D grid:
    the force block of tests/helpers.ibpm_system with the sign PetIBM solves (symmetric positive definite up to rounding)."""
    sub = [{"end": 0.6, "cells": n_side, "stretchRatio": 1.0 / 1.2}, {"end": 1.4, "cells": n_band, "stretchRatio": 1.0},
           {"end": 2.0, "cells": n_side, "stretchRatio": 1.2}]
    w = orc.axis_from_subdomains(0.0, sub)
    M, pN, _ = H.ibpm_system([w, w.copy()], dt=dt, nb=nb)
    F = (-M[pN:, pN:]).tocsr()
    F.sort_indices()
    return F


def _dense_solve(L, F, B, threads=128):
    L.emu_dense_solve.argtypes = [C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32), _dp, C.c_int, _dp, _dp, C.c_int]
    rp = np.ascontiguousarray(F.indptr, dtype=np.int64); col = np.ascontiguousarray(F.indices, dtype=np.int32)
    val = np.ascontiguousarray(F.data, dtype=np.float64)
    B = np.ascontiguousarray(np.atleast_2d(B), dtype=np.float64)
    X = np.empty_like(B)
    rc = L.emu_dense_solve(F.shape[0], rp.ctypes.data_as(C.POINTER(C.c_int64)), col.ctypes.data_as(C.POINTER(C.c_int32)),
                           val.ctypes.data_as(_dp), B.shape[0], B.ctypes.data_as(_dp), X.ctypes.data_as(_dp), threads)
    return rc, X


def test_emulated_direct_solve_of_the_forces_system(emu):
    """-forces_ksp_type preonly -forces_pc_type lu (every shipped decoupled-IBPM case): dense LU by one thread block,
    checked against numpy's LAPACK solve; a singular matrix is reported, not solved."""
    F = _forces_system()
    n = F.shape[0]
    assert abs(F - F.T).max() < 1e-12 * abs(F).max() and np.linalg.eigvalsh(F.toarray()).min() > 0
    rng = np.random.default_rng(3)
    B = rng.standard_normal((3, n))
    for threads in (128, 96):
        rc, X = _dense_solve(emu, F, B, threads)
        assert rc == 0
        ref = np.linalg.solve(F.toarray(), B.T).T
        np.testing.assert_allclose(X, ref, rtol=0, atol=1e-11 * np.abs(ref).max())
        assert np.abs(F @ X.T - B.T).max() <= 1e-12 * np.abs(B).max() * np.linalg.cond(F.toarray())
    import scipy.sparse as sp

    S = sp.csr_matrix(np.array([[1.0, 2.0, 0.0], [2.0, 4.0, 0.0], [0.0, 0.0, 1.0]]))
    rc, _ = _dense_solve(emu, S, np.ones(3), 32)
    assert rc == 2                                  # zero pivot in column 1 (1-based: 2): KSP_DIVERGED_PC_FAILED on the device path


@pytest.mark.parametrize("n", [3, 31, 32, 33, 75, 130, 257])
def test_emulated_direct_solve_pivots(emu, n):
    """Partial pivoting (what PETSc's LU / superlu_dist do): a matrix with a zero diagonal and rows of very different
    scale, sizes around the panel width of 32; unpivoted elimination fails or loses all digits on these."""
    import scipy.sparse as sp

    rng = np.random.default_rng(100 + n)
    A = rng.standard_normal((n, n))
    A[np.arange(n), np.arange(n)] = 0.0             # no usable diagonal pivot to start with
    A[rng.integers(0, n, 2)] *= 1e6
    assert np.linalg.cond(A) < 1e12
    B = rng.standard_normal((2, n))
    # 128 threads and more than two blocks of 32 unknowns: the multi-CTA wavefront substitution (k_dense_sweep);
    # otherwise the one-CTA kernel
    rc, X = _dense_solve(emu, sp.csr_matrix(A), B, 128 if n > 64 else 64)
    assert rc == 0
    ref = np.linalg.solve(A, B.T).T
    np.testing.assert_allclose(X, ref, rtol=0, atol=1e-13 * np.linalg.cond(A) * np.abs(ref).max())
    assert np.abs(A @ X.T - B.T).max() <= 1e-12 * (np.abs(A).max() * np.abs(X).max() * n)

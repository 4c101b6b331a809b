"""Multi-GPU parity check, one process per GPU (run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py

Every rank holds one z-slab (PETSc DMDA ownership rule).  Checks, for each transport combination
(in-kernel peer stores + mailbox all-reduce / cudaMemcpy halos + NCCL / peer stores + NCCL):
 * the distributed matrix-free SpMV is bit-identical to the oracle's MatMult on the assembled matrix,
 * the CG residual history matches the oracle (1e-10) and is bit-identical on every rank,
 * iteration counts / reasons agree with the single-process oracle solve."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import petibm_b200 as pb  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from petibm_b200.dist import Comm  # noqa: E402
from tests import helpers as H  # noqa: E402


def run_case(comm, shape, per, pc, reduce, halo):
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b, xs = H.consistent_rhs(A)
    c = Comm(comm.rank, comm.nranks, comm.device, reduce, halo)
    s = pb.LinSolverB200("poisson", "None", comm=c, device=comm.device)
    nit = 30
    s.setOptions(pc_type=pc, rtol=0.0, atol=0.0, max_it=nit)
    s.setStencil(H.grid_of(widths, per))
    s.setNullSpace(True)
    ok = True
    msgs = []
    # --- SpMV
    xl = c.local_block(xs, shape)
    yl = s.apply(xl)
    yo = c.local_block(A.spmv(xs), shape)
    if not np.array_equal(yl, yo):
        ok = False
        msgs.append(f"spmv max diff {np.abs(yl - yo).max():.3e}")
    # --- CG
    ref = orc.ksp_solve(A, b, pc_type=pc, rtol=0, atol=0, max_it=nit, const_nullspace=True)
    bl = c.local_block(b, shape)
    x = np.empty_like(bl)
    try:
        s.solve(x, bl)
    except pb.B200Error as e:
        if e.code != -5:
            raise
    hist = s.getHistory()
    if s.getIters() != nit or s.getReason() != -3 or hist.size != nit + 1:
        ok = False
        msgs.append(f"its {s.getIters()} reason {s.getReason()} nhist {hist.size}")
    else:
        rel = np.abs(hist - ref.history) / ref.history
        if rel.max() > 1e-10:
            ok = False
            msgs.append(f"history rel diff {rel.max():.3e}")
    xo = c.local_block(ref.x, shape)
    if np.abs(x - xo).max() > 1e-9 * np.abs(ref.x).max():
        ok = False
        msgs.append(f"x diff {np.abs(x - xo).max():.3e}")
    # --- every rank must hold the same scalars bit for bit
    allh = c.allgather_bytes(hist.tobytes())
    if any(h != allh[0] for h in allh):
        ok = False
        msgs.append("histories differ between ranks")
    # --- a converging solve ends on the same iteration everywhere
    s.setOptions(rtol=1e-8, atol=1e-50, max_it=2000)
    s.solve(x, bl)
    ref2 = orc.ksp_solve(A, b, pc_type=pc, rtol=1e-8, atol=1e-50, max_it=2000, const_nullspace=True)
    its = c.allgather_bytes(s.getIters())
    if len(set(its)) != 1 or abs(its[0] - ref2.its) > 1 or s.getReason() != 2:
        ok = False
        msgs.append(f"converging solve: its {its} vs oracle {ref2.its}, reason {s.getReason()}")
    s.destroy()
    allok = all(c.allgather_bytes(ok))
    if comm.rank == 0:
        print(f"[{'PASS' if allok else 'FAIL'}] shape {shape} per {per} pc {pc} reduce {reduce} halo {halo} {'; '.join(msgs)}",
              flush=True)
    return allok


def run_box_case(comm, shape, per, procs):
    """The solver behind a DMDA whose process grid is NOT 1 x 1 x P: matrix rows and vectors arrive as boxes in PETSc
    ordering; setMatrix finds the process grid from the matrix itself and solve() re-partitions into slabs and back."""
    dim = len(shape)
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b, xs = H.consistent_rhs(A)
    Aloc, bl, plan = H.box_local_system(A, b, dim, shape, procs, comm.rank)
    c = Comm(comm.rank, comm.nranks, comm.device, "p2p", "store")
    s = pb.LinSolverB200("poisson", "None", comm=c, device=comm.device)
    nit = 25
    s.setOptions(rtol=0.0, atol=0.0, max_it=nit)
    s.setGrid(H.grid_of(widths, per))
    s.setMatrix(Aloc.setNullSpace(True))
    msgs = []
    ok = s.operator == "stencil" and s._repart is not None and not s._repart.identity
    if not ok:
        msgs.append(f"operator {s.operator}, plan {s._repart and s._repart.procs}")
    else:
        ref = orc.ksp_solve(A, b, rtol=0, atol=0, max_it=nit, const_nullspace=True)
        x = np.empty_like(bl)
        try:
            s.solve(x, bl)
        except pb.B200Error as e:
            if e.code != -5:
                raise
        hist = s.getHistory()
        xo = ref.x[plan.box_rows()]
        if hist.size != nit + 1 or (np.abs(hist - ref.history) / ref.history).max() > 1e-10:
            ok = False
            msgs.append("history differs from the oracle")
        if np.abs(x - xo).max() > 1e-9 * np.abs(ref.x).max():
            ok = False
            msgs.append(f"x diff {np.abs(x - xo).max():.3e}")
    s.destroy()
    allok = all(c.allgather_bytes(ok))
    if comm.rank == 0:
        print(f"[{'PASS' if allok else 'FAIL'}] DMDA boxes: shape {shape} per {per} process grid {procs} {'; '.join(msgs)}",
              flush=True)
    return allok


def run_replicated_case(comm, kind):
    """Systems that are not the pressure stencil, on several ranks (the velocity system A = I/dt - c nu L with BiCGStab +
    Jacobi, navierstokes.cpp:342-344,524-537; IBPM's modified Poisson system with its explicit null vector, ibpm.cpp:164-194,
    251-267): rows are split over the ranks in uneven contiguous blocks with global column indices, as a parallel Mat
    arrives; every rank solves a replica on its GPU.  Checked against the oracle's solve of the whole system."""
    import scipy.sparse as sp

    if kind == "velocity":
        widths = H.make_widths((14, 12, 10))
        A, _ = H.velocity_system(widths, (0, 0, 0), dt=0.01, nu=0.01, c=0.5)
        nv, opts = None, dict(ksp_type="bcgs", pc_type="jacobi", rtol=0.0, atol=1e-9, max_it=500)
        grid = H.grid_of(widths, (0, 0, 0))
    else:
        w = orc.axis_from_subdomains(0.0, [{"end": 0.8, "cells": 10, "stretchRatio": 1.0 / 1.15}, {"end": 1.2, "cells": 16, "stretchRatio": 1.0},
                                           {"end": 2.0, "cells": 10, "stretchRatio": 1.15}])
        widths = [w, w.copy()]
        A, pN, nv = H.ibpm_system(widths, dt=0.01, nb=20)
        opts = dict(ksp_type="cg", pc_type="jacobi", rtol=0.0, atol=0.0, max_it=40)
        grid = H.grid_of(widths, (0, 0))
    n = A.shape[0]
    Ao = orc.Csr.from_arrays(n, n, A.indptr, A.indices, A.data)
    rng = np.random.default_rng(12)
    xs = rng.standard_normal(n)
    if nv is not None:
        xs -= (xs @ nv) * nv
    b = A @ xs
    cuts = np.linspace(0, n, comm.nranks + 1).astype(int)
    cuts[1:-1] += 3 * (np.arange(1, comm.nranks) % 2)
    lo, hi = int(cuts[comm.rank]), int(cuts[comm.rank + 1])
    loc = A[lo:hi].tocsr()
    loc.sort_indices()
    c = Comm(comm.rank, comm.nranks, comm.device, "p2p", "store")
    s = pb.LinSolverB200(kind, "None", comm=c, device=comm.device)
    s.setOptions(**opts)
    s.setGrid(grid)
    M = pb.Mat(loc.indptr.astype(np.int64), loc.indices.astype(np.int32), loc.data, n)
    M.setNullSpace(False, None if nv is None else nv[lo:hi])
    msgs = []
    ok = True
    for rep in range(2):                         # a second setMatrix (moving bodies) goes through the same path
        s.setMatrix(M)
        if s.operator != "csr" or s.nlocal != hi - lo:
            ok = False
            msgs.append(f"operator {s.operator} nlocal {s.nlocal}")
        x = np.empty(hi - lo)
        try:
            s.solve(x, b[lo:hi])
        except pb.B200Error as e:
            if e.code != -5:
                raise
        ref = orc.ksp_solve(Ao, b, nullvecs=nv, **{k: v for k, v in opts.items()})
        hist = s.getHistory()
        if kind == "ibpm":
            if hist.size != ref.history.size or (np.abs(hist - ref.history) / ref.history).max() > 1e-10:
                ok = False
                msgs.append("history differs from the oracle")
        else:
            if s.getReason() != ref.reason or abs(s.getIters() - ref.its) > 2:
                ok = False
                msgs.append(f"its {s.getIters()} vs {ref.its}, reason {s.getReason()} vs {ref.reason}")
        if np.abs(x - ref.x[lo:hi]).max() > 1e-7 * np.abs(ref.x).max():
            ok = False
            msgs.append(f"x diff {np.abs(x - ref.x[lo:hi]).max():.3e}")
        allh = c.allgather_bytes(hist.tobytes())
        if any(h != allh[0] for h in allh):
            ok = False
            msgs.append("histories differ between ranks")
    s.destroy()
    allok = all(c.allgather_bytes(ok))
    if comm.rank == 0:
        print(f"[{'PASS' if allok else 'FAIL'}] replicated solve of the {kind} system on {comm.nranks} ranks ({n} rows) {'; '.join(msgs)}",
              flush=True)
    return allok


def run_mg_case(comm, shape, per, procs=None):
    """-poisson_pc_type mg on several ranks: every rank solves a replica of the whole grid on its GPU (all-gather of b, own
    part of x back), for slab vectors and -- with `procs` -- for DMDA boxes through the box <-> slab exchange."""
    dim = len(shape)
    widths = H.make_widths(shape)
    A = H.oracle_matrix(widths, per)
    b, xs = H.consistent_rhs(A)
    c = Comm(comm.rank, comm.nranks, comm.device, "p2p", "store")
    s = pb.LinSolverB200("poisson", "None", comm=c, device=comm.device)
    s.setOptions(pc_type="mg", rtol=1e-10, atol=1e-50, max_it=200)
    msgs, ok = [], True
    if procs is None:
        s.setStencil(H.grid_of(widths, per))
        s.setNullSpace(True)
        bl = c.local_block(b, shape)
        want = c.local_block(xs, shape)
        yl = s.apply(c.local_block(xs, shape))
        if not np.array_equal(yl, bl):
            ok = False
            msgs.append("replicated apply differs")
    else:
        Aloc, bl, plan = H.box_local_system(A, b, dim, shape, procs, comm.rank)
        s.setGrid(H.grid_of(widths, per))
        s.setMatrix(Aloc.setNullSpace(True))
        want = xs[plan.box_rows()]
        if s.operator != "stencil" or s._mg_rep is None:
            ok = False
            msgs.append(f"operator {s.operator}")
    x = np.empty_like(bl)
    s.solve(x, bl)
    its = c.allgather_bytes((s.getIters(), s.getReason(), s.getHistory().tobytes()))
    if any(t != its[0] for t in its) or s.getReason() != 2 or s.getIters() > 25:
        ok = False
        msgs.append(f"its/reason {[t[:2] for t in its]}")
    # x is determined up to the constant: compare mean-free parts through the global mean of the gathered solution
    xg = np.frombuffer(b"".join(c.allgather_bytes(np.ascontiguousarray(x).tobytes())), dtype=np.float64)
    wg = np.frombuffer(b"".join(c.allgather_bytes(np.ascontiguousarray(want).tobytes())), dtype=np.float64)
    err = np.abs((xg - xg.mean()) - (wg - wg.mean())).max()
    if err > 1e-7 * np.abs(xs).max():
        ok = False
        msgs.append(f"x diff {err:.3e}")
    s.destroy()
    allok = all(c.allgather_bytes(ok))
    if comm.rank == 0:
        print(f"[{'PASS' if allok else 'FAIL'}] pc_type mg as replicas on {comm.nranks} ranks: shape {shape} per {per} "
              f"{'slabs' if procs is None else 'process grid ' + str(procs)}, {its[0][0]} iterations {'; '.join(msgs)}", flush=True)
    return allok


def main():
    import torch

    comm = Comm.from_env()
    torch.cuda.set_device(comm.device)
    cases = [((24, 20, 44), (0, 0, 0)), ((16, 12, 31), (0, 0, 1)), ((70, 9, 17), (1, 1, 0))]
    new_cases = os.environ.get("B200_MGPU_BOX", "0") == "1"   # paths that have not run on GPUs yet (emulation only)
    if new_cases:
        cases += [((40, 26), (0, 0)), ((33, 21), (1, 1))]      # 2-D grids: y-slabs through the (nx, 1, ny) mapping
    modes = [("p2p", "store"), ("nccl", "memcpy"), ("nccl", "store")]
    args = [a for a in sys.argv[1:] if a != "--c4"]
    if "--c4" in sys.argv:
        # BASELINE.json config 4: 256 x 128 x 128 stretched grid, domain-decomposed (run on 4 GPUs)
        cases = [((256, 128, 128), (0, 0, 0))]
        orc.set_fast(True, 0)
    if args:
        modes = [tuple(m.split("+")) for m in args]
    allok = True
    for reduce, halo in modes:
        for shape, per in cases:
            for pc in (("none",) if "--c4" in sys.argv else ("none", "jacobi")):
                allok &= run_case(comm, shape, per, pc, reduce, halo)
    if "--c4" not in sys.argv:
        # process grids PETSc's DMDA may pick instead of 1 x 1 x P.  These cases were written after round 1's GPU budget was
        # spent and have not run on GPUs yet: they run (and count) with B200_MGPU_BOX=1 -- a one-sided failure inside the
        # collective set-up would otherwise hang the validated checks above with it
        if new_cases:
            grids3 = {2: [(2, 1, 1), (1, 2, 1)], 4: [(2, 2, 1), (2, 1, 2)], 8: [(2, 2, 2)]}.get(comm.nranks, [])
            for procs in grids3:
                allok &= run_box_case(comm, (22, 18, 19), (0, 0, 0), procs)
                allok &= run_box_case(comm, (16, 14, 12), (1, 0, 1), procs)
            grids2 = {2: [(2, 1)], 4: [(2, 2)], 8: [(4, 2)]}.get(comm.nranks, [])
            for procs in grids2:
                allok &= run_box_case(comm, (30, 23), (0, 0), procs)
        elif comm.rank == 0:
            print("[SKIP] 2-D grids and DMDA box cases on several GPUs (set B200_MGPU_BOX=1)", flush=True)
        if new_cases:
            allok &= run_replicated_case(comm, "velocity")
            allok &= run_replicated_case(comm, "ibpm")
            allok &= run_mg_case(comm, (48, 40, 44), (0, 0, 0))
            allok &= run_mg_case(comm, (40, 36), (0, 0))
            boxes = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(comm.nranks)
            if boxes:
                allok &= run_mg_case(comm, (22, 18, 19), (0, 0, 0), boxes)
    comm.barrier()
    if comm.rank == 0:
        print("MGPU_CHECK", "PASS" if allok else "FAIL", flush=True)
    import torch.distributed as dist

    if dist.is_initialized():
        dist.destroy_process_group()
    sys.exit(0 if allok else 1)


if __name__ == "__main__":
    main()

"""Time to solution of one pressure solve on several ranks (run under torchrun): the multigrid-preconditioned solve as
replicas (every rank solves the whole grid; all-gather of b per solve) against the distributed plain / Jacobi CG.
Wall-clock per solve through LinSolverB200.solve with host vectors (what the LinSolver interface receives), max over ranks."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import petibm_b200 as pb
from petibm_b200.dist import Comm

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, nargs=3, default=[256, 256, 256])
ap.add_argument("--rtol", type=float, default=1e-8)
ap.add_argument("--pcs", nargs="+", default=["mg", "jacobi"])
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
comm = Comm.from_env()
torch.cuda.set_device(comm.device)
n = tuple(a.size)
grid = pb.Grid.uniform(n, dt=0.01)
rng = np.random.default_rng(20240521)
xs = rng.standard_normal(grid.size); xs -= xs.mean()
xl = comm.local_block(xs, n) if comm.nranks > 1 else xs
b = None
for pc in a.pcs:
    s = pb.LinSolverB200("poisson", "None", comm=comm if comm.nranks > 1 else None, device=comm.device)
    s.setOptions(pc_type=pc, rtol=a.rtol, atol=1e-50, max_it=20000)
    s.setStencil(grid)
    s.setNullSpace(True)
    if b is None:
        b = s.apply(xl)
    x = np.empty_like(b)
    best = None
    for _ in range(a.reps):
        comm.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.solve(x, b)
        dt = comm.allreduce_max(time.perf_counter() - t0)
        best = dt if best is None or dt < best else best
    err = float(np.abs((x - x.mean()) - (xl - xl.mean())).max() / np.abs(xs).max())
    if comm.rank == 0:
        print(json.dumps({"size": list(n), "ranks": comm.nranks, "pc": pc, "mode": "replicas" if s._mg_rep is not None else "distributed slabs",
                          "iterations": s.getIters(), "wall_ms_per_solve": round(best * 1e3, 2), "device_solve_ms": round(s.timing()["solve_ms"], 2),
                          "max_rel_error": err}), flush=True)
    s.destroy()
if comm.nranks > 1:
    import torch.distributed as dist
    dist.barrier(); dist.destroy_process_group()

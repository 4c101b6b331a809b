"""Device-side timeline of the CG kernels (any number of ranks, run under torchrun for N>1)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import petibm_b200 as pb
from petibm_b200.dist import Comm

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, nargs=3, default=[256, 256, 256])
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--reduce", default="p2p")
ap.add_argument("--halo", default="store")
ap.add_argument("--graph", type=int, default=1)
ap.add_argument("--tune", nargs="*", default=[])
a = ap.parse_args()
comm = Comm.from_env(reduce=a.reduce, halo=a.halo)
torch.cuda.set_device(comm.device)
n = tuple(a.size)
grid = pb.Grid.uniform(n, dt=0.01)
s = pb.LinSolverB200("poisson", "None", comm=comm if comm.nranks > 1 else None, device=comm.device)
s.setOptions(rtol=0.0, atol=0.0, max_it=a.iters)
s.setTuning("use_graph", a.graph)
for kv in a.tune:
    k, v = kv.split("="); s.setTuning(k, int(v))
s.setStencil(grid); s.setNullSpace(True)
rng = np.random.default_rng(1)
xs = rng.standard_normal(grid.size); xs -= xs.mean()
xl = comm.local_block(xs, n) if comm.nranks > 1 else xs
b = torch.from_numpy(s.apply(xl)).cuda()
x = torch.empty_like(b)
def solve():
    try: s.solve(x, b)
    except pb.B200Error as e: assert e.code == -5
for _ in range(3): solve()
s.setTrace(4 * a.iters + 16)
solve()
tm = s.timing()
tr = s.getTrace(4 * a.iters + 16).astype(np.int64)
s.setTrace(0)
solve()
tm2 = s.timing()
if comm.rank == 0 or comm.rank == comm.nranks - 1:
    k = tr[:, 4]
    comp, red, fin = tr[:, 1] - tr[:, 0], tr[:, 2] - tr[:, 1], tr[:, 3] - tr[:, 2]
    gap = tr[1:, 0] - tr[:-1, 3]
    print(f"[rank {comm.rank}/{comm.nranks}] traced solve {tm['solve_ms']:.2f} ms loop {tm['loop_ms']:.2f} ms ({tm['loop_ms']/a.iters*1e3:.1f} us/iter); untraced {tm2['loop_ms']/a.iters*1e3:.1f} us/iter; entries {len(tr)}")
    for kind, name in ((0, "k_spmv  "), (1, "k_update")):
        m = k == kind
        m2 = m[:-1]
        print(f"   {name}: compute {np.median(comp[m])/1e3:7.2f} us (p90 {np.percentile(comp[m],90)/1e3:7.2f})  all-reduce {np.median(red[m])/1e3:6.2f} us (p90 {np.percentile(red[m],90)/1e3:6.2f})  scalars {np.median(fin[m])/1e3:5.2f} us  gap-to-next-start {np.median(gap[m2])/1e3:6.2f} us (p90 {np.percentile(gap[m2],90)/1e3:6.2f})")
s.destroy()
if comm.nranks > 1:
    import torch.distributed as dist
    dist.barrier(); dist.destroy_process_group()

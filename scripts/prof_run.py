"""Short single-GPU run for ncu captures: one solve of a few CG iterations at the benchmark size."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import petibm_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, nargs=3, default=[256, 256, 256])
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--pc", default="none")
ap.add_argument("--tune", nargs="*", default=[])
a = ap.parse_args()
grid = pb.Grid.uniform(tuple(a.size), dt=0.01)
s = pb.LinSolverB200("poisson", "None")
s.setOptions(pc_type=a.pc, rtol=0.0, atol=0.0, max_it=a.iters, check_every=4)
s.setTuning("use_graph", 0)
for kv in a.tune:
    k, v = kv.split("=")
    s.setTuning(k, int(v))
s.setStencil(grid)
s.setNullSpace(True)
rng = np.random.default_rng(1)
xs = rng.standard_normal(grid.size)
xs -= xs.mean()
b = s.apply(xs)
x = np.empty_like(b)
for _ in range(2):
    try:
        s.solve(x, b)
    except pb.B200Error as e:
        assert e.code == -5
print("iters", s.getIters(), "timing", s.timing())

import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import petibm_b200 as pb
ap = argparse.ArgumentParser()
ap.add_argument("--sizes", nargs="*", default=["256x256x32", "256x256x64"])
ap.add_argument("--blocks", type=int, nargs="*", default=[0, 128, 256, 512, 1024, 2048, 4096])
a = ap.parse_args()
for sz in a.sizes:
    n = tuple(int(v) for v in sz.split("x")); N = n[0]*n[1]*n[2]
    grid = pb.Grid.uniform(n, dt=0.01)
    xs = np.random.default_rng(1).standard_normal(N); xs -= xs.mean()
    out = []
    for ub in a.blocks:
        s = pb.LinSolverB200("poisson", "None")
        s.setOptions(rtol=0.0, atol=0.0, max_it=4)
        s.setTuning("upd_blocks", ub)
        s.setStencil(grid); s.setNullSpace(True)
        b = s.apply(xs); x = np.empty_like(b)
        try: s.solve(x, b)
        except pb.B200Error as e: assert e.code == -5
        t = s.timeKernel(1, 20, N*8*5 > 100e6)
        out.append(f"b{ub}={t*1e3:.1f}us")
        s.destroy()
    print(sz, "k_update:", " ".join(out), flush=True)

"""Kernel-level sweep on one GPU: average launch time of the fused SpMV (class 0) and update (class 1)
kernels for several tile variants / z-chunk sizes, L2 flushed between launches."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import petibm_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, nargs=3, default=[256, 256, 256])
ap.add_argument("--tiles", type=int, nargs="*", default=[0, 10, 13, 15, 18, 30, 31, 32, 33])
ap.add_argument("--kz", type=int, nargs="*", default=[0])
ap.add_argument("--upd", type=int, nargs="*", default=[0])
ap.add_argument("--pc", default="none")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
n = tuple(a.size)
N = n[0] * n[1] * n[2]
grid = pb.Grid.uniform(n, dt=0.01)
rng = np.random.default_rng(1)
xs = rng.standard_normal(N); xs -= xs.mean()
for tile in a.tiles:
    for kz in a.kz:
        s = pb.LinSolverB200("poisson", "None")
        s.setOptions(pc_type=a.pc, rtol=0.0, atol=0.0, max_it=6)
        s.setTuning("tile", tile); s.setTuning("kz_chunk", kz)
        s.setStencil(grid); s.setNullSpace(True)
        b = s.apply(xs); x = np.empty_like(b)
        try: s.solve(x, b)
        except pb.B200Error as e: assert e.code == -5
        t0 = s.timeKernel(0, a.reps, True)
        t0w = s.timeKernel(0, a.reps, False)
        print(f"tile {tile:2d} kz {kz:3d}: k_spmv  {t0*1e3:7.1f} us flushed ({48*N/t0/1e6:7.0f} GB/s)   {t0w*1e3:7.1f} us warm ({48*N/t0w/1e6:7.0f} GB/s)", flush=True)
        s.destroy()
for ub in a.upd:
    s = pb.LinSolverB200("poisson", "None")
    s.setOptions(pc_type=a.pc, rtol=0.0, atol=0.0, max_it=6)
    s.setTuning("upd_blocks", ub)
    s.setStencil(grid); s.setNullSpace(True)
    b = s.apply(xs); x = np.empty_like(b)
    try: s.solve(x, b)
    except pb.B200Error as e: assert e.code == -5
    t1 = s.timeKernel(1, a.reps, True)
    t1w = s.timeKernel(1, a.reps, False)
    print(f"upd_blocks {ub:5d}: k_update {t1*1e3:7.1f} us flushed ({24*N/t1/1e6:7.0f} GB/s)   {t1w*1e3:7.1f} us warm ({24*N/t1w/1e6:7.0f} GB/s)", flush=True)
    s.destroy()

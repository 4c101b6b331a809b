"""Sweep (tile, kz_chunk) of the fused SpMV kernel for several local grid sizes on one GPU (warm, no L2 flush
for slabs that fit L2, flushed otherwise) -> table used to derive the launch heuristic."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import petibm_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", nargs="*", default=["256x256x32", "256x256x64", "256x256x128", "256x256x256", "128x128x128"])
ap.add_argument("--tiles", type=int, nargs="*", default=[10, 18])
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
for sz in a.sizes:
    n = tuple(int(v) for v in sz.split("x"))
    N = n[0] * n[1] * n[2]
    grid = pb.Grid.uniform(n, dt=0.01)
    xs = np.random.default_rng(1).standard_normal(N); xs -= xs.mean()
    res = []
    for tile in a.tiles:
        ty = {10: 6, 18: 10}.get(tile, 6)
        cands = sorted({max(4, (n[2] + c - 1) // c) for c in (1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16)})
        for kz in cands:
            s = pb.LinSolverB200("poisson", "None")
            s.setOptions(rtol=0.0, atol=0.0, max_it=4)
            s.setTuning("tile", tile); s.setTuning("kz_chunk", kz)
            s.setStencil(grid); s.setNullSpace(True)
            b = s.apply(xs); x = np.empty_like(b)
            try: s.solve(x, b)
            except pb.B200Error as e: assert e.code == -5
            flush = N * 8 * 5 > 100e6
            t = s.timeKernel(0, a.reps, flush)
            nch = (n[2] + kz - 1) // kz
            blocks = ((n[0] + 63) // 64) * ((n[1] + ty - 1) // ty) * nch
            res.append((t, tile, kz, nch, blocks))
            s.destroy()
    res.sort()
    print(f"== {sz}: best " + "  ".join(f"[t{r[1]} kz{r[2]} nch{r[3]} blk{r[4]}: {r[0]*1e3:.1f}us]" for r in res[:5]), flush=True)
    print("   all: " + " ".join(f"t{r[1]}/kz{r[2]}/b{r[4]}={r[0]*1e3:.1f}" for r in sorted(res, key=lambda r: (r[1], r[2]))), flush=True)

"""Time to solution of one pressure solve on the GPU for the preconditioners of the B200 backend (none / jacobi / mg):
iterations, device time of the solve (CUDA events inside libb200ls), kernel launches.  Not the headline metric of
bench.py (CG iterations/s); this is the number an application sees per time step.

    python scripts/tts_bench.py --size 256 256 256 --rtol 1e-8 [--stretched]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petibm_b200 as pb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, nargs="+", default=[256, 256, 256])
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--stretched", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--pcs", nargs="+", default=["none", "jacobi", "mg"])
    ap.add_argument("--smooth", type=int, default=2)
    ap.add_argument("--mg-graph", type=int, default=0, help="replay pairs of multigrid-preconditioned iterations as a CUDA graph")
    ap.add_argument("--mg-tail", type=int, default=0, help="coarse levels of the cycle as one launch (k_mg_tail)")
    ap.add_argument("--mg-fuse", type=int, default=0, help="residual update and reduction folded into the cycle's first/last step")
    args = ap.parse_args()
    n = tuple(args.size)
    rng = np.random.default_rng(20240521)
    if args.stretched:
        widths = [rng.uniform(0.7, 1.3, m) / m for m in n]
        grid = pb.Grid(widths, (False,) * 3, 0.01)
    else:
        grid = pb.Grid.uniform(n, dt=0.01)
    xs = rng.standard_normal(grid.size)
    xs -= xs.mean()
    b = None
    for pc in args.pcs:
        s = pb.LinSolverB200("poisson", "None")
        s.setOptions(pc_type=pc, rtol=args.rtol, atol=1e-50, max_it=20000, mg_smooth_its=args.smooth)
        s.setTuning("mg_graph", args.mg_graph)
        s.setTuning("mg_tail", args.mg_tail)
        s.setTuning("mg_fuse", args.mg_fuse)
        s.setStencil(grid)
        s.setNullSpace(True)
        if b is None:
            b = s.apply(xs)
        x = np.empty_like(b)
        best = None
        for _ in range(args.reps):
            s.solve(x, b)
            t = s.timing()
            best = t if best is None or t["solve_ms"] < best["solve_ms"] else best
        err = float(np.abs((x - x.mean()) - xs).max() / np.abs(xs).max())
        print(json.dumps({"size": list(n), "stretched": args.stretched, "pc": pc, "rtol": args.rtol, "iterations": s.getIters(),
                          "reason": s.getReason(), "solve_ms": round(best["solve_ms"], 3), "launches": best["launches"],
                          "max_rel_error": err, "mg_graph": args.mg_graph, "mg_tail": args.mg_tail, "mg_fuse": args.mg_fuse}), flush=True)
        s.destroy()


if __name__ == "__main__":
    main()

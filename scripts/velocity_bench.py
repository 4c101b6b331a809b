"""Velocity system A = I/dt - c nu L (createlaplacian.cpp:108-262, navierstokes.cpp:342-344), BiCGStab + Jacobi as in
every shipped velocity_solver.info: iterations/s and effective HBM GB/s of the line-coefficient operator (176 B/row per
iteration) and of the assembled CSR path (about 360 B/row), with the oracle's KSPSolve_BCGS on the host
cores beside them.

    python scripts/velocity_bench.py --size 128 128 128 [--iters 60] [--no-cpu]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import petibm_b200 as pb
from tests import helpers as H

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, nargs="+", default=[128, 128, 128])
ap.add_argument("--iters", type=int, default=60)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--stretched", action="store_true")
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--graph", type=int, default=0)
ap.add_argument("--tiles", type=int, nargs="+", default=[-1], help="tuning sep_tile values for the line-coefficient form (-1: the library's rule)")
ap.add_argument("--zchunks", type=int, nargs="+", default=[0], help="tuning sep_zchunk values tried with every tile > 0")
ap.add_argument("--stages", type=int, nargs="+", default=[3], help="tuning sep_stages values tried with every tile > 0")
ap.add_argument("--no-csr", action="store_true")
a = ap.parse_args()
n = tuple(a.size)
widths = H.make_widths(n, stretched=a.stretched)
per = (0,) * len(n)
t0 = time.perf_counter()
A = H.velocity_system_fast(widths, per, dt=0.01, nu=0.01, c=0.5)
t_asm = time.perf_counter() - t0
rows = A.shape[0]
rng = np.random.default_rng(3)
b = rng.standard_normal(rows)
M = pb.Mat.from_scipy(A)
variants = [("staggered", t, z, g) for t in a.tiles for z in (a.zchunks if t > 0 else [0]) for g in (a.stages if t > 0 else [3])] + \
           ([] if a.no_csr else [("csr", 0, 0, 3)])
for form, tile, zchunk, stages in variants:
    s = pb.LinSolverB200("velocity", "None")
    s.setOptions(ksp_type="bcgs", pc_type="jacobi", rtol=0.0, atol=0.0, max_it=a.iters)
    s.setTuning("csr_graph", a.graph)
    s.setTuning("sep_tile", tile)
    s.setTuning("sep_zchunk", zchunk)
    s.setTuning("sep_stages", stages)
    s.setGrid(pb.Grid([np.asarray(w) for w in widths], (False,) * 3, 0.01))
    s.setStaggered(form == "staggered")
    s.setMatrix(M)
    x = np.empty(rows)
    best = None
    for _ in range(a.reps):
        try:
            s.solve(x, b)
        except pb.B200Error as e:
            assert e.code == -5
        t = s.timing()
        best = t if best is None or t["loop_ms"] < best["loop_ms"] else best
    its = s.getIters()
    per_it = best["loop_ms"] / its * 1e-3
    model = 176.0 if s.operator == "staggered" else 360.0
    print(json.dumps({"system": "velocity A = I/dt - c nu L, BiCGStab + Jacobi", "size": list(n), "rows": rows, "operator": s.operator,
                      "sep_tile": tile, "sep_zchunk": zchunk, "sep_stages": stages,
                      "iterations": its, "loop_ms": round(best["loop_ms"], 3), "iterations_per_s": round(1.0 / per_it, 1),
                      "model_bytes_per_row": model, "model_GBs": round(model * rows / per_it / 1e9, 1),
                      "launches": best["launches"], "csr_graph": a.graph, "assembly_s": round(t_asm, 1),
                      "history_5": [float(v) for v in s.getHistory()[:5]]}), flush=True)
    hist_gpu = s.getHistory()
    s.destroy()
if not a.no_cpu:
    from oracle import oracle as orc
    Ao = orc.Csr.from_arrays(rows, rows, A.indptr, A.indices, A.data)
    threads = len(os.sched_getaffinity(0))
    orc.set_fast(True, threads)
    nit = min(a.iters, 30)
    t0 = time.perf_counter()
    ref = orc.ksp_solve(Ao, b, ksp_type="bcgs", pc_type="jacobi", rtol=0.0, atol=0.0, max_it=nit)
    dt = time.perf_counter() - t0
    orc.set_fast(False, 0)
    m = min(len(hist_gpu), len(ref.history), 12)
    rel = float(np.max(np.abs(hist_gpu[:m] - ref.history[:m]) / ref.history[:m]))
    print(json.dumps({"system": "velocity, oracle KSPSolve_BCGS port on the host", "cores": threads, "iterations": ref.its,
                      "iterations_per_s": round(ref.its / dt, 2), "history_max_rel_first_%d" % m: rel}), flush=True)

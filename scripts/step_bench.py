"""Device-resident fractional-step skeleton around the pressure solve: per time step
    N(q) (convection, createconvection.cpp)  ->  rhs2 = D u*  ->  KSPSolve(dP)  ->  u = u* - BNG dP, p += dP
with every vector resident in HBM, against the same step with rhs2 / dP crossing PCIe through host buffers the way the
LinSolver interface receives them inside PetIBM.  Reports ms per step and the per-kernel times.

    python scripts/step_bench.py --size 256 256 256 --pc mg
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import petibm_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, nargs="+", default=[256, 256, 256])
ap.add_argument("--pc", default="mg")
ap.add_argument("--rtol", type=float, default=1e-6)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
n = tuple(a.size)
dev = torch.device("cuda", 0)
grid = pb.Grid.uniform(n, dt=0.01)
s = pb.LinSolverB200("poisson", "None")
s.setOptions(pc_type=a.pc, rtol=a.rtol, atol=1e-50, max_it=20000)
if a.pc == "mg":
    s.setTuning("mg_tail", 1); s.setTuning("mg_fuse", 1); s.setTuning("mg_graph", 1)
s.setStencil(grid)
s.setNullSpace(True)
nv, npr = s.velocitySize()
sizes = s.ghostedSizes()
dim = len(n)
rng = np.random.default_rng(5)
u = torch.from_numpy(rng.standard_normal(nv)).to(dev)
p = torch.zeros(npr, dtype=torch.float64, device=dev)
q = [torch.zeros(sizes[f], dtype=torch.float64, device=dev) for f in range(dim)]
conv = torch.empty(nv, dtype=torch.float64, device=dev)
rhs = torch.empty(npr, dtype=torch.float64, device=dev)
dp = torch.empty(npr, dtype=torch.float64, device=dev)
rhs_pin = torch.empty(npr, dtype=torch.float64).pin_memory()
dp_pin = torch.empty(npr, dtype=torch.float64).pin_memory()

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def step(host_vectors):
    marks = [ev()]
    s.ghostedFromPacked(u, q); s.convection(q, out=conv); marks.append(ev())          # explicit convection term
    u.add_(conv, alpha=-0.01 * 1e-3); marks.append(ev())                                # stand-in for the velocity solve (u*)
    s.divergence(u, out=rhs); marks.append(ev())                                        # rhs2 = D u*
    if host_vectors:
        rhs_pin.copy_(rhs); torch.cuda.synchronize()
        s.solve(dp_pin, rhs_pin)                                                        # LinSolver interface: host Vec in / out
        dp.copy_(dp_pin, non_blocking=True)
    else:
        s.solve(dp, rhs)
    marks.append(ev())
    s.project(u, p, dp); marks.append(ev())                                             # u = u* - BNG dP ; p += dP
    torch.cuda.synchronize()
    return [marks[i].elapsed_time(marks[i + 1]) for i in range(len(marks) - 1)], s.getIters()

for host in (False, True):
    step(host)
    acc, its = np.zeros(5), 0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        t, it = step(host); acc += t; its += it
    wall = (time.perf_counter() - t0) / a.steps * 1e3
    acc /= a.steps
    print(json.dumps({"size": list(n), "pc": a.pc, "rtol": a.rtol, "vectors": "host (PCIe each step)" if host else "device-resident",
                      "ms_per_step_wall": round(wall, 3), "iterations_per_solve": its / a.steps,
                      "ms": {"convection(+ghost fill)": round(acc[0], 3), "axpy": round(acc[1], 3), "divergence": round(acc[2], 3),
                             "solve": round(acc[3], 3), "projection": round(acc[4], 3)},
                      "GBs": {"convection": round(16.0 * nv / acc[0] / 1e6, 0), "divergence": round((8.0 * nv + 8.0 * npr) / acc[2] / 1e6, 0),
                              "projection": round((24.0 * nv + 24.0 * npr) / acc[4] / 1e6, 0)}}), flush=True)
s.destroy()

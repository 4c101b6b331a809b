"""Forces-system direct solve (preonly + lu, dense_kernels.cuh) on the GPU: factorisation and substitution times for
n = 500 / 2000 / 4096 against numpy's LAPACK on the host cores."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import petibm_b200 as pb

for n in [int(a) for a in sys.argv[1:]] or [500, 2000, 4096]:
    rng = np.random.default_rng(n)
    # SPD like E BN H: banded coupling of neighbouring Lagrangian points plus a weak dense part
    B = sp.random(n, n, density=min(1.0, 60.0 / n), random_state=7, format="csr")
    A = (B @ B.T + sp.identity(n) * 0.5).toarray() + 1e-3 * np.ones((n, n))
    M = sp.csr_matrix(A)
    b = rng.standard_normal(n)
    s = pb.LinSolverB200("forces", "None")
    s.setOptions(ksp_type="preonly", pc_type="lu")
    x = np.empty(n)
    out = {"n": n, "nnz": int(M.nnz)}
    for rep in range(2):                       # second round: kernels warm
        s.setMatrix(pb.Mat.from_scipy(M))      # moving bodies: every time step (rigidkinematics.cpp:119-140)
        s.solve(x, b)
        t = s.timing()
        out["factor_plus_solve_ms"] = round(t["solve_ms"], 3)
        out["launches_factor"] = t["launches"]
        s.solve(x, b)
        t = s.timing()
        out["solve_ms"] = round(t["solve_ms"], 3)
        out["substitution_ms"] = round(t["loop_ms"], 3)
    t0 = time.perf_counter(); ref = np.linalg.solve(A, b); out["lapack_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
    out["max_rel_err"] = float(np.abs(x - ref).max() / np.abs(ref).max())
    print(json.dumps(out), flush=True)
    s.destroy()

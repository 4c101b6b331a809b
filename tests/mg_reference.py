"""Checker (test infrastructure) for the geometric multigrid preconditioner of petibm_b200/csrc/mg_kernels.cuh: an
independent numpy/scipy restatement with ASSEMBLED level operators, explicit restriction matrices and the textbook
Chebyshev recurrence (Saad, Iterative Methods for Sparse Linear Systems, alg. 12.1) -- nothing here shares code with the
kernels or with mg_schedule.h.  The fine-level operator built here (a sum over cell faces) is tied to the pinned
oracle's literal D (dt I) G assembly in tests/test_emulated_mg.py.

PetIBM itself has no geometric multigrid (its shipped Poisson configurations use PETSc GAMG / AmgX classical AMG:
examples/*/config/poisson_solver.info), so this is the only oracle that preconditioner can have; what it is measured
against in the end is plain CG: same converged solution, far fewer iterations."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def faces(d, periodic, dt):
    """g[s] = dt * (1 / (0.5 (d[s-1] + d[s]))) on the minus face of cell s; 0 at walls; wrap value at both ends."""
    m = len(d)
    g = np.zeros(m + 1)
    for s in range(1, m):
        g[s] = dt * (1.0 / (0.5 * (d[s] + d[s - 1])))
    if periodic:
        g[0] = g[m] = dt * (1.0 / (0.5 * (d[0] + d[m - 1])))
    return g


def assemble(d, per, dt, active):
    """A = - sum over faces c_f (e_a - e_b)(e_a - e_b)^T, c_f = (face area) * dt / h: the pressure operator D (dt I) G of
    the cells with widths d[0..2], including wrap faces of periodic axes."""
    n = [len(a) for a in d]
    N = int(np.prod(n))
    idx = np.arange(N).reshape(n[2], n[1], n[0])          # [k, j, i]
    rows, cols, vals = [], [], []
    area = [np.multiply.outer(d[2], d[1])[:, :, None] * np.ones(n[0]),      # x faces: dy*dz  -> [k, j, i]
            np.multiply.outer(d[2], np.ones(n[1]))[:, :, None] * d[0],     # y faces: dx*dz
            np.multiply.outer(np.ones(n[2]), d[1])[:, :, None] * d[0]]     # z faces: dx*dy
    for ax in range(3):
        if not active[ax]:
            continue
        g = faces(d[ax], per[ax], dt)
        axis = 2 - ax                                          # numpy axis of direction ax
        for s in range(n[ax] + (0 if per[ax] else -1)):        # face between cell s and s+1 (wrap: n-1 and 0)
            a = np.take(idx, s, axis=axis).ravel()
            b = np.take(idx, (s + 1) % n[ax], axis=axis).ravel()
            c = np.take(area[ax], s, axis=axis).ravel() * g[s + 1]     # g[n] is the wrap face of a periodic axis
            for r_, c_, v_ in ((a, a, -c), (b, b, -c), (a, b, c), (b, a, c)):
                rows.append(r_); cols.append(c_); vals.append(v_)
    if not rows:
        return sp.csr_matrix((N, N))
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    A.sum_duplicates()
    return A


def merge_map(d, tau):
    """Greedy pairing from the left: cells i, i+1 form one coarse cell iff d[i] + d[i+1] <= tau."""
    out, i, I = [], 0, 0
    while i < len(d):
        if i + 1 < len(d) and d[i] + d[i + 1] <= tau:
            out += [I, I]
            i += 2
        else:
            out += [I]
            i += 1
        I += 1
    return np.array(out)


def hierarchy(widths, per, dt, max_levels=0, ratio=2.0):
    """Levels [{d, n, A, R}].  Width-equalising coarsening: on every level the smallest sum of two neighbouring widths
    over all axes with at least four cells sets the scale; pairs whose merged width is within `ratio` of it are merged,
    everything wider waits (a uniform grid is halved, a stretched one loses its fine band first).  R sums the merged
    cells (restriction), R^T is the prolongation."""
    dim = len(widths)
    d = [np.asarray(w, dtype=np.float64) for w in widths] + [np.ones(1)] * (3 - dim)
    per = list(per) + [0] * (3 - len(per))
    active = [len(a) > 1 or bool(per[q]) for q, a in enumerate(d)]
    levels = []
    cap = max_levels if max_levels > 0 else 32
    while True:
        n = [len(a) for a in d]
        lev = {"d": d, "n": n, "A": assemble(d, per, dt, active), "R": None}
        levels.append(lev)
        cand = [q for q in range(3) if n[q] >= 4]
        if not cand or len(levels) >= cap:
            break
        tau = ratio * min(float((d[q][:-1] + d[q][1:]).min()) for q in cand)
        maps = [merge_map(d[q], tau) if q in cand else np.arange(n[q]) for q in range(3)]
        dc = [np.bincount(maps[q], weights=d[q]) for q in range(3)]       # widths add up
        nc = [len(a) for a in dc]
        K, J, I = np.meshgrid(maps[2], maps[1], maps[0], indexing="ij")
        coarse = (I + nc[0] * (J + nc[1] * K)).ravel()
        N = int(np.prod(n))
        lev["R"] = sp.csr_matrix((np.ones(N), (coarse, np.arange(N))), shape=(int(np.prod(nc)), N))
        d = dc
    return levels


def chebyshev(A, dinv, b, x, its, lmax, ratio):
    """its Chebyshev updates for D^-1 A with spectrum in [lmax/ratio, lmax], starting from x (None = zero)."""
    lmin = lmax / ratio
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    if x is None:
        x = np.zeros_like(b)
        r = b.copy()
    else:
        r = b - A @ x
    dvec = dinv * r / theta
    for k in range(its):
        x = x + dvec
        if k == its - 1:
            break
        r = b - A @ x
        rho1 = 1.0 / (2.0 * sigma - rho)
        dvec = rho1 * rho * dvec + (2.0 * rho1 / delta) * (dinv * r)
        rho = rho1
    return x


class VCycle:
    def __init__(self, widths, per, dt, max_levels=0, smooth_its=2, coarse_its=16, lmax=2.0, smooth_ratio=5.0, coarse_ratio=40.0):
        self.levels = hierarchy(widths, per, dt, max_levels)
        for lev in self.levels:
            dg = lev["A"].diagonal()
            lev["dinv"] = np.where(dg != 0.0, 1.0 / np.where(dg != 0.0, dg, 1.0), 0.0)
        self.smooth_its, self.coarse_its = smooth_its, coarse_its
        self.lmax, self.smooth_ratio, self.coarse_ratio = lmax, smooth_ratio, coarse_ratio

    def apply(self, b, l=0):
        lev = self.levels[l]
        if l == len(self.levels) - 1:
            return chebyshev(lev["A"], lev["dinv"], b, None, self.coarse_its, self.lmax, self.coarse_ratio)
        x = chebyshev(lev["A"], lev["dinv"], b, None, self.smooth_its, self.lmax, self.smooth_ratio)
        bc = lev["R"] @ (b - lev["A"] @ x)
        x = x + lev["R"].T @ self.apply(bc, l + 1)
        return chebyshev(lev["A"], lev["dinv"], b, x, self.smooth_its, self.lmax, self.smooth_ratio)


def pcg(A, b, M, const_nullspace=True, rtol=0.0, atol=0.0, max_it=50, nullvec=None):
    """KSPSolve_CG semantics (zero guess, left PC, preconditioned norm, MatNullSpaceRemove after PCApply -- the constant
    or one explicit orthonormal vector --, KSPConvergedDefault) with z = M(r).  Returns (x, history, its, reason)."""
    n = b.size
    x = np.zeros(n)
    r = b.copy()

    def pc(v):
        z = M(v)
        if nullvec is not None:
            z = z + (-(z @ nullvec)) * nullvec
        elif const_nullspace:
            z = z + z.sum() / (-1.0 * n)
        return z

    z = pc(r)
    dp = np.sqrt(z @ z)
    hist = [dp]
    ttol = max(rtol * dp, atol)
    if dp <= ttol:
        return x, np.array(hist), 0, (3 if dp < atol else 2)
    beta = z @ r
    p = None
    betaold = 1.0
    for i in range(max_it):
        p = z.copy() if i == 0 else z + (beta / betaold) * p
        w = A @ p
        a = beta / (p @ w)
        x = x + a * p
        r = r - a * w
        z = pc(r)
        dp = np.sqrt(z @ z)
        hist.append(dp)
        if dp <= ttol:
            return x, np.array(hist), i + 1, (3 if dp < atol else 2)
        betaold = beta
        beta = z @ r
    return x, np.array(hist), max_it, -3

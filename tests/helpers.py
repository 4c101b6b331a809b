"""Shared problem builders for the parity tests (oracle side = checker, petibm_b200 = product)."""
from __future__ import annotations

import numpy as np

from oracle import oracle as orc

SEED = 20240521


def make_widths(shape, stretched=True, seed=SEED):
    rng = np.random.default_rng(seed)
    return [(rng.uniform(0.7, 1.3, n) if stretched else np.ones(n)) / n for n in shape]


def oracle_matrix(widths, periodic, dt=0.01):
    per = list(periodic) + [0] * (3 - len(periodic))
    return orc.assemble_dbng(widths, per[: len(widths)] + [0] * (3 - len(widths)), dt)


def consistent_rhs(A, seed=SEED):
    rng = np.random.default_rng(seed + 1)
    xs = rng.standard_normal(A.shape[0])
    xs -= xs.mean()
    return A.spmv(xs), xs


def grid_of(widths, periodic, dt=0.01):
    from petibm_b200 import Grid

    per = tuple(bool(p) for p in periodic) + (False,) * (3 - len(periodic))
    return Grid([np.asarray(w, dtype=np.float64) for w in widths], per, dt)


def mat_of(A, const_nullspace=True):
    from petibm_b200 import Mat

    rp, col, val = A.arrays()
    return Mat(rp, col, val, A.shape[1]).setNullSpace(const_nullspace)


def velocity_system(widths, periodic, dt=0.01, nu=0.01, c=0.5, vmin=0.0):
    """Test-side assembly of PetIBM's implicit velocity operator A = I/dt - c*nu*L in packed [u|v|w]
    ordering: createLaplacian (src/operators/createlaplacian.cpp:108-162: 1/(dLNeg*dLSelf), 1/(dLPos*dLSelf),
    diagonal = -sum incl. ghost-side terms; :232-243 ghost coefficient * a0 folded into the row's diagonal) with
    all-Dirichlet walls (a0 = 0 normal, -1 tangential, singleboundarydirichlet.cpp:33-42), then MatScale(-c*nu)
    and MatShift(1/dt) (navierstokes.cpp:342-344).  Returns a scipy CSR matrix with sorted columns."""
    import scipy.sparse as sp

    dim = len(widths)
    per = list(periodic) + [0] * (3 - dim)
    ext = [float(np.sum(w)) for w in widths]
    ax = [[orc.velocity_axis(widths[d], vmin, vmin + ext[d], same_dir=(f == d), periodic=bool(per[d])) for d in range(dim)]
          for f in range(dim)]
    nn = [[ax[f][d][0] for d in range(dim)] + [1] * (3 - dim) for f in range(dim)]
    off = np.concatenate([[0], np.cumsum([int(np.prod(nn[f])) for f in range(dim)])])
    rows, cols, vals = [], [], []
    for f in range(dim):
        n0, n1, n2 = nn[f]
        for k in range(n2):
            for j in range(n1):
                for i in range(n0):
                    idx = (i, j, k)
                    row = off[f] + i + n0 * (j + n1 * k)
                    v = []
                    for d in range(dim):
                        _, dL, co = ax[f][d]
                        s = idx[d]
                        dls = dL[s + 1]
                        v.append(1.0 / ((co[s + 1] - co[s]) * dls))
                        v.append(1.0 / ((co[s + 2] - co[s + 1]) * dls))
                    acc = 0.0
                    for t in v:
                        acc = acc + t
                    diag = -acc
                    entries = {}
                    for d in range(dim):
                        for side, coef in ((-1, v[2 * d]), (+1, v[2 * d + 1])):
                            nb = list(idx)
                            nb[d] += side
                            nd = nn[f][d]
                            if 0 <= nb[d] < nd:
                                pass
                            elif per[d]:
                                nb[d] %= nd
                            else:
                                a0 = 0.0 if d == f else -1.0     # Dirichlet: normal / tangential ghost
                                diag = diag + coef * a0
                                continue
                            col = off[f] + nb[0] + n0 * (nb[1] + n1 * nb[2])
                            entries[col] = entries.get(col, 0.0) + coef
                    entries[row] = diag
                    for col in sorted(entries):
                        rows.append(row)
                        cols.append(col)
                        vals.append(entries[col])
    L = sp.csr_matrix((vals, (rows, cols)), shape=(off[-1], off[-1]))
    A = L.copy()
    A.data = (-(c * nu)) * A.data
    A = (A + sp.identity(off[-1], format="csr") * (1.0 / dt)).tocsr()
    A.sort_indices()
    return A, L


def box_local_system(A, b, dim, n, procs, rank):
    """What one MPI rank of PetIBM holds when the DMDA uses the process grid `procs` (cartesianmesh.cpp:500-538,
    709-721): the rows of its box in PETSc ordering with PETSc global column indices, and its part of b.
    A: oracle matrix in natural ordering.  Returns (Mat, b_local, plan)."""
    from petibm_b200 import Mat
    from petibm_b200.dist import Repart

    plan = Repart(dim, n, procs, rank)
    total = int(np.prod(n))
    to_nat = plan.petsc_to_natural(np.arange(total))
    to_petsc = np.empty(total, dtype=np.int64)
    to_petsc[to_nat] = np.arange(total)
    rp, col, val = A.arrays()
    rows = plan.box_rows()
    cnt = rp[rows + 1] - rp[rows]
    indptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    take = np.concatenate([np.arange(rp[r], rp[r + 1]) for r in rows]) if rows.size else np.zeros(0, dtype=np.int64)
    cols_p = to_petsc[col[take]]
    vals = val[take].copy()
    # PETSc keeps every row sorted by (PETSc) column index
    for q in range(rows.size):
        a, e = indptr[q], indptr[q + 1]
        o = np.argsort(cols_p[a:e], kind="stable")
        cols_p[a:e] = cols_p[a:e][o]
        vals[a:e] = vals[a:e][o]
    return Mat(indptr, cols_p.astype(np.int32), vals, total), np.ascontiguousarray(b[rows]), plan

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


def velocity_system_fast(widths, periodic, dt=0.01, nu=0.01, c=0.5, vmin=0.0):
    """velocity_system vectorised with numpy (same entries bit for bit, checked on small grids by
    tests/test_velocity_operator.py); for the sizes of scripts/velocity_bench.py.  Periodic axes need >= 3 cells."""
    import scipy.sparse as sp

    dim = len(widths)
    per = list(periodic) + [0] * (3 - dim)
    ext = [float(np.sum(w)) for w in widths]
    ax = [[orc.velocity_axis(widths[d], vmin, vmin + ext[d], same_dir=(f == d), periodic=bool(per[d])) for d in range(dim)]
          for f in range(dim)]
    nn = [[ax[f][d][0] for d in range(dim)] + [1] * (3 - dim) for f in range(dim)]
    off = np.concatenate([[0], np.cumsum([int(np.prod(nn[f])) for f in range(dim)])])
    R, Cc, V = [], [], []
    for f in range(dim):
        n0, n1, n2 = nn[f]
        idx = np.meshgrid(np.arange(n0), np.arange(n1), np.arange(n2), indexing="ij")   # [d][i, j, k]
        row = off[f] + idx[0] + n0 * (idx[1] + n1 * idx[2])
        vm, vp = [], []
        for d in range(dim):
            _, dL, co = ax[f][d]
            s = np.arange(nn[f][d])
            dls = dL[s + 1]
            vm.append(1.0 / ((co[s + 1] - co[s]) * dls))
            vp.append(1.0 / ((co[s + 2] - co[s + 1]) * dls))
        acc = np.zeros(row.shape)
        for d in range(dim):
            acc = acc + vm[d][idx[d]]
            acc = acc + vp[d][idx[d]]
        diag = -acc
        for d in range(dim):
            nd = nn[f][d]
            for side, coef1 in ((-1, vm[d]), (+1, vp[d])):
                coef = coef1[idx[d]]
                nb = idx[d] + side
                inside = (nb >= 0) & (nb < nd)
                if per[d]:
                    assert nd >= 3
                    nb = nb % nd
                    inside = np.ones_like(inside)
                else:
                    a0 = 0.0 if d == f else -1.0
                    diag = np.where(inside, diag, diag + coef * a0)
                nbi = [idx[0], idx[1], idx[2]]
                nbi[d] = nb
                col = off[f] + nbi[0] + n0 * (nbi[1] + n1 * nbi[2])
                R.append(row[inside]); Cc.append(col[inside]); V.append((-(c * nu)) * coef[inside])
        R.append(row.ravel()); Cc.append(row.ravel()); V.append(((-(c * nu)) * diag + 1.0 / dt).ravel())
    A = sp.csr_matrix((np.concatenate(V), (np.concatenate(R), np.concatenate(Cc))), shape=(off[-1], off[-1]))
    A.sort_indices()
    return A


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


def ibpm_system(widths, dt=0.01, nb=24, radius=0.3, centre=None):
    """Test-side assembly of IBPM's modified Poisson system [D;E] BN [G,-H] (applications/ibpm/ibpm.cpp:164-194) on a
    (stretched) grid with non-periodic walls: D and G from the oracle's literal assembly (createdivergence.cpp:140-223,
    creategradient.cpp:70-128), E a regularised-delta interpolation of each velocity component onto nb Lagrangian points
    on a circle (createdelta.cpp:32-164 in spirit: a 4 x 4(x 4) patch of the Roma kernel), H = W E^T with
    W = diag(1 / (face area * face spacing)) so that the product is symmetric like PetIBM's, BN = dt I, and both
    MatMatMult calls through the oracle's restatement (operand and accumulation order of navierstokes.cpp:351-356).
    Returns (M as scipy CSR with sorted rows, pN, null vector: constant on the pressure block, zero on the forces)."""
    import scipy.sparse as sp

    dim = len(widths)
    per = [0] * dim
    D = orc.assemble_divergence(widths, per + [0] * (3 - dim))
    G = orc.assemble_gradient(widths, per + [0] * (3 - dim))
    Ds, Gs = D.to_scipy().tocsc(), G.to_scipy().tocsr()
    UN, pN = Gs.shape
    n = [len(w) for w in widths]
    edges = [np.concatenate([[0.0], np.cumsum(w)]) for w in widths]
    cen = [0.5 * (e[:-1] + e[1:]) for e in edges]
    if centre is None:
        centre = [0.5 * e[-1] for e in edges]
    hmin = min(float(np.min(w)) for w in widths)

    def roma(r):
        r = np.abs(r) / hmin
        return np.where(r <= 0.5, (1 + np.sqrt(np.maximum(1 - 3 * r * r, 0))) / 3,
                        np.where(r <= 1.5, (5 - 3 * r - np.sqrt(np.maximum(1 - 3 * (1 - r) ** 2, 0))) / 6, 0.0)) / hmin

    th = np.linspace(0.0, 2 * np.pi, nb, endpoint=False)
    pts = np.zeros((nb, dim))
    pts[:, 0] = centre[0] + radius * np.cos(th)
    pts[:, 1] = centre[1] + radius * np.sin(th)
    if dim == 3:
        pts[:, 2] = centre[2] + 0.2 * radius * np.sin(3 * th)
    rows, cols, vals = [], [], []
    off = 0
    for f in range(dim):
        coords = [edges[d][1:-1] if d == f else cen[d] for d in range(dim)]     # field f sits on its own faces
        nf = [len(c) for c in coords]
        for p in range(nb):
            i0 = [int(np.searchsorted(coords[d], pts[p, d])) for d in range(dim)]
            rng_ = [range(max(i0[d] - 2, 0), min(i0[d] + 2, nf[d])) for d in range(dim)]
            for idx in np.ndindex(*[len(r) for r in rng_]):
                ii = [rng_[d][idx[d]] for d in range(dim)]
                wgt = float(np.prod([roma(coords[d][ii[d]] - pts[p, d]) * hmin for d in range(dim)]))
                if wgt > 0.0:
                    lin = ii[0] + nf[0] * (ii[1] + (nf[1] * ii[2] if dim == 3 else 0))
                    rows.append(f * nb + p); cols.append(off + lin); vals.append(wgt)
        off += int(np.prod(nf))
    assert off == UN
    E = sp.csr_matrix((vals, (rows, cols)), shape=(dim * nb, UN))
    area = np.abs(Ds).max(axis=0).toarray().ravel()
    invh = np.abs(Gs).max(axis=1).toarray().ravel()
    W = sp.diags(invh / area)
    H = (W @ E.T).tocsr()
    DE = sp.vstack([Ds.tocsr(), E]).tocsr(); DE.sort_indices()
    GH = sp.hstack([Gs, -H]).tocsr(); GH.sort_indices()
    BNGH = GH.copy(); BNGH.data = dt * BNGH.data                      # BN = dt I (createbn.cpp:49-53): MatMatMult(BN, GH)
    A1 = orc.Csr.from_arrays(DE.shape[0], DE.shape[1], DE.indptr, DE.indices, DE.data)
    A2 = orc.Csr.from_arrays(BNGH.shape[0], BNGH.shape[1], BNGH.indptr, BNGH.indices, BNGH.data)
    M = orc.matmatmult(A1, A2).to_scipy().tocsr()
    M.sort_indices()
    nv = np.zeros(M.shape[0])
    nv[:pN] = 1.0 / np.sqrt(pN)
    return M, pN, nv


def ghosted_fields(shape, periodic, seed=3):
    """Random ghosted local arrays of the velocity fields (one ghost layer on every side; on periodic axes the ghost
    layers hold the wrap values, as DMGlobalToLocal leaves them), the interior packed vector, and the field sizes."""
    rng = np.random.default_rng(seed)
    dim = len(shape)
    nf = orc.field_sizes(shape, periodic)
    q, packed = [], []
    for f in range(dim):
        a = rng.standard_normal(tuple(m + 2 for m in reversed(nf[f])))
        for d in range(dim):
            if periodic[d]:
                ax = dim - 1 - d
                lo = [slice(None)] * dim; hi = [slice(None)] * dim; first = [slice(None)] * dim; last = [slice(None)] * dim
                lo[ax], hi[ax], first[ax], last[ax] = 0, -1, 1, -2
                a[tuple(lo)] = a[tuple(last)]
                a[tuple(hi)] = a[tuple(first)]
        q.append(a)
        packed.append(a[tuple([slice(1, -1)] * dim)].ravel())
    return q, np.concatenate(packed), nf

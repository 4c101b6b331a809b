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

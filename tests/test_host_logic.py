"""Host-side logic of the plugin mirror that needs no GPU: grid description from the YAML settings, DMDA slab
ownership, factory dispatch (linsolver.cpp:57-91)."""
import os

import numpy as np
import pytest

import petibm_b200 as pb
from petibm_b200.mesh import parse_periodic


def _node(bc=("DIRICHLET", "DIRICHLET", "DIRICHLET"), dim=3, stype="B200"):
    mesh = [{"direction": d, "start": -1.0, "subDomains": [{"end": 0.0, "cells": 6, "stretchRatio": 0.9},
                                                            {"end": 1.0, "cells": 5, "stretchRatio": 1.0}]}
            for d in "zyx"[3 - dim:]]   # order in the file is not guaranteed (parser.cpp:253-255)
    locs = ["xMinus", "xPlus", "yMinus", "yPlus", "zMinus", "zPlus"][: 2 * dim]
    bcs = [{"location": loc, **{f: [bc[i // 2], 0.0] for f in "uvw"[:dim]}} for i, loc in enumerate(locs)]
    return {"directory": "/tmp", "mesh": mesh, "flow": {"boundaryConditions": bcs},
            "parameters": {"dt": 0.005, "poissonSolver": {"type": stype, "config": "None"}}}


def test_grid_from_config():
    g = pb.Grid.from_config(_node())
    assert g.dim == 3 and g.n == (11, 11, 11) and g.dt == 0.005 and g.periodic == (False, False, False)
    assert abs(g.widths[0].sum() - 2.0) < 1e-14 and np.allclose(g.widths[0][6:], 0.2)
    g2 = pb.Grid.from_config(_node(dim=2))
    assert g2.dim == 2 and g2.n == (11, 11)


def test_periodic_detection_follows_checkPeriodicBC():
    assert parse_periodic(_node(bc=("PERIODIC", "DIRICHLET", "PERIODIC")), 3) == (True, False, True)
    bad = _node(bc=("PERIODIC", "DIRICHLET", "DIRICHLET"))
    bad["flow"]["boundaryConditions"][1]["u"] = ["DIRICHLET", 0.0]      # xPlus not periodic
    with pytest.raises(ValueError):
        parse_periodic(bad, 3)


@pytest.mark.parametrize("M,m", [(256, 8), (31, 2), (17, 5), (8, 8), (100, 3)])
def test_slab_range_is_the_dmda_ownership_rule(M, m):
    ranges = [pb.slab_range(M, r, m) for r in range(m)]
    assert ranges[0][0] == 0 and ranges[-1][1] == M
    sizes = [hi - lo for lo, hi in ranges]
    assert all(ranges[r][1] == ranges[r + 1][0] for r in range(m - 1))
    assert sizes == [M // m + (1 if r < M % m else 0) for r in range(m)]


def test_factory_dispatch():
    with pytest.raises(ValueError, match="PETSc KSP / AmgX"):
        pb.createLinSolver("poisson", _node(stype="CPU"))
    with pytest.raises(ValueError, match="Unrecognized"):
        pb.createLinSolver("poisson", _node(stype="FPGA"))
    # default type is "CPU" (linsolver.cpp:65)
    n = _node()
    del n["parameters"]["poissonSolver"]["type"]
    with pytest.raises(ValueError, match="PETSc KSP / AmgX"):
        pb.createLinSolver("poisson", n)


def test_mat_wrapper_keeps_nullspace_and_types():
    import scipy.sparse as sp

    A = sp.random(20, 20, 0.2, format="csr", random_state=1) + sp.identity(20, format="csr")
    M = pb.Mat.from_scipy(A).setNullSpace(True)
    assert M.nrows == 20 and M.indptr.dtype == np.int64 and M.indices.dtype == np.int32 and M.data.dtype == np.float64
    assert M.null_has_const and M.null_vecs is None


def test_vcycle_schedule_invariants_for_every_degree(tmp_path):
    """tests/cpp/mg_schedule_check.cpp: the shared V-cycle schedule (petibm_b200/csrc/mg_schedule.h) never lets a kernel
    write a buffer it reads, reads only what was written, and issues (levels - 1) * (2 m + 1) + m_coarse launches, for
    1..6 levels x smoothing degree 1..6 x coarse degree 1..6."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "mg_schedule_check")
    res = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "petibm_b200", "csrc"),
                          os.path.join(root, "tests", "cpp", "mg_schedule_check.cpp"), "-o", exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0 and "216 combinations, 0 bad" in run.stdout, run.stdout + run.stderr

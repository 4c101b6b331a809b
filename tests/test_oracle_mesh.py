"""Pin the oracle's grid restatement against the reference's own mesh golden vectors
(tests/mesh/cartesianmesh2d_dirichlet.cpp:171-284, cartesianmesh2d_yperiodic.cpp:180-290),
committed as tests/golden/*.json by tests/golden/make_mesh_golden.py."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc


@pytest.mark.parametrize("name", ["cartesianmesh2d_dirichlet", "cartesianmesh2d_yperiodic"])
def test_mesh_golden(golden_dir, name):
    g = json.load(open(os.path.join(golden_dir, name + ".json")))
    tol = g["tolerance"]
    widths, mins, maxs = [], [], []
    for ax in g["mesh"]:
        w = orc.axis_from_subdomains(ax["start"], ax["subDomains"])
        widths.append(w)
        mins.append(ax["start"])
        maxs.append(ax["subDomains"][-1]["end"])
    # pressure mesh (field 3) and vertex mesh (field 4)
    for d in range(2):
        np.testing.assert_allclose(widths[d], g["dLTrue"][3][d], rtol=0, atol=tol)
        np.testing.assert_allclose(orc.pressure_coord(widths[d], mins[d]), g["coordTrue"][3][d], rtol=0, atol=tol)
        np.testing.assert_allclose(orc.vertex_coord(widths[d], mins[d]), g["coordTrue"][4][d], rtol=0, atol=tol)
    # velocity meshes incl. ghosts
    UN = 0
    for comp in range(2):
        nvalid = []
        for d in range(2):
            n, dl, co = orc.velocity_axis(widths[d], mins[d], maxs[d], comp == d, g["periodic"][d])
            nvalid.append(n)
            np.testing.assert_allclose(dl, g["dLTrue"][comp][d], rtol=0, atol=tol)
            np.testing.assert_allclose(co, g["coordTrue"][comp][d], rtol=0, atol=tol)
        UN += nvalid[0] * nvalid[1]
    assert UN == g["UN"]
    assert len(widths[0]) * len(widths[1]) == g["pN"]


def test_stretch_grid_sums_to_length():
    # misc.h:148-163: geometric progression filling [bg, ed]
    w = orc.stretch_grid(0.1, 1.6, 4, 0.5)
    np.testing.assert_allclose(w, [0.8, 0.4, 0.2, 0.1], rtol=0, atol=1e-15)
    for r in (0.9, 1.01, 1.3, 2.0):
        w = orc.stretch_grid(-2.0, 3.0, 37, r)
        assert abs(w.sum() - 5.0) < 1e-12
        np.testing.assert_allclose(w[1:] / w[:-1], r, rtol=1e-14)


def test_uniform_threshold():
    # parser.cpp:350: |r-1| <= 1e-12 is uniform
    w = orc.axis_from_subdomains(0.0, [{"end": 1.0, "cells": 8, "stretchRatio": 1.0 + 5e-13}])
    assert np.all(w == 1.0 / 8)
    w = orc.axis_from_subdomains(0.0, [{"end": 1.0, "cells": 8, "stretchRatio": 1.0 + 1e-9}])
    assert not np.all(w == w[0])

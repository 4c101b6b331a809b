"""CPU-side checks of the drop-in boundary: libb200ls.so loads, exports every symbol include/b200ls.h
declares, and its host-only entry points (options parser, grid helper, error strings) behave like the
reference's (linsolverksp.cpp:48-69 options handling, parser.cpp:297-356 sub-domain arithmetic).
No compute entry point is called here (no GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle as orc
from petibm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200ls.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ls_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/b200ls.h but not exported"
    # the ctypes table (what a binding author copies) covers the header and nothing else
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_strings():
    L = _lib.lib()
    assert L.b200ls_version() == 100
    assert L.b200ls_error_string(0) == b"ok"
    for code in range(-7, 0):
        assert len(L.b200ls_error_string(code)) > 3
    assert L.b200ls_last_error(None) == b"null handle"


def test_no_cpu_fallback_without_a_device():
    L = _lib.lib()
    n = C.c_int(-1)
    rc = L.b200ls_device_count(C.byref(n))
    h = C.c_void_p()
    rc2 = L.b200ls_create(C.byref(h), 0)
    if n.value <= 0:
        assert rc == _lib.ERR_CUDA and rc2 == _lib.ERR_CUDA and not h.value
    else:  # running on a GPU box: creation works and nothing is solved here
        assert rc2 == _lib.OK
        L.b200ls_destroy(h)


def _parse(text, prefix="poisson_"):
    L = _lib.lib()
    o = _lib.Options()
    L.b200ls_default_options(C.byref(o))
    err = C.create_string_buffer(256)
    rc = L.b200ls_parse_options(text.encode(), prefix.encode(), C.byref(o), err, 256)
    return rc, o, err.value.decode()


def test_default_options_match_ksp_defaults():
    rc, o, _ = _parse("")
    assert rc == 0
    assert (o.ksp_type, o.pc_type, o.norm_type, o.max_it) == (_lib.KSP_CG, _lib.PC_NONE, _lib.NORM_PRECONDITIONED, 10000)
    assert (o.rtol, o.atol, o.divtol) == (1e-5, 1e-50, 1e4)


def test_parse_reference_style_options_file():
    text = """# Poisson solver: prefix `-poisson_`
-poisson_ksp_type cg
-poisson_ksp_atol 1.0E-06   # absolute
-poisson_ksp_rtol 0.0
-poisson_ksp_max_it 20000
-poisson_pc_type jacobi
-poisson_pc_jacobi_type diagonal
-velocity_ksp_type bcgs
-velocity_pc_type gamg
-options_left
"""
    rc, o, msg = _parse(text)
    assert rc == 0, msg
    assert (o.ksp_type, o.pc_type, o.max_it, o.atol, o.rtol) == (0, 1, 20000, 1e-6, 0.0)
    rc, o, msg = _parse(text, "velocity_")
    assert rc == _lib.ERR_UNSUPPORTED and "gamg" in msg      # velocity_pc_type gamg: loud failure, no fallback
    rc, o, msg = _parse("-velocity_ksp_type bcgs -velocity_pc_type jacobi -velocity_ksp_norm_type unpreconditioned", "velocity_")
    assert rc == 0 and (o.ksp_type, o.pc_type, o.norm_type) == (_lib.KSP_BCGS, _lib.PC_JACOBI, _lib.NORM_UNPRECONDITIONED)


@pytest.mark.parametrize("text", [
    "-poisson_pc_type gamg", "-poisson_pc_type hypre", "-poisson_ksp_type gmres", "-poisson_pc_gamg_type agg",
    "-poisson_mg_levels_ksp_type cg", "-poisson_ksp_initial_guess_nonzero", "-poisson_ksp_norm_type none"])
def test_unsupported_options_fail_loudly(text):
    rc, _, msg = _parse(text)
    assert rc == _lib.ERR_UNSUPPORTED and msg


def test_multigrid_options_use_the_pcmg_names():
    """pc_type mg is an extension of this backend (PetIBM's shipped configs precondition with GAMG / AmgX AMG, which is
    refused above); it takes PETSc's PCMG option names, and only the combination that is implemented."""
    rc, o, msg = _parse("-poisson_pc_type mg")
    assert rc == 0 and o.pc_type == _lib.PC_MG and (o.mg_levels, o.mg_smooth_its, o.mg_coarse_its) == (0, 2, 16), msg
    rc, o, msg = _parse("-poisson_ksp_type cg\n-poisson_pc_type mg\n-poisson_pc_mg_levels 4\n-poisson_mg_levels_ksp_type chebyshev\n"
                        "-poisson_mg_levels_ksp_max_it 3\n-poisson_mg_levels_pc_type jacobi\n-poisson_mg_coarse_ksp_max_it 24\n"
                        "-poisson_pc_mg_cycle_type v\n")
    assert rc == 0 and (o.pc_type, o.mg_levels, o.mg_smooth_its, o.mg_coarse_its) == (_lib.PC_MG, 4, 3, 24), msg
    for bad in ("-poisson_mg_levels_pc_type sor", "-poisson_pc_mg_cycle_type w", "-poisson_mg_coarse_ksp_type preonly",
                "-poisson_pc_mg_galerkin"):
        rc, _, msg = _parse(bad)
        assert rc == _lib.ERR_UNSUPPORTED and msg


def test_direct_solve_options_of_the_forces_system():
    """examples/decoupledibpm/*/config/forces_solver.info as shipped: preonly + lu (+ the factorisation package, which is
    PETSc's business); the two halves are only accepted together."""
    text = "# forces solver: prefix `-forces_`\n-forces_ksp_type preonly\n-forces_pc_type lu\n-forces_pc_factor_mat_solver_type superlu_dist\n"
    rc, o, msg = _parse(text, "forces_")
    assert rc == 0 and (o.ksp_type, o.pc_type) == (_lib.KSP_PREONLY, _lib.PC_LU), msg
    for bad in ("-forces_ksp_type preonly", "-forces_pc_type lu", "-forces_ksp_type preonly -forces_pc_type jacobi"):
        rc, _, msg = _parse(bad, "forces_")
        assert rc == _lib.ERR_UNSUPPORTED and msg


def test_example_option_files_parse():
    for name, prefix, want in (("poisson_solver.info", "poisson_", (_lib.KSP_CG, _lib.PC_MG)),
                               ("velocity_solver.info", "velocity_", (_lib.KSP_BCGS, _lib.PC_JACOBI)),
                               ("forces_solver.info", "forces_", (_lib.KSP_PREONLY, _lib.PC_LU))):
        rc, o, msg = _parse(open(os.path.join(ROOT, "examples", "config", name)).read(), prefix)
        assert rc == 0 and (o.ksp_type, o.pc_type) == want, msg
        assert name.startswith("forces") or (o.atol == 1e-6 and o.rtol == 0.0)


@pytest.mark.parametrize("text", ["-poisson_ksp_rtol abc", "-poisson_ksp_max_it", "stray -poisson_ksp_type cg"])
def test_malformed_options(text):
    rc, _, msg = _parse(text)
    assert rc == _lib.ERR_PARSE and msg


def test_axis_from_subdomains_matches_oracle_and_golden(golden_dir):
    import json

    from petibm_b200 import axis_from_subdomains

    sub = [{"end": 1.5, "cells": 4, "stretchRatio": 0.5}, {"end": 2.0, "cells": 5, "stretchRatio": 1.0},
           {"end": 5.0, "cells": 4, "stretchRatio": 2.0}]
    mine = axis_from_subdomains(0.0, sub)
    assert np.array_equal(mine, orc.axis_from_subdomains(0.0, sub))
    # the reference's own golden vector (tests/mesh/cartesianmesh2d_dirichlet.cpp:171-270)
    np.testing.assert_allclose(mine[:4], [0.8, 0.4, 0.2, 0.1], rtol=1e-14)
    g = json.load(open(os.path.join(golden_dir, "cartesianmesh2d_dirichlet.json")))
    assert g  # the oracle itself is pinned against this file in test_oracle_mesh.py
    rng = np.random.default_rng(3)
    for _ in range(20):
        nsub = int(rng.integers(1, 4))
        ends = np.cumsum(rng.uniform(0.3, 2.0, nsub))
        sub = [{"end": float(e), "cells": int(rng.integers(1, 40)), "stretchRatio": float(rng.choice([1.0, 0.9, 1.1, 1.03]))}
               for e in ends]
        assert np.array_equal(axis_from_subdomains(-0.25, sub), orc.axis_from_subdomains(-0.25, sub))


def test_the_header_is_plain_c_and_the_example_links(tmp_path):
    """include/b200ls.h compiles as C99 and examples/poisson_cabi.c (the binding calls from plain C) links against the
    library; without a CUDA device the program says so and exits with 2 -- there is no CPU path to fall back to."""
    import subprocess

    exe = str(tmp_path / "poisson_cabi")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                          os.path.join(ROOT, "examples", "poisson_cabi.c"), "-L", os.path.join(ROOT, "petibm_b200"), "-lb200ls",
                          "-Wl,-rpath," + os.path.join(ROOT, "petibm_b200"), "-lm", "-o", exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    L = _lib.lib()
    n = C.c_int(-1)
    have_gpu = L.b200ls_device_count(C.byref(n)) == 0 and n.value > 0
    if not have_gpu:                              # on a GPU box the program is a device run: not this suite's business
        run = subprocess.run([exe, "16"], capture_output=True, text=True, timeout=300)
        assert run.returncode == 2 and "no CPU path" in run.stderr

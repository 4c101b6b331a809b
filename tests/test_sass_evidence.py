"""What the compiler made of the hot kernels (no GPU needed: nvcc cross-compiles, cuobjdump disassembles).  A guard for the
claims of DESIGN.md / profiles/r02_sass_evidence.txt: the default fused SpMV moves its tiles with TMA (UTMALDG + mbarrier
transaction waits), the tiled line-coefficient kernels feed themselves through cp.async (LDGSTS + DEPBAR) and keep local memory
out of their plane loop, and every kernel is sm_100a code."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "petibm_b200", "libb200ls.so")


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not found")
    from petibm_b200 import build

    build.build()
    out = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name is not None and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            tok = line.split()
            funcs[name].append(tok[2] if tok[1].startswith("@") else tok[1])
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    return funcs, archs


def _count(ops, prefix):
    return sum(1 for o in ops if o.startswith(prefix))


def test_every_kernel_is_sm_100a_code(sass):
    funcs, archs = sass
    assert archs == {"sm_100a"} and len(funcs) > 150


def test_default_fused_spmv_uses_tma_and_mbarrier_transactions(sass):
    funcs, _ = sass
    names = [n for n in funcs if "k_spmv4" in n]
    assert names
    for n in names:
        ops = funcs[n]
        assert _count(ops, "UTMALDG") >= 2, n                       # r / p' (/ x) boxes
        assert _count(ops, "SYNCS.PHASECHK.TRANS64.TRYWAIT") >= 1   # mbarrier try_wait.parity


def test_tiled_line_coefficient_kernels_use_the_cp_async_queue(sass):
    funcs, _ = sass
    names = [n for n in funcs if "k_sep_tile_bcgs_spmv" in n or "k_sep_tile_cg_spmv" in n or "k_sep_tile_apply" in n]
    assert len(names) >= 8
    for n in names:
        ops = funcs[n]
        assert _count(ops, "LDGSTS") >= 8, n          # own cells + rim cells, every operand array
        assert _count(ops, "LDGDEPBAR") >= 3 and _count(ops, "DEPBAR") >= 2, n   # commit_group / wait_group
        assert _count(ops, "BAR.SYNC") >= 2, n
        # the launch geometry (SepTilePlan, indexed by the field found at run time) is copied to local memory once, before
        # the march: stores only in the prologue, a handful of loads, nothing per plane
        assert _count(ops, "LDL") <= 8, (n, _count(ops, "LDL"))

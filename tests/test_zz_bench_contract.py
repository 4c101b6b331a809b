"""bench.py keeps the driver's JSON contract: the reference arm on CPU (tiny size), the B200 arm on a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(args, timeout=600):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--size", "24", "24", "24", "--steps", "2", "--warmup", "1", "--cpu-iters", "15"])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"] == "poisson_cg_iterations_per_second" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["value"] > 0 and d["steps"] == 2 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "15 CG iterations" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


@pytest.mark.gpu
def test_b200_arm_line():
    d = _run(["--size", "96", "96", "96", "--iters", "40", "--steps", "2", "--warmup", "3", "--cpu-iters", "10", "--no-other-paths"])
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "strong"
    assert d["value"] > 0 and d["gpu_launches"] >= 2 * 2 * 40
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert r["launches_timed"] == 2 * 40 and r["algorithmic_bytes_per_launch"] == 48.0 * 96 ** 3
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == e["d2h_bytes_per_step"] == 8 * 96 ** 3 and 0 < e["value"] <= d["value"] * 1.25
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


@pytest.mark.gpu
def test_other_paths_block_reports_the_velocity_and_multigrid_solves():
    """The side measurements bench.py appends at N = 1 (velocity BiCGStab, multigrid time to solution): numbers, not
    errors, on small systems."""
    sys.path.insert(0, ROOT)
    import bench

    out = bench.other_paths_block(vel_size=(64, 24, 16), mg_size=(48, 48, 48), timeout_s=300)
    v, m = out["velocity_bicgstab_jacobi"], out["poisson_cg_mg_time_to_solution"]
    assert "error" not in v and "error" not in m, out
    assert v["iterations_per_s"] > 0 and v["row_per_thread_kernels_iterations_per_s"] > 0 and v["rows"] == 63 * 24 * 16 + 64 * 23 * 16 + 64 * 24 * 15
    assert m["iterations"] <= 20 and m["solve_ms"] > 0 and m["max_rel_error"] < 1e-6


def test_other_paths_block_survives_a_box_without_a_device():
    """No GPU here: both sub-processes fail, the block says so and bench.py would still print its line."""
    sys.path.insert(0, ROOT)
    import bench
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    out = bench.other_paths_block(vel_size=(8, 8, 8), mg_size=(8, 8, 8), timeout_s=300)
    assert set(out) == {"velocity_bicgstab_jacobi", "poisson_cg_mg_time_to_solution"}
    assert all("error" in v for v in out.values())

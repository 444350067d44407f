"""bench.py keeps the driver's JSON contract: one line on stdout with the agreed keys, for the CPU reference arm
(runs anywhere) and for the B200 arm (GPU tier, small cavity)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def run_bench(*flags):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--n", "6")
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "ipcs_timesteps_per_sec_p2p1_3d_cavity" and d["unit"] == "steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["dtype"] == "f64" and d["gpu_launches"] == 0
    assert d["scaling"] == "strong"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["openmp_spmv"]["jacobian_bsr3_GBs"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm steps the STATED mesh; what it printed is what it ran
    assert "UnitCubeMesh(6)" in d["config"]["workload"] and "DIFFERENT ALGORITHM" in d["config"]["algorithm"]
    assert d["steps"] == 2 and d["warmup"] == 1 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6
    assert d["checksum"]["steps_total"] == 3 and d["checksum"]["u_l2"] > 0


def test_reference_arm_respects_its_time_budget_and_says_so():
    d = run_bench("--impl", "reference", "--steps", "50", "--warmup", "3", "--n", "6", "--cpu-budget-s", "0")
    assert d["steps"] == 1 and d["steps_requested"] == 50  # stopped after the first timed step, and printed that


def test_reference_arm_uses_all_cores_under_torchrun_environment():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm sets its thread count itself."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--n", "4"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.strip().splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


@pytest.mark.gpu
def test_b200_arm_line(gpu_ctx):
    d = run_bench("--steps", "2", "--warmup", "3", "--n", "12", "--no-cpu")
    assert BASE_KEYS | {"roofline", "clocks", "iterations", "phase_ms"} <= set(d)
    assert d["n_gpus"] == 1 and d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    assert d["gpu_launches"] > 100
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["variants"]["momentum_rtol_1e-5"]["value"] > 0 and d["variants"]["semi_implicit"]["value"] > 0
    assert d["iterations"]["newton"] <= 10
    assert d["scaling"] == "strong" and d["e2e"]["steps"] == 2
    assert d["checksum"]["steps_total"] == 5 and d["checksum"]["u_l2"] > 0 and d["checksum"]["p_l2_mean_free"] > 0

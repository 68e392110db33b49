"""bench.py's output contract: one JSON line with the keys the driver reads.  The reference arm runs without a GPU
(CPU oracle + the reference's ALGLIB); the B200 arm is checked on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(args, timeout=600):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_runs_on_the_cpu_alone(have_ref):
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--per-gpu", "96"])
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "wbc_control_cycle_solves_per_sec" and d["unit"] == "solves/s"
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


@pytest.mark.gpu
def test_b200_arm_json_line():
    d = _run(["--steps", "3", "--warmup", "3", "--no-cpu-baseline"])
    assert BASE_KEYS | {"roofline", "clocks", "p50_ms"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["dtype"] == "f64" and d["scaling"] == "weak"
    assert d["value"] > 1e5 and d["e2e"]["value"] > 1e5 and d["gpu_launches"] == 6
    assert d["e2e"]["h2d_bytes_per_step"] == 4096 * (8 * 93 + 4) and d["e2e"]["d2h_bytes_per_step"] == 4096 * 8 * 18
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0 < r["frac"] < 1
    assert d["stats"]["solver_failures"] == 0
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
    assert "4096" in d["config"]["workload"]

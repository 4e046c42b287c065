"""bench.py prints ONE JSON line with the keys the driver reads, for both arms."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference: the CPU port of the reference algorithm on the host cores (no GPU needed)."""
    d = run_bench("--impl", "reference", "--n-data", "20000", "--steps", "2", "--warmup", "1")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "param_updates_per_sec" and d["unit"] == "param-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


@pytest.mark.gpu
def test_our_arm_line():
    d = run_bench("--steps", "2", "--warmup", "3", "--n-data", "200000", "--chains-per-gpu", "512", "--cpu-iters", "1",
                  "--e2e-iters", "2", "--no-sweeps")
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks"} <= set(d) and "impl" not in d
    assert d["metric"] == "param_updates_per_sec" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["value"] - 512 * 32 * 2 * 2 / (d["ms_per_step"] * 2e-3)) / d["value"] < 1e-6
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and 0 < r["frac"] < 1.2 and r["unit"] == "TFLOP/s"
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] >= 200000 * 8 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["gpu_launches"] == 2 * 4  # per step: 1 sweep kernel + 3 output transposes
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert d["clocks"]["sm_max_mhz"] is None or d["clocks"]["sm_max_mhz"] > 0

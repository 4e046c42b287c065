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
    d = run_bench("--impl", "reference", "--n-data", "20000", "--steps", "2", "--warmup", "1", "--configs", "K1,K2")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "param_updates_per_sec" and d["unit"] == "param-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # the shapes the host can finish are run WHOLE (same config as the GPU arm's configs.K1 / K2 / K3 entries)
    for k in ("K1", "K2"):
        c = d["configs"][k]
        assert c["same_config"] is True and c["value"] > 0 and c["kind"] == "port" and "burnin 1000 + samples 10000" in c["sample"]


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
    assert e["value"] > 0 and e["h2d_bytes_per_call"] >= 200000 * 8 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert e["steps_per_call"] == 2 and e["h2d_bytes_per_step"] == e["h2d_bytes_per_call"] // 2
    assert d["gpu_launches"] == 2 * 4  # per step: 1 sweep kernel + 3 output transposes
    assert 0 < r["nominal_frac"] <= r["frac"] * 1.1
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert d["clocks"]["sm_max_mhz"] is None or d["clocks"]["sm_max_mhz"] > 0


@pytest.mark.gpu
def test_named_configs_in_the_line():
    """The other named shapes ride in the same line (here: the two cheapest), each with value, roofline (incl. the
    fraction of the nominal peak) and the CPU port beside it; plus the one-process multi-device leg."""
    d = run_bench("--steps", "1", "--warmup", "3", "--n-data", "200000", "--chains-per-gpu", "256", "--cpu-iters", "1",
                  "--e2e-iters", "1", "--e2e-calls", "1", "--configs", "K1,K2", "--hbm-n", "0")
    assert set(d["configs"]) == {"K1", "K2"}
    for k, c in d["configs"].items():
        assert c["value"] > 0 and c["full_run"] is True and c["iterations_timed"] == 10999
        assert {"bound", "achieved", "peak", "frac", "nominal_frac"} <= set(c["roofline"])
        assert c["cpu_baseline"]["same_config"] is True and c["cpu_baseline"]["value"] > 0
    assert d["configs"]["K2"]["chains_total"] == 64 and d["configs"]["K2"]["rungs"] == 20 and d["configs"]["K2"]["free_params"] == 1
    m = d["multi_abi"]
    assert m["shards_bitwise_equal"] is True and m["value"] > 0 and d["diag"]["shards_bitwise_equal"] is True
    assert d["ess"]["value"] > 0 and len(d["ess"]["with_quantiles"]["mu"]) == 3

"""bench.py's reference arm runs on the host cores alone: check its JSON contract here (no GPU needed).
The b200 arm's line is produced and checked on the GPU box (tests/test_bench_contract_gpu below is marked gpu)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(*args, timeout=600):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_json_contract():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "cfg1_single_224")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"].startswith("crops/sec") and d["unit"] == "crops/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["config"]["workload"] == "cfg1_single_224" and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "crops" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None


@pytest.mark.gpu
def test_b200_arm_json_contract():
    d = _run("--steps", "5", "--warmup", "3", "--workload", "cfg2_multitask_256")
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks", "serial_value"} <= set(d)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["gpu_launches"] == 2 * 5                      # K1 + the fused heads step (K2 + K3 + finalize) per step
    assert d["e2e"]["h2d_bytes_per_step"] > 256 * 256 * 256 * 3 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] < d["value"] and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["clocks"]["sm_mhz"] is None or d["clocks"]["sm_mhz"] > 0

"""bench.py output contract: the reference arm (the reference's own processor on the CPU) on the tiny contract-test
workload must print ONE JSON line with every key the driver reads; the GPU arm is checked the same way on a B200."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(extra):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "tiny", *extra],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    d = _run(["--impl", "reference", "--steps", "2", "--warmup", "1"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "dit_denoise_step_ms" and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"] == d["ms_per_step"]          # measured, never extrapolated
    assert d["e2e"] == {"value": d["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "tiny", "--impl", "reference"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, RANK="1"))
    assert res.returncode == 0 and res.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    d = _run(["--steps", "2", "--warmup", "3"])
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks"} <= set(d)
    assert d["metric"] == "dit_denoise_step_ms" and d["dtype"] == "bf16" and d["n_gpus"] == 1
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] > 0
    lfl = d["like_for_like"]
    assert lfl["gpu_ms"] > 0 and lfl["gpu_e2e_ms"] > 0 and lfl["reference_cpu_ms"] == d["cpu_baseline"]["value"]
    assert lfl["parity"]["cosine"] >= 0.99
    rr = d["roofline"]
    assert rr["cross_attention"]["launches_per_step"] > 0 and rr["launches_per_step"] > 0
    fl = d["attn_flops_per_step"]
    assert abs(fl["library_counter"] - fl["closed_form"]) <= 1e-6 * fl["closed_form"]
    # both arms print the SAME config (the reference arm runs "on your arm's config")
    ref = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert ref["config"] == d["config"] and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])

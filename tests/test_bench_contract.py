"""bench.py's contract with the driver: one JSON line per run with the agreed keys.  The reference arm runs on the CPU
(here: the small cfg1 workload, one step); the B200 arms are GPU tests on reduced sizes of each workload kind."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(*flags, timeout=600):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    d = _run("--impl", "reference", "--workload", "cfg1_512_fp32", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "msdeformattn_fwd_bwd_sampled_points_per_sec" and d["unit"] == "points/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "cfg1_512_fp32" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "cfg1_512_fp32" and "model" not in d["config"]


@pytest.mark.gpu
def test_operator_arm_prints_the_contract_line(built_library):
    d = _run("--workload", "cfg1_512_fp32", "--steps", "2", "--warmup", "3", "--layers", "1", "--cpu-sample-batch", "1")
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks"} <= set(d) and "impl" not in d
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["dtype"] == "f32" and d["scaling"] == "weak"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1.5 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [
    ("--workload", "cfg3_train_step_1024", "--batch", "1", "--layers", "2", "--fused-layers"),
    ("--workload", "cfg4_decoder_step_300q", "--batch", "1", "--layers", "2", "--shared-value-proj"),
    ("--workload", "cfg4_decoder_step_300q", "--batch", "1", "--layers", "2", "--shared-value-proj", "--cuda-graph"),
])
def test_step_workloads_print_the_contract_line(built_library, flags):
    d = _run(*flags, "--steps", "2", "--warmup", "3", "--no-e2e")
    assert BASE_KEYS <= set(d)
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["scaling"] == "strong" and d["dtype"] == "bf16"
    assert d["config"]["workload"] == flags[1] and d["config"]["final_loss"] == d["config"]["final_loss"]   # not NaN

"""bench.py prints ONE JSON line with the keys the round driver reads.  The reference arm (CPU port of the reference
path) runs anywhere; the B200 arm needs the GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def run_bench(*argv, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], cwd=ROOT, capture_output=True, text=True,
                         timeout=timeout)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    line = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-images", "2")
    assert BASE_KEYS <= set(line)
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["gpu_launches"] == 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "BASELINE.json configs[2]" in line["config"]["workload"] and "model" not in line["config"]


@pytest.mark.gpu
def test_b200_arm_line():
    line = run_bench("--steps", "3", "--warmup", "3", "--batch", "32", "--no-cpu", "--no-extras")
    assert BASE_KEYS | {"roofline", "clocks", "cpu_baseline"} <= set(line)
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["dtype"] == "bf16" and line["scaling"] == "strong"
    assert line["value"] > 1000 and line["gpu_launches"] > 0
    roof = line["roofline"]
    assert roof["bound"] in ("hbm", "tensor") and roof["unit"] in ("GB/s", "TFLOP/s")
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9 and 0 < roof["frac"] < 1.2
    assert set(roof["classes"]) >= {"conv3x3_halo", "sep_fused_cout512", "gemm_conv1x1"}
    e2e = line["e2e"]
    assert e2e["value"] > 1000 and e2e["h2d_bytes_per_step"] == 32 * 368 * 368 * 3 and e2e["d2h_bytes_per_step"] > 0
    assert e2e["value"] != line["value"]                        # measured through the host-buffer API, not copied
    clk = line["clocks"]
    assert clk["sm_mhz"] and clk["sm_max_mhz"] and isinstance(clk["reasons"], list)
    assert line["config"]["e2e_gathered_on_rank0"]["images"] >= 3 * 32

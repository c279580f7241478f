"""bench.py's own arm on a GPU: one small invocation must print the contract's JSON line with every object the
round's measurement rules ask for (roofline, e2e with byte counts, pcie, cpu_baseline, clocks, launch count)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", *extra],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_default_line_is_cfg3_with_nested_secondaries():
    # (the aerial secondary is left out here: its 256 images of 8192 x 8192 take a minute to set up)
    line = run_bench("--batch", "64", "--cpu-sample", "2", "--also", "supervised")
    assert line["metric"] == "gaze_steps_per_sec" and line["unit"] == "gaze-steps/s" and line["higher_is_better"] is True
    assert "cfg3 reinforce" in line["config"]["workload"] and line["config"]["episodes_per_gpu"] == 64
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 3 and line["scaling"] == "weak"
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["gpu_launches"] == 2 * (1 + 1 + 1 + 2 * 20 + 1)
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and 0 < roof["frac"] < 1.2 and roof["launches_timed"] == 2 * 21
    assert roof["algorithmic_bytes_per_launch"] == 64 * 3 * 448 * 448 * 5
    assert roof["traffic"] is None  # the committed capture is for the config batch (1024), not for this one
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert e2e["h2d_bytes_per_step"] % (3 * 448 * 448) == 0  # whole uint8 tiles, each first visit once
    assert line["pcie"]["bound"] == "pcie" and 0 < line["pcie"]["frac"] <= 1.5 and line["pcie"]["peak"] > 1
    cpu = line["cpu_baseline"]
    assert cpu["kind"] in ("reference", "port") and cpu["cores"] >= 1 and cpu["value"] > 0
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert set(line["also"]) == {"cfg2"}
    for key, sub in line["also"].items():
        assert sub["value"] > 0 and sub["roofline"]["frac"] > 0 and sub["gpu_launches"] > 0 and sub["e2e"]["value"] > 0
        assert key[-1] in sub["config"]["workload"][:4]


def test_other_primary_workload_without_secondaries():
    line = run_bench("--workload", "supervised", "--batch", "32", "--also", "none", "--no-cpu-baseline", "--no-e2e")
    assert "cfg2 supervised" in line["config"]["workload"] and line["also"] == {} and line["e2e"] is None
    assert line["roofline"]["kernel"].startswith("gather_xform_kernel") and line["cpu_baseline"] is None

"""bench.py's reference arm runs on host cores only: one tiny invocation must print the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference(*extra):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "2", *extra], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_prints_one_contract_line():
    line = run_reference()
    assert line["impl"] == "reference" and line["metric"] == "gaze_steps_per_sec" and line["unit"] == "gaze-steps/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 0
    # the unmodified reference (baseline/_ref, installed by __graft_entry__.build()) when it is there, else the port
    from baseline import ref_env

    assert line["cpu_baseline"]["kind"] == ("reference" if ref_env.available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # default workload = BASELINE configs[2], the config the metric is quoted on; same `config` object as our arm
    assert line["gpu_launches"] == 0 and "cfg3 reinforce" in line["config"]["workload"]
    assert line["config"]["episodes_per_gpu"] == 1024


def test_reference_arm_supervised_workload():
    line = run_reference("--workload", "supervised")
    assert "cfg2 supervised" in line["config"]["workload"] and line["value"] > 0


def test_reference_and_port_agree_on_the_bench_workloads():
    """The two CPU arms of bench.py (unmodified reference / oracle port) count the same gaze-steps."""
    import bench
    from baseline import ref_env

    if not ref_env.available():
        import pytest

        pytest.skip("no reference copy in this checkout")
    wl = bench.SupervisedWorkload(2, 0, "cpu", "f32")
    units_ref, _ = wl.cpu_sample(2, 5)
    real = bench.reference_kind
    try:
        bench.reference_kind = lambda: "port"
        units_port, _ = wl.cpu_sample(2, 5)
    finally:
        bench.reference_kind = real
    assert units_ref == units_port > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "2"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""

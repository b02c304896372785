"""bench.py's reference arm (CPU only): the JSON contract the driver reads, and that it runs the SAME configuration the GPU
arm declares (the driver compares the two `config` objects), with an explicit OpenMP team size."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_config():
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = "1"                      # what torchrun exports: the arm must override it
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    sys.path.insert(0, ROOT)
    import bench
    assert line["impl"] == "reference" and line["metric"] == "icp_iterations_per_s" and line["higher_is_better"] is True
    assert line["config"] == bench.workload_config(1)                      # same_config
    assert "50 fixed ICP iterations per step" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] == line["e2e"]["value"] and line["value"] > 0
    usable = bench.host_cpu_budget()[0]
    assert 1 <= cb["cores"] <= usable and (cb["cores"] > 1 or usable == 1)  # not torchrun's single thread
    assert cb["step_s"]["min"] <= cb["step_s"]["median"] <= cb["step_s"]["max"]
    assert cb["reference_faithful_8_threads"]["cores"] == min(8, usable)
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""

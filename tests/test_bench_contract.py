"""bench.py's contract where it can be checked without a GPU: the reference arm prints one JSON line with the agreed
keys (it times the CPU oracle port on the host cores), and the GPU arm refuses to run without CUDA instead of falling
back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT, env=env)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"].startswith("audio-sec") and line["unit"] == "audio-s/s"
    assert line["value"] > 0 and line["steps"] == 1 and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and line["dtype"] == "f32"
    assert line["config"]["workload"].startswith("cfg2") and line["config"]["sample_clips_per_step"] >= 1
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] == (os.cpu_count() or 1) and cpu["value"] == line["value"] and cpu["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env_rank = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env_rank)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the GPU arm runs for real in the driver's bench step")
    r = _run("--steps", "1", "--warmup", "0", timeout=180)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)

"""bench.py contract pieces that run without a GPU: the reference arm (`--impl reference`) prints ONE JSON line with the keys
the driver reads, on the same metric / unit / config as the GPU arm; the GPU arm refuses to run without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line(ob, demo_index):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--reads-per-step", "256"],
                       capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = [l for l in r.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mbases/s classified" and d["unit"] == "Mbases/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly(ob, demo_index):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == b""


def test_gpu_arm_fails_loudly_without_a_device(ob, demo_index):
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--reads-per-step", "64"], capture_output=True, timeout=600)
    assert r.returncode != 0 and b"no CPU fallback" in r.stderr

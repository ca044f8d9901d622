"""CPU: the bench.py contract that does not need a GPU -- the reference arm prints one JSON line with
the contract's keys (it times the C port of the reference algorithm on the host cores), and the GPU
arm refuses to run without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "20", "--warmup", "3")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "agent-steps/sec" and d["unit"] == "agent-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == 20 and d["warmup"] == 3 and d["value"] > 0
    assert d["dtype"] == "f64" and d["config"]["workload"] == "config3" and d["config"]["n_agents"] == 10
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_only_rank0_works():
    """N > 1: rank 0 alone runs and prints; the other ranks exit 0 without work."""
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "5", "--warmup", "1",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run("--steps", "5", "--warmup", "1", "--no-cpu", "--no-e2e")
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")], "no bench line may be printed without a GPU"

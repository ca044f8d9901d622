"""CPU: the bench.py contract that does not need a GPU -- the reference arm prints one JSON line with
the contract's keys (it times the unmodified NumPy reference when it is available on the box, else
the C port of its algorithm, on the host cores), and the GPU arm refuses to run without a device
instead of falling back to anything."""
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


def _line(r):
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def _check_contract(d, steps, warmup):
    assert d["impl"] == "reference" and d["metric"] == "agent-steps/sec" and d["unit"] == "agent-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == steps and d["warmup"] == warmup and d["value"] > 0
    assert d["dtype"] == "f64" and d["config"]["workload"] == "config3" and d["config"]["n_agents"] == 10
    assert d["config"]["env_steps_per_step"] == 200
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline_port"]["kind"] == "port" and d["cpu_baseline_port"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_port_prints_the_contract_line():
    d = _line(_run("--impl", "reference", "--ref-kind", "port", "--steps", "3", "--warmup", "1", "--cpu-budget", "1"))
    _check_contract(d, 3, 1)
    assert d["cpu_baseline"]["kind"] == "port"


def test_reference_arm_times_the_unmodified_reference_when_available():
    from oracle import ref_cpu
    if not ref_cpu.available():
        pytest.skip("neither /root/reference nor the staged oracle/_ref is on this box")
    d = _line(_run("--impl", "reference", "--steps", "3", "--warmup", "1", "--cpu-budget", "1",
                   "--ref-step-seconds", "0.5"))
    _check_contract(d, 3, 1)
    assert d["cpu_baseline"]["kind"] == "reference"
    # the Python reference is orders of magnitude slower than the C port of its algorithm
    assert d["value"] < d["cpu_baseline_port"]["value"]


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


@pytest.mark.gpu
def test_gpu_arm_small_steps_is_a_warm_kernel_measurement():
    """The driver's command line shape (`--steps 20 --warmup 5`): the timed region must hold warm,
    same-shape rollout launches, so a short run and a longer run agree and the per-launch kernel time
    explains the whole step (round-1 regression: a cold-shape single launch timed the allocator)."""
    a = _line(_run("--steps", "4", "--warmup", "3", "--no-cpu", "--no-e2e", "--no-extra"))
    b = _line(_run("--steps", "20", "--warmup", "5", "--no-cpu", "--no-e2e", "--no-extra"))
    for d in (a, b):
        assert d["config"]["env_steps_per_step"] == 200 and d["gpu_launches"] == 3 * d["steps"]   # reset+observe, rollout, reduce
        assert d["roofline"]["launches_timed"] == d["steps"]
        # the rollout launch is (nearly) the whole step
        assert d["roofline"]["avg_launch_ms"] > 0.7 * d["ms_per_step"]
        assert d["agg_check"]["steps_ok"]
    assert abs(a["value"] - b["value"]) / b["value"] < 0.25
    assert b["value"] > 5e9          # far above anything an allocator-bound timed region could show

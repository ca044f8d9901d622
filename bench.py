#!/usr/bin/env python
"""bench.py -- agent-steps/s of the drone_env.step() hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload config3] [--dtype f64|f32]

A "step" is one environment step of the whole batch (E environments x n agents), i.e.
one pass of the hot path.  Steps are executed as fused rollout launches of one episode
(<= 200 steps, reference drone_env.py:30) each; every episode starts from a fresh random lattice
reset drawn on the device (ds_reset_random), reads its action stream from HBM and writes the full
per-step outputs (state, rewards, observations, neighbour lists, collision counts,
finished flags) to HBM trajectory buffers.

Default workload: BASELINE config 3 (n=10, E=4096, Delta=1.0, grid [5,5]) -- the
configuration the north-star target is quoted on.  Multi-GPU: environments are sharded
E per rank (weak scaling), one NCCL all-reduce of the episode aggregates per episode.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {   # BASELINE.json configs[1..4]
    "config2": dict(n=5, E=4096, grid=[5, 5], delta=1.0),
    "config3": dict(n=10, E=4096, grid=[5, 5], delta=1.0),
    "config4": dict(n=32, E=8192, grid=[32, 32], delta=2.5),
    "config5": dict(n=128, E=1024, grid=[64, 64], delta=1.0),
    # HBM-resident point for the ncu capture (working set per step >> L2)
    "hbm": dict(n=10, E=1 << 20, grid=[5, 5], delta=1.0, episode=20),
}
EPISODE = 200
N_ACTIONS = 16
K_CLOSEST = 2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# dram__bytes_read.sum + dram__bytes_write.sum of ONE rollout launch from the committed
# `ncu --set full` captures (profiles/r01/*_summary.md), keyed by (workload, dtype): bytes per launch.
NCU_TRAFFIC = {
    ("config3", "f64"): 133.5e6 + 836.9e6,     # profiles/r01/final_prof_config3_summary.md
    ("hbm", "f64"): 3738.1e6 + 23873.4e6,      # profiles/r01/final_prof_hbm_summary.md
    ("config5", "f64"): 424.1e6 + 2788.5e6,    # profiles/r01/e_prof_config5_summary.md
}


def bytes_per_agent_step(rb, n, k=K_CLOSEST, cols=2, mode="rollout"):
    """Algorithmic HBM bytes per agent-step.
    rollout: read action 2*rb; write pos 2*rb, vel 2*rb, r rb, true_r rb, z (k+1)*cols*rb,
    Ni 4(k+1); per env-step 5 B (ncoll i32 + finished u8).  step mode adds the pos read."""
    b = 2 * rb + 2 * rb + 2 * rb + rb + rb + (k + 1) * cols * rb + 4 * (k + 1) + 5.0 / n
    if mode == "step":
        b += 2 * rb
    return b


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML every ~2 ms DURING the timed region
    (nvidia-smi's own start-up time is longer than a short timed region)."""
    HW_SLOWDOWN, SW_POWER_CAP, SW_THERMAL, HW_THERMAL = 0x8, 0x4, 0x20, 0x40

    def __init__(self, index=0):
        self.index, self.rows, self.thread, self.stop_flag, self.h = index, [], None, False, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # torch device ordinals follow CUDA_VISIBLE_DEVICES; NVML sees every device of the box
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except ValueError:
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def _pump(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.h is None:
            return
        self.stop_flag = False
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.h is None or self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        sm = [r[0] for r in self.rows]
        bits = 0
        for _, rs in self.rows:
            bits |= int(rs)
        reasons = [name for name, bit in (("hw_slowdown", self.HW_SLOWDOWN), ("hw_thermal_slowdown", self.HW_THERMAL),
                                          ("sw_thermal_slowdown", self.SW_THERMAL), ("sw_power_cap", self.SW_POWER_CAP))
                   if bits & bit]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_sm),
                "samples": len(sm), "reasons": reasons}


# ----------------------------------------------------------------------------- CPU legs
def cpu_oracle_rate(wl, budget_s, nthreads, steps_per_call=20):
    """C-oracle port of the reference algorithm on the host cores: agent-steps/s on a bounded
    sample of the same workload (same n, grid, Delta, action set)."""
    from oracle import c_oracle
    from scalable_collision_avoidance_rl_b200 import formation
    n, grid = wl["n"], wl["grid"]
    E = min(wl["E"], 4096)
    rng = np.random.default_rng(1234)
    xF = formation.end_formation("O", n, grid)
    ds = formation.safety_distances(xF, np.ones(n) * 0.1)
    deltas, _ = formation.clip_deltas(np.ones(n) * wl["delta"], ds)
    env = c_oracle.OracleEnv(E, n, xF, ds, deltas, None, K_CLOSEST, True, c_oracle.default_params(0.2),
                             nthreads=nthreads)
    tab = formation.unit_action_table(N_ACTIONS)
    env.set_state(formation.sample_start_batched(E, n, grid, rng))
    act = tab[rng.integers(0, N_ACTIONS, (steps_per_call, E, n))]
    env.rollout(act[:2], record=False)          # warm-up
    done_steps, t0 = 0, time.perf_counter()
    while True:
        env.set_state(formation.sample_start_batched(E, n, grid, rng), None, 0)
        t1 = time.perf_counter()
        env.rollout(act, record=True)
        dt_call = time.perf_counter() - t1
        done_steps += steps_per_call
        if time.perf_counter() - t0 + dt_call > budget_s:
            break
    # time only the rollout calls: recompute with a clean loop of the same count
    reps = max(1, done_steps // steps_per_call)
    t1 = time.perf_counter()
    for _ in range(reps):
        env.t[...] = 0
        env.rollout(act, record=True)
    el = time.perf_counter() - t1
    rate = reps * steps_per_call * E * n / el
    return rate, f"{reps * steps_per_call} steps x {E} envs x {n} agents ({el:.1f} s)", E


def run_reference(args, wl):
    """--impl reference: the reference algorithm's CPU implementation (C oracle port; the Python
    reference itself cannot travel to the GPU box) on all host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle
    from scalable_collision_avoidance_rl_b200 import formation
    cores = os.cpu_count() or 1
    n, grid = wl["n"], wl["grid"]
    E = min(wl["E"], 4096)
    K, W = args.steps, args.warmup
    rng = np.random.default_rng(1234)
    xF = formation.end_formation("O", n, grid)
    ds = formation.safety_distances(xF, np.ones(n) * 0.1)
    deltas, _ = formation.clip_deltas(np.ones(n) * wl["delta"], ds)
    env = c_oracle.OracleEnv(E, n, xF, ds, deltas, None, K_CLOSEST, True, c_oracle.default_params(0.2),
                             nthreads=cores)
    tab = formation.unit_action_table(N_ACTIONS)
    chunk = 50
    act = tab[rng.integers(0, N_ACTIONS, (chunk, E, n))]

    def run(steps):
        left = steps
        while left > 0:
            c = min(chunk, left)
            if int(env.t[0]) + c > EPISODE:
                env.set_state(formation.sample_start_batched(E, n, grid, rng), None, 0)
            env.rollout(act[:c], record=True)
            left -= c

    env.set_state(formation.sample_start_batched(E, n, grid, rng), None, 0)
    run(W)
    t0 = time.perf_counter()
    run(K)
    el = time.perf_counter() - t0
    val = K * E * n / el
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": val, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": el / K * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "n_agents": n, "n_envs": E, "grid": grid,
                   "delta": wl["delta"], "k_closest": K_CLOSEST,
                   "note": "reference algorithm as its C port (oracle/drone_oracle.c) on host cores; "
                           "the Python reference cannot run on the GPU box"},
        "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{K} steps x {E} envs x {n} agents"},
        "e2e": {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    from scalable_collision_avoidance_rl_b200 import dist as dsdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, E, grid = wl["n"], wl["E"], wl["grid"]
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    rb = 8 if args.dtype == "f64" else 4
    K, W = args.steps, args.warmup
    T = min(args.episode_steps or wl.get("episode", EPISODE), K)
    rec = ("pos", "vel", "reward", "true_reward", "obs", "ncoll", "finished")

    env = BatchedDrones(E, n, grid, "O", K_CLOSEST, np.ones(n) * wl["delta"], True, dtype=dtype,
                        device=dev, seed=1234 + rank, warn=False)
    env.log_mode = args.log_mode
    rng = np.random.default_rng(1234 + rank)
    tab = formation.unit_action_table(N_ACTIONS)
    # synthetic inputs resident in HBM before the timed region: action stream + episode starts
    n_ep_bufs = 2
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    idx = torch.randint(0, N_ACTIONS, (n_ep_bufs, T, E, n), device=dev, generator=gen, dtype=torch.uint8)
    actions = torch.as_tensor(tab, dtype=dtype, device=dev)[idx.long()]   # [bufs,T,E,n,2] Real
    del idx
    starts = torch.as_tensor(formation.sample_start_batched(n_ep_bufs * E, n, grid, rng)
                             .reshape(n_ep_bufs, E, n, 2), dtype=dtype, device=dev)
    out = {}
    launches = [0]
    per_launch_ms = []
    stream = torch.cuda.current_stream(dev)

    ep_base = [0]

    def episode(ep, steps, timed):
        # env.reset() on the device (drone_env.py:98-102,193-210): fresh distinct lattice nodes per
        # environment (Philox stream = episode number), zero velocity, t = 0, initial observation
        b = ep % n_ep_bufs
        env.reset_random(seed=1234 + rank, stream=ep_base[0] + ep); launches[0] += 2
        if timed:
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        env.rollout(actions=actions[b][:steps], record=rec, out=out); launches[0] += 1
        if timed:
            e1.record(stream)
            per_launch_ms.append((e0, e1, steps))
        agg = env.episode_aggregates(); launches[0] += 1     # device-side reduce, no host sync
        if world > 1:
            dsdist.allreduce_episode_aggregates(agg)
        return agg

    def run(total, timed):
        ep, left = 0, total
        while left > 0:
            s = min(T, left)
            episode(ep, s, timed)
            left -= s; ep += 1
        ep_base[0] += ep

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    run(max(W, 3), False)
    barrier()
    launches[0] = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
    barrier()
    g0.record(stream)
    run(K, True)
    g1.record(stream)
    barrier()
    ms = g0.elapsed_time(g1)
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = world * E * n * K / (ms * 1e-3)

    # dominant kernel: the fused rollout launch
    kms = [a.elapsed_time(b) for a, b, _ in per_launch_ms]
    ksteps = [s for _, _, s in per_launch_ms]
    bpas = bytes_per_agent_step(rb, n)
    alg_bytes_launch = bpas * E * n * float(np.mean(ksteps))
    avg_ms = float(np.mean(kms))
    hbm_peak, peak_src = peaks()
    achieved = alg_bytes_launch / (avg_ms * 1e-3) / 1e9

    # end to end through the public host API: pinned host action stream in, pinned host
    # trajectories of the reference's 6-tuple out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        h_act = torch.empty((T, E, n, 2), dtype=dtype, pin_memory=True)
        h_act.copy_(actions[0].cpu())
        h_start = starts.cpu().numpy()
        hout = {}
        e2e_steps = min(K, 2 * T)

        def e2e_run(total):
            ep, left = 0, total
            while left > 0:
                s = min(T, left)
                env.reset(h_start[ep % n_ep_bufs])               # H2D of the start state + observe
                env.rollout_host(actions=h_act[:s], record=rec, out=hout)
                left -= s; ep += 1

        e2e_run(T)
        barrier()
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        barrier()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        A = E * n
        zc = (K_CLOSEST + 1) * 2
        e2e = {"value": world * E * n * e2e_steps / float(el.item()), "unit": "agent-steps/s",
               "h2d_bytes_per_step": A * 2 * rb,
               # pos, r, true_r, z, Ni, ncoll, finished; vel (= the action, drone_env.py:238) is returned
               # as a view of the host action stream and does not cross PCIe
               "d2h_bytes_per_step": A * (2 * rb + rb + rb + zc * rb + 4 * (K_CLOSEST + 1)) + E * 5,
               "steps": e2e_steps,
               "api": "BatchedDrones.rollout_host -> ds_rollout_host (pinned host buffers, "
                      "H2D/D2H pipelined against the kernel)",
               "note": "state, observations, rewards, neighbour lists, collision counts and finished flags of "
                       "every step come back; the velocity columns equal the supplied actions "
                       "(drone_env.py:238) and are returned as a view of the host action stream"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        rate, sample, _ = cpu_oracle_rate(wl, args.cpu_budget, cores)
        cpu = {"value": rate, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample}
    line = {
        "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world,
        "steps": K, "warmup": max(W, 3), "ms_per_step": ms / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": args.workload, "n_agents": n, "n_envs_per_gpu": E, "grid": grid,
                   "delta": wl["delta"], "k_closest": K_CLOSEST, "simplify_zstate": True,
                   "actions": f"uniform over {N_ACTIONS} unit directions, streamed from HBM",
                   "episode_steps": T, "reset": "fresh lattice start per episode (ds_reset_random on the device)",
                   "l2": f"inputs larger than L2: {alg_bytes_launch / 1e6:.0f} MB streamed per launch",
                   "log_mode": args.log_mode},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak,
                     "traffic": NCU_TRAFFIC.get((args.workload, args.dtype)) if T == wl.get("episode", EPISODE) else None,
                     "traffic_unit": "DRAM bytes per launch (ncu capture, profiles/r01)",
                     "algorithmic_bytes_per_launch": alg_bytes_launch, "peak_source": peak_src,
                     "kernel": "ds::rollout_kernel", "bytes_per_agent_step": bpas,
                     "avg_launch_ms": avg_ms, "launches_timed": len(kms),
                     "note": "the kernel is issue/latency bound, not HBM bound (DESIGN.md section 5)"},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches[0], "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=600)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--log-mode", type=int, default=0)
    ap.add_argument("--episode-steps", type=int, default=0, help="steps per rollout launch (default 200)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()

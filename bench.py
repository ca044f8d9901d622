#!/usr/bin/env python
"""bench.py -- agent-steps/s of the drone_env.step() hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload config3] [--dtype f64|f32]

A bench "step" is ONE PASS OF THE HOT PATH OVER ONE BATCH: the per-episode step() loop of the
reference (train_problem.py:82-107; drone_env.py:214-258 called max_time_steps = 200 times,
drone_env.py:30) for the whole batch of E environments x n agents, i.e. E * n * 200 agent-steps.
On the GPU that is: a fresh random lattice reset drawn on the device (ds_reset_random), ONE fused
rollout launch that reads the episode's action stream from HBM and writes every per-step output
of step() (state, both rewards, observations, neighbour lists, collision counts, finished flags)
to HBM trajectory buffers, and the device-side reduction of the episode aggregates (+ one NCCL
all-reduce of 5 doubles when N > 1).  `--steps K` times exactly K such episodes; `value` is
agent-steps/s = N * E * n * 200 * K / time.  One step streams ~1 GB (config 3, f64): larger than
L2, and consecutive steps alternate between two action buffers.

Default workload: BASELINE config 3 (n=10, E=4096, Delta=1.0, grid [5,5]) -- the configuration
the north-star target is quoted on.  Multi-GPU: the headline keeps E environments per rank (weak
scaling); BASELINE configs 4 and 5 are added to the same line under `extra`, strong-sharded over
the N ranks with dist.shard_envs.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
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
    # HBM-resident point for the ncu capture (working set per env-step >> L2)
    "hbm": dict(n=10, E=1 << 20, grid=[5, 5], delta=1.0, episode=20),
}
EPISODE = 200          # drone_env.py:30 max_time_steps
N_ACTIONS = 16         # DiscreteSoftmaxNN(n_actions=16), SAC_agents.py:143
K_CLOSEST = 2
RECORD = ("pos", "vel", "reward", "true_reward", "obs", "ncoll", "finished")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, dtype):
    """DRAM bytes per rollout launch from the committed `ncu --set full` capture of this kernel
    build (profiles/ncu_traffic.json: keyed by workload:dtype, each entry names its capture and the
    commit of the kernel source it was taken from)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            ent = json.load(f).get(f"{workload}:{dtype}")
    except Exception:
        ent = None
    if not ent:
        return None, None
    return float(ent["dram_bytes_per_launch"]), ent


def bytes_per_agent_step(rb, n, k=K_CLOSEST, cols=2):
    """Algorithmic HBM bytes per agent-step of a recorded rollout: read action
    2*rb; write pos 2*rb, vel 2*rb, r rb, true_r rb, z (k+1)*cols*rb, Ni 4(k+1); per env-step 5 B
    (ncoll i32 + finished u8)."""
    return 2 * rb + 2 * rb + 2 * rb + rb + rb + (k + 1) * cols * rb + 4 * (k + 1) + 5.0 / n


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
def _port_env(wl, E, nthreads):
    from oracle import c_oracle
    from scalable_collision_avoidance_rl_b200 import formation
    n, grid = wl["n"], wl["grid"]
    xF = formation.end_formation("O", n, grid)
    ds = formation.safety_distances(xF, np.ones(n) * 0.1)
    deltas, _ = formation.clip_deltas(np.ones(n) * wl["delta"], ds)
    return c_oracle.OracleEnv(E, n, xF, ds, deltas, None, K_CLOSEST, True, c_oracle.default_params(0.2),
                              nthreads=nthreads)


def cpu_port_rate(wl, budget_s, nthreads, T=EPISODE):
    """C port of the reference algorithm (oracle/drone_oracle.c) on `nthreads` host threads:
    agent-steps/s over whole recorded episodes of a bounded sample of the workload's environments."""
    from scalable_collision_avoidance_rl_b200 import formation
    n, grid = wl["n"], wl["grid"]
    E = min(wl["E"], 1024)
    T = min(T, wl.get("episode", EPISODE))
    rng = np.random.default_rng(1234)
    env = _port_env(wl, E, nthreads)
    tab = formation.unit_action_table(N_ACTIONS)
    act = tab[rng.integers(0, N_ACTIONS, (T, E, n))]
    env.set_state(formation.sample_start_batched(E, n, grid, rng), None, 0)
    env.rollout(act[:5], record=False)          # warm-up
    eps, el = 0, 0.0
    while True:
        env.set_state(formation.sample_start_batched(E, n, grid, rng), None, 0)
        t1 = time.perf_counter()
        env.rollout(act, record=True)
        dt_call = time.perf_counter() - t1
        el += dt_call; eps += 1
        if el + dt_call > budget_s:
            break
    return eps * T * E * n / el, f"{eps} episodes x {T} steps x {E} envs x {n} agents ({el:.1f} s in rollout calls)"


def cpu_reference_rates(wl, budget_s):
    """The UNMODIFIED NumPy reference (oracle/_ref staged copy on the GPU box), 1 process and nproc
    processes.  None when the reference is not available on this box."""
    try:
        from oracle import ref_cpu
        if not ref_cpu.available():
            return None
        cores = os.cpu_count() or 1
        n = wl["n"]
        T = EPISODE if n <= 32 else 20
        r1, s1 = ref_cpu.rate(n, wl["grid"], wl["delta"], 1, budget_s, T)
        rN, sN = ref_cpu.rate(n, wl["grid"], wl["delta"], cores, budget_s, T)
        return {"kind": "reference", "unit": "agent-steps/s", "value": rN, "cores": cores, "sample": sN,
                "one_process": {"value": r1, "cores": 1, "sample": s1},
                "note": "unmodified reference drone_env.drones.step (NumPy, single threaded), one environment per "
                        "process, same action set, reset every episode"}
    except Exception as ex:   # the baseline must never take the GPU line down
        return {"kind": "reference", "unavailable": f"{type(ex).__name__}: {ex}"}


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation of the path on all host cores.
    The unmodified NumPy reference when it is available on this box (build container:
    /root/reference; GPU box: the staged copy oracle/_ref), else the C port of its algorithm.
    A step = one episode (<= 200 env-steps) of a BOUNDED sample of the workload's environments."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n, grid = wl["n"], wl["grid"]
    K, W = args.steps, args.warmup
    T = min(wl.get("episode", EPISODE), EPISODE)
    use_ref = False
    if args.ref_kind in ("auto", "reference"):
        try:
            from oracle import ref_cpu
            use_ref = ref_cpu.available()
        except Exception:
            use_ref = False
        if args.ref_kind == "reference" and not use_ref:
            print(json.dumps({"impl": "reference", "unavailable": "unmodified reference not staged on this box"}))
            return
    port_rate, port_sample = cpu_port_rate(wl, min(6.0, args.cpu_budget), cores)
    port = {"kind": "port", "value": port_rate, "unit": "agent-steps/s", "cores": cores, "sample": port_sample}
    if use_ref:
        from oracle import ref_cpu
        Ts = T if n <= 32 else 20
        pool = ref_cpu.Pool(cores)
        try:
            pool.run(n, grid, wl["delta"], 1, min(Ts, 10))                    # imports
            _, w1 = pool.run(n, grid, wl["delta"], 1, Ts, seed0=7)            # calibration
            # size a step so that W + K steps end within ~2.5 minutes
            m = int(max(1, min(64, min(args.ref_step_seconds, 150.0 / max(1, K + W)) / max(w1, 1e-3))))
            for i in range(W):
                pool.run(n, grid, wl["delta"], m, Ts, seed0=100 + i)
            steps_done, t0 = 0, time.perf_counter()
            for i in range(K):
                s, _ = pool.run(n, grid, wl["delta"], m, Ts, seed0=1000 + i)
                steps_done += s
            el = time.perf_counter() - t0
        finally:
            pool.close()
        val = steps_done / el
        E_s = cores * m
        kind, sample = "reference", (f"each step: one {Ts}-step episode of {E_s} environments ({cores} processes x {m}), "
                                     f"unmodified NumPy reference; {K} steps in {el:.1f} s")
        note = "UNMODIFIED reference drone_env.drones.step (NumPy) on all host cores, one environment per process"
    else:
        from scalable_collision_avoidance_rl_b200 import formation
        E_s = min(wl["E"], 1024)
        rng = np.random.default_rng(1234)
        env = _port_env(wl, E_s, cores)
        tab = formation.unit_action_table(N_ACTIONS)
        act = tab[rng.integers(0, N_ACTIONS, (T, E_s, n))]

        def run(k):
            for _ in range(k):
                env.set_state(formation.sample_start_batched(E_s, n, grid, rng), None, 0)
                env.rollout(act, record=True)
        run(W)
        t0 = time.perf_counter()
        run(K)
        el = time.perf_counter() - t0
        val = K * T * E_s * n / el
        kind, sample = "port", f"each step: one {T}-step episode of {E_s} environments; {K} steps in {el:.1f} s"
        note = ("reference algorithm as its C port (oracle/drone_oracle.c) on all host cores; the unmodified "
                "reference is not staged on this box")
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": val, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": el / K * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "n_agents": n, "n_envs": wl["E"], "grid": grid,
                   "delta": wl["delta"], "k_closest": K_CLOSEST, "simplify_zstate": True,
                   "env_steps_per_step": T, "sample_envs_per_step": E_s, "note": note},
        "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "cpu_baseline_port": port,
        "e2e": {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def pin_to_gpu_numa_node(index):
    """Run this rank on the CPU cores NVML reports as local to its GPU, so that the pinned host
    buffers of the end-to-end leg (allocated afterwards, first-touch) sit on the GPU's NUMA node and
    the D2H copies do not cross the socket interconnect.  Returns what was done, for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = index
        if vis:
            try:
                phys = int(vis.split(",")[index])
            except ValueError:
                phys = index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and 64 * w + b < ncpu]
        orig = sorted(os.sched_getaffinity(0))
        allowed = sorted(set(cpus) & set(orig))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"cpus": len(allowed), "of": ncpu, "_orig": orig}
        return {"cpus": 0, "of": ncpu, "note": "no local cores reported"}
    except Exception as ex:
        return {"error": f"{type(ex).__name__}: {ex}"}


class GpuLeg:
    """W warm-up + K timed episodes of one workload on this rank's GPU.  Everything the timed
    region touches (action streams, trajectory buffers, the library's launch plan) exists before
    the first event: the warm-up runs launches of the IDENTICAL shape."""

    def __init__(self, wl, E_local, dtype_name, dev, rank, world, log_mode=0, T=None, seed=1234, strong=False):
        import torch
        from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
        self.torch = torch
        self.wl, self.E, self.n, self.rank, self.world, self.dev = wl, E_local, wl["n"], rank, world, dev
        self.strong = strong                                 # E_local is this rank's shard of wl["E"]
        self.T = T or wl.get("episode", EPISODE)
        self.dtype = torch.float64 if dtype_name == "f64" else torch.float32
        self.rb = 8 if dtype_name == "f64" else 4
        n, E, T = self.n, self.E, self.T
        self.env = BatchedDrones(E, n, wl["grid"], "O", K_CLOSEST, np.ones(n) * wl["delta"], True, dtype=self.dtype,
                                 device=dev, seed=seed + rank, warn=False)
        self.env.log_mode = log_mode
        tab = formation.unit_action_table(N_ACTIONS)
        self.n_bufs = 2
        gen = torch.Generator(device=dev); gen.manual_seed(seed + rank)
        self.actions = []
        ttab = torch.as_tensor(tab, dtype=self.dtype, device=dev)
        for _ in range(self.n_bufs):   # synthetic inputs resident in HBM before the timed region
            idx = torch.randint(0, N_ACTIONS, (T, E, n), device=dev, generator=gen, dtype=torch.uint8)
            self.actions.append(ttab[idx.long()].contiguous())   # [T,E,n,2] Real
            del idx
        self.out = {}
        # ds_reset_random: draws + start observation in one launch for n <= 32, k = 2 (else two launches)
        self.reset_launches = 1 if (n <= 32 and K_CLOSEST == 2 and os.environ.get("DS_RESET_FUSED", "1") != "0") else 2
        self.agg_rows = torch.zeros((1024, 5), dtype=torch.float64, device=dev)   # one row per episode of a run
        self.seed = seed + rank
        self.ep = 0
        self.launches = 0
        self.stream = torch.cuda.current_stream(dev)
        self.events = []

    def episode(self, timed):
        """env.reset() on the device (drone_env.py:98-102,193-210: fresh distinct lattice nodes per
        environment, zero velocity, t = 0, initial observation), one fused rollout of the episode,
        device-side reduction of the episode aggregates (+ the all-reduce)."""
        torch = self.torch
        b = self.ep % self.n_bufs
        self.env.reset_random(seed=self.seed, stream=self.ep); self.launches += self.reset_launches
        if timed:
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
        self.env.rollout(actions=self.actions[b], record=RECORD, out=self.out); self.launches += 1
        if timed:
            e1.record(self.stream)
            self.events.append((e0, e1))
        # device-side reduce of this episode's aggregates into its row of the run's table, no host
        # sync; the rows are all-reduced together once, at the end of the timed episodes (run())
        row = self.agg_rows[self.ep % self.agg_rows.shape[0]]
        agg = self.env.episode_aggregates(out=row); self.launches += 1
        self.ep += 1
        return agg

    def allreduce_rows(self, first_ep, count):
        """ONE all-reduce (NCCL) of the per-episode aggregate rows of `count` episodes: the single
        exchange of the sharded path ("a single all-reduce of episode returns at the end"), enqueued
        behind the last rollout, inside the timed region."""
        if self.world > 1 and count > 0:
            import torch.distributed as dist
            R = self.agg_rows.shape[0]
            lo = first_ep % R
            if lo + count <= R:
                dist.all_reduce(self.agg_rows[lo:lo + count], op=dist.ReduceOp.SUM)
            else:                                            # (wrapped: more than 1024 episodes in flight)
                dist.all_reduce(self.agg_rows, op=dist.ReduceOp.SUM)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def run(self, K, W, sampler=None):
        """Returns (ms of the K timed episodes: max over ranks, per-launch kernel ms list)."""
        torch = self.torch
        for _ in range(max(W, 3)):
            self.episode(False)
        self.barrier()
        self.launches = 0
        self.events = []
        if sampler is not None:
            sampler.start()
        g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
        self.barrier()
        g0.record(self.stream)
        ep0 = self.ep
        for _ in range(K):
            self.episode(True)
        self.allreduce_rows(ep0, K)
        g1.record(self.stream)
        self.timed_rows = (ep0, K)
        self.barrier()
        ms = g0.elapsed_time(g1)
        tms = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        kms = [a.elapsed_time(b) for a, b in self.events]
        return float(tms.item()), kms

    def agg_check(self):
        """One more episode WITHOUT the in-line all-reduce: the all-reduced 5-vector must equal the
        sum of the per-rank vectors gathered separately, and the step count N * E * T (no environment
        of a random walk reaches its goal formation)."""
        torch = self.torch
        self.barrier()
        # the rows the timed run reduced: every episode must account for every environment-step of every rank
        ep0, K = getattr(self, "timed_rows", (0, 0))
        R = self.agg_rows.shape[0]
        rows = self.agg_rows[[(ep0 + q) % R for q in range(min(K, R))]].cpu().numpy() if K else np.zeros((0, 5))
        tot_envs = float(self.wl["E"]) if self.strong else float(self.world * self.E)
        timed_ok = bool(len(rows) == 0 or (np.abs(rows[:, 3] - tot_envs * self.T) < 0.5).all() and (rows[:, 4] == tot_envs).all())
        agg = self.episode(False).clone()
        want_steps = float(self.E * self.T)
        ok_local = abs(float(agg[3].item()) - want_steps) < 0.5 and float(agg[4].item()) == float(self.E)
        res = {"steps_per_rank_ok": bool(ok_local), "timed_episodes_reduced_ok": timed_ok}
        if self.world > 1:
            import torch.distributed as dist
            from scalable_collision_avoidance_rl_b200 import dist as dsdist
            parts = [torch.empty_like(agg) for _ in range(self.world)]
            dist.all_gather(parts, agg)
            red = dsdist.allreduce_episode_aggregates(agg.clone())
            tot = torch.stack(parts, 0).sum(0)
            rel = float(((red - tot).abs() / tot.abs().clamp_min(1e-300)).max().item())
            res.update({"allreduce_vs_gather_sum_max_rel": rel, "allreduce_equals_gather_sum": bool(rel < 1e-12),
                        "steps": float(red[3].item()), "steps_expected": float(self.world) * want_steps,
                        "n_envs": float(red[4].item())})
            # N * E_local * T only holds with equal shards; compare against the gathered truth
            res["steps_ok"] = bool(abs(float(red[3].item()) - float(tot[3].item())) < 0.5)
        else:
            res.update({"steps": float(agg[3].item()), "steps_expected": want_steps, "n_envs": float(agg[4].item()),
                        "steps_ok": bool(ok_local)})
        return res

    def free(self):
        self.actions = None
        self.out = None
        self.env = None
        self.torch.cuda.empty_cache()


def roofline_of(kms, bpas, units_per_launch, workload, dtype_name, kernel):
    hbm_peak, peak_src = peaks()
    alg = bpas * units_per_launch
    avg_ms, med_ms = float(np.mean(kms)), float(np.median(kms))
    achieved = alg / (avg_ms * 1e-3) / 1e9
    traffic, ent = ncu_traffic(workload, dtype_name)
    return {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu --set full)",
            "traffic_source": None if ent is None else {k: ent.get(k) for k in ("capture", "kernel_commit")},
            "algorithmic_bytes_per_launch": alg, "peak_source": peak_src, "kernel": kernel,
            "bytes_per_agent_step": bpas, "avg_launch_ms": avg_ms, "median_launch_ms": med_ms,
            "min_launch_ms": float(np.min(kms)), "launches_timed": len(kms),
            "note": "timed with CUDA events around every rollout launch of the timed region, on the launch stream"}


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from scalable_collision_avoidance_rl_b200 import dist as dsdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = pin_to_gpu_numa_node(local)     # before any pinned allocation: host buffers land next to the GPU
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, E, grid = wl["n"], wl["E"], wl["grid"]
    K, W = args.steps, args.warmup
    T = args.episode_steps or wl.get("episode", EPISODE)
    rb = 8 if args.dtype == "f64" else 4

    leg = GpuLeg(wl, E, args.dtype, dev, rank, world, args.log_mode, T)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, kms = leg.run(K, W, sampler)
    clocks = sampler.stop() if rank == 0 else None
    launches = leg.launches
    value = world * E * n * T * K / (ms * 1e-3)
    bpas = bytes_per_agent_step(rb, n)
    roof = roofline_of(kms, bpas, E * n * T, args.workload, args.dtype, leg.env.rollout_kernel)
    agg_check = leg.agg_check()

    # end to end through the public host API: pinned host action stream in, pinned host
    # trajectories of the reference's 6-tuple out, copies inside the timed region
    e2e = e2e_full = None
    if not args.no_e2e:
        from scalable_collision_avoidance_rl_b200 import formation
        env = leg.env
        h_act = [torch.empty((T, E, n, 2), dtype=leg.dtype, pin_memory=True) for _ in range(2)]
        for b in range(2):
            h_act[b].copy_(leg.actions[b])
        rng = np.random.default_rng(99 + rank)
        h_start = formation.sample_start_batched(2 * E, n, grid, rng).reshape(2, E, n, 2)
        e2e_steps = max(10, min(K, 50)) if T >= EPISODE else max(2, min(K, 10))
        A = E * n
        zc = (K_CLOSEST + 1) * 2

        def e2e_leg(compact):
            hout = {}

            def e2e_run(total):
                for ep in range(total):
                    env.reset(h_start[ep % 2])                       # H2D of the start state + observe
                    env.rollout_host(actions=h_act[ep % 2], record=RECORD, out=hout, compact=compact)

            e2e_run(2)
            leg.barrier()
            t0 = time.perf_counter()
            e2e_run(e2e_steps)
            leg.barrier()
            el_local = time.perf_counter() - t0
            el = torch.tensor([el_local], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(el, op=dist.ReduceOp.MAX)
            zb, nib = (4, 1) if compact else (rb, 4)
            d2h = A * (2 * rb + rb + rb + zc * zb + nib * (K_CLOSEST + 1)) + E * 5
            h2d = A * 2 * rb
            return {"value": world * E * n * T * e2e_steps / float(el.item()), "unit": "agent-steps/s",
                    # per bench step (= one episode of T env-steps)
                    "h2d_bytes_per_step": h2d * T + A * 2 * 8,
                    # pos, r, true_r, z, Ni, ncoll, finished; vel (= the action, drone_env.py:238) is returned
                    # as a view of the host action stream and does not cross PCIe
                    "d2h_bytes_per_step": d2h * T + E * 4 * 8,
                    "steps": e2e_steps, "env_steps_per_step": T,
                    "pcie_gbs_this_rank": {"d2h": d2h * T * e2e_steps / el_local / 1e9,
                                           "h2d": h2d * T * e2e_steps / el_local / 1e9},
                    "api": "BatchedDrones.reset + BatchedDrones.rollout_host -> ds_reset + ds_rollout_host (pinned host "
                           "buffers, H2D/D2H pipelined against the kernel)",
                    "observations": ("z as float32 (what the reference's actors cast it to, utils.py:305) and Ni as u8 "
                                     "on the host side (DS_HOST_COMPACT_OBS)") if compact else "z and Ni as the kernel writes them (float64 / int32)",
                    "note": "state, observations, rewards, neighbour lists, collision counts and finished flags of "
                            "every step come back; the velocity columns equal the supplied actions "
                            "(drone_env.py:238) and are returned as a view of the host action stream"}

        e2e = e2e_leg(not args.e2e_full)
        if not args.e2e_full:
            e2e_full = e2e_leg(False)
    leg.free()

    # BASELINE configs 4 and 5, strong-sharded over the ranks of this run
    extra = {}
    if not args.no_extra and args.workload == "config3":
        for name in ("config4", "config5"):
            w2 = WORKLOADS[name]
            lo, hi = dsdist.shard_envs(w2["E"], rank, world)
            l2 = GpuLeg(w2, hi - lo, args.dtype, dev, rank, world, args.log_mode, EPISODE, seed=4321, strong=True)
            K2 = 10
            ms2, kms2 = l2.run(K2, 3)
            chk = l2.agg_check()
            ctas = None
            ex = {"n_agents": w2["n"], "n_envs_total": w2["E"], "n_envs_this_rank": hi - lo, "scaling": "strong",
                  "grid": w2["grid"], "delta": w2["delta"], "steps": K2, "env_steps_per_step": EPISODE,
                  "value": w2["E"] * w2["n"] * EPISODE * K2 / (ms2 * 1e-3), "unit": "agent-steps/s",
                  "ms_per_step": ms2 / K2,
                  "roofline": roofline_of(kms2, bytes_per_agent_step(rb, w2["n"]), (hi - lo) * w2["n"] * EPISODE, name,
                                          args.dtype, l2.env.rollout_kernel),
                  "agg_check": chk}
            if name == "config5":
                ex["note"] = (f"{hi - lo} environments = {hi - lo} CTAs per GPU" +
                              (" < 148 SMs: sub-wave, latency bound" if hi - lo < 148 else "") +
                              "; n = 128 is CUDA-core bound (pass 1 over 16 256 pairs per frame), not HBM bound")
            extra[name] = ex
            l2.free()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if isinstance(numa, dict) and "_orig" in numa:          # the CPU baselines use every core of the box
        os.sched_setaffinity(0, numa.pop("_orig"))
    cpu = cpu_ref = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        rate, sample = cpu_port_rate(wl, args.cpu_budget, cores)
        cpu = {"value": rate, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample}
        cpu_ref = cpu_reference_rates(wl, args.cpu_budget)
    line = {
        "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world,
        "steps": K, "warmup": max(W, 3), "ms_per_step": ms / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": args.workload, "n_agents": n, "n_envs": E, "n_envs_per_gpu": E, "grid": grid,
                   "delta": wl["delta"], "k_closest": K_CLOSEST, "simplify_zstate": True,
                   "step": f"one episode = {T} env-steps of the whole batch in one fused rollout launch",
                   "env_steps_per_step": T, "agent_steps_per_step": E * n * T,
                   "actions": f"uniform over {N_ACTIONS} unit directions, streamed from HBM",
                   "reset": "fresh lattice start per episode (ds_reset_random on the device)",
                   "l2": (f"{roof['algorithmic_bytes_per_launch'] / 1e6:.0f} MB streamed per step "
                          f"({'larger' if roof['algorithmic_bytes_per_launch'] > 126e6 else 'SMALLER'} than the 126 MB L2), "
                          "two alternating action buffers; no flush"),
                   "log_mode": args.log_mode},
        "roofline": roof, "agg_check": agg_check,
        "cpu_baseline": cpu, "cpu_baseline_reference": cpu_ref, "e2e": e2e, "e2e_full_precision_obs": e2e_full,
        "gpu_launches": launches, "numa_pinning": numa,
        "clocks": clocks, "extra": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50, help="timed episodes (one fused rollout launch each)")
    ap.add_argument("--warmup", type=int, default=5, help="warm-up episodes (at least 3 are run)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--log-mode", type=int, default=0)
    ap.add_argument("--episode-steps", type=int, default=0, help="env-steps per episode / rollout launch (default 200)")
    ap.add_argument("--envs", type=int, default=0, help="override the workload's environment count (tuning runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the config4 / config5 legs")
    ap.add_argument("--e2e-full", action="store_true", help="e2e with float64 z / int32 Ni on the host side only")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "reference", "port"])
    ap.add_argument("--ref-step-seconds", type=float, default=6.0,
                    help="--impl reference: upper bound on the CPU time of one step's sample")
    ap.add_argument("--cpu-budget", type=float, default=10.0)
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.envs:
        wl["E"] = args.envs
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()

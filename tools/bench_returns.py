#!/usr/bin/env python
"""Throughput of ds_returns (Monte-Carlo returns + Delta-neighbourhood advantage sums)
on trajectories produced by ds_rollout: agent-steps/s, fraction of the HBM roofline, and the C
oracle's rate on the host cores beside it.  Prints one JSON line per workload.
Usage: python tools/bench_returns.py [config3|hbm ...]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation   # noqa: E402

WL = {"config3": dict(n=10, E=4096, grid=[5, 5], T=200, reps=20),
      "config4": dict(n=32, E=8192, grid=[32, 32], T=200, reps=5),
      "hbm": dict(n=10, E=1 << 20, grid=[5, 5], T=20, reps=5)}


def main(names):
    dev = torch.device("cuda", 0)
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    src = "fallback (B200_PROFILING.md)"
    if os.path.exists(p):
        peak, src = float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    for name in names:
        w = WL[name]
        n, E, T = w["n"], w["E"], w["T"]
        env = BatchedDrones(E, n, w["grid"], "O", 2, np.ones(n), True, device=dev, seed=1, warn=False)
        tab = torch.as_tensor(formation.unit_action_table(16), device=dev)
        gen = torch.Generator(device=dev); gen.manual_seed(1)
        act = tab[torch.randint(0, 16, (T, E, n), device=dev, generator=gen)]
        ro = env.rollout(actions=act, record=("reward", "obs", "finished"))
        del act
        base = torch.randn((T, E, n), dtype=torch.float64, device=dev)
        out = {}
        for _ in range(3):
            env.returns(ro["reward"], ro["Ni"], ro["finished"], 0.99, base, out=out)
        torch.cuda.synchronize()
        ms = []
        for _ in range(w["reps"]):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); env.returns(ro["reward"], ro["Ni"], ro["finished"], 0.99, base, out=out); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        m = float(np.median(ms))
        bpas = 8 + 12 + 1.0 / n + 8 + 8 + 8 + 1      # r, Ni, finished, baseline | returns, advantage, count
        units = n * E * T
        cpu = None
        if name == "config3":
            from oracle import c_oracle
            r, Ni, fin, b = (x.cpu().numpy() for x in (ro["reward"], ro["Ni"], ro["finished"], base))
            t0 = time.perf_counter(); c_oracle.returns(r, Ni, fin, 0.99, b); el = time.perf_counter() - t0
            cpu = {"value": units / el, "unit": "agent-steps/s", "cores": 1, "kind": "port",
                   "sample": f"{T} steps x {E} envs x {n} agents ({el:.2f} s)"}
        print(json.dumps({"metric": "agent-steps/sec (returns + advantage gather)", "value": units / (m * 1e-3),
                          "unit": "agent-steps/s", "dtype": "f64", "ms_per_launch": m,
                          "config": {"workload": name, "n_agents": n, "n_envs": E, "T": T,
                                     "l2": f"{bpas * units / 1e6:.0f} MB streamed per launch"},
                          "roofline": {"bound": "hbm", "achieved": bpas * units / (m * 1e-3) / 1e9, "peak": peak,
                                       "unit": "GB/s", "frac": bpas * units / (m * 1e-3) / 1e9 / peak,
                                       "bytes_per_agent_step": bpas, "peak_source": src, "kernel": "ds::returns_kernel"},
                          "cpu_baseline": cpu}), flush=True)
        del env, ro, base, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main(sys.argv[1:] or ["config3", "config4", "hbm"])

#!/usr/bin/env python
"""Tuning aid: rollout-kernel time on BASELINE config 3 as a function of what is recorded."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation

def main():
    n, E, T = int(os.environ.get("N", 10)), int(os.environ.get("E", 4096)), 200
    grid = [5, 5] if n <= 10 else ([32, 32] if n <= 32 else [64, 64])
    delta = 1.0 if n != 32 else 2.5
    dev = torch.device("cuda", 0)
    env = BatchedDrones(E, n, grid, "O", 2, np.ones(n) * delta, True, device=dev, seed=1, warn=False)
    env.log_mode = int(os.environ.get("LOG_MODE", 0))
    tab = torch.as_tensor(formation.unit_action_table(16), device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    idx = torch.randint(0, 16, (T, E, n), device=dev, generator=gen)
    actions = tab[idx]
    start = env.pos.clone()
    sets = [(), ("reward",), ("pos", "vel"), ("reward", "true_reward", "ncoll", "finished"),
            ("pos", "vel", "reward", "true_reward", "ncoll", "finished"), ("obs",),
            ("pos", "vel", "reward", "true_reward", "obs", "ncoll", "finished")]
    for rec in sets:
        out = {}
        ms = []
        for it in range(6):
            env.pos.copy_(start); env.vel.zero_(); env.internal_t.zero_(); env.done.zero_(); env.agg.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); env.rollout(actions=actions, record=rec, out=out); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        m = float(np.median(ms[2:]))
        print(f"record={','.join(rec) or '-':55s} {m:7.3f} ms  {n * E * T / m * 1e3:.3e} agent-steps/s", flush=True)

if __name__ == "__main__":
    main()

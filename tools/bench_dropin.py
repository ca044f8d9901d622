#!/usr/bin/env python
"""Latency of the drop-in class for ONE environment (BASELINE config 1: n = 5, E = 1 -- the
reference's own use, train_problem.py:82-104 / benchmark_agent.py): `drone_env.drones.step(actions)`
with host arrays in and out (set_state + ds_step_host + unpacking), per step.  The reference's step
takes 0.35 ms at n = 5, 1.09 ms at n = 10 and 8.2 ms at n = 32 on one host core."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drone_env

def main():
    for n in (5, 10, 32):
        grid = [5, 5] if n <= 10 else [32, 32]
        env = drone_env.drones(n, 0, grid, "O", 2, np.ones(n), True)
        rng = np.random.default_rng(0)
        tab = np.array([[np.cos(a / 16 * 2 * np.pi), np.sin(a / 16 * 2 * np.pi)] for a in range(16)])
        steps = 600
        idx = rng.integers(0, 16, (steps, n))
        for t in range(50):
            env.step([tab[idx[t, i]] for i in range(n)])
        env.reset()
        t0 = time.perf_counter()
        done = 0
        for t in range(steps):
            _, z, r, nc, fin, tr = env.step([tab[idx[t, i]] for i in range(n)])
            done += 1
            if fin: env.reset()
        dt = time.perf_counter() - t0
        print(json.dumps({"metric": "drop-in drones.step latency, one environment", "n_agents": n,
                          "us_per_step": dt / done * 1e6, "agent_steps_per_s": n * done / dt}), flush=True)

if __name__ == "__main__":
    main()

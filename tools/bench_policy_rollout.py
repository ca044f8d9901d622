#!/usr/bin/env python
"""Closed-loop episodes with the actors on the device (ds_rollout_policy, SURVEY 8f row 1 feeding the
step): agent-steps/s of the loop `actions = agents.forward(z, Ni); env.step(actions)`
(train_problem.py:82-104) for E environments, 200 steps per episode, device reset per episode;
enqueued call by call and as one captured CUDA graph per episode."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation

def main():
    A, T = 16, 200
    for n, E, grid in ((10, 4096, [5, 5]), (5, 64, [5, 5])):
        rng = np.random.default_rng(0)
        u = lambda shape, fan: rng.uniform(-1, 1, shape).astype(np.float32) / np.float32(np.sqrt(fan))
        W = (u((n, 300, 6), 6), u((n, 300), 6), u((n, 300, 300), 300), u((n, 300), 300), u((n, A, 300), 300), u((n, A), 300))
        env = BatchedDrones(E, n, grid, "O", 2, np.ones(n), True, seed=1, warn=False)
        env.load_policy(*W, formation.unit_action_table(A))
        seed = torch.tensor([1], dtype=torch.uint64, device=env.device)
        out = {}
        rec = ("reward", "true_reward", "obs", "ncoll", "finished", "action_idx")

        def episode(ep):
            env.reset_random(seed=7, stream=ep)
            env.rollout_policy(T, stream0=0, record=rec, out=out, seed_tensor=seed)

        def timed(fn, reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for r in range(reps): fn(r)
            e1.record(); torch.cuda.synchronize()
            steps = float(out["agg"][:, 3].sum().item())          # executed env-steps of the LAST episode
            return e0.elapsed_time(e1) / reps, steps

        for ep in range(2): episode(ep)
        ms_eager, steps = timed(episode, 5)
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                episode(0)
        def replay(r):
            seed.fill_(r + 10); g.replay()
        replay(0); torch.cuda.synchronize()
        ms_graph, steps_g = timed(replay, 5)
        print(json.dumps({"metric": "agent-steps/s, closed loop with per-agent actors (6-300-300-16) on the device",
                          "n_agents": n, "n_envs": E, "episode_steps": T,
                          "ms_per_episode_enqueued": ms_eager, "ms_per_episode_graph": ms_graph,
                          "executed_env_steps_last_episode": steps_g,
                          "value_enqueued": n * steps / (ms_eager * 1e-3), "value_graph": n * steps_g / (ms_graph * 1e-3)}), flush=True)
        del env, g

if __name__ == "__main__":
    main()

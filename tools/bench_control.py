#!/usr/bin/env python
"""Closed-loop throughput: ds_rollout_control (one launch per episode) vs one
ds_step_control launch per step, BASELINE config 3 batch, gradient controller.  Tuning aid."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scalable_collision_avoidance_rl_b200 import BatchedDrones

n, E, T = 10, 4096, 200
env = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=1, warn=False)
start = env.pos.clone()
rec = ("pos", "vel", "reward", "true_reward", "obs", "ncoll", "finished")
res = {}
for mode in ("fused", "rollout_control", "per-step"):
    ms, steps = [], 0
    for it in range(5):
        env.pos.copy_(start); env.vel.zero_(); env.internal_t.zero_(); env.done.zero_(); env.agg.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode in ("fused", "rollout_control"):
            os.environ["DS_CTRL_FUSED"] = "1" if mode == "fused" else "0"   # fused: one launch; default: T one-step launches
            out = env.rollout_control(T, "gradient", record=rec)
        else:
            for t in range(T):
                env.step_control("gradient")
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        steps = float(env.agg[:, 3].sum()) if mode != "per-step" else float(env.internal_t.sum())
    m = float(np.median(ms[1:]))
    res[mode] = {"ms_per_episode": m, "executed_agent_steps": steps * n, "agent_steps_per_s": steps * n / (m * 1e-3)}
print(json.dumps({"workload": "config3 closed loop, gradient_control, all outputs recorded (fused)", **res}))

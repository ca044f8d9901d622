#!/usr/bin/env python
"""Throughput of ds_policy_forward (batched per-agent actor MLPs on tcgen05):
agent-forwards/s, TFLOP/s of the 300 x 300 layer (useful flops = 2*300*300 per forward; the 3xTF32
split issues 3 MMAs per product), against torch fp32 matmul on the same GPU and on the host cores."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation

def main():
    n, A = 10, 16
    rng = np.random.default_rng(0)
    u = lambda shape, fan: rng.uniform(-1, 1, shape).astype(np.float32) / np.float32(np.sqrt(fan))
    W = (u((n, 300, 6), 6), u((n, 300), 6), u((n, 300, 300), 300), u((n, 300), 300), u((n, A, 300), 300), u((n, A), 300))
    for E in (4096, 1 << 16):
        env = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=1, warn=False)
        env.load_policy(*W, formation.unit_action_table(A))
        for it in range(3):
            env.policy_forward(seed=1, stream=it)
        torch.cuda.synchronize()
        ms = []
        for it in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); env.policy_forward(seed=1, stream=it, want_probs=False); e1.record()
            torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        m = float(np.median(ms))
        flops = 2.0 * E * n * (6 * 300 + 300 * 300 + 300 * A)
        # torch fp32 on the same GPU (TF32 off): bmm over agents
        z = env.z_states.reshape(E, n, 6).float().transpose(0, 1).contiguous()      # [n,E,6]
        tW = [torch.as_tensor(w, device=env.device) for w in W]
        torch.backends.cuda.matmul.allow_tf32 = False
        def tfwd():
            h = torch.relu(torch.baddbmm(tW[1][:, None], z, tW[0].transpose(1, 2)))
            h = torch.relu(torch.baddbmm(tW[3][:, None], h, tW[2].transpose(1, 2)))
            return torch.softmax(torch.baddbmm(tW[5][:, None], h, tW[4].transpose(1, 2)), -1)
        for _ in range(3): tfwd()
        torch.cuda.synchronize(); tms = []
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); tfwd(); e1.record(); torch.cuda.synchronize(); tms.append(e0.elapsed_time(e1))
        cpu = None
        if E == 4096:
            zc = z.cpu(); cW = [w.cpu() for w in tW]
            t0 = time.perf_counter()
            h = torch.relu(torch.baddbmm(cW[1][:, None], zc, cW[0].transpose(1, 2)))
            h = torch.relu(torch.baddbmm(cW[3][:, None], h, cW[2].transpose(1, 2)))
            torch.softmax(torch.baddbmm(cW[5][:, None], h, cW[4].transpose(1, 2)), -1)
            cpu = E * n / (time.perf_counter() - t0)
        print(json.dumps({"metric": "agent-forwards/s (policy inference, 6-300-300-16, one network per agent)",
                          "n_agents": n, "n_envs": E, "ms": m, "value": E * n / (m * 1e-3),
                          "useful_tflops": flops / (m * 1e-3) / 1e12, "issued_tf32_tflops": 3 * 2.0 * E * n * 304 * 320 / (m * 1e-3) / 1e12,
                          "torch_fp32_same_gpu_ms": float(np.median(tms)), "torch_cpu_forwards_per_s": cpu}), flush=True)
        del env

if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- golden vectors for the Monte-Carlo return scan and the
Delta-neighbourhood advantage gather, recorded from the reference.

The reference computes both inside its learners:
  * returns      G_i(t) = G_i(t+1) * discount + r_i(t), G_i(T-1) = r_i(T-1)
                 (SAC_agents.py:304-310, and verbatim in TrainedAgent.benchmark_cirtic :108-113,
                 which RETURNS them -- that function is executed here, unmodified);
  * advantages   sum_{j in N_i(t)} (G_j(t) - V_i(z_i(t)))  (SAC_agents.py:333-345, inside
                 SA2CAgents.train_NN, not callable on its own: the loop is re-typed below and
                 runs on the reference's own objects -- its ExperienceBuffers, its critics).
One <= 200-step episode of BASELINE config 1 (softmax8_n5 policy) and one with n = 8 is stored as
tests/golden/returns_*.npz.  Re-run with:  python oracle/make_golden_returns.py
"""
from __future__ import annotations

import contextlib
import io
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_harness import import_reference, import_reference_policy_stack, REFERENCE_ROOT  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def episode(name, model, n, seed, discount=0.99):
    import torch
    ref = import_reference("drone_env")
    utils, sac = import_reference_policy_stack()
    load = lambda f: torch.load(os.path.join(REFERENCE_ROOT, "models", "final", f), weights_only=False)
    agent = object.__new__(sac.TrainedAgent)          # TrainedAgent.__init__ uses torch.load defaults of torch 1.10
    agent.criticsNN, agent.actors = load(model + "-A2Ccritics.pth"), load(model + "-A2Cactors.pth")
    agent.n_agents, agent.discount = n, discount
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        env = ref.drones(n_agents=n, n_obstacles=0, grid=[5, 5], end_formation="O",
                         deltas=np.ones(n) * 1.0, simplify_zstate=True)
    env.collision_weight = 0.2
    buffers = utils.ExperienceBuffers(n)
    finished = False
    while not finished:                                # benchmark_agent.py:69-94
        z_states, Ni = env.z_states, env.Ni
        actions = agent.forward(z_states, Ni)
        _, new_z, rewards, _, finished, _ = env.step(actions)
        buffers.append(z_states, actions, rewards, new_z, Ni, finished)
    Gts, V_approxs = agent.benchmark_cirtic(buffers)   # reference code: the returns
    T, k = len(buffers), env.k_closest
    adv = np.zeros((T, n)); base = np.zeros((T, n))
    for i in range(n):                                 # SAC_agents.py:333-345, re-typed
        for t in range(T):
            zit = buffers.buffers[i][t].z_state
            Nit = buffers.buffers[i][t].Ni
            Advantage_j_sum = 0
            input_tensor = torch.tensor(zit, dtype=torch.float32)
            Vi_baseline = agent.criticsNN[i](input_tensor).detach().numpy()[0]
            for j in Nit:
                Advantage_j_sum += (Gts[j][t] - Vi_baseline)
            adv[t, i] = Advantage_j_sum; base[t, i] = Vi_baseline
    r = np.array([[buffers.buffers[i][t].reward for i in range(n)] for t in range(T)], np.float64)
    Ni = np.full((T, n, k + 1), -1, np.int32)
    for t in range(T):
        for i in range(n):
            lst = buffers.buffers[i][t].Ni
            Ni[t, i, :len(lst)] = lst
    fin = np.array([buffers.buffers[0][t].finished for t in range(T)], np.uint8)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), n=n, k=k, discount=discount, reward=r, Ni=Ni,
                        finished=fin, returns=np.array(Gts, np.float64).T.copy(), baseline=base, advantage=adv)
    print(f"{name}: T={T} n={n} mean |G|={np.abs(np.array(Gts)).mean():.3f} neighbours/agent={np.mean((Ni >= 0).sum(-1)):.2f}")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    episode("returns_n5_seed0", "softmax8_n5", 5, 0)
    episode("returns_n5_seed3_g0.9", "softmax8_n5", 5, 3, discount=0.9)
    episode("returns_n8_seed1", "softmax8_n8", 8, 1)

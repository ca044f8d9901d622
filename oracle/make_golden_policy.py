#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- golden vectors for batched policy inference (SURVEY.md section 8f
row 1), recorded from the reference: the pretrained `DiscreteSoftmaxNN` actors of BASELINE config 1
(models/final/softmax8_n5-A2Cactors.pth, utils.py:255-318) evaluated by the reference's own
`forward` on observations of a recorded episode.  Stored: the fp32 weights of the first two agents
(the networks are 6 -> 300 -> 300 -> 8), observations z, and the probabilities the reference
returns.  Re-run with:  python oracle/make_golden_policy.py"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_harness import import_reference_policy_stack, REFERENCE_ROOT  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    import torch
    utils, _ = import_reference_policy_stack()
    actors = torch.load(os.path.join(REFERENCE_ROOT, "models", "final", "softmax8_n5-A2Cactors.pth"), weights_only=False)
    g = np.load(os.path.join(OUT, "policy_n5_seed0.npz"))
    z = g["z"]                                             # [T, n, k+1, 2] observations of the recorded episode
    T = min(z.shape[0], 96)
    agents = (0, 1)
    W = {}
    probs = np.zeros((T, len(agents), actors[0].n_actions), np.float32)
    for a, i in enumerate(agents):
        net = actors[i]
        assert isinstance(net, utils.DiscreteSoftmaxNN)
        for name, layer in (("1", net.input_layer), ("2", net.hidden_layer1), ("3", net.out_1)):
            W[f"W{name}_{a}"] = layer.weight.detach().numpy().astype(np.float32)
            W[f"b{name}_{a}"] = layer.bias.detach().numpy().astype(np.float32)
        for t in range(T):
            state_tensor = torch.tensor(z[t, i].flatten(), dtype=torch.float32)       # utils.py:305
            probs[t, a] = net.forward(state_tensor).detach().numpy()                     # utils.py:306
    np.savez_compressed(os.path.join(OUT, "policynet_n5_agents01.npz"), z=z[:T, list(agents)].reshape(T, len(agents), -1),
                        probs=probs, action_list=actors[0].action_list, n_actions=actors[0].n_actions, **W)
    print("policynet_n5_agents01:", T, "steps,", len(agents), "agents, probs range", probs.min(), probs.max(),
          os.path.getsize(os.path.join(OUT, "policynet_n5_agents01.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()

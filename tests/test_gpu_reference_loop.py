"""GPU: the acceptance surface the north star names -- the reference's OWN episode loop
(train_problem.py:82-121: `SA2CAgents.forward` -> `env.step` -> `ExperienceBuffers.append`, then
`train_NN`) running UNCHANGED against this repository's drop-in `drone_env.drones` (every step a
launch of the sm_100a step kernel through the C ABI), in lock-step with the unmodified reference
environment fed the same actions.

The reference travels to the GPU box as the git-ignored staging copy `oracle/_ref` (unmodified
files, `oracle/make_ref.py`); without it the test is skipped with that reason.  The reference's
agents run on the CPU as the reference prescribes (torch cpuonly).
"""
import random

import numpy as np
import pytest
import torch

from helpers import assert_close, compare_obs

pytestmark = pytest.mark.gpu

FP64_TOL = 1e-9


def _ref():
    from oracle import ref_harness
    if not ref_harness.reference_available():
        pytest.skip("the unmodified reference is not on this box (neither /root/reference nor oracle/_ref)")
    utils, sac = ref_harness.import_reference_policy_stack()
    return ref_harness.import_reference("drone_env"), utils, sac


def _tie_rows(env_ref, state):
    """Rows whose k + 2 smallest clipped distances contain an exact tie (np.argsort is unstable there)."""
    d_ij, _, _, _ = env_ref.distance_data(state, env_ref.deltas, env_ref.d_safety)
    k = env_ref.k_closest
    srt = np.sort(d_ij, axis=1)[:, :k + 2]
    return (np.diff(srt, axis=1) == 0).any(axis=1)


@pytest.mark.parametrize("n,delta,seed", [(5, 1.0, 0), (4, 2.43, 3)])
def test_reference_training_loop_runs_unchanged_on_the_dropin_env(n, delta, seed, capsys):
    ref_env_mod, utils, sac = _ref()
    import drone_env as dropin                               # this repository's module (repo root)
    assert dropin.drones is not ref_env_mod.drones

    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    mk = dict(n_agents=n, n_obstacles=0, grid=[5, 5], end_formation="O", deltas=np.ones(n) * delta, simplify_zstate=True)
    env = dropin.drones(**mk)                                 # train_problem.py:30-31
    env.collision_weight = 0.2
    ref = ref_env_mod.drones(**{**mk, "deltas": np.ones(n) * delta})
    ref.collision_weight = 0.2
    agents = sac.SA2CAgents(n_agents=env.n_agents, dim_local_state=env.local_state_space,
                            dim_local_action=env.local_action_space, discount=0.99, epochs=10,
                            learning_rate_critic=1e-3, learning_rate_actor=1e-3)          # train_problem.py:59
    assert env.local_state_space == ref.local_state_space == 6 and env.local_action_space == 2

    steps_total = 0
    for episode in range(2):                                  # two whole episodes, the reference's loop verbatim
        ref.state = env.state.copy()                          # same start (env.reset() drew it with `random`)
        ref.internal_t = 0
        _, _, ref.z_states, ref.Ni, _ = ref.rewards(ref.state, ref.end_points, ref.n_agents, ref.d_safety, ref.deltas)
        buffers = utils.ExperienceBuffers(env.n_agents)
        finished, t_iter = False, 0
        tot_r = tot_tr = 0.0
        tot_c = 0
        while not finished:                                   # train_problem.py:82-107
            state, z_states, Ni = env.state, env.z_states, env.Ni
            actions = agents.forward(z_states, Ni)            # reference actors, CPU torch, np.random.choice
            new_state, new_z, rewards, n_collisions, finished, true_rewards = env.step(actions)
            buffers.append(z_states, actions, rewards, new_z, Ni, finished)
            tot_r += np.mean(rewards); tot_tr += np.mean(true_rewards); tot_c += n_collisions
            # the unmodified reference environment, same actions
            r_state, r_z, r_rew, r_nc, r_fin, r_trw = ref.step([np.asarray(a) for a in actions])
            assert new_state is env.state                     # aliasing contract (drone_env.py:258)
            assert np.array_equal(new_state[:, :4], r_state[:, :4]), f"state, episode {episode} step {t_iter}"
            assert_close(rewards, r_rew, FP64_TOL, "reward")
            assert_close(true_rewards, r_trw, FP64_TOL, "true reward")
            assert int(n_collisions) == int(r_nc) and bool(finished) == bool(r_fin)
            tie = _tie_rows(ref, r_state)
            compare_obs(np.stack(new_z), _pad(env.Ni, env.k_closest), np.stack(r_z), _pad(ref.Ni, ref.k_closest), tie,
                        FP64_TOL, f"observation, episode {episode} step {t_iter}")
            t_iter += 1
        steps_total += t_iter
        assert 1 <= t_iter <= 200 and len(buffers) == t_iter
        agents.train_NN(buffers, actor_lr=1e-3)               # train_problem.py:115: consumes z / Ni / rewards as stored
        env.reset(renew_obstacles=False)                      # train_problem.py:132
        assert env.internal_t == 0
    capsys.readouterr()
    assert steps_total >= 2


def _pad(Ni, k):
    out = np.full((len(Ni), k + 1), -1, np.int32)
    for i, lst in enumerate(Ni):
        out[i, :len(lst)] = lst
    return out

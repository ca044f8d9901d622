"""CPU: pin the oracles (C and NumPy restatements) to the golden vectors that
oracle/make_golden.py recorded from the unmodified reference drone_env.py."""
import numpy as np
import pytest

from helpers import FP64_TOL, assert_close, compare_obs, golden_names, load_golden
from oracle import c_oracle, np_oracle


def _c_env(g, E=1):
    p = c_oracle.default_params(g["collision_weight"], g["dt"], g["max_time_steps"])
    return c_oracle.OracleEnv(E, g["n"], g["end_points"], g["d_safety"], g["deltas"], g["radius"],
                              g["k"], bool(g["simplify"]), p)


@pytest.mark.parametrize("name", golden_names())
def test_c_oracle_teacher_forced(name):
    """Every recorded step, fed the reference's own input state (batched as E=T envs)."""
    g = load_golden(name)
    T = len(g["ncoll"])
    env = _c_env(g, E=T)
    env.set_state(g["state_in"][:, :, 0:2], g["state_in"][:, :, 2:4], g["t_in"])
    out = env.step(g["actions"])
    assert_close(out.pos, g["state"][:, :, 0:2], 0.0, "pos (bit-exact)")
    assert_close(out.vel, g["state"][:, :, 2:4], 0.0, "vel (bit-exact)")
    assert_close(out.r, g["r"], FP64_TOL, "reward")
    assert_close(out.true_r, g["true_r"], FP64_TOL, "true reward")
    assert np.array_equal(out.ncoll.astype(np.int64), g["ncoll"]), "collision counts"
    assert np.array_equal(out.finished, g["finished"]), "finished"
    assert np.array_equal(out.t, g["t_in"] + 1)
    assert np.array_equal(out.tie.astype(bool), g["tie"]), "tie mask"
    compare_obs(out.z, out.Ni, g["z"], g["Ni"], g["tie"], FP64_TOL, name)


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith(("free", "policy"))])
def test_c_oracle_free_running(name):
    """Whole episode from state0 with only the action stream shared."""
    g = load_golden(name)
    env = _c_env(g)
    env.set_state(g["state0"][None, :, 0:2], g["state0"][None, :, 2:4], 0)
    o0 = env.observe()
    compare_obs(o0.z[0], o0.Ni[0], g["z0"], g["Ni0"], g["tie0"], FP64_TOL, name + " init")
    ret = 0.0
    for t in range(len(g["ncoll"])):
        out = env.step(g["actions"][t][None])
        assert_close(out.pos[0], g["state"][t][:, 0:2], 0.0, f"pos t={t}")
        assert_close(out.r[0], g["r"][t], FP64_TOL, f"r t={t}")
        assert int(out.ncoll[0]) == int(g["ncoll"][t])
        assert int(out.finished[0]) == int(g["finished"][t])
        ret += out.r[0].mean()
    assert abs(ret - g["r"].mean(1).sum()) < 1e-9


def test_c_oracle_rollout_matches_stepping():
    g = load_golden("free_n10_g5_d1.0")
    T = len(g["ncoll"])
    env = _c_env(g)
    env.set_state(g["state0"][None, :, 0:2], None, 0)
    res = env.rollout(g["actions"][:, None])
    assert_close(res["r"][:, 0], g["r"], FP64_TOL, "rollout r")
    assert np.array_equal(res["ncoll"][:, 0].astype(np.int64), g["ncoll"])
    assert np.array_equal(res["finished"][:, 0], g["finished"])
    assert_close(res["agg"][0], [g["r"].mean(1).sum(), g["true_r"].mean(1).sum(), g["ncoll"].sum(), T],
                 1e-9, "episode aggregates")


def test_policy_episode_known_answer():
    """seed 0, softmax8_n5 policy -> 200 steps, return -33.5050, 0 collisions."""
    g = load_golden("policy_n5_seed0")
    assert len(g["ncoll"]) == 200 and int(g["ncoll"].sum()) == 0
    assert abs(g["r"].mean(1).sum() - (-33.5050)) < 5e-4


@pytest.mark.parametrize("name", golden_names())
def test_np_oracle_teacher_forced(name):
    g = load_golden(name)
    pos = g["state_in"][:, :, 0:2].copy(); vel = g["state_in"][:, :, 2:4].copy()
    t = g["t_in"].copy()
    out = np_oracle.step(pos, vel, t, g["actions"], g["radius"], g["end_points"].reshape(-1, 2),
                         g["d_safety"], g["deltas"], g["k"], bool(g["simplify"]),
                         g["collision_weight"], g["dt"], g["max_time_steps"])
    assert_close(pos, g["state"][:, :, 0:2], 0.0, "pos (bit-exact)")
    assert_close(out["r"], g["r"], FP64_TOL, "reward")
    assert_close(out["true_r"], g["true_r"], FP64_TOL, "true reward")
    assert np.array_equal(out["ncoll"].astype(np.int64), g["ncoll"])
    assert np.array_equal(out["finished"], g["finished"])
    compare_obs(out["z"], out["Ni"], g["z"], g["Ni"], g["tie"], FP64_TOL, name)


def test_known_answer_facts():
    """SURVEY section 4 edge cases, read back from the reference's own outputs."""
    g = load_golden("edge_cases_n4")
    # step 0: agents 0,1 are 0.15 apart -> both directions collide -> 2; +99.9 each
    assert g["ncoll"][0] == 2
    # step 1: exactly 0.2 apart -> d == 0 -> -1e-6 -> counts as a collision
    assert g["ncoll"][1] == 2
    # step 2: agent 2 exactly on its goal, fewer than k neighbours -> NaN ghost (0/0)
    assert np.isnan(g["z"][2][2, 2, 0:2]).all()
    # step 4: everyone within 0.2 of goal -> finished
    assert g["finished"][4] == 1
    # last two steps: t=198 not finished, t=199 finished by time
    assert g["finished"][-2] == 0 and g["finished"][-1] == 1
    env = _c_env(g, E=1)
    env.set_state(g["state_in"][0][None, :, 0:2], None, 5)
    out = env.step(g["actions"][0][None])
    q, b = 2 * g["dt"], g["collision_weight"] * g["dt"]
    assert abs(b * 9.99e3 - 99.9) < 1e-12
    goal0 = q * np.sum((g["end_points"].reshape(-1, 2)[0] - out.pos[0, 0]) ** 2)
    assert -out.r[0, 0] >= goal0 + 99.9 - 1e-9


RETURNS = ["returns_n5_seed0", "returns_n5_seed3_g0.9", "returns_n8_seed1"]


@pytest.mark.parametrize("name", RETURNS)
def test_returns_oracles_match_reference(name):
    """Monte-Carlo returns (computed by the reference's own TrainedAgent.benchmark_cirtic) and the
    Delta-neighbourhood advantage sums (reference loop SAC_agents.py:333-345 on the reference's
    buffers and critics), recorded by oracle/make_golden_returns.py: both restatements bit-exact."""
    import os
    from oracle import np_oracle
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    r, Ni, fin = g["reward"][:, None, :], g["Ni"][:, None], g["finished"][:, None]
    assert fin[-1, 0] == 1 and fin[:-1].sum() == 0                       # one whole episode
    for mod in (c_oracle, np_oracle):
        G, adv, cnt = mod.returns(r, Ni, fin, float(g["discount"]), g["baseline"][:, None, :])
        assert np.array_equal(G[:, 0], g["returns"]), mod.__name__
        assert np.array_equal(adv[:, 0], g["advantage"]), mod.__name__
        assert np.array_equal(cnt[:, 0], (g["Ni"] >= 0).sum(-1))
        G0, adv0, _ = mod.returns(r, Ni, fin, float(g["discount"]))       # no baseline: plain neighbour sums
        assert np.array_equal(G0, G)
        want = np.zeros_like(adv0[:, 0])
        for m in range(g["Ni"].shape[-1]):
            j = g["Ni"][..., m]
            want = np.where(j >= 0, want + np.take_along_axis(g["returns"], np.where(j >= 0, j, 0), 1), want)
        assert np.array_equal(adv0[:, 0], want)


def test_returns_oracles_agree_on_ragged_batch():
    """Batch of environments with different episode ends (finished codes 0/1/2 as ds_rollout
    writes them, including environments that never ran): C and NumPy restatements agree."""
    from oracle import np_oracle
    rng = np.random.default_rng(5)
    T, E, n, k = 37, 23, 6, 2
    r = -rng.uniform(0, 30, (T, E, n))
    Ni = rng.integers(-1, n, (T, E, n, k + 1)).astype(np.int32)
    Ni[..., 0] = np.arange(n)
    fin = np.zeros((T, E), np.uint8)
    ends = rng.integers(0, T + 8, E)
    for e in range(E):
        if ends[e] < T:
            fin[ends[e], e] = 1
            fin[ends[e] + 1:, e] = 2
    fin[:, 3] = 2                                                        # done before the rollout
    base = rng.standard_normal((T, E, n))
    a = c_oracle.returns(r, Ni, fin, 0.97, base)
    b = np_oracle.returns(r, Ni, fin, 0.97, base)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert (a[0][:, 3] == 0).all() and (a[2][:, 3] == 0).all()
    e = int(np.argmax(ends < T))
    assert (a[0][ends[e] + 1:, e] == 0).all() and np.array_equal(a[0][ends[e], e], r[ends[e], e])


@pytest.mark.parametrize("name,key,mode", [("control_gradient_n5", "action", 2), ("control_gradient_n10", "action", 2),
                                           ("control_proportional_n8", "action", 1),
                                           ("control_dense_n7", "gradient", 2), ("control_dense_n7", "proportional", 1)])
def test_control_oracles_match_reference(name, key, mode):
    """Baseline controllers: C and NumPy restatements against the actions recorded from the
    reference's gradient_control / proportional_control (drone_env.py:612-679), bit-exact."""
    import os
    from oracle import np_oracle
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    u_max = float(g["u_max"]) if mode == 2 else 1.0
    for mod in (c_oracle, np_oracle):
        act = mod.control(mode, g["state_in"][:, :, 0:2], g["end_points"], g["d_safety"], None, u_max)
        assert np.array_equal(act, g[key], equal_nan=True), mod.__name__


def test_philox_known_answers_and_reset_restatement():
    """Philox4x32-10 of the device-reset restatement against the Random123 known-answer vectors;
    the sampler returns distinct nodes of the reference's lattice (drone_env.py:193-205)."""
    from oracle import np_oracle
    from scalable_collision_avoidance_rl_b200 import formation
    kat = [((0, 0, 0, 0, 0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 6, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for args, want in kat:
        assert tuple(np_oracle._philox4x32_10(*args)) == want
    d0, d1 = formation.lattice_shape([5, 5])
    pos = np_oracle.reset_random(200, 10, d0, d1, formation.LATTICE_PITCH, seed=7, stream=3)
    nodes = np.rint(pos / formation.LATTICE_PITCH).astype(int)
    assert np.array_equal(pos, nodes * formation.LATTICE_PITCH)               # exactly idx * pitch
    assert nodes.min() >= 0 and nodes[..., 0].max() < d0 and nodes[..., 1].max() < d1
    flat = nodes[..., 0] * d1 + nodes[..., 1]
    assert all(len(set(row)) == 10 for row in flat)
    assert not np.array_equal(pos, np_oracle.reset_random(200, 10, d0, d1, formation.LATTICE_PITCH, 7, 4))


def test_policy_restatement_matches_reference():
    """fp32 restatement of DiscreteSoftmaxNN.forward against the probabilities the reference's own
    pretrained actors return (oracle/make_golden_policy.py): <= 5e-7; sampling restatement sane."""
    import os
    from oracle import np_oracle
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "policynet_n5_agents01.npz"))
    for a in (0, 1):
        p = np_oracle.policy_probs(g["z"][:, a], *[g[f"{w}_{a}"] for w in ("W1", "b1", "W2", "b2", "W3", "b3")])
        assert np.abs(p - g["probs"][:, a]).max() <= 5e-7
        assert np.abs(p.sum(1) - 1).max() < 1e-6
    idx = np_oracle.policy_sample(g["probs"][:, :1], g["probs"].shape[0], 1, seed=3, stream=0)
    assert idx.shape == (g["probs"].shape[0], 1) and idx.max() < int(g["n_actions"])

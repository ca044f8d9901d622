"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes ->
libdronestep.so), against (a) the golden vectors recorded from the unmodified
reference and (b) the C oracle on seeded random inputs.

Tolerances: BASELINE.json's contract is |delta| <= 1e-5 on float state/reward and
bit-exact integer collision counts.  The float64 instantiation is asserted at
1e-9 (and positions bit-exact); the float32 throughput mode has its own stated
relative tolerance.
"""
import numpy as np
import pytest
import torch

from helpers import CONTRACT_TOL, FP64_TOL, assert_close, compare_obs, golden_names, load_golden
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _mk(g, E, dtype=torch.float64, log_mode=0):
    from scalable_collision_avoidance_rl_b200 import BatchedDrones
    deltas_in = g["deltas_in"] if g["deltas_in"].size else None
    env = BatchedDrones(E, g["n"], list(g["grid"]), "O", g["k"], deltas_in, bool(g["simplify"]),
                        dtype=dtype, seed=0, warn=False)
    env.collision_weight = g["collision_weight"]
    env.log_mode = log_mode
    # host-side setup must reproduce the reference's constructor outputs bit for bit
    assert np.array_equal(env.end_points, g["end_points"])
    assert np.array_equal(env.d_safety, g["d_safety"])
    assert np.array_equal(np.asarray(env.deltas, np.float64), g["deltas"])
    return env


@pytest.mark.parametrize("log_mode", [0, 1])
@pytest.mark.parametrize("name", golden_names())
def test_step_vs_reference_golden(name, log_mode):
    """Every recorded reference step, teacher-forced, as one batch of E = T environments."""
    g = load_golden(name)
    T = len(g["ncoll"])
    env = _mk(g, T, log_mode=log_mode)
    env.set_state(g["state_in"], g["t_in"])
    act = torch.as_tensor(g["actions"], device=env.device)
    (pos, vel), z, r, ncoll, fin, true_r = env.step(act)
    torch.cuda.synchronize()
    assert_close(pos.cpu().numpy(), g["state"][:, :, 0:2], 0.0, "pos (bit-exact)")
    assert_close(vel.cpu().numpy(), g["state"][:, :, 2:4], 0.0, "vel (bit-exact)")
    assert_close(r.cpu().numpy(), g["r"], FP64_TOL, "reward")
    assert_close(true_r.cpu().numpy(), g["true_r"], FP64_TOL, "true reward")
    assert np.array_equal(ncoll.cpu().numpy().astype(np.int64), g["ncoll"]), "collision counts"
    assert np.array_equal(fin.cpu().numpy(), g["finished"]), "finished"
    assert np.array_equal(env.internal_t.cpu().numpy(), g["t_in"] + 1)
    compare_obs(z.cpu().numpy(), env.Ni.cpu().numpy(), g["z"], g["Ni"], g["tie"], FP64_TOL, name)


@pytest.mark.parametrize("name", [f"policy_n5_seed{s}" for s in range(10)] +
                         ["free_n10_g5_d1.0", "free_n8_g5_d1.0_k3_cw0.5", "free_n5_g5_dNone_full"])
def test_dropin_class_lockstep(name):
    """The reference-facing class `drone_env.drones`, free-running a whole recorded episode
    (BASELINE config 1 for the policy_* fixtures) with only the action stream shared."""
    import random
    import drone_env
    g = load_golden(name)
    deltas_in = g["deltas_in"] if g["deltas_in"].size else None
    random.seed(0)
    env = drone_env.drones(n_agents=g["n"], n_obstacles=0, grid=[float(x) for x in g["grid"]],
                           end_formation="O", k_closest=g["k"], deltas=deltas_in,
                           simplify_zstate=bool(g["simplify"]))
    env.collision_weight = g["collision_weight"]      # mutated after construction (train_problem.py:31)
    assert np.array_equal(env.end_points, g["end_points"]) and np.array_equal(env.d_safety, g["d_safety"])
    assert env.local_state_space == (2 if g["simplify"] else 5) * (g["k"] + 1)
    env.state[:, :] = g["state0"]                     # inject the reference's start state
    _, _, z0, Ni0, _ = env.rewards(env.state, env.end_points, env.n_agents, env.d_safety, env.deltas)
    pad0 = np.full((g["n"], g["k"] + 1), -1)
    for i, l in enumerate(Ni0):
        pad0[i, :len(l)] = l
    compare_obs(np.array(z0), pad0, g["z0"], g["Ni0"], g["tie0"], FP64_TOL, name + " init")
    ret = coll = 0
    for t in range(len(g["ncoll"])):
        state, z, r, ncoll, fin, true_r = env.step([g["actions"][t][i] for i in range(g["n"])])
        assert state is env.state                                  # aliasing contract (drone_env.py:258)
        assert isinstance(fin, bool) and isinstance(ncoll, np.integer)
        assert_close(state, g["state"][t], 0.0, f"state t={t} (bit-exact)")
        assert_close(r, g["r"][t], FP64_TOL, f"r t={t}")
        assert_close(true_r, g["true_r"][t], FP64_TOL, f"true_r t={t}")
        assert int(ncoll) == int(g["ncoll"][t]) and fin == bool(g["finished"][t])
        pad = np.full((g["n"], g["k"] + 1), -1)
        for i, l in enumerate(env.Ni):
            assert l[0] == i
            pad[i, :len(l)] = l
        compare_obs(np.array(z), pad, g["z"][t], g["Ni"][t], g["tie"][t], FP64_TOL, f"{name} t={t}")
        ret += float(np.mean(r)); coll += int(ncoll)
    assert abs(ret - g["r"].mean(1).sum()) < 1e-9 and coll == int(g["ncoll"].sum())
    env.reset(renew_obstacles=False)
    assert env.internal_t == 0 and np.all(env.state[:, 2:4] == 0)


def _random_case(n, E, k, simplify, grid, delta, seed, box=None, hetero=False):
    from scalable_collision_avoidance_rl_b200 import BatchedDrones
    rng = np.random.default_rng(seed)
    deltas = (rng.uniform(0.2, delta, n) if hetero else np.ones(n) * delta)
    env = BatchedDrones(E, n, grid, "O", k, deltas, simplify, seed=seed, warn=False)
    box = box or grid[0]
    pos = rng.uniform(0, box, (E, n, 2))
    vel = rng.standard_normal((E, n, 2))
    t = rng.integers(0, 205, E).astype(np.int32)
    act = rng.uniform(-1, 1, (E, n, 2))
    state = np.concatenate([pos, vel, np.full((E, n, 1), 0.1)], 2)
    orc = c_oracle.OracleEnv(E, n, env.end_points, env.d_safety, env.deltas, None, k, simplify,
                             c_oracle.default_params(env.collision_weight), nthreads=8)
    orc.set_state(pos, vel, t)
    env.set_state(state, t)
    return env, orc, act


CASES = [  # n, E, k, simplify, grid, delta, box, hetero
    (5, 777, 2, True, [5, 5], 1.0, 2.0, False),
    (10, 512, 2, True, [5, 5], 1.0, 3.0, False),
    (10, 300, 2, False, [5, 5], 1.0, 2.0, True),
    (7, 301, 0, True, [5, 5], 1.0, 2.0, False),
    (9, 257, 1, False, [5, 5], 1.0, 2.0, False),
    (12, 129, 4, True, [8, 8], 1.5, 3.0, True),
    (20, 65, 6, False, [16, 16], 2.0, 5.0, False),    # dynamic-k kernel
    (32, 200, 2, True, [32, 32], 2.5, 8.0, False),
    (33, 100, 2, True, [32, 32], 2.5, 8.0, False),    # first CTA-group size
    (64, 50, 3, False, [32, 32], 1.0, 6.0, True),
    (128, 24, 2, True, [64, 64], 1.0, 12.0, False),
    (300, 5, 2, True, [128, 128], 1.0, 20.0, False),  # > 256 threads per environment
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c[0]}_E{c[1]}_k{c[2]}_{'s' if c[3] else 'f'}")
def test_step_vs_oracle_random(case):
    """Seeded dense random states (collisions guaranteed) against the C oracle."""
    n, E, k, simplify, grid, delta, box, hetero = case
    env, orc, act = _random_case(n, E, k, simplify, grid, delta, 1000 + n, box, hetero)
    ref = orc.step(act)
    (pos, vel), z, r, ncoll, fin, true_r = env.step(torch.as_tensor(act, device=env.device))
    torch.cuda.synchronize()
    assert ref.ncoll.sum() > 0, "test must exercise collisions"
    assert_close(pos.cpu().numpy(), ref.pos, 0.0, "pos")
    assert_close(r.cpu().numpy(), ref.r, FP64_TOL, "reward")
    assert_close(true_r.cpu().numpy(), ref.true_r, FP64_TOL, "true reward")
    assert np.array_equal(ncoll.cpu().numpy(), ref.ncoll)
    assert np.array_equal(fin.cpu().numpy(), ref.finished)
    compare_obs(z.cpu().numpy(), env.Ni.cpu().numpy(), ref.z, ref.Ni, ref.tie, FP64_TOL, "obs")
    # observe(): same evaluation without integrating, t/finished untouched
    t_before = env.internal_t.clone()
    env.observe()
    torch.cuda.synchronize()
    assert torch.equal(env.internal_t, t_before)
    assert_close(env.rewards.cpu().numpy(), ref.r, FP64_TOL, "observe reward")


@pytest.mark.parametrize("n,E,grid,delta", [(5, 96, [5, 5], 1.0), (10, 64, [5, 5], 1.0),
                                            (32, 16, [32, 32], 2.5), (40, 9, [32, 32], 2.5)])
def test_rollout_vs_oracle_and_stepping(n, E, grid, delta):
    """ds_rollout (T fused steps) == T x ds_step bit for bit, and == the oracle's episode loop,
    including early termination, frozen finished envs and the episode aggregates."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones
    T = 60
    rng = np.random.default_rng(n)
    tab = np.stack([np.cos(np.arange(16) / 16 * 2 * np.pi), np.sin(np.arange(16) / 16 * 2 * np.pi)], 1)
    idx = rng.integers(0, 16, (T, E, n)).astype(np.uint8)
    act = tab[idx]
    envA = BatchedDrones(E, n, grid, "O", 2, np.ones(n) * delta, True, seed=5, warn=False)
    start = envA.pos.cpu().numpy().copy()
    t0 = rng.integers(150, 199, E).astype(np.int32)      # some envs hit the 200-step limit inside T
    state = np.concatenate([start, np.zeros((E, n, 2)), np.full((E, n, 1), 0.1)], 2)
    envA.set_state(state, t0)
    envB = BatchedDrones(E, n, grid, "O", 2, np.ones(n) * delta, True, seed=5, warn=False)
    envB.set_state(state, t0)
    envC = BatchedDrones(E, n, grid, "O", 2, np.ones(n) * delta, True, seed=5, warn=False)
    envC.set_state(state, t0)
    orc = c_oracle.OracleEnv(E, n, envA.end_points, envA.d_safety, envA.deltas, None, 2, True,
                             c_oracle.default_params(envA.collision_weight))
    orc.set_state(start, None, t0)
    ref = orc.rollout(act)
    rec = ("pos", "vel", "reward", "true_reward", "obs", "ncoll", "finished")
    out = envA.rollout(actions=torch.as_tensor(act, device=envA.device), record=rec)
    outC = envC.rollout(action_idx=torch.as_tensor(idx, device=envC.device), action_table=tab, record=rec)
    torch.cuda.synchronize()
    fin_tr = out["finished"].cpu().numpy()
    assert np.array_equal(fin_tr, ref["finished"])
    live = fin_tr != 2
    assert (fin_tr == 2).any() and (fin_tr == 1).any()
    assert_close(out["reward"].cpu().numpy()[live], ref["r"][live], FP64_TOL, "rollout r")
    assert_close(out["true_reward"].cpu().numpy()[live], ref["true_r"][live], FP64_TOL, "rollout true r")
    assert np.array_equal(out["ncoll"].cpu().numpy()[live], ref["ncoll"][live])
    assert_close(out["agg"].cpu().numpy(), ref["agg"], 1e-9, "episode aggregates")
    assert np.array_equal(out["done"].cpu().numpy(), ref["done"])
    assert_close(envA.pos.cpu().numpy(), ref["pos"], 0.0, "final pos")
    assert np.array_equal(envA.internal_t.cpu().numpy(), ref["t"])
    lv_t = torch.as_tensor(live, device=envA.device)   # index mode == real-valued mode, bit for bit
    for key in ("pos", "reward", "true_reward", "z", "Ni", "ncoll"):
        assert torch.equal(out[key][lv_t], outC[key][lv_t]), key
    # stepping envB one launch at a time reproduces the fused trajectory exactly
    done = np.zeros(E, bool)
    for t in range(T):
        before = (envB.pos.clone(), envB.vel.clone(), envB.internal_t.clone())
        envB.step(torch.as_tensor(act[t], device=envB.device))
        torch.cuda.synchronize()
        # freeze finished environments like the rollout does
        d = torch.as_tensor(done, device=envB.device)
        envB.pos[d] = before[0][d]; envB.vel[d] = before[1][d]; envB.internal_t[d] = before[2][d]
        lv = ~done
        assert torch.equal(envB.rewards[torch.as_tensor(lv, device=envB.device)],
                           out["reward"][t][torch.as_tensor(lv, device=envB.device)])
        assert torch.equal(envB.z_states[torch.as_tensor(lv, device=envB.device)],
                           out["z"][t][torch.as_tensor(lv, device=envB.device)])
        done |= envB.finished.cpu().numpy().astype(bool) & lv
    assert torch.equal(envB.pos, envA.pos)
    # device-side reduction of the aggregates (the vector a rank all-reduces)
    s = envA.episode_aggregates().cpu().numpy()
    assert_close(s[:4], ref["agg"].sum(0), 1e-6, "reduced aggregates")
    assert s[4] == E


def test_host_entries_match_device_entries():
    """ds_step_host / ds_rollout_host (host buffers, copies inside) == device-pointer entries."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones
    n, E, T = 10, 200, 37
    rng = np.random.default_rng(3)
    act = rng.uniform(-1, 1, (T, E, n, 2))
    a = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=9, warn=False)
    b = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=9, warn=False)
    assert torch.equal(a.pos, b.pos)
    o = b.step_host(act[0])
    a.step(torch.as_tensor(act[0], device=a.device))
    torch.cuda.synchronize()
    assert np.array_equal(o["pos"], a.pos.cpu().numpy()) and np.array_equal(o["r"], a.rewards.cpu().numpy())
    assert np.array_equal(o["z"], a.z_states.cpu().numpy()) and np.array_equal(o["Ni"], a.Ni.cpu().numpy())
    assert np.array_equal(o["nc"], a.n_collisions.cpu().numpy())
    # one transfer of the result block (ds_step_host_block) == one copy per array (ds_step_host)
    c = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=9, warn=False)
    o2 = c.step_host(act[0], block=False)
    for key in ("pos", "vel", "z", "r", "tr", "Ni", "nc", "fin"):
        assert np.array_equal(o[key], o2[key]), key
    rec = ("pos", "vel", "reward", "true_reward", "obs", "ncoll", "finished")
    da = a.rollout(actions=torch.as_tensor(act[1:], device=a.device), record=rec)
    for chunk in (0, 5, 36, 100):
        b2 = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=9, warn=False)
        b2.step_host(act[0])
        hb = b2.rollout_host(actions=act[1:], record=rec, chunk=chunk)
        torch.cuda.synchronize()
        for key in ("pos", "vel", "reward", "true_reward", "z", "Ni", "ncoll", "finished"):
            assert torch.equal(hb[key], da[key].cpu()), f"{key} chunk={chunk}"
        # episode sums are reduced per call (rows first, then agents): chunking only reorders additions
        assert_close(hb["agg"].numpy(), da["agg"].cpu().numpy(), FP64_TOL, f"agg chunk={chunk}")
    # index mode: u8 indices + table in; the recorded velocity is table[idx], written on the host
    from scalable_collision_avoidance_rl_b200 import formation
    tab = formation.unit_action_table(16)
    idx = rng.integers(0, 16, (T, E, n)).astype(np.uint8)
    c1 = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=9, warn=False)
    c2 = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=9, warn=False)
    dd = c1.rollout(action_idx=torch.as_tensor(idx, device=c1.device), action_table=tab, record=rec)
    hh = c2.rollout_host(action_idx=idx, action_table=tab, record=rec, chunk=7)
    torch.cuda.synchronize()
    for key in ("pos", "vel", "reward", "true_reward", "z", "Ni", "ncoll", "finished"):
        assert torch.equal(hh[key], dd[key].cpu()), f"index mode {key}"
    assert np.array_equal(hh["vel"].numpy(), tab[idx])
    # compact host observations (DS_HOST_COMPACT_OBS): z as float32, Ni as u8 (255 = none); the rest unchanged
    c3 = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=9, warn=False)
    hc = c3.rollout_host(action_idx=idx, action_table=tab, record=rec, chunk=7, compact=True)
    torch.cuda.synchronize()
    assert hc["z"].dtype == torch.float32 and hc["Ni"].dtype == torch.uint8
    assert torch.equal(hc["z"], dd["z"].cpu().float())
    ni = dd["Ni"].cpu()
    assert torch.equal(hc["Ni"], torch.where(ni < 0, torch.full_like(ni, 255), ni).to(torch.uint8))
    for key in ("pos", "reward", "true_reward", "ncoll", "finished"):
        assert torch.equal(hc[key], dd[key].cpu()), f"compact mode {key}"
    # host tensors that cannot be used in place (wrong dtype / not contiguous) are staged, not misread
    c4 = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=9, warn=False)
    wide = torch.zeros((T, E, n, 4), dtype=torch.float64).pin_memory()
    wide[..., :2] = torch.as_tensor(tab[idx])
    bad = wide[..., :2]                                                  # pinned, right dtype, NOT contiguous
    assert bad.is_pinned() and not bad.is_contiguous()
    hb = c4.rollout_host(actions=bad, record=("pos", "reward"), chunk=7)
    assert torch.equal(hb["pos"], dd["pos"].cpu())
    with pytest.raises(ValueError):
        c4.rollout_host(actions=np.zeros((3, E, n + 1, 2)), record=("pos",))


@pytest.mark.parametrize("n,grid,delta", [(10, [5, 5], 1.0), (32, [32, 32], 2.5)])
def test_float32_throughput_mode(n, grid, delta):
    """float32 instantiation, fed the oracle's state: stated tolerance is RELATIVE,
    |dr| <= 2e-6 * max(1, |r|) per step on collision-free states (fp32 cannot meet 1e-5 absolute
    on |r| ~ 100), collision counts compared where no pair sits within
    1e-5 of contact."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones
    E = 256
    rng = np.random.default_rng(11)
    env = BatchedDrones(E, n, grid, "O", 2, np.ones(n) * delta, True, dtype=torch.float32, seed=1, warn=False)
    pos = env.pos.cpu().numpy().astype(np.float64)      # lattice start: exactly representable path
    act = rng.uniform(-1, 1, (E, n, 2)).astype(np.float32)
    orc = c_oracle.OracleEnv(E, n, env.end_points, env.d_safety, env.deltas, None, 2, True,
                             c_oracle.default_params(env.collision_weight))
    orc.set_state(pos)
    ref = orc.step(act.astype(np.float64))
    (p32, _), z, r, ncoll, fin, tr = env.step(torch.as_tensor(act, device=env.device))
    torch.cuda.synchronize()
    assert np.abs(p32.cpu().numpy() - ref.pos).max() < 1e-5
    err = np.abs(r.cpu().numpy() - ref.r) / np.maximum(1.0, np.abs(ref.r))
    assert err.max() < 2e-6, err.max()
    assert np.array_equal(ncoll.cpu().numpy(), ref.ncoll)


@pytest.mark.parametrize("n,grid,delta,box", [(10, [5, 5], 1.0, None), (10, [5, 5], 1.0, 1.6), (32, [32, 32], 2.5, None)])
def test_float32_free_running_episode_drift(n, grid, delta, box):
    """float32 throughput mode FREE-RUNNING a whole 200-step episode (fused rollout, no teacher forcing)
    against the float64 oracle on the same action stream, lattice starts and dense (colliding) starts
    (box: all agents in a box of that side).  The stated tolerance of this mode, asserted here:
      positions   |dx| <= 4e-6 * grid  (200 roundings of ~1 float32 ulp at the coordinate's magnitude,
                  random-walk accumulation: 2e-5 on the [5,5] grid, 1.3e-4 on [32,32])
      rewards     |dr| <= 2e-5 * max(1, |r|) on 99.9 % of the rows whose two runs agree on the collision
                  count, and <= 2e-2 * max(1, |r|) on all of them: the barrier term b log(d_safety / d) has
                  slope b / d, so a pair a few 1e-4 from contact turns a 1e-5 position error into a 1e-3
                  reward error (a pair within ~1e-5 of contact flips a 99.9 term: not a rounding error)
      collisions  per-step totals differ on < 0.1 % of the steps.
    The measured maxima are printed (pytest -s) and recorded in profiles/r02/f32_drift.txt by
    tools/gpu_r02_record.sh."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    E, T = 512, 200
    rng = np.random.default_rng(5)
    env = BatchedDrones(E, n, grid, "O", 2, np.ones(n) * delta, True, dtype=torch.float32, seed=2, warn=False)
    if box is not None:
        st, _ = env.get_state()
        st[:, :, 0:2] = rng.uniform(1.0, 1.0 + box, (E, n, 2)).astype(np.float32)
        env.set_state(st, np.zeros(E, np.int32)); env.observe()
    pos0 = env.pos.cpu().numpy().astype(np.float64)
    tab = formation.unit_action_table(16)
    idx = rng.integers(0, 16, (T, E, n))
    act32 = tab.astype(np.float32)[idx]
    orc = c_oracle.OracleEnv(E, n, env.end_points, env.d_safety, env.deltas, None, 2, True,
                             c_oracle.default_params(env.collision_weight), nthreads=8)
    orc.set_state(pos0)
    ref = orc.rollout(act32.astype(np.float64), record_obs=True)
    out = env.rollout(actions=torch.as_tensor(act32, device=env.device), record=("pos", "reward", "ncoll", "finished"))
    torch.cuda.synchronize()
    fin = ref["finished"]
    assert np.array_equal(out["finished"].cpu().numpy(), fin)
    live = fin != 2
    dpos = np.abs(out["pos"].cpu().numpy().astype(np.float64) - ref["pos_tr"])[live].max()
    nc32, nc64 = out["ncoll"].cpu().numpy(), ref["ncoll"]
    same = (nc32 == nc64) & live
    r32, r64 = out["reward"].cpu().numpy().astype(np.float64), ref["r"]
    rel = (np.abs(r32 - r64) / np.maximum(1.0, np.abs(r64)))[same]
    mism = 1.0 - same.sum() / live.sum()
    q999 = float(np.quantile(rel, 0.999))
    print(f"f32 drift n={n} box={box}: max |dpos| {dpos:.2e}, rel |dr| median {np.median(rel):.2e} p99.9 {q999:.2e} "
          f"max {rel.max():.2e}, collision-count mismatches {100 * mism:.4f} % of steps, "
          f"collisions seen {int(nc64[live].sum())}")
    assert dpos <= 4e-6 * max(5.0, float(grid[0]))
    assert q999 <= 2e-5 * max(1.0, grid[0] / 5.0) and rel.max() <= 2e-2
    assert mism < 1e-3
    if box is not None:
        assert nc64[live].sum() > 0


@pytest.mark.parametrize("n,grid,delta,box,segs", [(4, [5, 5], 1.0, None, 0), (5, [5, 5], 1.0, 1.2, 0), (8, [5, 5], 0.8, None, 2),
                                                   (10, [5, 5], 1.0, 1.5, 0), (16, [12, 12], 1.5, None, 4), (20, [12, 12], 1.0, 4.0, 1),
                                                   (32, [32, 32], 2.5, None, 0), (32, [32, 32], 2.5, 6.0, 2), (10, [5, 5], None, None, 0),
                                                   (10, [5, 5], 1.0, None, 1), (10, [5, 5], 1.0, 1.5, 2), (10, [5, 5], 1.0, None, 4),
                                                   (10, [5, 5], 1.0, None, 8), (5, [5, 5], 1.0, None, 1)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_rollout2_every_instantiation_equals_general_kernel(n, grid, delta, box, segs, dtype, monkeypatch):
    """Every agent count the warp-per-segment kernel is instantiated for, in both arithmetic types (the
    float32 instances of n = 16 and 20 use the pair table, the float64 ones the segment layout), every
    segment count, sparse and dense (the dense
    cases overflow the pair list of the segment layout / fill the pair table), delta=None (every agent
    inside every Delta disk), T not a multiple of the chunk, environments that start late in their
    episode and finish inside the call, a second call continuing the first: bit-identical to the
    general rollout_kernel (itself pinned to the oracle and the reference's vectors) in EVERY output."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    E, T = 67, 53
    rng = np.random.default_rng(n)
    deltas = None if delta is None else np.ones(n) * delta
    if segs:                                                       # time segments per environment (0: the plan's choice)
        monkeypatch.setenv("DS_RO2_SEGS", str(segs))
    new = BatchedDrones(E, n, grid, "O", 2, deltas, True, dtype=dtype, seed=3, warn=False)
    monkeypatch.delenv("DS_RO2_SEGS", raising=False)
    monkeypatch.setenv("DS_RO2", "0")
    old = BatchedDrones(E, n, grid, "O", 2, deltas, True, dtype=dtype, seed=3, warn=False)
    monkeypatch.delenv("DS_RO2")
    assert new.rollout_kernel == "ds::rollout2_kernel" and old.rollout_kernel == "ds::rollout_kernel"
    st, _ = new.get_state()
    if box is not None:
        st[:, :, 0:2] = rng.uniform(1.0, 1.0 + box, (E, n, 2))
        st[0, 1, 0:2] = st[0, 0, 0:2]                              # coincident agents
    tt = rng.integers(0, 200, E).astype(np.int32)
    tt[:8] = 199 - rng.integers(0, T, 8)                           # these hit the time limit inside the call
    act = torch.as_tensor(formation.unit_action_table(16)[rng.integers(0, 16, (2 * T, E, n))], device=new.device).to(dtype)
    rec = ("pos", "vel", "reward", "true_reward", "obs", "ncoll", "finished")
    outs = []
    for env in (new, old):
        env.set_state(st, tt); env.observe()
        a = {k: v.clone() for k, v in env.rollout(actions=act[:T], record=rec).items()}
        b = {k: v.clone() for k, v in env.rollout(actions=act[T:], record=rec).items()}     # continues: done / t / agg carry over
        live = {name: getattr(env, name).clone() for name in ("pos", "vel", "rewards", "true_rewards", "z_states", "Ni",
                                                               "n_collisions", "finished", "internal_t", "done")}
        outs.append((a, b, live, env.agg.clone()))
    torch.cuda.synchronize()
    (a1, b1, l1, g1), (a0, b0, l0, g0) = outs
    fin = a0["finished"].cpu().numpy()
    assert (fin == 1).any() and (fin == 2).any()
    for x, y, what in ((a1, a0, "first call"), (b1, b0, "second call")):
        ex = (y["finished"] != 2)                                  # [T,E]: executed steps
        for key in ("pos", "vel", "reward", "true_reward", "z", "Ni"):
            xs, ys = x[key][ex], y[key][ex]
            assert torch.equal(torch.nan_to_num(xs.double()), torch.nan_to_num(ys.double())), f"{what}: {key}"
        assert torch.equal(x["ncoll"][ex], y["ncoll"][ex]) and torch.equal(x["finished"], y["finished"]), what
    for name in l0:
        assert torch.equal(torch.nan_to_num(l1[name].double()), torch.nan_to_num(l0[name].double())), f"live {name}"
    # episode sums: same terms, different reduction order (float32 mode keeps the per-row running sums in float32)
    assert torch.equal(g1[:, 2:], g0[:, 2:])                       # collisions, steps
    if dtype == torch.float64:
        assert torch.allclose(g1, g0, rtol=0, atol=1e-9)
    else:
        assert torch.allclose(g1, g0, rtol=3e-5, atol=1e-3)
    if box is not None:
        assert (a0["ncoll"] > 0).any()


def test_log_mode_rcp_within_stated_difference():
    """DS_LOG_RCP (log(d_safety / d) as -log(d * (1 / d_safety)), no division): rewards within 1e-13 of
    the DS_LOG_DIV path (stated: <= 3.4e-16 per barrier term, times b = 0.01, times <= n terms), every
    integer output identical, on a dense 200-step rollout; step kernel and rollout kernel agree bit
    for bit in this mode too."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation, _lib
    n, E, T = 10, 256, 200
    rng = np.random.default_rng(6)
    act = torch.as_tensor(formation.unit_action_table(16)[rng.integers(0, 16, (T, E, n))], device="cuda")
    outs = {}
    for mode in (_lib.DS_LOG_DIV, _lib.DS_LOG_RCP):
        env = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=8, warn=False)
        st, _ = env.get_state()
        st[:, :, 0:2] = np.random.default_rng(7).uniform(1.0, 3.0, (E, n, 2))
        env.set_state(st, np.zeros(E, np.int32)); env.log_mode = mode; env.observe()
        outs[mode] = {k: v.clone() for k, v in env.rollout(actions=act, record=("pos", "reward", "true_reward", "obs",
                                                                              "ncoll", "finished")).items()}
        if mode == _lib.DS_LOG_RCP:            # stepping reproduces the fused rollout bit for bit
            env.set_state(st, np.zeros(E, np.int32)); env.observe()
            for t in range(3):
                _, _, r, nc, _, tr = env.step(act[t])
                assert torch.equal(r, outs[mode]["reward"][t]) and torch.equal(tr, outs[mode]["true_reward"][t])
    a, b = outs[_lib.DS_LOG_DIV], outs[_lib.DS_LOG_RCP]
    for key in ("pos", "Ni", "ncoll", "finished"):
        assert torch.equal(a[key], b[key]), key
    assert (a["ncoll"] > 0).any()
    assert (a["reward"] - b["reward"]).abs().max().item() <= 1e-13
    assert (a["true_reward"] - b["true_reward"]).abs().max().item() <= 1e-13


@pytest.mark.parametrize("n,k,simplify", [(3, 2, True), (5, 2, True), (5, 3, False), (7, 1, True), (9, 2, False),
                                          (11, 3, True), (33, 2, True), (33, 2, False), (129, 2, True)])
def test_float32_mode_every_kernel_on_odd_shapes(n, k, simplify):
    """float32 instantiations of every kernel on agent counts / neighbour counts / layouts that make the
    shared-memory pieces odd-sized (an 8-byte float2 block in front of 16-byte data was a real bug at
    n = 5): observe, step, both rollout forms, both controllers (one step and fused), device reset --
    all must run (a misaligned access poisons the context) and track the float64 twin on the same
    inputs to float32 accuracy."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    E, T = 37, 11
    grid = [5, 5] if n <= 11 else [32, 32] if n <= 33 else [64, 64]
    rng = np.random.default_rng(n * 10 + k)
    envs = [BatchedDrones(E, n, grid, "O", k, np.ones(n), simplify, dtype=dt, seed=4, warn=False)
            for dt in (torch.float32, torch.float64)]
    tab = formation.unit_action_table(16)
    idx = torch.as_tensor(rng.integers(0, 16, (T, E, n)).astype(np.uint8), device=envs[0].device)
    res = []
    for env in envs:
        env.reset_random(seed=9, stream=1)
        env.observe()
        z0 = env.z_states.double().clone()
        (pos1, _), z1, r1, nc1, *_ = env.step(torch.as_tensor(tab[idx[0].cpu().numpy()], device=env.device).to(env.dtype))
        pos1, r1 = pos1.double().clone(), r1.double().clone()
        a = env.rollout(action_idx=idx, action_table=tab, record=("pos", "reward", "true_reward", "obs", "ncoll", "finished"))
        b = env.rollout(actions=torch.as_tensor(tab[idx.cpu().numpy()], device=env.device).to(env.dtype), record=("pos", "reward"))
        env.reset_random(seed=9, stream=2)
        (pc, vc), *_ = env.step_control("gradient", u_max=0.9)
        pc = pc.double().clone()
        c = env.rollout_control(5, "proportional", record=("pos", "reward"))
        torch.cuda.synchronize()
        res.append((z0, pos1, r1, a["pos"].double(), a["reward"].double(), a["ncoll"], a["finished"], b["pos"].double(),
                    pc, c["pos"].double()))
    f32, f64 = res
    g = float(max(grid))
    # (the observations themselves are not compared: on lattice starts the k nearest are exact ties that
    # float32 rounding breaks differently; positions and rewards do not depend on the choice)
    assert torch.isfinite(f32[0]).all()
    for name, x, y, tol in (("pos after one step", f32[1], f64[1], 1e-6 * g),
                            ("rollout pos", f32[3], f64[3], 2e-5 * g), ("second rollout pos", f32[7], f64[7], 4e-5 * g),
                            ("controller step", f32[8], f64[8], 1e-5 * g), ("controller rollout", f32[9], f64[9], 4e-5 * g)):
        assert torch.isfinite(x).all(), name
        assert (x - y).abs().max().item() <= tol, (name, (x - y).abs().max().item())
    assert torch.equal(f32[6], f64[6])                             # nobody finishes in 11 steps: same flags
    rel = ((f32[4] - f64[4]).abs() / f64[4].abs().clamp_min(1.0))
    assert rel.median().item() <= 1e-5


def test_float32_handle_side_paths():
    """The float32 instantiations of everything around the step (device reset, actors, closed loops,
    returns): consistent with each other bit for bit, and with the fp64 restatements to fp32
    accuracy.  float32 is the throughput mode, not the parity path."""
    from oracle import np_oracle
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    n, E, A, T = 10, 200, 16, 40
    rng = np.random.default_rng(8)
    W = _torch_like_weights(rng, n, 6, A)
    tab = formation.unit_action_table(A)
    a = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, dtype=torch.float32, seed=1, warn=False)
    b = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, dtype=torch.float32, seed=1, warn=False)
    d = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=1, warn=False)          # fp64 twin
    for env in (a, b, d):
        env.load_policy(*W, tab)
        env.reset_random(seed=3, stream=2)
    torch.cuda.synchronize()
    assert torch.equal(a.pos, b.pos) and np.array_equal(a.pos.cpu().numpy(), d.pos.cpu().numpy().astype(np.float32))
    # actors on the f32 observation
    act, idx, probs = a.policy_forward(seed=2, stream=0)
    torch.cuda.synchronize()
    z = a.z_states.cpu().numpy().reshape(E, n, 6)
    for i in range(n):
        assert np.abs(probs[:, i].cpu().numpy() - np_oracle.policy_probs(z[:, i], *[w[i] for w in W])).max() <= 1e-5
    assert act.dtype == torch.float32 and np.array_equal(act.cpu().numpy(), tab.astype(np.float32)[idx.cpu().numpy()])
    # actor-driven episode == forward; step, and == the open-loop rollout on the recorded indices
    out = a.rollout_policy(T, seed=2, stream0=0, record=("pos", "reward", "obs", "finished", "action_idx"))
    for t in range(3):
        act, idx, _ = b.policy_forward(seed=2, stream=t)
        (pos, _), zt, r, *_ = b.step(act)
        torch.cuda.synchronize()
        assert torch.equal(out["pos"][t], pos) and torch.equal(out["reward"][t], r) and torch.equal(out["z"][t], zt)
        assert torch.equal(out["action_idx"][t], idx)
    b.reset_random(seed=3, stream=2)
    ref = b.rollout(action_idx=out["action_idx"], action_table=tab, record=("pos", "reward", "finished"))
    torch.cuda.synchronize()
    assert torch.equal(ref["pos"], out["pos"]) and torch.equal(ref["reward"], out["reward"])
    # returns in f32 against the fp64 restatement on the same (f32) rewards
    base = torch.zeros((T, E, n), dtype=torch.float32, device=a.device)
    ret = a.returns(out["reward"], out["Ni"], out["finished"], discount=0.99, baseline=base)
    torch.cuda.synchronize()
    G, adv, cnt = c_oracle.returns(out["reward"].cpu().numpy().astype(np.float64), out["Ni"].cpu().numpy(),
                                   out["finished"].cpu().numpy(), 0.99, np.zeros((T, E, n)))
    assert np.abs(ret["returns"].cpu().numpy() - G).max() <= 2e-5 * max(1.0, np.abs(G).max())
    assert np.array_equal(ret["count"].cpu().numpy(), cnt)
    # closed loop with the reference's controller: fused == stepped, and close to the fp64 twin
    a.reset_random(seed=3, stream=5); b.reset_random(seed=3, stream=5); d.reset_random(seed=3, stream=5)
    fused = a.rollout_control(5, "gradient", record=("pos", "reward"))
    for t in range(5):
        (pos, _), _, r, *_ = b.step_control("gradient")
        (pos64, _), *_ = d.step_control("gradient")
        torch.cuda.synchronize()
        assert torch.equal(fused["pos"][t], pos) and torch.equal(fused["reward"][t], r)
        assert np.abs(pos.cpu().numpy() - pos64.cpu().numpy()).max() < 1e-4


def test_full_size_properties():
    """BASELINE config 3 (n = 10, E = 4096) at full size: size-independent invariants."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones
    n, E = 10, 4096
    rng = np.random.default_rng(0)
    env = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=0, warn=False)
    p0 = env.pos.cpu().numpy()
    # reset: distinct lattice nodes, no initial overlap, zero velocity
    d = np.linalg.norm(p0[:, :, None] - p0[:, None, :], axis=-1) + np.eye(n) * 9
    assert d.min() >= 0.22 - 1e-12 and int(env.n_collisions.sum()) == 0
    act = rng.uniform(-1, 1, (E, n, 2))
    perm = rng.permutation(E)
    env2 = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=0, warn=False, start_positions=p0[perm])
    for _ in range(3):
        env.step(torch.as_tensor(act, device=env.device))
        env2.step(torch.as_tensor(act[perm], device=env.device))
    torch.cuda.synchronize()
    # environments are independent: permuting the batch permutes the outputs bit for bit
    pt = torch.as_tensor(perm, device=env.device)
    assert torch.equal(env.rewards[pt], env2.rewards) and torch.equal(env.z_states[pt], env2.z_states)
    assert torch.equal(env.n_collisions[pt], env2.n_collisions)
    r, tr = env.rewards.cpu().numpy(), env.true_rewards.cpu().numpy()
    assert (r <= 0).all() and (tr <= r + 1e-12).all()       # barrier terms are >= 0; global has more of them
    assert (env.n_collisions.cpu().numpy() % 2 == 0).all()  # ordered pairs, uniform radius (README.md:46)
    Ni = env.Ni.cpu().numpy()
    assert (Ni[:, :, 0] == np.arange(n)[None]).all() and (Ni >= -1).all() and (Ni < n).all()
    # checksum of checksums against the oracle on the same inputs
    orc = c_oracle.OracleEnv(E, n, env.end_points, env.d_safety, env.deltas, None, 2, True,
                             c_oracle.default_params(env.collision_weight), nthreads=8)
    orc.set_state(p0)
    for _ in range(3):
        ref = orc.step(act)
    assert_close(r, ref.r, FP64_TOL, "full-size reward")
    assert int(env.n_collisions.sum()) == int(ref.ncoll.sum())


FULL = [  # BASELINE.json configs 2-5 at their full batch sizes, whole 200-step episodes
    ("config2", 5, 4096, [5, 5], 1.0, 200),
    ("config3", 10, 4096, [5, 5], 1.0, 200),
    ("config4", 32, 8192, [32, 32], 2.5, 200),
    ("config5", 128, 1024, [64, 64], 1.0, 200),
]


@pytest.mark.parametrize("name,n,E,grid,delta,T", FULL, ids=[c[0] for c in FULL])
def test_rollout_full_size_vs_oracle(name, n, E, grid, delta, T):
    """The benchmarked launch itself -- one fused rollout of the whole BASELINE batch, every output
    recorded -- against the C oracle stepping the same environments with the same action stream."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    rng = np.random.default_rng(42)
    env = BatchedDrones(E, n, grid, "O", 2, np.ones(n) * delta, True, seed=3, warn=False)
    start = env.pos.cpu().numpy().copy()
    tab = formation.unit_action_table(16)
    idx = rng.integers(0, 16, (T, E, n)).astype(np.uint8)
    act = tab[idx]
    orc = c_oracle.OracleEnv(E, n, env.end_points, env.d_safety, env.deltas, None, 2, True,
                             c_oracle.default_params(env.collision_weight), nthreads=16)
    orc.set_state(start)
    ref = orc.rollout(act, record_obs=True)
    rec = ("pos", "vel", "reward", "true_reward", "obs", "ncoll", "finished")
    out = env.rollout(actions=torch.as_tensor(act, device=env.device), record=rec)
    torch.cuda.synchronize()
    assert np.array_equal(out["finished"].cpu().numpy(), ref["finished"])
    live = ref["finished"] != 2
    # EVERY step of the fused rollout against the oracle: positions bit-exact, observations and
    # neighbour lists tie-aware (in slabs of steps to bound the host memory of the comparison)
    for t0 in range(0, T, 25):
        sl = slice(t0, min(T, t0 + 25))
        lv = live[sl]
        assert np.array_equal(out["pos"][sl].cpu().numpy()[lv], ref["pos_tr"][sl][lv]), f"positions, steps {t0}.."
        compare_obs(out["z"][sl].cpu().numpy()[lv], out["Ni"][sl].cpu().numpy()[lv], ref["z_tr"][sl][lv],
                    ref["Ni_tr"][sl][lv], ref["tie_tr"][sl][lv], FP64_TOL, f"{name} steps {t0}..")
    assert_close(out["reward"].cpu().numpy()[live], ref["r"][live], FP64_TOL, "reward")
    assert_close(out["true_reward"].cpu().numpy()[live], ref["true_r"][live], FP64_TOL, "true reward")
    assert np.array_equal(out["ncoll"].cpu().numpy()[live], ref["ncoll"][live]), "collision counts"
    assert_close(out["agg"].cpu().numpy(), ref["agg"], 1e-8, "episode aggregates")
    assert_close(env.pos.cpu().numpy(), ref["pos"], 0.0, "final positions (bit-exact)")
    assert np.array_equal(env.internal_t.cpu().numpy(), ref["t"])
    # last recorded observation == the oracle's (tie aware) == live buffers == a fresh observe()
    z_last, Ni_last = out["z"][T - 1].clone(), out["Ni"][T - 1].clone()
    compare_obs(z_last.cpu().numpy(), Ni_last.cpu().numpy(), ref["z"], ref["Ni"], orc.tie, FP64_TOL, name)
    assert torch.equal(z_last, env.z_states) and torch.equal(Ni_last, env.Ni)
    env.observe()
    torch.cuda.synchronize()
    assert torch.equal(z_last, env.z_states) and torch.equal(Ni_last, env.Ni)
    # size-independent invariants of the recorded trajectory
    assert torch.equal(out["vel"], torch.as_tensor(act, device=env.device))         # v <- u (:238)
    Ni = out["Ni"].cpu().numpy()
    assert (Ni[..., 0] == np.arange(n)).all() and (Ni >= -1).all() and (Ni < n).all()
    assert (out["ncoll"].cpu().numpy() % 2 == 0).all()


@pytest.mark.parametrize("n,E,grid,delta,box,k,hetero", [
    (10, 300, [5, 5], 1.0, 1.0, 2, False),     # every pair near: the work list overflows, rows evaluate themselves
    (24, 64, [8, 8], 1.5, 2.0, 3, True),       # dense, heterogeneous Delta, k = 3
    (200, 6, [128, 128], 1.0, 30.0, 2, False), # 128 < n <= 256: near masks in local memory
    (300, 4, [128, 128], 1.0, 30.0, 2, False), # n > 256: 1024-thread CTAs
])
def test_rollout_dense_and_large(n, E, grid, delta, box, k, hetero):
    """Rollout paths the BASELINE configs do not reach: list overflow, dense frames with many
    collisions and ties, large n.  Checked against the oracle's episode loop."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones
    T = 12
    rng = np.random.default_rng(7 * n)
    deltas = rng.uniform(0.3, delta, n) if hetero else np.ones(n) * delta
    env = BatchedDrones(E, n, grid, "O", k, deltas, False, seed=2, warn=False)
    pos = rng.uniform(0, box, (E, n, 2))
    state = np.concatenate([pos, np.zeros((E, n, 2)), np.full((E, n, 1), 0.1)], 2)
    env.set_state(state, np.zeros(E, np.int32))
    act = rng.uniform(-1, 1, (T, E, n, 2))
    orc = c_oracle.OracleEnv(E, n, env.end_points, env.d_safety, env.deltas, None, k, False,
                             c_oracle.default_params(env.collision_weight), nthreads=8)
    orc.set_state(pos)
    ref = orc.rollout(act)
    out = env.rollout(actions=torch.as_tensor(act, device=env.device),
                      record=("reward", "true_reward", "obs", "ncoll", "finished"))
    torch.cuda.synchronize()
    assert ref["ncoll"].sum() > 0 or box > 20
    assert np.array_equal(out["finished"].cpu().numpy(), ref["finished"])
    assert_close(out["reward"].cpu().numpy(), ref["r"], FP64_TOL, "reward")
    assert_close(out["true_reward"].cpu().numpy(), ref["true_r"], FP64_TOL, "true reward")
    assert np.array_equal(out["ncoll"].cpu().numpy(), ref["ncoll"])
    assert_close(env.pos.cpu().numpy(), ref["pos"], 0.0, "final positions")
    # the fused rollout's observations == the oracle's (tie aware) == the step kernel's on the final state
    z_last, Ni_last = out["z"][T - 1].clone(), out["Ni"][T - 1].clone()
    compare_obs(z_last.cpu().numpy(), Ni_last.cpu().numpy(), ref["z"], ref["Ni"], orc.tie, FP64_TOL, "final obs")
    env.observe()
    torch.cuda.synchronize()
    assert torch.equal(z_last, env.z_states) and torch.equal(Ni_last, env.Ni)


@pytest.mark.parametrize("name", ["returns_n5_seed0", "returns_n5_seed3_g0.9", "returns_n8_seed1"])
def test_returns_vs_reference_golden(name):
    """ds_returns on the reference's own episode: returns recorded from the reference's
    TrainedAgent.benchmark_cirtic and advantage sums of its train_NN loop, bit-exact."""
    import os
    from scalable_collision_avoidance_rl_b200 import BatchedDrones
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    n, T = int(g["n"]), g["reward"].shape[0]
    env = BatchedDrones(1, n, [5, 5], "O", int(g["k"]), np.ones(n), True, seed=0, warn=False)
    dev = env.device
    out = env.returns(torch.as_tensor(g["reward"][:, None, :], device=dev),
                      torch.as_tensor(g["Ni"][:, None], device=dev),
                      torch.as_tensor(g["finished"][:, None], device=dev),
                      discount=float(g["discount"]), baseline=torch.as_tensor(g["baseline"][:, None, :], device=dev))
    torch.cuda.synchronize()
    assert np.array_equal(out["returns"].cpu().numpy()[:, 0], g["returns"])
    assert np.array_equal(out["advantage"].cpu().numpy()[:, 0], g["advantage"])
    assert np.array_equal(out["count"].cpu().numpy()[:, 0], (g["Ni"] >= 0).sum(-1))


@pytest.mark.parametrize("n,E,k,T", [(10, 4096, 2, 200), (5, 333, 2, 61), (32, 65, 2, 40), (7, 50, 4, 33), (128, 9, 2, 10)])
def test_returns_on_rollout_vs_oracle(n, E, k, T):
    """ds_rollout -> ds_returns on the device, against the oracle's restatement of the reference's
    host-side walk over its ExperienceBuffers; includes episodes that end inside the rollout."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    grid = [5, 5] if n <= 10 else ([32, 32] if n <= 32 else [64, 64])
    rng = np.random.default_rng(n + E)
    env = BatchedDrones(E, n, grid, "O", k, np.ones(n), True, seed=4, warn=False)
    st, _ = env.get_state()
    env.set_state(st, rng.integers(max(0, 200 - 2 * T), 200, E).astype(np.int32))   # some hit the time limit
    tab = formation.unit_action_table(16)
    act = tab[rng.integers(0, 16, (T, E, n))]
    ro = env.rollout(actions=torch.as_tensor(act, device=env.device), record=("reward", "obs", "finished"))
    base = torch.as_tensor(rng.standard_normal((T, E, n)), device=env.device)
    out = env.returns(ro["reward"], ro["Ni"], ro["finished"], discount=0.95, baseline=base)
    torch.cuda.synchronize()
    fin = ro["finished"].cpu().numpy()
    assert (fin == 2).any() and (fin == 1).any()
    G, adv, cnt = c_oracle.returns(ro["reward"].cpu().numpy(), ro["Ni"].cpu().numpy(), fin, 0.95,
                                   base.cpu().numpy())
    assert np.array_equal(out["returns"].cpu().numpy(), G)
    assert np.array_equal(out["advantage"].cpu().numpy(), adv)
    assert np.array_equal(out["count"].cpu().numpy(), cnt)


def test_returns_recipe_pairs_steps_with_pre_step_neighbour_lists():
    """The documented recipe (rollout with "obs_pre" -> returns_from_rollout) against the reference's
    episode loop restated step by step: `Ni = env.Ni` is read BEFORE `env.step` (train_problem.py:84-96)
    and train_NN sums the advantages of step t over that list (SAC_agents.py:333-346).  The stepping
    side records the pre-step lists itself; the rollout side must rebuild them from the post-step
    trajectory and the pre-call snapshot."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    n, E, T, k = 6, 37, 40, 2
    rng = np.random.default_rng(11)
    mk = lambda: BatchedDrones(E, n, [5, 5], "O", k, np.ones(n) * 0.9, True, seed=9, warn=False)
    a, b = mk(), mk()
    st, _ = a.get_state()
    st[:, :, 0:2] = rng.uniform(1.0, 3.5, (E, n, 2))                 # dense: neighbour lists change from step to step
    tt = rng.integers(150, 200, E).astype(np.int32)                  # some episodes end inside the rollout
    for env in (a, b):
        env.set_state(st, tt); env.observe()
    act = formation.unit_action_table(16)[rng.integers(0, 16, (T, E, n))]
    act_d = torch.as_tensor(act, device=a.device)
    # the reference's loop, one step at a time: lists read before the step
    Ni_pre, z_pre, rew, fin = [], [], [], []
    done = np.zeros(E, bool)
    for t in range(T):
        Ni_pre.append(a.Ni.cpu().numpy().copy()); z_pre.append(a.z_states.cpu().numpy().copy())
        _, _, r, _, f, _ = a.step(act_d[t])
        torch.cuda.synchronize()
        rew.append(r.cpu().numpy().copy())
        code = np.where(done, 2, f.cpu().numpy())
        fin.append(code.astype(np.uint8)); done |= f.cpu().numpy().astype(bool)
    Ni_pre, z_pre, rew, fin = map(np.stack, (Ni_pre, z_pre, rew, fin))
    V = rng.standard_normal((T, E, n))
    G, adv, cnt = c_oracle.returns(rew, Ni_pre, fin, 0.97, V)
    # the fused rollout + the documented recipe
    ro = b.rollout(actions=act_d, record=("reward", "obs", "obs_pre", "finished"))
    zp, Np = b.pre_step_observations(ro)
    out = b.returns_from_rollout(ro, discount=0.97, critic=lambda z: torch.as_tensor(V, device=b.device))
    torch.cuda.synchronize()
    ex = fin != 2
    assert (fin == 2).any() and np.array_equal(ro["finished"].cpu().numpy(), fin)
    assert np.array_equal(Np.cpu().numpy()[ex], Ni_pre[ex]) and np.array_equal(zp.cpu().numpy()[ex], z_pre[ex], equal_nan=True)
    assert not np.array_equal(ro["Ni"].cpu().numpy()[ex], Ni_pre[ex]), "post-step lists differ: the shift matters here"
    assert np.array_equal(out["returns"].cpu().numpy()[ex], G[ex])
    assert np.array_equal(out["advantage"].cpu().numpy()[ex], adv[ex])
    assert np.array_equal(out["count"].cpu().numpy()[ex], cnt[ex])


@pytest.mark.parametrize("name,ctrl", [("control_gradient_n5", "gradient"), ("control_gradient_n10", "gradient"),
                                       ("control_proportional_n8", "proportional")])
def test_closed_loop_controller_vs_reference_golden(name, ctrl):
    """ds_step_control free-running a whole closed-loop episode that the reference's own controller
    drove in the reference's environment: actions and state bit-exact at every step, rewards to
    1e-9, collision counts and the finishing step exact.  E copies of the episode run side by side."""
    import os
    from scalable_collision_avoidance_rl_b200 import BatchedDrones
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    n, E, T = int(g["n"]), 7, g["action"].shape[0]
    env = BatchedDrones(E, n, [float(x) for x in g["grid"]], "O", 2, np.ones(n) * float(g["delta"]), True,
                        seed=0, warn=False)
    assert np.array_equal(env.end_points, g["end_points"]) and np.array_equal(env.d_safety, g["d_safety"])
    env.set_state(np.broadcast_to(g["state_in"][0], (E, n, 5)).copy(), np.zeros(E, np.int32))
    for t in range(T):
        (pos, vel), z, r, ncoll, fin, tr = env.step_control(ctrl, u_max=float(g["u_max"]))
        torch.cuda.synchronize()
        for e in (0, E - 1):
            assert np.array_equal(vel[e].cpu().numpy(), g["action"][t]), f"action t={t}"
            assert np.array_equal(pos[e].cpu().numpy(), g["state"][t][:, 0:2]), f"state t={t}"
            assert_close(r[e].cpu().numpy(), g["r"][t], FP64_TOL, f"r t={t}")
            assert int(ncoll[e]) == int(g["ncoll"][t]) and bool(fin[e]) == bool(g["finished"][t])
    assert bool(g["finished"][-1])


@pytest.mark.parametrize("ctrl,n,E,grid", [("gradient", 10, 4096, [5, 5]), ("proportional", 5, 300, [5, 5]),
                                          ("gradient", 32, 64, [32, 32])])
def test_closed_loop_rollout_equals_stepping(ctrl, n, E, grid):
    """ds_rollout_control (T closed-loop steps in one launch) == T x ds_step_control, bit for bit:
    trajectories, finished codes (episodes end inside the call: the controllers reach the goals),
    final state, done flags; episode sums to 1e-9."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones
    T = 130
    a = BatchedDrones(E, n, grid, "O", 2, np.ones(n), True, seed=6, warn=False)
    b = BatchedDrones(E, n, grid, "O", 2, np.ones(n), True, seed=6, warn=False)
    assert torch.equal(a.pos, b.pos)
    if n == 32:                                   # far goals: let the 200-step limit end the episodes
        st, _ = a.get_state()
        t0 = np.random.default_rng(0).integers(100, 190, E).astype(np.int32)
        a.set_state(st, t0); b.set_state(st, t0)
    rec = ("pos", "vel", "reward", "true_reward", "obs", "ncoll", "finished")
    out = a.rollout_control(T, ctrl, u_max=1.0, record=rec)
    torch.cuda.synchronize()
    fin_tr = out["finished"].cpu().numpy()
    done = np.zeros(E, bool)
    sums = np.zeros((E, 4))
    for t in range(T):
        before = (b.pos.clone(), b.vel.clone(), b.internal_t.clone())
        (pos, vel), z, r, nc, fin, tr = b.step_control(ctrl, u_max=1.0)
        torch.cuda.synchronize()
        d = torch.as_tensor(done, device=b.device)
        b.pos[d] = before[0][d]; b.vel[d] = before[1][d]; b.internal_t[d] = before[2][d]   # finished envs freeze
        lv = torch.as_tensor(~done, device=b.device)
        assert np.array_equal(fin_tr[t][~done], fin.cpu().numpy()[~done]) and (fin_tr[t][done] == 2).all()
        for key, val in (("pos", pos), ("vel", vel), ("reward", r), ("true_reward", tr), ("z", z), ("Ni", b.Ni), ("ncoll", nc)):
            assert torch.equal(out[key][t][lv], val[lv]), f"{key} t={t}"
        live = ~done
        sums[live, 0] += r.cpu().numpy()[live].mean(1); sums[live, 1] += tr.cpu().numpy()[live].mean(1)
        sums[live, 2] += nc.cpu().numpy()[live]; sums[live, 3] += 1
        done |= fin.cpu().numpy().astype(bool) & live
    assert done.any()
    assert torch.equal(a.pos, b.pos) and torch.equal(a.internal_t, b.internal_t)
    assert np.array_equal(out["done"].cpu().numpy().astype(bool), done)
    assert_close(out["agg"].cpu().numpy(), sums, 1e-9, "episode sums")


@pytest.mark.parametrize("name,key,fn", [("control_gradient_n5", "action", "gradient"),
                                         ("control_proportional_n8", "action", "proportional"),
                                         ("control_dense_n7", "gradient", "gradient"),
                                         ("control_dense_n7", "proportional", "proportional")])
def test_dropin_controller_functions_vs_reference_golden(name, key, fn):
    """drone_env.gradient_control / proportional_control of the drop-in module (ds_control on the
    device) return, bit for bit, what the reference's functions returned on the recorded states
    (NaN patterns at exact contact / on the goal included)."""
    import os
    import drone_env
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    n = int(g["n"])
    env = drone_env.drones(n, 0, [float(x) for x in g["grid"]], "O", 2, np.ones(n) * float(g["delta"]), True)
    assert np.array_equal(env.end_points, g["end_points"]) and np.array_equal(env.d_safety, g["d_safety"])
    frames = range(0, g["state_in"].shape[0], max(1, g["state_in"].shape[0] // 12))
    for f in frames:
        if fn == "gradient":
            act = drone_env.gradient_control(g["state_in"][f], env, u_max=float(g["u_max"]))
        else:
            act = drone_env.proportional_control(g["state_in"][f], env)
        assert len(act) == n and act[0].shape == (2,)
        assert np.array_equal(np.array(act), g[key][f], equal_nan=True), f"frame {f}"


def test_dropin_controllers_take_the_radii_from_the_state():
    """The reference's controllers read every agent's radius from state[:, 4] (drone_env.py:632,643,664),
    not from env.drone_radius: a state with an edited radius column is served with those radii
    (oracle with the same radii: bit-exact; and different from the answer for the default radii)."""
    import drone_env
    n = 6
    rng = np.random.default_rng(5)
    env = drone_env.drones(n, 0, [5, 5], "O", 2, np.ones(n), True)
    state = env.state.copy()
    state[:, 0:2] = rng.uniform(1.0, 2.2, (n, 2))
    radii = rng.uniform(0.05, 0.25, n)
    state[:, 4] = radii
    for mode, um in ((c_oracle.CTRL_GRADIENT, 0.7), (c_oracle.CTRL_PROPORTIONAL, 1.0)):
        want = c_oracle.control(mode, state[None, :, 0:2], env.end_points, env.d_safety, radii, um)[0]
        base = c_oracle.control(mode, state[None, :, 0:2], env.end_points, env.d_safety, None, um)[0]
        if mode == c_oracle.CTRL_GRADIENT:
            got = np.array(drone_env.gradient_control(state, env, u_max=um))
            assert not np.array_equal(want, base)                  # the radii matter for the barrier gradient
        else:
            got = np.array(drone_env.proportional_control(state, env))
        assert np.array_equal(got, want, equal_nan=True), mode


def test_controllers_vs_oracle_dense_batch():
    """Both controllers on dense random batches (many pairs inside d_safety) against the oracle,
    through the closed-loop step: the action taken is left in vel."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones
    for n, E, box in ((10, 513, 2.0), (33, 40, 6.0), (128, 6, 20.0)):
        rng = np.random.default_rng(n)
        grid = [5, 5] if n <= 10 else [32, 32] if n <= 33 else [64, 64]
        env = BatchedDrones(E, n, grid, "O", 2, np.ones(n), True, seed=1, warn=False)
        pos = rng.uniform(0, box, (E, n, 2))
        state = np.concatenate([pos, np.zeros((E, n, 2)), np.full((E, n, 1), 0.1)], 2)
        for mode, ctrl, um in ((c_oracle.CTRL_GRADIENT, "gradient", 0.8), (c_oracle.CTRL_PROPORTIONAL, "proportional", 1.0)):
            env.set_state(state, np.zeros(E, np.int32))
            want = c_oracle.control(mode, pos, env.end_points, env.d_safety, None, um)
            (p2, vel), *_ = env.step_control(ctrl, u_max=um)
            torch.cuda.synchronize()
            assert np.array_equal(vel.cpu().numpy(), want, equal_nan=True), (n, ctrl)
            assert np.array_equal(p2.cpu().numpy(), pos + 0.05 * want, equal_nan=True)
    with pytest.raises(ValueError):
        env.step_control("pid")


@pytest.mark.parametrize("n,E,grid", [(10, 64, [5, 5]), (33, 20, [32, 32]), (128, 5, [64, 64]), (300, 3, [128, 128])])
def test_device_reset_matches_restatement(n, E, grid):
    """ds_reset_random against oracle/np_oracle.reset_random (same Philox counters, same unbiased
    index, same redraw rule): start positions bit-exact; zero velocity, t = 0; the observation of
    the start state is the one ds_observe / the oracle computes."""
    from oracle import np_oracle
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    env = BatchedDrones(E, n, grid, "O", 2, np.ones(n), True, seed=0, warn=False)
    env.step(torch.zeros((E, n, 2), dtype=torch.float64, device=env.device) + 0.3)      # dirty the state
    env.reset_random(seed=0x1234567890ABCDEF, stream=5)
    torch.cuda.synchronize()
    d0, d1 = formation.lattice_shape(grid)
    want = np_oracle.reset_random(E, n, d0, d1, formation.LATTICE_PITCH, 0x1234567890ABCDEF, 5)
    assert np.array_equal(env.pos.cpu().numpy(), want)
    assert not env.vel.any() and not env.internal_t.any() and not env.finished.any() and not env.done.any()
    orc = c_oracle.OracleEnv(E, n, env.end_points, env.d_safety, env.deltas, None, 2, True,
                             c_oracle.default_params(env.collision_weight))
    orc.set_state(want)
    ref = orc.observe()
    assert_close(env.rewards.cpu().numpy(), ref.r, FP64_TOL, "reward of the start state")
    compare_obs(env.z_states.cpu().numpy(), env.Ni.cpu().numpy(), ref.z, ref.Ni, ref.tie, FP64_TOL, "start obs")


@pytest.mark.parametrize("n,grid,simplify,dtype", [(10, [5, 5], True, torch.float64), (32, [32, 32], False, torch.float64),
                                                   (5, [5, 5], True, torch.float32)])
def test_device_reset_fused_with_observation_equals_two_launches(n, grid, simplify, dtype, monkeypatch):
    """ds_reset_random draws the start and evaluates its observation in ONE launch (n <= 32, k = 2):
    every buffer bit-identical to the two-launch form (reset_random_kernel, then ds_observe)."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones
    E = 300
    mk = lambda: BatchedDrones(E, n, grid, "O", 2, np.ones(n), simplify, dtype=dtype, seed=1, warn=False)
    a, b = mk(), mk()
    a.reset_random(seed=77, stream=5)
    monkeypatch.setenv("DS_RESET_FUSED", "0")
    b.reset_random(seed=77, stream=5)
    torch.cuda.synchronize()
    for name in ("pos", "vel", "rewards", "true_rewards", "z_states", "Ni", "n_collisions", "finished", "internal_t"):
        x, y = getattr(a, name), getattr(b, name)
        assert torch.equal(x, y) or (x.is_floating_point() and torch.equal(torch.nan_to_num(x), torch.nan_to_num(y))), name


def test_device_reset_statistics():
    """The distribution random.sample gives (drone_env.py:204): every ordered pick is uniform over
    the lattice and the picks of an environment are distinct -- chi-square over 2^17 environments;
    streams are independent of each other and reproducible."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    n, E, grid = 5, 1 << 17, [5, 5]
    env = BatchedDrones(E, n, grid, "O", 2, np.ones(n), True, seed=0, warn=False)
    d0, d1 = formation.lattice_shape(grid)
    L = d0 * d1
    env.reset_random(seed=11, stream=0)
    p0 = env.pos.clone()
    nodes = np.rint(p0.cpu().numpy() / formation.LATTICE_PITCH).astype(np.int64)
    flat = nodes[..., 0] * d1 + nodes[..., 1]
    assert flat.min() >= 0 and flat.max() < L
    srt = np.sort(flat, axis=1)
    assert (srt[:, 1:] != srt[:, :-1]).all()                                   # distinct within an environment
    assert int(env.n_collisions.sum()) == 0                                    # pitch 0.22 > 2 radii: no initial overlap
    for col in range(n):                                                        # each ordered pick uniform
        cnt = np.bincount(flat[:, col], minlength=L)
        chi2 = ((cnt - E / L) ** 2 / (E / L)).sum()
        assert chi2 < L + 6 * np.sqrt(2 * L), (col, chi2)                       # ~6 sigma of chi2(L-1)
    # first two picks jointly: P(second == first + 1) = 1/L * (L-1)/(L-1)... simple independence proxy
    same_row = (flat[:, 0] // d1 == flat[:, 1] // d1).mean()
    assert abs(same_row - (d1 - 1) / (L - 1)) < 5 * np.sqrt(0.05 / E) + 2e-3
    env.reset_random(seed=11, stream=0)
    assert torch.equal(env.pos, p0)                                            # reproducible
    env.reset_random(seed=11, stream=1)
    assert (env.pos != p0).any(dim=-1).float().mean() > 0.9                    # a new stream moves (almost) everyone
    env.reset_random(seed=12, stream=0)
    assert (env.pos != p0).any(dim=-1).float().mean() > 0.9


def _torch_like_weights(rng, n, in_dim, A):
    u = lambda shape, fan: rng.uniform(-1, 1, shape).astype(np.float32) / np.float32(np.sqrt(fan))
    return (u((n, 300, in_dim), in_dim), u((n, 300), in_dim), u((n, 300, 300), 300), u((n, 300), 300),
            u((n, A, 300), 300), u((n, A), 300))


def test_policy_forward_vs_reference_golden():
    """ds_policy_forward (tcgen05, 3xTF32) with the reference's own pretrained actors
    (softmax8_n5, agents 0 and 1) on observations of a recorded episode: probabilities within 1e-5
    of what the reference's DiscreteSoftmaxNN.forward returned (fp32), sampled index = inverse CDF
    of the Philox uniform on the device's own probabilities, action = action_list[index]."""
    import os
    from oracle import np_oracle
    from scalable_collision_avoidance_rl_b200 import BatchedDrones
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "policynet_n5_agents01.npz"))
    T, A = g["probs"].shape[0], int(g["n_actions"])
    use = [0, 1, 0]                                        # three agents (k = 2 needs n >= 3): networks 0, 1, 0
    n = len(use)
    env = BatchedDrones(T, n, [5, 5], "O", 2, np.ones(n), True, seed=0, warn=False)
    W = [np.stack([g[f"{w}_{a}"] for a in use]) for w in ("W1", "b1", "W2", "b2", "W3", "b3")]
    env.load_policy(*W, g["action_list"])
    z = torch.as_tensor(g["z"][:, use].reshape(T, n, 3, 2), device=env.device).contiguous()
    act, idx, probs = env.policy_forward(seed=99, stream=7, z=z)
    torch.cuda.synchronize()
    probs = probs.cpu().numpy(); idx = idx.cpu().numpy()
    assert np.abs(probs - g["probs"][:, use]).max() <= 1e-5
    want = np.stack([np_oracle.policy_probs(g["z"][:, a], *[g[f"{w}_{a}"] for w in ("W1", "b1", "W2", "b2", "W3", "b3")])
                     for a in use], 1)
    assert np.abs(probs - want).max() <= 2e-6
    assert np.array_equal(idx, np_oracle.policy_sample(probs, T, n, 99, 7))
    assert np.array_equal(act.cpu().numpy(), g["action_list"][idx])


@pytest.mark.parametrize("n,E,k,simplify,A", [(10, 1000, 2, True, 16), (5, 129, 2, False, 8), (32, 300, 1, True, 5),
                                              (5, 1, 4, True, 16), (9, 7, 7, True, 3), (4, 128, 0, True, 1),
                                              # more tiles than SMs: persistent CTAs walk several tiles and
                                              # cross from one agent's network to the next
                                              (10, 2567, 2, True, 16), (5, 15360, 2, True, 8)])
def test_policy_forward_vs_fp32_restatement(n, E, k, simplify, A):
    """Random torch-style weights (layers 2 and 3 scaled x3 to spread the logits), one network per
    agent, E not a multiple of the 128-row tile: probabilities against the fp32 NumPy restatement
    (3e-5: two fp32 summation orders over 300 terms on logits of magnitude ~10 differ by that much;
    the unscaled reference networks are asserted at 1e-5 / 2e-6 above), index and action as
    specified; the action then drives a step."""
    from oracle import np_oracle
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    rng = np.random.default_rng(n * 7 + A)
    grid = [5, 5] if n <= 10 else [32, 32]
    env = BatchedDrones(E, n, grid, "O", k, np.ones(n), simplify, seed=2, warn=False)
    in_dim = (k + 1) * (2 if simplify else 5)
    W = _torch_like_weights(rng, n, in_dim, A)
    W = tuple(w * np.float32(3) if j in (2, 4) else w for j, w in enumerate(W))       # spread the logits
    tab = formation.unit_action_table(A)
    env.load_policy(*W, tab)
    act, idx, probs = env.policy_forward(seed=5, stream=1)                           # on the live observation
    torch.cuda.synchronize()
    z = env.z_states.cpu().numpy().reshape(E, n, in_dim)
    probs = probs.cpu().numpy(); idx = idx.cpu().numpy()
    for i in range(n):
        want = np_oracle.policy_probs(z[:, i], *[w[i] for w in W])
        assert np.abs(probs[:, i] - want).max() <= 3e-5, i
    assert np.abs(probs.sum(-1) - 1).max() < 1e-5
    assert np.array_equal(idx, np_oracle.policy_sample(probs, E, n, 5, 1))
    assert np.array_equal(act.cpu().numpy(), tab[idx])
    p0 = env.pos.clone()
    env.step(act)
    torch.cuda.synchronize()
    assert np.array_equal(env.pos.cpu().numpy(), p0.cpu().numpy() + 0.05 * tab[idx])


@pytest.mark.parametrize("n,E,grid,simplify,A", [(5, 300, [5, 5], True, 8), (10, 129, [5, 5], False, 16)])
def test_rollout_policy_equals_forward_then_step_and_oracle(n, E, grid, simplify, A):
    """ds_rollout_policy (the episode loop of train_problem.py:82-104 with the actors on the device)
    == T x (policy_forward; step) bit for bit on every executed step, with ds_rollout's finished
    codes / done / episode sums; the open-loop rollout on the recorded indices reproduces it; the
    C oracle, driven with the recorded actions, agrees on states and rewards (fp64 tolerance)."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    T, k = 70, 2
    rng = np.random.default_rng(n + A)
    in_dim = (k + 1) * (2 if simplify else 5)
    W = _torch_like_weights(rng, n, in_dim, A)
    W = tuple(w * np.float32(3) if j in (2, 4) else w for j, w in enumerate(W))
    tab = formation.unit_action_table(A)
    envs = [BatchedDrones(E, n, grid, "O", k, np.ones(n), simplify, seed=9, warn=False) for _ in range(3)]
    a, b, c = envs
    t0 = np.random.default_rng(1).integers(150, 199, E).astype(np.int32)     # episodes end inside the call
    st, _ = a.get_state()
    for env in envs:
        env.load_policy(*W, tab)
        env.set_state(st, t0)
        env.observe()
    rec = ("pos", "vel", "reward", "true_reward", "obs", "ncoll", "finished", "action_idx", "probs")
    out = a.rollout_policy(T, seed=11, stream0=100, record=rec)
    torch.cuda.synchronize()
    fin_tr = out["finished"].cpu().numpy()
    idx_tr = out["action_idx"].cpu().numpy()
    done = np.zeros(E, bool)
    sums = np.zeros((E, 4))
    for t in range(T):
        act, idx, probs = b.policy_forward(seed=11, stream=100 + t)
        before = (b.pos.clone(), b.vel.clone(), b.internal_t.clone(), b.z_states.clone())
        (pos, vel), z, r, nc, fin, tr = b.step(act)
        torch.cuda.synchronize()
        lv = torch.as_tensor(~done, device=b.device)
        assert np.array_equal(fin_tr[t][~done], fin.cpu().numpy()[~done]) and (fin_tr[t][done] == 2).all()
        for key, val in (("pos", pos), ("vel", vel), ("reward", r), ("true_reward", tr), ("z", z), ("Ni", b.Ni),
                         ("ncoll", nc), ("action_idx", idx), ("probs", probs)):
            assert torch.equal(out[key][t][lv], val[lv]), f"{key} t={t}"
        assert torch.equal(out["vel"][t][lv], torch.as_tensor(tab[idx_tr[t]], device=b.device)[lv])
        live = ~done
        sums[live, 0] += r.cpu().numpy()[live].mean(1); sums[live, 1] += tr.cpu().numpy()[live].mean(1)
        sums[live, 2] += nc.cpu().numpy()[live]; sums[live, 3] += 1
        done |= fin.cpu().numpy().astype(bool) & live
        dprev = torch.as_tensor(~live, device=b.device)                       # finished before this step: freeze
        b.pos[dprev] = before[0][dprev]; b.vel[dprev] = before[1][dprev]; b.internal_t[dprev] = before[2][dprev]
        b.z_states[dprev] = before[3][dprev]
    assert done.all()
    assert np.array_equal(out["done"].cpu().numpy().astype(bool), done)
    assert_close(out["agg"].cpu().numpy(), sums, 1e-9, "episode sums")
    assert torch.equal(a.pos, b.pos) and torch.equal(a.internal_t, b.internal_t)
    # open loop on the recorded indices: the time-parallel rollout kernel gives the same episode
    ref = c.rollout(action_idx=out["action_idx"], action_table=tab, record=("pos", "reward", "true_reward", "ncoll", "finished"))
    torch.cuda.synchronize()
    ex = torch.as_tensor(fin_tr != 2, device=a.device)
    assert torch.equal(ref["finished"], out["finished"])
    for key in ("pos", "reward", "true_reward", "ncoll"):
        assert torch.equal(ref[key][ex], out[key][ex]), key
    # the C oracle on the same actions
    orc = c_oracle.OracleEnv(E, n, a.end_points, a.d_safety, np.asarray(a.deltas, np.float64).reshape(-1),
                             a.drone_radius, k, simplify, c_oracle.default_params(a.collision_weight))
    orc.set_state(st[..., 0:2], st[..., 2:4], t0)
    executed = fin_tr != 2
    for t in range(T):
        res = orc.step(tab[idx_tr[t]].astype(np.float64))
        m = executed[t]
        assert_close(out["pos"][t].cpu().numpy()[m], res.pos[m], FP64_TOL, f"oracle pos t={t}")
        assert_close(out["reward"][t].cpu().numpy()[m], res.r[m], FP64_TOL, f"oracle r t={t}")
        assert np.array_equal(out["ncoll"][t].cpu().numpy()[m], res.ncoll[m])
        assert np.array_equal(fin_tr[t][m], res.finished[m].astype(np.uint8))


def test_rollout_policy_cuda_graph_replay():
    """The whole actor-driven episode captured in one CUDA graph; replays with a new device-resident
    seed give new draws, the same seed reproduces the episode bit for bit."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    n, E, A, T = 5, 64, 8, 20
    rng = np.random.default_rng(3)
    W = _torch_like_weights(rng, n, 6, A)
    tab = formation.unit_action_table(A)
    env = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=4, warn=False)
    env.load_policy(*W, tab)
    st, t0 = env.get_state()
    seed = torch.tensor([5], dtype=torch.uint64, device=env.device)
    out = {}
    rec = ("pos", "reward", "finished", "action_idx")

    def restart():
        env.set_state(st, t0); env.observe(); env.done.zero_(); env.agg.zero_()

    restart()
    env.rollout_policy(T, stream0=0, record=rec, out=out, seed_tensor=seed)      # warm-up: attributes, scratch
    torch.cuda.synchronize()
    first = {k_: v.clone() for k_, v in out.items()}
    restart()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            env.rollout_policy(T, stream0=0, record=rec, out=out, seed_tensor=seed)
    g.replay(); torch.cuda.synchronize()
    for key in rec:
        assert torch.equal(out[key], first[key]), key
    restart()
    seed.fill_(6)
    g.replay(); torch.cuda.synchronize()
    assert not torch.equal(out["action_idx"], first["action_idx"])
    restart()
    seed.fill_(5)
    g.replay(); torch.cuda.synchronize()
    for key in rec:
        assert torch.equal(out[key], first[key]), key


def test_policy_sampling_statistics():
    """Same observation in 2^15 environments: the sampled indices follow the probabilities
    (chi-square), different streams give different draws, same stream reproduces."""
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, formation
    n, E, A = 3, 1 << 15, 8
    rng = np.random.default_rng(1)
    env = BatchedDrones(E, n, [5, 5], "O", 2, np.ones(n), True, seed=0, warn=False)
    env.load_policy(*_torch_like_weights(rng, n, 6, A), formation.unit_action_table(A))
    z = torch.as_tensor(np.broadcast_to(rng.uniform(-2, 2, (1, n, 3, 2)), (E, n, 3, 2)).copy(), device=env.device)
    _, idx, probs = env.policy_forward(seed=8, stream=0, z=z)
    torch.cuda.synchronize()
    idx0 = idx.clone()
    p = probs[0].cpu().numpy().astype(np.float64)
    for i in range(n):
        cnt = np.bincount(idx[:, i].cpu().numpy(), minlength=A)
        exp = p[i] * E
        keep = exp > 5
        chi2 = ((cnt[keep] - exp[keep]) ** 2 / exp[keep]).sum()
        assert chi2 < keep.sum() + 6 * np.sqrt(2 * keep.sum()) + 5, (i, chi2)
    _, idx1, _ = env.policy_forward(seed=8, stream=0, z=z)
    assert torch.equal(idx1, idx0)
    _, idx2, _ = env.policy_forward(seed=8, stream=1, z=z)
    assert (idx2 != idx0).float().mean() > 0.3


def test_error_behaviour():
    from scalable_collision_avoidance_rl_b200 import BatchedDrones, DroneStepError
    with pytest.raises(DroneStepError):
        BatchedDrones(4, 3, [5, 5], "O", 3, None, True, warn=False)        # k >= n
    with pytest.raises(ValueError):
        BatchedDrones(4, 3, [5, 5], "X", 2, None, True, warn=False)        # unknown formation
    env = BatchedDrones(4, 3, [5, 5], "O", 2, None, True, warn=False)
    with pytest.raises(TypeError):
        env.step(np.zeros((4, 3, 2)))
    env.log_mode = 7
    with pytest.raises(DroneStepError):
        env.observe()

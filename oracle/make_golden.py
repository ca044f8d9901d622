#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- record golden vectors from the UNMODIFIED reference.

Runs ``/root/reference/drone_env.py`` (shimmed import, ``oracle/ref_harness.py``)
in the build container and writes small ``.npz`` fixtures to ``tests/golden/``.
The reference ships no tests or golden vectors of its own,
so these recordings are what pins the oracle -- and through it the CUDA path --
to the reference.  Re-run with:  python oracle/make_golden.py

Every fixture stores, per recorded step t:
  state_in[t] (n,5), t_in[t], actions[t] (n,2)  ->  state[t] (n,5), z[t], Ni[t]
  (padded -1), r[t], true_r[t], ncoll[t], finished[t], tie[t] (rows of d_ij whose
  k+2 smallest entries contain an exact tie: np.argsort is unstable there).
plus the constructor outputs end_points, d_safety, deltas and the observation
computed by init_agents (z0, Ni0, tie0) for state0.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_harness import import_reference, import_reference_policy_stack, REFERENCE_ROOT  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ref = import_reference("drone_env")


def pad_ni(Ni, k):
    out = np.full((len(Ni), k + 1), -1, np.int32)
    for i, lst in enumerate(Ni):
        out[i, : len(lst)] = lst
    return out


def tie_rows(env):
    d_ij, _, _, _ = env.distance_data(env.state, env.deltas, env.d_safety)
    k = env.k_closest
    s = np.sort(d_ij, axis=1)[:, : min(env.n_agents, k + 2)]
    return (s[:, 1:] == s[:, :-1]).any(1)


def unit_actions(rng, n, n_dirs=16):
    a = rng.integers(0, n_dirs, size=n)
    return np.stack([np.cos(a / n_dirs * 2 * np.pi), np.sin(a / n_dirs * 2 * np.pi)], 1)


class Recorder:
    def __init__(self, env, name, meta):
        self.env, self.name, self.meta = env, name, meta
        self.rows = {k: [] for k in ("state_in", "t_in", "actions", "state", "z", "Ni", "r",
                                     "true_r", "ncoll", "finished", "tie")}
        self.state0 = env.state.copy()
        self.z0 = np.array(env.z_states)
        self.Ni0 = pad_ni(env.Ni, env.k_closest)
        self.tie0 = tie_rows(env)

    def step(self, actions):
        env = self.env
        R = self.rows
        R["state_in"].append(env.state.copy())
        R["t_in"].append(env.internal_t)
        R["actions"].append(np.array(actions, dtype=np.float64))
        state, z, r, ncoll, fin, true_r = env.step([np.asarray(a) for a in actions])
        assert isinstance(ncoll, (int, np.integer))
        R["state"].append(state.copy())
        R["z"].append(np.array(z))
        R["Ni"].append(pad_ni(env.Ni, env.k_closest))
        R["r"].append(np.array(r)); R["true_r"].append(np.array(true_r))
        R["ncoll"].append(int(ncoll)); R["finished"].append(bool(fin))
        R["tie"].append(tie_rows(env))
        return fin

    def save(self):
        env = self.env
        arrs = {k: np.array(v) for k, v in self.rows.items()}
        arrs["ncoll"] = arrs["ncoll"].astype(np.int64)
        arrs["finished"] = arrs["finished"].astype(np.uint8)
        arrs["t_in"] = arrs["t_in"].astype(np.int32)
        np.savez_compressed(
            os.path.join(OUT, self.name + ".npz"),
            end_points=env.end_points.copy(), d_safety=env.d_safety.copy(),
            deltas=np.asarray(env.deltas, np.float64).copy(), radius=env.drone_radius.copy(),
            state0=self.state0, z0=self.z0, Ni0=self.Ni0, tie0=self.tie0,
            n=env.n_agents, k=env.k_closest, simplify=int(env.simplify_zstate),
            collision_weight=float(env.collision_weight), grid=np.array(env.grid, np.float64),
            dt=ref.dt, max_time_steps=ref.max_time_steps,
            deltas_in=self.meta.get("deltas_in", np.array([])),
            note=self.meta.get("note", ""), **arrs)
        T = len(self.rows["ncoll"])
        print(f"{self.name:34s} n={env.n_agents:3d} k={env.k_closest} T={T:3d} "
              f"coll={int(arrs['ncoll'].sum()):5d} ties={int(arrs['tie'].sum()):4d} "
              f"fin={int(arrs['finished'].sum())}")


def make_env(n, grid, deltas, k=2, simplify=True, seed=0, cw=None):
    random.seed(seed); np.random.seed(seed)
    env = ref.drones(n_agents=n, n_obstacles=0, grid=list(grid), end_formation="O",
                     k_closest=k, deltas=deltas, simplify_zstate=simplify)
    if cw is not None:
        env.collision_weight = cw      # mutated after construction, train_problem.py:31
    return env


def free_run(name, n, grid, delta, T, k=2, simplify=True, seed=0, cw=None, n_dirs=16, note=""):
    deltas = None if delta is None else np.ones(n) * delta
    env = make_env(n, grid, deltas, k, simplify, seed, cw)
    rec = Recorder(env, name, dict(deltas_in=np.array([]) if deltas is None else deltas, note=note))
    rng = np.random.default_rng(1234 + seed)
    for _ in range(T):
        if rec.step(unit_actions(rng, n, n_dirs)):
            break
    rec.save()


def dense_teacher_forced(name, n, grid, deltas, T, box, k=2, simplify=True, seed=0, cw=None, note=""):
    """Fresh random dense state before every step -> many collisions."""
    env = make_env(n, grid, deltas, k, simplify, seed, cw)
    rec = Recorder(env, name, dict(deltas_in=deltas, note=note))
    rng = np.random.default_rng(99 + seed)
    for t in range(T):
        env.state[:, 0:2] = rng.uniform(0, box, size=(n, 2))
        env.state[:, 2:4] = rng.standard_normal((n, 2))
        env.internal_t = int(rng.integers(0, 205))
        rec.step(rng.uniform(-1, 1, size=(n, 2)))
    rec.save()


def edge_cases(name):
    """Known-answer facts of SURVEY.md section 4, one injected state per step."""
    n = 4
    env = make_env(n, [5, 5], np.ones(n) * 1.0, 2, False, 3)
    rec = Recorder(env, name, dict(deltas_in=np.ones(n), note="edge cases"))
    zero = np.zeros((n, 2))
    xF = env.end_points.reshape(n, 2)
    far = np.array([[0.5, 0.5], [0.5, 3.0], [3.0, 0.5], [3.0, 3.0]])
    cases = []
    s = far.copy(); s[1] = s[0] + [0.15, 0.0]; cases.append(s)        # 0.15 apart: d=-0.05
    s = far.copy(); s[0] = [0, 0]; s[1] = [0.2, 0.0]; cases.append(s)  # exactly 0.2 apart -> -1e-6
    s = far.copy(); s[2] = xF[2]; cases.append(s)                      # on goal -> NaN ghost
    s = far.copy(); s[3] = s[0]; cases.append(s)                       # coincident agents
    s = xF + 0.05; cases.append(s)                                     # all within 0.2 of goal -> finished
    s = far.copy(); s[1] = s[0] + [0.5, 0.0]; s[2] = s[0] + [0.0, 0.7]; cases.append(s)  # 2 in range
    s = far.copy(); s[1] = s[0] + [1.2, 0.0]; cases.append(s)          # exactly on Delta boundary-ish
    for s in cases:
        env.state[:, 0:2] = s
        env.state[:, 2:4] = 0.0
        env.internal_t = 5
        rec.step(zero)
    env.internal_t = 198
    rec.step(zero)   # t=198 -> not finished by time
    rec.step(zero)   # t=199 -> finished by time
    rec.save()


def policy_episode(name, seed):
    """BASELINE config 1: n=5, trained softmax8_n5 policy, one <=200-step episode."""
    import torch
    utils, _ = import_reference_policy_stack()
    actors = torch.load(os.path.join(REFERENCE_ROOT, "models", "final", "softmax8_n5-A2Cactors.pth"),
                        weights_only=False)
    n = 5
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    env = ref.drones(n_agents=n, n_obstacles=0, grid=[5, 5], end_formation="O",
                     deltas=np.ones(n) * 1.0, simplify_zstate=True)
    env.collision_weight = 0.2
    rec = Recorder(env, name, dict(deltas_in=np.ones(n), note="softmax8_n5 policy, seed %d" % seed))
    fin = False
    ret = 0.0
    while not fin:
        acts = [actors[i].sample_action(env.z_states[i].flatten(), env.Ni[i]) for i in range(n)]
        fin = rec.step(acts)
        ret += float(np.mean(rec.rows["r"][-1]))
    print(f"   policy episode seed {seed}: return {ret:.4f}")
    rec.save()


def ctor_table():
    """d_safety / end_points / clipped deltas for the BASELINE grids."""
    rows = {}
    for n, grid, delta in [(5, [5, 5], 1.0), (10, [5, 5], 1.0), (32, [32, 32], 2.5),
                           (128, [64, 64], 1.0), (32, [5, 5], 2.5), (128, [5, 5], 1.0),
                           (7, [4, 9], 3.0), (5, [5, 5], None)]:
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            env = make_env(n, grid, None if delta is None else np.ones(n) * delta, seed=n)
        key = f"n{n}_g{grid[0]}x{grid[1]}_d{delta}"
        rows[key + "_end_points"] = env.end_points.copy()
        rows[key + "_d_safety"] = env.d_safety.copy()
        rows[key + "_deltas"] = np.asarray(env.deltas).copy()
        rows[key + "_state0"] = env.state.copy()
        rows[key + "_lss"] = np.array([env.local_state_space, env.local_action_space])
    np.savez_compressed(os.path.join(OUT, "ctor_table.npz"), **rows)
    print("ctor_table", len(rows))


def main():
    os.makedirs(OUT, exist_ok=True)
    free_run("free_n5_g5_d1.0", 5, [5, 5], 1.0, 200, seed=0)
    free_run("free_n10_g5_d1.0", 10, [5, 5], 1.0, 200, seed=1)
    free_run("free_n32_g32_d2.5", 32, [32, 32], 2.5, 40, seed=2)
    free_run("free_n128_g64_d1.0", 128, [64, 64], 1.0, 3, seed=3)
    free_run("free_n5_g5_dNone_full", 5, [5, 5], None, 30, simplify=False, seed=4,
             note="deltas=None -> Delta == d_safety, mass ties")
    free_run("free_n8_g5_d1.0_k3_cw0.5", 8, [5, 5], 1.0, 60, k=3, simplify=False, seed=5, cw=0.5)
    free_run("free_n6_g5_d0.8_k1", 6, [5, 5], 0.8, 60, k=1, simplify=True, seed=6)
    dense_teacher_forced("dense_n6_g5", 6, [5, 5], np.ones(6) * 1.0, 60, 1.5, seed=7)
    dense_teacher_forced("dense_n10_g5_full", 10, [5, 5], np.ones(10) * 0.9, 40, 2.0,
                         simplify=False, seed=8)
    dense_teacher_forced("dense_n6_hetero", 6, [6, 6], np.array([0.3, 0.6, 0.9, 1.2, 1.5, 5.0]),
                         60, 2.0, simplify=False, seed=9,
                         note="heterogeneous deltas: N_delta uses deltas[j]")
    dense_teacher_forced("dense_n33_g32", 33, [32, 32], np.ones(33) * 2.5, 6, 6.0, seed=10)
    edge_cases("edge_cases_n4")
    policy_episodes()
    ctor_table()


def policy_episodes(seeds=range(10)):
    """BASELINE config 1 as SURVEY section 8(d) states it: seeds 0..9."""
    os.makedirs(OUT, exist_ok=True)
    for seed in seeds:
        policy_episode(f"policy_n5_seed{seed}", seed)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--policy-only":      # add seeds without touching the other fixtures
        policy_episodes([int(x) for x in sys.argv[2:]] or range(10))
    else:
        main()

#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- golden vectors for the baseline controllers (SURVEY.md section 8f
row 3), recorded from the UNMODIFIED reference: closed-loop episodes in which the reference's own
`gradient_control` / `proportional_control` (drone_env.py:612-679) drive the reference's `drones`
environment.  Stored per step: state_in, action, and the step's outputs; plus dense teacher-forced
states (many pairs inside d_safety, exact contacts) for the controllers alone.
Re-run with:  python oracle/make_golden_control.py"""
from __future__ import annotations

import contextlib
import io
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_harness import import_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ref = import_reference("drone_env")


def make_env(n, grid, delta, seed):
    random.seed(seed); np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        env = ref.drones(n_agents=n, n_obstacles=0, grid=list(grid), end_formation="O",
                         deltas=np.ones(n) * delta, simplify_zstate=True)
    env.collision_weight = 0.2
    return env


def closed_loop(name, ctrl, n, grid, delta, seed, T=200):
    env = make_env(n, grid, delta, seed)
    fn = getattr(ref, ctrl)
    rows = {k: [] for k in ("state_in", "action", "state", "r", "true_r", "ncoll", "finished")}
    fin = False
    while not fin and len(rows["r"]) < T:
        state = env.state.copy()
        act = fn(env.state, env)                                   # the reference's controller
        rows["state_in"].append(state); rows["action"].append(np.array(act, np.float64))
        st, _, r, nc, fin, tr = env.step(act)
        rows["state"].append(st.copy()); rows["r"].append(np.array(r)); rows["true_r"].append(np.array(tr))
        rows["ncoll"].append(int(nc)); rows["finished"].append(bool(fin))
    arrs = {k: np.array(v) for k, v in rows.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), n=n, grid=np.array(grid, np.float64), delta=delta,
                        controller=ctrl, end_points=env.end_points.copy(), d_safety=env.d_safety.copy(),
                        deltas=np.asarray(env.deltas, np.float64), u_max=1.0, **arrs)
    print(f"{name}: T={len(rows['r'])} coll={int(arrs['ncoll'].sum())} finished={bool(arrs['finished'][-1])} "
          f"|u|max={np.abs(arrs['action']).max():.3f}")


def dense_states(name, n, grid, delta, seed, frames=60, box=1.6):
    env = make_env(n, grid, delta, seed)
    rng = np.random.default_rng(seed)
    S, G, Pc = [], [], []
    for f in range(frames):
        env.state[:, 0:2] = rng.uniform(0, box, (n, 2))
        if f == 0:
            env.state[1, 0:2] = env.state[0, 0:2] + [0.2, 0.0]      # exact contact: dij = 0 -> division by zero
        if f == 1:
            env.state[2, 0:2] = env.end_points.reshape(n, 2)[2]     # on goal
        S.append(env.state.copy())
        with np.errstate(all="ignore"):
            G.append(np.array(ref.gradient_control(env.state, env, u_max=0.7)))
            Pc.append(np.array(ref.proportional_control(env.state, env)))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), n=n, grid=np.array(grid, np.float64), delta=delta,
                        end_points=env.end_points.copy(), d_safety=env.d_safety.copy(),
                        deltas=np.asarray(env.deltas, np.float64), u_max=0.7, state_in=np.array(S),
                        gradient=np.array(G), proportional=np.array(Pc))
    print(f"{name}: frames={frames} nan={int(np.isnan(np.array(G)).sum())}")


if __name__ == "__main__":
    closed_loop("control_gradient_n5", "gradient_control", 5, [5, 5], 1.0, 0)
    closed_loop("control_gradient_n10", "gradient_control", 10, [5, 5], 1.0, 1)
    closed_loop("control_proportional_n8", "proportional_control", 8, [5, 5], 1.0, 2)
    dense_states("control_dense_n7", 7, [5, 5], 1.0, 3)

"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of ``oracle/drone_oracle.c``.

Builds (``gcc -O2 -ffp-contract=off``) and loads the plain-C float64
restatement of the reference's ``drones.step()`` path.  Consumers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs -- as the checker / CPU baseline, never as the
product path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "drone_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_OUT = os.path.join(_OUT_DIR, "libdroneoracle.so")

_lib = None


class OracleParams(ctypes.Structure):
    _fields_ = [
        ("dt", ctypes.c_double),
        ("collision_weight", ctypes.c_double),
        ("goal_tol", ctypes.c_double),
        ("sentinel", ctypes.c_double),
        ("zero_eps", ctypes.c_double),
        ("ghost_factor", ctypes.c_double),
        ("max_time_steps", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
    ]


def default_params(collision_weight: float = 0.2, dt: float = 0.05,
                   max_time_steps: int = 200) -> OracleParams:
    """Constants of the reference (drone_env.py:29-30,72,251,320,330,386)."""
    return OracleParams(dt, collision_weight, 0.2, 9.99e3, -10 ** -6, 1.1, max_time_steps, 0)


def build(force: bool = False) -> str:
    os.makedirs(_OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(_OUT)
            and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC)):
        return _OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-Wall", "-Wextra", "-o", _OUT, _SRC, "-lm", "-lpthread"]
    subprocess.run(cmd, check=True)
    return _OUT


def lib():
    global _lib
    if _lib is None:
        path = _OUT if (os.path.exists(_OUT) and not os.path.exists(_SRC)) else build()
        _lib = ctypes.CDLL(path)
        _lib.oracle_step_batch.restype = ctypes.c_int
        _lib.oracle_rollout_batch.restype = ctypes.c_int
        _lib.oracle_rollout_batch_obs.restype = ctypes.c_int
        _lib.oracle_returns_batch.restype = ctypes.c_int
        _lib.oracle_control_batch.restype = ctypes.c_int
    return _lib


def _p(a, typ=None):
    if a is None:
        return ctypes.c_void_p(0)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


@dataclass
class StepOut:
    pos: np.ndarray        # [E,n,2]
    vel: np.ndarray        # [E,n,2]
    r: np.ndarray          # [E,n]
    true_r: np.ndarray     # [E,n]
    z: np.ndarray          # [E,n,k+1,cols]
    Ni: np.ndarray         # [E,n,k+1] int32, -1 padded
    ncoll: np.ndarray      # [E] int32
    finished: np.ndarray   # [E] uint8
    t: np.ndarray          # [E] int32
    tie: np.ndarray        # [E,n] uint8


class OracleEnv:
    """E independent environments stepped by the C oracle (float64)."""

    def __init__(self, n_envs, n_agents, end_points, d_safety, deltas, radius=None,
                 k_closest=2, simplify_zstate=True, params: OracleParams | None = None,
                 nthreads=1):
        self.E, self.n, self.k = int(n_envs), int(n_agents), int(k_closest)
        self.simplify = bool(simplify_zstate)
        self.cols = 2 if self.simplify else 5
        self.xF = np.ascontiguousarray(np.asarray(end_points, np.float64).reshape(self.n, 2))
        self.d_safety = np.ascontiguousarray(np.asarray(d_safety, np.float64).reshape(self.n))
        self.deltas = np.ascontiguousarray(np.asarray(deltas, np.float64).reshape(self.n))
        self.radius = (np.full(self.n, 0.1) if radius is None
                       else np.ascontiguousarray(np.asarray(radius, np.float64).reshape(self.n)))
        self.params = params or default_params()
        self.nthreads = nthreads
        E, n, k = self.E, self.n, self.k
        self.pos = np.zeros((E, n, 2)); self.vel = np.zeros((E, n, 2))
        self.r = np.zeros((E, n)); self.true_r = np.zeros((E, n))
        self.z = np.zeros((E, n, k + 1, self.cols)); self.Ni = np.full((E, n, k + 1), -1, np.int32)
        self.ncoll = np.zeros(E, np.int32); self.finished = np.zeros(E, np.uint8)
        self.t = np.zeros(E, np.int32); self.tie = np.zeros((E, n), np.uint8)

    def set_state(self, pos, vel=None, t=None):
        self.pos[...] = np.asarray(pos, np.float64).reshape(self.E, self.n, 2)
        self.vel[...] = 0.0 if vel is None else np.asarray(vel, np.float64).reshape(self.E, self.n, 2)
        if t is not None:
            self.t[...] = t

    def _call(self, act):
        rc = lib().oracle_step_batch(
            self.E, self.n, self.k, int(self.simplify), ctypes.byref(self.params),
            _p(self.pos), _p(self.vel), _p(self.radius), _p(act), _p(self.xF),
            _p(self.d_safety), _p(self.deltas), _p(self.r), _p(self.true_r), _p(self.z),
            _p(self.Ni), _p(self.ncoll), _p(self.finished), _p(self.t), _p(self.tie),
            int(self.nthreads))
        if rc != 0:
            raise ValueError(f"oracle_step_batch failed rc={rc}")
        return StepOut(self.pos.copy(), self.vel.copy(), self.r.copy(), self.true_r.copy(),
                       self.z.copy(), self.Ni.copy(), self.ncoll.copy(), self.finished.copy(),
                       self.t.copy(), self.tie.copy())

    def observe(self) -> StepOut:
        """rewards() on the current state (reference drone_env.py:208)."""
        return self._call(None)

    def step(self, actions) -> StepOut:
        act = np.ascontiguousarray(np.asarray(actions, np.float64).reshape(self.E, self.n, 2))
        return self._call(act)

    def rollout(self, act_stream, record=True, record_obs=False):
        """T steps from act_stream[T,E,n,2]; returns dict of aggregates/trajectories.  record_obs:
        also the observation every executed step returned (z_tr, Ni_tr, tie_tr, pos_tr)."""
        act = np.ascontiguousarray(np.asarray(act_stream, np.float64))
        T = act.shape[0]
        assert act.shape == (T, self.E, self.n, 2)
        E, n = self.E, self.n
        done = np.zeros(E, np.uint8)
        agg = np.zeros((E, 4))
        r_tr = np.zeros((T, E, n)) if record else None
        t_tr = np.zeros((T, E, n)) if record else None
        c_tr = np.zeros((T, E), np.int32) if record else None
        f_tr = np.zeros((T, E), np.uint8) if record else None
        k = self.k
        z_tr = np.zeros((T, E, n, k + 1, self.cols)) if record_obs else None
        Ni_tr = np.full((T, E, n, k + 1), -1, np.int32) if record_obs else None
        tie_tr = np.zeros((T, E, n), np.uint8) if record_obs else None
        pos_tr = np.zeros((T, E, n, 2)) if record_obs else None
        rc = lib().oracle_rollout_batch_obs(
            E, n, self.k, int(self.simplify), T, ctypes.byref(self.params),
            _p(self.pos), _p(self.vel), _p(self.radius), _p(act), _p(self.xF),
            _p(self.d_safety), _p(self.deltas), _p(self.r), _p(self.true_r), _p(self.z),
            _p(self.Ni), _p(self.ncoll), _p(self.finished), _p(self.t), _p(done),
            _p(agg), _p(r_tr), _p(t_tr), _p(c_tr), _p(f_tr), _p(z_tr), _p(Ni_tr), _p(tie_tr), _p(pos_tr),
            int(self.nthreads))
        if rc != 0:
            raise ValueError(f"oracle_rollout_batch failed rc={rc}")
        return dict(agg=agg, r=r_tr, true_r=t_tr, ncoll=c_tr, finished=f_tr, done=done,
                    pos=self.pos.copy(), vel=self.vel.copy(), z=self.z.copy(), Ni=self.Ni.copy(),
                    t=self.t.copy(), z_tr=z_tr, Ni_tr=Ni_tr, tie_tr=tie_tr, pos_tr=pos_tr)


def returns(reward_tr, Ni_tr, finished_tr, discount, baseline=None):
    """Monte-Carlo returns and Delta-neighbourhood advantage sums of a recorded rollout
    (reference SAC_agents.py:304-310,333-345): reward_tr [T,E,n], Ni_tr [T,E,n,k+1] (-1 padded),
    finished_tr [T,E] (0/1/2).  Returns (G [T,E,n], adv [T,E,n], cnt [T,E,n] int32)."""
    r = np.ascontiguousarray(np.asarray(reward_tr, np.float64))
    Ni = np.ascontiguousarray(np.asarray(Ni_tr, np.int32))
    fin = np.ascontiguousarray(np.asarray(finished_tr, np.uint8))
    T, E, n = r.shape
    k = Ni.shape[-1] - 1
    assert Ni.shape == (T, E, n, k + 1) and fin.shape == (T, E)
    base = None if baseline is None else np.ascontiguousarray(np.asarray(baseline, np.float64))
    G = np.zeros((T, E, n)); adv = np.zeros((T, E, n)); cnt = np.zeros((T, E, n), np.int32)
    rc = lib().oracle_returns_batch(E, n, k, T, ctypes.c_double(discount), _p(r), _p(Ni), _p(fin), _p(base),
                                    _p(G), _p(adv), _p(cnt))
    if rc != 0:
        raise ValueError(f"oracle_returns_batch failed rc={rc}")
    return G, adv, cnt


CTRL_PROPORTIONAL, CTRL_GRADIENT = 1, 2


def control(mode, pos, end_points, d_safety, radius=None, u_max=1.0):
    """Baseline controllers of the reference (drone_env.py:612-679) for a batch of environments:
    pos [E,n,2] -> actions [E,n,2].  mode: CTRL_PROPORTIONAL | CTRL_GRADIENT."""
    pos = np.ascontiguousarray(np.asarray(pos, np.float64))
    E, n, _ = pos.shape
    xF = np.ascontiguousarray(np.asarray(end_points, np.float64).reshape(n, 2))
    ds = np.ascontiguousarray(np.asarray(d_safety, np.float64).reshape(n))
    rad = np.full(n, 0.1) if radius is None else np.ascontiguousarray(np.asarray(radius, np.float64).reshape(n))
    act = np.zeros((E, n, 2))
    rc = lib().oracle_control_batch(int(mode), E, n, _p(pos), _p(rad), _p(xF), _p(ds), ctypes.c_double(u_max), _p(act))
    if rc != 0:
        raise ValueError(f"oracle_control_batch failed rc={rc}")
    return act

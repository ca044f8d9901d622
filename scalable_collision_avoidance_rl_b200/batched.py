"""BatchedDrones: E independent drone environments stepped by one fused CUDA kernel.

Host-side mirror of the reference's ``drone_env.drones`` for a batch of
environments (reference drone_env.py:53-401).  Python/PyTorch here only own the
device memory and the CUDA stream; every step / observation / rollout goes
through the C ABI of ``libdronestep.so`` (include/dronestep.h).  There is no CPU
implementation: constructing a BatchedDrones without a CUDA device raises.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, formation

_REAL = {torch.float64: 8, torch.float32: 4}


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


class BatchedDrones:
    """E environments x n agents, env-major SoA tensors on one GPU.

    Tensors (all caller-visible, updated in place by every call):
      pos, vel [E,n,2] | rewards, true_rewards [E,n] | z_states [E,n,k+1,cols]
      Ni [E,n,k+1] int32 (-1 padded) | n_collisions [E] int32 | finished [E] uint8
      internal_t [E] int32
    """

    def __init__(self, n_envs: int, n_agents: int, grid, end_formation: str = "O", k_closest: int = 2,
                 deltas=None, simplify_zstate: bool = False, dtype=torch.float64, device=None,
                 seed: int | None = None, start_positions=None, warn=True, constants=None):
        if not torch.cuda.is_available():
            raise _lib.DroneStepError(
                "BatchedDrones needs a CUDA device: the step path exists only as sm_100a kernels "
                "(libdronestep.so); there is no CPU fallback")
        if dtype not in _REAL:
            raise ValueError("dtype must be torch.float64 (parity) or torch.float32 (throughput)")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = dtype
        self.n_envs, self.n_agents, self.k_closest = int(n_envs), int(n_agents), int(k_closest)
        self.simplify_zstate = bool(simplify_zstate)
        self.grid = list(grid)
        self.collision_weight = 0.2                       # drone_env.py:72 (mutable attribute)
        self.dt = 0.05                                    # drone_env.py:29
        self.max_time_steps = 200                         # drone_env.py:30
        self.log_mode = _lib.DS_LOG_DIV
        self.drone_radius = np.ones(self.n_agents) * formation.DRONE_RADIUS
        if constants is not None:
            # (end_points, d_safety, deltas[, radii]) given verbatim, e.g. by drones.rewards()
            if len(constants) > 3 and constants[3] is not None:
                self.drone_radius = np.asarray(constants[3], np.float64).reshape(-1).copy()
                if self.drone_radius.shape[0] != self.n_agents:
                    raise ValueError("radii must have one entry per agent")
            self.end_points = np.asarray(constants[0], np.float64).reshape(-1, 1)
            self.d_safety = np.asarray(constants[1], np.float64).reshape(-1)
            self.deltas, clipped = np.asarray(constants[2], np.float64).reshape(-1), False
        else:
            self.end_points = formation.end_formation(end_formation, self.n_agents, self.grid)
            self.d_safety = formation.safety_distances(self.end_points, self.drone_radius)
            self.deltas, clipped = formation.clip_deltas(deltas, self.d_safety)
        if clipped and warn:
            print("Some deltas are greater than the final minimum distance between end positions. "
                  "Using minimum distance between end positions for those cases instead.",
                  f"deltas = {self.deltas}")
        self.local_state_space = formation.local_state_space(self.k_closest, self.simplify_zstate)
        self.local_action_space = formation.DIM
        self.cols = 2 if self.simplify_zstate else 5
        self._rng = np.random.default_rng(seed)

        E, n, k = self.n_envs, self.n_agents, self.k_closest
        dev, dt_ = self.device, self.dtype
        # the step's state and results live in ONE device allocation (256-byte aligned pieces), so
        # that step_host() brings the whole 6-tuple back in a single transfer (ds_step_host_block)
        self._layout, off = {}, 0
        for name, shape, dtype in (("pos", (E, n, 2), dt_), ("vel", (E, n, 2), dt_),
                                   ("z", (E, n, k + 1, self.cols), dt_), ("r", (E, n), dt_), ("tr", (E, n), dt_),
                                   ("Ni", (E, n, k + 1), torch.int32), ("nc", (E,), torch.int32),
                                   ("fin", (E,), torch.uint8)):
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            self._layout[name] = (off, nbytes, shape, dtype)
            off += (nbytes + 255) // 256 * 256
        self._block = torch.zeros(off, dtype=torch.uint8, device=dev)
        self._hblock = None

        def piece(block, name):
            o, nb, shape, dtype = self._layout[name]
            return block[o:o + nb].view(dtype).view(shape)

        self._piece = piece
        self.pos, self.vel = piece(self._block, "pos"), piece(self._block, "vel")
        self.rewards, self.true_rewards = piece(self._block, "r"), piece(self._block, "tr")
        self.z_states = piece(self._block, "z")
        self.Ni = piece(self._block, "Ni"); self.Ni.fill_(-1)
        self.n_collisions, self.finished = piece(self._block, "nc"), piece(self._block, "fin")
        self.internal_t = torch.zeros(E, dtype=torch.int32, device=dev)
        # per-episode bookkeeping in ONE allocation: a reset clears it with a single fill
        self._episode = torch.zeros(E * 32 + E, dtype=torch.uint8, device=dev)
        self.agg = self._episode[:E * 32].view(torch.float64).view(E, 4)
        self.done = self._episode[E * 32:]
        self._agg_sum = torch.zeros(5, dtype=torch.float64, device=dev)

        # constants -> device (ds_create)
        self._c_xF = np.ascontiguousarray(self.end_points.reshape(-1), np.float64)
        self._c_ds = np.ascontiguousarray(self.d_safety, np.float64)
        self._c_dl = np.ascontiguousarray(np.asarray(self.deltas, np.float64).reshape(-1))
        self._c_rad = np.ascontiguousarray(self.drone_radius, np.float64)
        if self._c_dl.shape[0] != n:
            raise ValueError("deltas must have one entry per agent")
        cfg = _lib.ds_config(E, n, k, int(self.simplify_zstate), _REAL[dt_], self.device.index or 0,
                             self._c_xF.ctypes.data, self._c_ds.ctypes.data, self._c_dl.ctypes.data,
                             self._c_rad.ctypes.data)
        self._h = ctypes.c_void_p()
        _lib.check(self.lib.ds_create(ctypes.byref(cfg), ctypes.byref(self._h)), "ds_create")
        self._io = _lib.ds_buffers(self.pos.data_ptr(), self.vel.data_ptr(), self.rewards.data_ptr(),
                                   self.true_rewards.data_ptr(), self.z_states.data_ptr(),
                                   self.Ni.data_ptr(), self.n_collisions.data_ptr(),
                                   self.finished.data_ptr(), self.internal_t.data_ptr())
        self._pinned = {}
        self.reset(start_positions)

    # ------------------------------------------------------------------ plumbing
    @property
    def rollout_kernel(self) -> str:
        """The kernel rollout() launches for this configuration (ds_rollout_kernel_name)."""
        return self.lib.ds_rollout_kernel_name(self._h).decode()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                pol = getattr(self, "_pol", None)
                if pol is not None and pol.value:
                    self.lib.ds_policy_destroy(pol)
                self.lib.ds_destroy(h)
            except Exception:
                pass
            self._h = None

    def _params(self):
        return _lib.ds_params(self.dt, float(self.collision_weight), 0.2, 9.99e3, -10 ** -6, 1.1,
                              int(self.max_time_steps), int(self.log_mode))

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _pin(self, name, shape, dtype):
        buf = self._pinned.get(name)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._pinned[name] = buf
        return buf

    # ------------------------------------------------------------------ reset / state
    def reset(self, start_positions=None):
        """drones.reset() (drone_env.py:98-102) for all E environments: distinct lattice
        nodes, zero velocity, t = 0, then the observation of the start state."""
        E, n = self.n_envs, self.n_agents
        if start_positions is None:
            start_positions = formation.sample_start_batched(E, n, self.grid, self._rng)
        sp = np.ascontiguousarray(np.asarray(start_positions, np.float64).reshape(E, n, 2))
        p = self._params()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ds_reset(self._h, sp.ctypes.data_as(ctypes.c_void_p), ctypes.byref(p),
                                         ctypes.byref(self._io), self._stream()), "ds_reset")
        self._episode.zero_()

    def reset_random(self, seed=0, stream=0):
        """drones.reset() with the start drawn on the device (ds_reset_random): distinct lattice
        nodes per environment from Philox(seed; environment, stream) -- the distribution of the
        reference's random.sample (drone_env.py:193-205), not its stream.  No host round trip;
        pass a new `stream` (e.g. the episode number) for every reset."""
        d0, d1 = formation.lattice_shape(self.grid)
        p = self._params()
        _lib.check(self.lib.ds_reset_random(self._h, ctypes.c_uint64(int(seed)), ctypes.c_uint32(int(stream)),
                                            d0, d1, ctypes.c_double(formation.LATTICE_PITCH), ctypes.byref(p),
                                            ctypes.byref(self._io), self._stream()), "ds_reset_random")
        self._episode.zero_()

    def observe(self):
        """rewards() on the current state (drone_env.py:208): refresh z_states/Ni/rewards."""
        p = self._params()
        _lib.check(self.lib.ds_observe(self._h, ctypes.byref(p), ctypes.byref(self._io), self._stream()),
                   "ds_observe")

    def observe_host(self):
        """observe(), then the whole result block (state, z, rewards, Ni, collision count) in ONE
        device->host transfer and one synchronise; returns the numpy views step_host() returns."""
        self.observe()
        if self._hblock is None:
            self._hblock = torch.empty(self._block.numel(), dtype=torch.uint8, pin_memory=True)
            self._hviews = {name: self._piece(self._hblock, name).numpy() for name in self._layout}
        with torch.cuda.device(self.device):
            self._hblock.copy_(self._block, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        return self._hviews

    def set_state(self, state, internal_t=None):
        """state[E,n,5] rows [x,y,vx,vy,l] (host float64) -> device; observation NOT refreshed."""
        st = np.ascontiguousarray(np.asarray(state, np.float64).reshape(self.n_envs, self.n_agents, 5))
        tt = None
        if internal_t is not None:
            tt = np.ascontiguousarray(np.broadcast_to(np.asarray(internal_t, np.int32), (self.n_envs,)))
        _lib.check(self.lib.ds_set_state(self._h, st.ctypes.data_as(ctypes.c_void_p),
                                         None if tt is None else tt.ctypes.data_as(ctypes.c_void_p),
                                         ctypes.byref(self._io), self._stream()), "ds_set_state")

    def get_state(self):
        st = np.empty((self.n_envs, self.n_agents, 5))
        tt = np.empty(self.n_envs, np.int32)
        _lib.check(self.lib.ds_get_state(self._h, st.ctypes.data_as(ctypes.c_void_p),
                                         tt.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self._io),
                                         self._stream()), "ds_get_state")
        return st, tt

    @property
    def state(self):
        """[E,n,5] device tensor [x,y,vx,vy,l] assembled from the SoA buffers."""
        rad = torch.as_tensor(self.drone_radius, dtype=self.dtype, device=self.device)
        return torch.cat([self.pos, self.vel, rad.expand(self.n_envs, -1).unsqueeze(-1)], dim=2)

    # ------------------------------------------------------------------ step
    def step(self, actions):
        """drones.step(actions) for every environment (drone_env.py:214-258).

        actions: device tensor [E,n,2] of this env's dtype.  Returns the reference's
        6-tuple as device tensors (views of the live buffers, like the reference's aliasing):
        (state-as-(pos,vel), z_states, rewards, n_collisions, finished, true_rewards).
        """
        if not (isinstance(actions, torch.Tensor) and actions.is_cuda):
            raise TypeError("step() takes a CUDA tensor; use step_host() for host arrays")
        a = actions.to(self.dtype).reshape(self.n_envs, self.n_agents, 2).contiguous()
        p = self._params()
        _lib.check(self.lib.ds_step(self._h, _ptr(a), ctypes.byref(p), ctypes.byref(self._io),
                                    self._stream()), "ds_step")
        return ((self.pos, self.vel), self.z_states, self.rewards, self.n_collisions, self.finished,
                self.true_rewards)

    def control(self, controller="gradient", u_max=1.0, out=None):
        """Actions of one of the reference's baseline controllers on the current state, computed on
        the device (ds_control; drone_env.py:612-679) -> [E,n,2]; nothing is stepped."""
        mode = {"proportional": _lib.DS_CTRL_PROPORTIONAL, "gradient": _lib.DS_CTRL_GRADIENT}.get(controller)
        if mode is None:
            raise ValueError("controller must be 'proportional' or 'gradient'")
        if out is None:
            out = torch.empty((self.n_envs, self.n_agents, 2), dtype=self.dtype, device=self.device)
        _lib.check(self.lib.ds_control(self._h, mode, ctypes.c_double(u_max), ctypes.byref(self._io), _ptr(out),
                                       self._stream()), "ds_control")
        return out

    def step_control(self, controller="gradient", u_max=1.0):
        """One closed-loop step: actions computed on the device by one of the reference's baseline
        controllers (drone_env.py:612-679: "proportional" | "gradient"), then drones.step()
        (ds_step_control).  Returns the same 6-tuple as step(); the action taken is self.vel."""
        mode = {"proportional": _lib.DS_CTRL_PROPORTIONAL, "gradient": _lib.DS_CTRL_GRADIENT}.get(controller)
        if mode is None:
            raise ValueError("controller must be 'proportional' or 'gradient'")
        p = self._params()
        _lib.check(self.lib.ds_step_control(self._h, mode, ctypes.c_double(u_max), ctypes.byref(p),
                                            ctypes.byref(self._io), self._stream()), "ds_step_control")
        return ((self.pos, self.vel), self.z_states, self.rewards, self.n_collisions, self.finished,
                self.true_rewards)

    def step_host(self, actions, block=True):
        """Same step with HOST arrays in and out: one H2D, one kernel, the reference's 6-tuple back
        into pinned memory, one synchronise.  block=True (ds_step_host_block): ONE D2H of the whole
        result block; block=False (ds_step_host): one copy per array.  Returns numpy views (keys
        pos, vel, z, r, tr, Ni, nc, fin) that the next call overwrites."""
        E, n, k = self.n_envs, self.n_agents, self.k_closest
        a = self._pin("act", (E, n, 2), self.dtype)
        a.numpy()[...] = np.asarray(actions).reshape(E, n, 2)
        p = self._params()
        if not block:
            o = {name: self._pin(name, shape, dtype) for name, (_, _, shape, dtype) in self._layout.items()}
            out = _lib.ds_host_step_out(o["pos"].data_ptr(), o["vel"].data_ptr(), o["z"].data_ptr(),
                                        o["r"].data_ptr(), o["tr"].data_ptr(), o["Ni"].data_ptr(),
                                        o["nc"].data_ptr(), o["fin"].data_ptr())
            _lib.check(self.lib.ds_step_host(self._h, _ptr(a), ctypes.byref(p), ctypes.byref(self._io),
                                             ctypes.byref(out), self._stream()), "ds_step_host")
            return {k_: v.numpy() for k_, v in o.items()}
        if self._hblock is None:
            self._hblock = torch.empty(self._block.numel(), dtype=torch.uint8, pin_memory=True)
            self._hviews = {name: self._piece(self._hblock, name).numpy() for name in self._layout}
        _lib.check(self.lib.ds_step_host_block(self._h, _ptr(a), ctypes.byref(p), ctypes.byref(self._io),
                                               _ptr(self._block), _ptr(self._hblock),
                                               ctypes.c_size_t(self._block.numel()), self._stream()),
                   "ds_step_host_block")
        return self._hviews

    # ------------------------------------------------------------------ rollout
    def _snapshot_obs0(self, out, rec):
        """`record` name "obs_pre": keep the observation the FIRST step of the call is taken from
        (z_states / Ni of the state before the call).  The trajectory buffers hold the observation
        AFTER each step (what `env.step` returns, drone_env.py:258); the reference's learners pair
        step t with the observation BEFORE it (train_problem.py:84-96 stores `z_states`, `Ni` read
        before `env.step`; SAC_agents.py:333-346) -- see pre_step_observations()."""
        if "obs_pre" in rec:
            out["z0"] = self.z_states.clone()
            out["Ni0"] = self.Ni.clone()

    @staticmethod
    def pre_step_observations(out):
        """(z_pre, Ni_pre) [T,E,n,...]: the observation each recorded step was TAKEN FROM, i.e.
        z_pre[t] = z(s_t), Ni_pre[t] = N(s_t): the pre-call snapshot followed by the recorded
        post-step observations shifted by one step.  This -- not out["Ni"] -- is what the reference's
        train_NN sums its advantages over (buffers[i][t].Ni = N_i(s_t), SAC_agents.py:333-346) and what
        the actor that produced action t was fed (log_p_of_a, utils.py:311-318).  Needs a rollout
        recorded with "obs" and "obs_pre"."""
        if "z0" not in out or "z" not in out:
            raise KeyError('record the rollout with ("obs", "obs_pre") to rebuild the pre-step observations')
        z_pre = torch.cat([out["z0"].unsqueeze(0), out["z"][:-1]], 0)
        Ni_pre = torch.cat([out["Ni0"].unsqueeze(0), out["Ni"][:-1]], 0)
        return z_pre, Ni_pre

    def rollout(self, actions=None, action_idx=None, action_table=None, record=("reward", "true_reward",
                "ncoll", "finished"), out=None):
        """T fused steps in ONE launch (ds_rollout).

        actions [T,E,n,2] device tensor, or action_idx [T,E,n] uint8 + action_table [A,2].
        record: names among pos, vel, reward, true_reward, obs (z + Ni), ncoll, finished.
        Returns a dict of device trajectory tensors plus 'agg' [E,4] and 'done' [E].
        """
        E, n, k = self.n_envs, self.n_agents, self.k_closest
        if actions is not None:
            actions = actions.to(self.dtype).contiguous()
            T = actions.shape[0]
            assert tuple(actions.shape) == (T, E, n, 2)
        else:
            action_idx = action_idx.contiguous()
            T = action_idx.shape[0]
            assert action_idx.dtype == torch.uint8 and tuple(action_idx.shape) == (T, E, n)
            action_table = torch.as_tensor(action_table, dtype=self.dtype, device=self.device).contiguous()
        out = {} if out is None else out
        dev, dt_ = self.device, self.dtype

        def buf(name, shape, dtype):
            t = out.get(name)
            if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
                t = torch.empty(shape, dtype=dtype, device=dev)
                out[name] = t
            return t

        rec = set(record)
        self._snapshot_obs0(out, rec)
        ro = _lib.ds_rollout_io()
        ro.T = T
        ro.n_actions = 0 if action_table is None else int(action_table.shape[0])
        ro.actions = None if actions is None else actions.data_ptr()
        ro.action_idx = None if action_idx is None else action_idx.data_ptr()
        ro.action_table = None if action_table is None else action_table.data_ptr()
        if "pos" in rec: ro.pos_tr = buf("pos", (T, E, n, 2), dt_).data_ptr()
        if "vel" in rec: ro.vel_tr = buf("vel", (T, E, n, 2), dt_).data_ptr()
        if "reward" in rec: ro.reward_tr = buf("reward", (T, E, n), dt_).data_ptr()
        if "true_reward" in rec: ro.true_reward_tr = buf("true_reward", (T, E, n), dt_).data_ptr()
        if "obs" in rec:
            ro.z_tr = buf("z", (T, E, n, k + 1, self.cols), dt_).data_ptr()
            ro.Ni_tr = buf("Ni", (T, E, n, k + 1), torch.int32).data_ptr()
        if "ncoll" in rec: ro.ncoll_tr = buf("ncoll", (T, E), torch.int32).data_ptr()
        if "finished" in rec: ro.finished_tr = buf("finished", (T, E), torch.uint8).data_ptr()
        ro.agg = self.agg.data_ptr()
        ro.done = self.done.data_ptr()
        p = self._params()
        _lib.check(self.lib.ds_rollout(self._h, ctypes.byref(p), ctypes.byref(self._io), ctypes.byref(ro),
                                       self._stream()), "ds_rollout")
        out["agg"], out["done"] = self.agg, self.done
        return out

    def rollout_control(self, T, controller="gradient", u_max=1.0,
                        record=("reward", "true_reward", "ncoll", "finished"), out=None):
        """T closed-loop steps in ONE launch (ds_rollout_control): actions from the reference's
        baseline controller on the current state at every step (train_problem.py:89-90 with
        drone_env.py:612-679), state resident on chip.  Same outputs as rollout()."""
        mode = {"proportional": _lib.DS_CTRL_PROPORTIONAL, "gradient": _lib.DS_CTRL_GRADIENT}.get(controller)
        if mode is None:
            raise ValueError("controller must be 'proportional' or 'gradient'")
        E, n, k = self.n_envs, self.n_agents, self.k_closest
        out = {} if out is None else out
        dev, dt_ = self.device, self.dtype

        def buf(name, shape, dtype):
            t = out.get(name)
            if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
                t = torch.empty(shape, dtype=dtype, device=dev)
                out[name] = t
            return t

        rec = set(record)
        self._snapshot_obs0(out, rec)
        ro = _lib.ds_rollout_io()
        ro.T = int(T)
        if "pos" in rec: ro.pos_tr = buf("pos", (T, E, n, 2), dt_).data_ptr()
        if "vel" in rec: ro.vel_tr = buf("vel", (T, E, n, 2), dt_).data_ptr()
        if "reward" in rec: ro.reward_tr = buf("reward", (T, E, n), dt_).data_ptr()
        if "true_reward" in rec: ro.true_reward_tr = buf("true_reward", (T, E, n), dt_).data_ptr()
        if "obs" in rec:
            ro.z_tr = buf("z", (T, E, n, k + 1, self.cols), dt_).data_ptr()
            ro.Ni_tr = buf("Ni", (T, E, n, k + 1), torch.int32).data_ptr()
        if "ncoll" in rec: ro.ncoll_tr = buf("ncoll", (T, E), torch.int32).data_ptr()
        if "finished" in rec: ro.finished_tr = buf("finished", (T, E), torch.uint8).data_ptr()
        ro.agg = self.agg.data_ptr()
        ro.done = self.done.data_ptr()
        p = self._params()
        _lib.check(self.lib.ds_rollout_control(self._h, mode, ctypes.c_double(u_max), ctypes.byref(p),
                                               ctypes.byref(self._io), ctypes.byref(ro), self._stream()),
                   "ds_rollout_control")
        out["agg"], out["done"] = self.agg, self.done
        return out

    def rollout_host(self, actions=None, action_idx=None, action_table=None,
                     record=("pos", "vel", "reward", "true_reward", "obs", "ncoll", "finished"),
                     chunk=0, out=None, compact=False):
        """End-to-end episode loop with HOST buffers (ds_rollout_host): pinned host action stream
        in, pinned host trajectories out, copies pipelined against the kernel.  `actions` /
        `action_idx` are used in place when they are pinned, contiguous CPU tensors of this
        environment's dtype and shape [T,E,n,2] / uint8 [T,E,n]; anything else is copied into pinned
        staging first.  out["vel"] is a view of the action stream when explicit actions are given.
        compact=True: out["z"] comes back as float32 (what the reference's actors cast it to,
        utils.py:305) and out["Ni"] as uint8 (255 = none): less than half the observation bytes
        across PCIe.
        One call = one episode segment from the CURRENT state with done = 0 and fresh episode sums
        (returned in out["agg"]); self.done / self.agg are not touched: an environment that finishes
        stops for the rest of the call only (its vel rows past the end still echo the actions)."""
        E, n, k = self.n_envs, self.n_agents, self.k_closest
        hr = _lib.ds_host_rollout()
        keep = []

        def usable(t, dtype, tail):
            return (isinstance(t, torch.Tensor) and not t.is_cuda and t.is_pinned() and t.dtype == dtype and
                    t.is_contiguous() and t.dim() == len(tail) + 1 and tuple(t.shape[1:]) == tail)

        if actions is not None:
            a = actions if usable(actions, self.dtype, (E, n, 2)) else None
            if a is None:
                src = torch.as_tensor(np.asarray(actions.cpu() if isinstance(actions, torch.Tensor) else actions))
                if tuple(src.shape[1:]) != (E, n, 2):
                    raise ValueError(f"actions must have shape [T,{E},{n},2]")
                a = self._pin("ro_act", tuple(src.shape), self.dtype)
                a.copy_(src)
            T = a.shape[0]
            hr.actions = a.data_ptr(); keep.append(a)
        else:
            a = action_idx if usable(action_idx, torch.uint8, (E, n)) else None
            if a is None:
                src = torch.as_tensor(np.asarray(action_idx.cpu() if isinstance(action_idx, torch.Tensor) else action_idx))
                if tuple(src.shape[1:]) != (E, n):
                    raise ValueError(f"action_idx must have shape [T,{E},{n}]")
                a = self._pin("ro_aidx", tuple(src.shape), torch.uint8)
                a.copy_(src)
            T = a.shape[0]
            tab = torch.as_tensor(np.asarray(action_table), dtype=self.dtype).contiguous()
            hr.action_idx = a.data_ptr(); hr.action_table = tab.data_ptr(); hr.n_actions = tab.shape[0]
            keep += [a, tab]
        hr.T, hr.chunk = T, int(chunk)
        hr.flags = _lib.DS_HOST_COMPACT_OBS if compact else 0
        out = {} if out is None else out
        rec = set(record)

        def hbuf(name, shape, dtype):
            t = out.get(name)
            if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
                t = torch.empty(shape, dtype=dtype, pin_memory=True)
                out[name] = t
            return t.data_ptr()

        dt_ = self.dtype
        if "pos" in rec: hr.pos_tr = hbuf("pos", (T, E, n, 2), dt_)
        if "vel" in rec:
            # state[:,2:4] = u (drone_env.py:238): with explicit actions the recorded velocity IS the
            # action stream -- returned as a view of it, nothing copied; index mode: table[idx] on the host
            if actions is not None:
                out["vel"] = keep[0].view(T, E, n, 2); hr.vel_tr = hr.actions
            else:
                hr.vel_tr = hbuf("vel", (T, E, n, 2), dt_)
        if "reward" in rec: hr.reward_tr = hbuf("reward", (T, E, n), dt_)
        if "true_reward" in rec: hr.true_reward_tr = hbuf("true_reward", (T, E, n), dt_)
        if "obs" in rec:
            hr.z_tr = hbuf("z", (T, E, n, k + 1, self.cols), torch.float32 if compact else dt_)
            hr.Ni_tr = hbuf("Ni", (T, E, n, k + 1), torch.uint8 if compact else torch.int32)
        if "ncoll" in rec: hr.ncoll_tr = hbuf("ncoll", (T, E), torch.int32)
        if "finished" in rec: hr.finished_tr = hbuf("finished", (T, E), torch.uint8)
        hr.agg = hbuf("agg", (E, 4), torch.float64)
        p = self._params()
        _lib.check(self.lib.ds_rollout_host(self._h, ctypes.byref(p), ctypes.byref(self._io),
                                            ctypes.byref(hr), self._stream()), "ds_rollout_host")
        return out

    def returns(self, reward, Ni, finished, discount=0.99, baseline=None, out=None):
        """Monte-Carlo returns G_i(t) and Delta-neighbourhood advantage sums
        sum_{j in N_i(t)} (G_j(t) - V_i(t)) of a recorded rollout, on the device (ds_returns;
        reference SAC_agents.py:304-310, 333-345).  reward [T,E,n] and finished [T,E] are the
        trajectory tensors rollout() records.  Ni [T,E,n,k+1] must be the neighbour lists of the
        state each step was TAKEN FROM, N_i(s_t) -- the reference stores `Ni = env.Ni` BEFORE
        `env.step` (train_problem.py:84-96) -- i.e. pre_step_observations(out)[1], NOT out["Ni"]
        (which holds N_i(s_{t+1}), the observation step t returned); likewise baseline [T,E,n] is
        the critic's V_i(z_i(s_t)) on pre_step_observations(out)[0] (None = 0).
        returns_from_rollout() does the shift.  Returns dict(returns, advantage, count)."""
        E, n, k = self.n_envs, self.n_agents, self.k_closest
        T = reward.shape[0]
        assert tuple(reward.shape) == (T, E, n) and reward.dtype == self.dtype and reward.is_cuda
        assert tuple(Ni.shape) == (T, E, n, k + 1) and Ni.dtype == torch.int32
        assert tuple(finished.shape) == (T, E) and finished.dtype == torch.uint8
        reward, Ni, finished = reward.contiguous(), Ni.contiguous(), finished.contiguous()
        if baseline is not None:
            baseline = baseline.to(self.dtype).contiguous()
            assert tuple(baseline.shape) == (T, E, n)
        out = {} if out is None else out

        def buf(name, dtype):
            t = out.get(name)
            if t is None or tuple(t.shape) != (T, E, n) or t.dtype != dtype:
                t = torch.empty((T, E, n), dtype=dtype, device=self.device)
                out[name] = t
            return t

        io = _lib.ds_returns_io()
        io.T, io.discount = T, float(discount)
        io.reward_tr, io.Ni_tr, io.finished_tr = reward.data_ptr(), Ni.data_ptr(), finished.data_ptr()
        io.baseline = None if baseline is None else baseline.data_ptr()
        io.returns = buf("returns", self.dtype).data_ptr()
        io.advantage = buf("advantage", self.dtype).data_ptr()
        io.count = buf("count", torch.uint8).data_ptr()
        _lib.check(self.lib.ds_returns(self._h, ctypes.byref(io), self._stream()), "ds_returns")
        return out

    def returns_from_rollout(self, ro, discount=0.99, critic=None, out=None):
        """train_NN's returns and advantage sums (SAC_agents.py:304-310, 333-346) of a rollout
        recorded with ("reward", "obs", "obs_pre", "finished"): neighbour lists and critic inputs
        are the PRE-step observations.  critic: callable z_pre [T,E,n,(k+1)*cols] -> V [T,E,n], or None."""
        z_pre, Ni_pre = self.pre_step_observations(ro)
        base = None if critic is None else critic(z_pre.flatten(-2))
        return self.returns(ro["reward"], Ni_pre, ro["finished"], discount=discount, baseline=base, out=out)

    # ------------------------------------------------------------------ policy
    def load_policy(self, W1, b1, W2, b2, W3, b3, action_table):
        """Per-agent actors of the reference (utils.DiscreteSoftmaxNN: in -> 300 -> 300 -> A) for
        policy_forward(): fp32 arrays stacked over agents, W* as torch stores nn.Linear.weight
        ([n,300,in], [n,300,300], [n,A,300]); action_table [A,2] = action_list (ds_policy_create)."""
        f = lambda x: np.ascontiguousarray(np.asarray(x, np.float32))
        W1, b1, W2, b2, W3, b3 = map(f, (W1, b1, W2, b2, W3, b3))
        n, A = self.n_agents, W3.shape[1]
        in_dim = W1.shape[2]
        assert W1.shape == (n, 300, in_dim) and W2.shape == (n, 300, 300) and W3.shape == (n, A, 300)
        assert b1.shape == (n, 300) and b2.shape == (n, 300) and b3.shape == (n, A)
        tab = np.ascontiguousarray(np.asarray(action_table, np.float64).reshape(A, 2))
        cfg = _lib.ds_policy_config(n, in_dim, A, _REAL[self.dtype], self.device.index or 0, 0,
                                    W1.ctypes.data, b1.ctypes.data, W2.ctypes.data, b2.ctypes.data,
                                    W3.ctypes.data, b3.ctypes.data, tab.ctypes.data)
        old = getattr(self, "_pol", None)
        if old is not None and old.value:
            self.lib.ds_policy_destroy(old)
        self._pol = ctypes.c_void_p()
        _lib.check(self.lib.ds_policy_create(ctypes.byref(cfg), ctypes.byref(self._pol)), "ds_policy_create")
        self._pol_A = A
        self.actions = torch.zeros((self.n_envs, n, 2), dtype=self.dtype, device=self.device)
        self.action_idx = torch.zeros((self.n_envs, n), dtype=torch.uint8, device=self.device)
        self.action_probs = torch.zeros((self.n_envs, n, A), dtype=torch.float32, device=self.device)

    def policy_forward(self, seed=0, stream=0, z=None, want_probs=True):
        """`actions = agents.forward(z_states, Ni)` (SAC_agents.py:170-180) for every environment on
        the device (ds_policy_forward, tcgen05): probabilities of each agent's own network on its
        observation, one action index drawn per (environment, agent) from Philox(seed; e, i, stream).
        Returns (actions [E,n,2], action_idx [E,n], probs [E,n,A] or None); feed `actions` to step()."""
        z = self.z_states if z is None else z
        assert z.is_cuda and z.dtype == self.dtype and z.is_contiguous()
        io = _lib.ds_policy_io(z.data_ptr(), self.actions.data_ptr(), self.action_idx.data_ptr(),
                               self.action_probs.data_ptr() if want_probs else None, int(seed), int(stream), 0)
        _lib.check(self.lib.ds_policy_forward(self._h, self._pol, ctypes.byref(io), self._stream()),
                   "ds_policy_forward")
        return self.actions, self.action_idx, (self.action_probs if want_probs else None)

    def rollout_policy(self, T, seed=0, stream0=0, record=("reward", "true_reward", "obs", "ncoll", "finished",
                                                          "action_idx"), out=None, seed_tensor=None):
        """A closed-loop episode with the loaded actors, entirely on the device (ds_rollout_policy):
        for t < T: `actions = agents.forward(z_states, Ni)` then `env.step(actions)`
        (train_problem.py:82-104) for every environment -- 2 T launches on the current stream, no
        host round trip.  Semantics and outputs as rollout(): finished environments stop, `finished`
        codes 0 / 1 / 2, agg / done; "vel" records the actions taken, "action_idx" the drawn
        indices [T,E,n] (what log_p_of_a needs, utils.py:311-318), "probs" the distributions.
        Step t samples from Philox(seed; e, i, stream0 + t); seed_tensor (uint64 device tensor of
        one element) overrides seed, so that a captured CUDA graph can be replayed with new draws."""
        assert getattr(self, "_pol", None) is not None and self._pol.value, "load_policy() first"
        E, n, k = self.n_envs, self.n_agents, self.k_closest
        out = {} if out is None else out
        dev, dt_ = self.device, self.dtype

        def buf(name, shape, dtype):
            t = out.get(name)
            if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
                t = torch.empty(shape, dtype=dtype, device=dev)
                out[name] = t
            return t

        rec = set(record)
        self._snapshot_obs0(out, rec)
        ro = _lib.ds_rollout_io()
        ro.T = int(T)
        if "pos" in rec: ro.pos_tr = buf("pos", (T, E, n, 2), dt_).data_ptr()
        if "vel" in rec: ro.vel_tr = buf("vel", (T, E, n, 2), dt_).data_ptr()
        if "reward" in rec: ro.reward_tr = buf("reward", (T, E, n), dt_).data_ptr()
        if "true_reward" in rec: ro.true_reward_tr = buf("true_reward", (T, E, n), dt_).data_ptr()
        if "obs" in rec:
            ro.z_tr = buf("z", (T, E, n, k + 1, self.cols), dt_).data_ptr()
            ro.Ni_tr = buf("Ni", (T, E, n, k + 1), torch.int32).data_ptr()
        if "ncoll" in rec: ro.ncoll_tr = buf("ncoll", (T, E), torch.int32).data_ptr()
        if "finished" in rec: ro.finished_tr = buf("finished", (T, E), torch.uint8).data_ptr()
        ro.agg = self.agg.data_ptr()
        ro.done = self.done.data_ptr()
        pio = _lib.ds_policy_rollout_io()
        if "action_idx" in rec: pio.action_idx_tr = buf("action_idx", (T, E, n), torch.uint8).data_ptr()
        if "probs" in rec: pio.probs_tr = buf("probs", (T, E, n, self._pol_A), torch.float32).data_ptr()
        pio.seed = int(seed); pio.stream0 = int(stream0)
        if seed_tensor is not None:
            assert seed_tensor.is_cuda and seed_tensor.dtype == torch.uint64 and seed_tensor.numel() == 1
            pio.seed_dev = seed_tensor.data_ptr()
        p = self._params()
        _lib.check(self.lib.ds_rollout_policy(self._h, self._pol, ctypes.byref(p), ctypes.byref(self._io),
                                              ctypes.byref(ro), ctypes.byref(pio), self._stream()),
                   "ds_rollout_policy")
        out["agg"], out["done"] = self.agg, self.done
        return out

    def episode_aggregates(self, out=None):
        """Device-side sum over this rank's environments of the per-env episode accumulators
        -> float64 [5] = (sum_t mean_i r, sum_t mean_i true_r, sum_t collisions, steps, #envs):
        the vector a rank all-reduces (train_problem.py:98-100,118-121).  out: a caller-owned
        float64 [5] device tensor (e.g. one of a ring, when the all-reduce runs asynchronously)."""
        dst = self._agg_sum if out is None else out
        assert dst.is_cuda and dst.dtype == torch.float64 and dst.numel() == 5
        _lib.check(self.lib.ds_reduce_aggregates(self._h, _ptr(self.agg), _ptr(dst),
                                                 self._stream()), "ds_reduce_aggregates")
        return dst

"""ctypes binding of libdronestep.so (C ABI: include/dronestep.h).

The library is the ONLY implementation of the step path in this package: there
is no Python/NumPy/torch fallback.  If the shared object is missing, or no CUDA
device is visible, the product API raises instead of computing on the CPU.
"""
from __future__ import annotations

import ctypes
import os

from . import build as _build

c_void_p, c_int32, c_double = ctypes.c_void_p, ctypes.c_int32, ctypes.c_double

DS_OK, DS_ERR_ARG, DS_ERR_CUDA, DS_ERR_NO_DEVICE, DS_ERR_INTERNAL = 0, -1, -2, -3, -4
DS_LOG_DIV, DS_LOG_DIFF, DS_LOG_RCP = 0, 1, 2
DS_HOST_COMPACT_OBS = 1
DS_CTRL_PROPORTIONAL, DS_CTRL_GRADIENT = 1, 2
DS_MAX_AGENTS, DS_MAX_K = 1024, 16

EXPORTED_SYMBOLS = (
    "ds_abi_version", "ds_last_error", "ds_device_count", "ds_create", "ds_destroy",
    "ds_default_params", "ds_step", "ds_observe", "ds_rollout", "ds_reduce_aggregates",
    "ds_set_state", "ds_get_state", "ds_reset", "ds_step_host", "ds_step_host_block", "ds_rollout_host", "ds_returns",
    "ds_step_control", "ds_rollout_control", "ds_reset_random",
    "ds_policy_create", "ds_policy_destroy", "ds_policy_forward", "ds_rollout_policy", "ds_control",
    "ds_rollout_kernel_name",
)


class DroneStepError(RuntimeError):
    pass


class ds_config(ctypes.Structure):
    _fields_ = [("n_envs", c_int32), ("n_agents", c_int32), ("k_closest", c_int32),
                ("simplify_zstate", c_int32), ("real_bytes", c_int32), ("device", c_int32),
                ("end_points", c_void_p), ("d_safety", c_void_p), ("deltas", c_void_p),
                ("radius", c_void_p)]


class ds_params(ctypes.Structure):
    _fields_ = [("dt", c_double), ("collision_weight", c_double), ("goal_tol", c_double),
                ("sentinel", c_double), ("zero_eps", c_double), ("ghost_factor", c_double),
                ("max_time_steps", c_int32), ("log_mode", c_int32)]


class ds_buffers(ctypes.Structure):
    _fields_ = [("pos", c_void_p), ("vel", c_void_p), ("reward", c_void_p),
                ("true_reward", c_void_p), ("z", c_void_p), ("Ni", c_void_p), ("ncoll", c_void_p),
                ("finished", c_void_p), ("t", c_void_p)]


class ds_rollout_io(ctypes.Structure):
    _fields_ = [("T", c_int32), ("n_actions", c_int32), ("actions", c_void_p),
                ("action_idx", c_void_p), ("action_table", c_void_p), ("pos_tr", c_void_p),
                ("vel_tr", c_void_p), ("reward_tr", c_void_p), ("true_reward_tr", c_void_p),
                ("z_tr", c_void_p), ("Ni_tr", c_void_p), ("ncoll_tr", c_void_p),
                ("finished_tr", c_void_p), ("agg", c_void_p), ("done", c_void_p)]


class ds_returns_io(ctypes.Structure):
    _fields_ = [("T", c_int32), ("_pad", c_int32), ("discount", c_double), ("reward_tr", c_void_p),
                ("Ni_tr", c_void_p), ("finished_tr", c_void_p), ("baseline", c_void_p),
                ("returns", c_void_p), ("advantage", c_void_p), ("count", c_void_p)]


class ds_policy_config(ctypes.Structure):
    _fields_ = [("n_agents", c_int32), ("in_dim", c_int32), ("n_actions", c_int32), ("real_bytes", c_int32),
                ("device", c_int32), ("_pad", c_int32), ("W1", c_void_p), ("b1", c_void_p), ("W2", c_void_p),
                ("b2", c_void_p), ("W3", c_void_p), ("b3", c_void_p), ("action_table", c_void_p)]


class ds_policy_io(ctypes.Structure):
    _fields_ = [("z", c_void_p), ("actions", c_void_p), ("action_idx", c_void_p), ("probs", c_void_p),
                ("seed", ctypes.c_uint64), ("stream", ctypes.c_uint32), ("_pad", ctypes.c_uint32)]


class ds_policy_rollout_io(ctypes.Structure):
    _fields_ = [("action_idx_tr", c_void_p), ("probs_tr", c_void_p), ("seed", ctypes.c_uint64),
                ("seed_dev", c_void_p), ("stream0", ctypes.c_uint32), ("_pad", ctypes.c_uint32)]


class ds_host_step_out(ctypes.Structure):
    _fields_ = [("pos", c_void_p), ("vel", c_void_p), ("z", c_void_p), ("reward", c_void_p),
                ("true_reward", c_void_p), ("Ni", c_void_p), ("ncoll", c_void_p),
                ("finished", c_void_p)]


class ds_host_rollout(ctypes.Structure):
    _fields_ = [("T", c_int32), ("chunk", c_int32), ("n_actions", c_int32), ("flags", c_int32),
                ("actions", c_void_p), ("action_idx", c_void_p), ("action_table", c_void_p),
                ("pos_tr", c_void_p), ("vel_tr", c_void_p), ("reward_tr", c_void_p),
                ("true_reward_tr", c_void_p), ("z_tr", c_void_p), ("Ni_tr", c_void_p),
                ("ncoll_tr", c_void_p), ("finished_tr", c_void_p), ("agg", c_void_p)]


_lib = None


def lib_path() -> str:
    # DS_LIB_OVERRIDE: another build of the same sources (kernel tuning experiments only)
    return os.environ.get("DS_LIB_OVERRIDE") or _build.LIB_PATH


def load():
    """Load libdronestep.so; raise DroneStepError (never fall back) when absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise DroneStepError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). This package has no CPU implementation of drone_env.step().")
    lib = ctypes.CDLL(path)
    lib.ds_last_error.restype = ctypes.c_char_p
    lib.ds_rollout_kernel_name.restype = ctypes.c_char_p
    lib.ds_rollout_kernel_name.argtypes = [ctypes.c_void_p]
    lib.ds_destroy.restype = None
    lib.ds_default_params.restype = None
    lib.ds_policy_destroy.restype = None
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)   # AttributeError if the ABI is incomplete
        if name not in ("ds_last_error", "ds_destroy", "ds_default_params", "ds_policy_destroy", "ds_rollout_kernel_name"):
            fn.restype = ctypes.c_int
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != DS_OK:
        msg = load().ds_last_error().decode("utf-8", "replace")
        raise DroneStepError(f"{what} failed (status {rc}): {msg}")


def default_params() -> ds_params:
    p = ds_params()
    load().ds_default_params(ctypes.byref(p))
    return p

"""CPU: the C-ABI library loads and exports every symbol include/dronestep.h declares; the
product path fails loudly without a GPU; host-side setup matches the reference's
constructor outputs (golden ctor table)."""
import ctypes
import os
import random
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dronestep.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ds_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from scalable_collision_avoidance_rl_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(build.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in dronestep.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert _lib.load().ds_abi_version() == 1


def test_struct_layouts_match_header():
    """ctypes mirrors must have the C sizes (x86-64 SysV): guards silent ABI drift."""
    from scalable_collision_avoidance_rl_b200 import _lib
    assert ctypes.sizeof(_lib.ds_config) == 6 * 4 + 4 * 8
    assert ctypes.sizeof(_lib.ds_params) == 6 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.ds_buffers) == 9 * 8
    assert ctypes.sizeof(_lib.ds_rollout_io) == 2 * 4 + 13 * 8
    assert ctypes.sizeof(_lib.ds_host_step_out) == 8 * 8
    assert ctypes.sizeof(_lib.ds_host_rollout) == 4 * 4 + 12 * 8
    assert ctypes.sizeof(_lib.ds_returns_io) == 2 * 4 + 8 + 7 * 8
    assert ctypes.sizeof(_lib.ds_policy_config) == 6 * 4 + 7 * 8
    assert ctypes.sizeof(_lib.ds_policy_io) == 4 * 8 + 8 + 2 * 4
    assert ctypes.sizeof(_lib.ds_policy_rollout_io) == 2 * 8 + 8 + 8 + 2 * 4
    p = _lib.default_params()
    assert (p.dt, p.collision_weight, p.goal_tol, p.sentinel, p.zero_eps, p.ghost_factor,
            p.max_time_steps) == (0.05, 0.2, 0.2, 9.99e3, -1e-6, 1.1, 200)


def test_no_gpu_means_loud_failure():
    """No CPU fallback: without a device the ABI reports DS_ERR_NO_DEVICE and the Python
    product API raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scalable_collision_avoidance_rl_b200 import _lib, DroneStepError, BatchedDrones
    lib = _lib.load()
    assert lib.ds_device_count() == 0
    n = 3
    arr = (ctypes.c_double * (2 * n))()
    cfg = _lib.ds_config(1, n, 2, 1, 8, 0, ctypes.addressof(arr), ctypes.addressof(arr),
                         ctypes.addressof(arr), ctypes.addressof(arr))
    h = ctypes.c_void_p()
    assert lib.ds_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.DS_ERR_NO_DEVICE
    assert b"no CPU path" in lib.ds_last_error()
    with pytest.raises(DroneStepError):
        BatchedDrones(1, 3, [5, 5])
    import drone_env
    with pytest.raises(DroneStepError):
        drone_env.drones(n_agents=5, n_obstacles=0, grid=[5, 5], end_formation="O")


def test_argument_validation_without_device():
    from scalable_collision_avoidance_rl_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.ds_create(None, ctypes.byref(h)) == _lib.DS_ERR_ARG
    arr = (ctypes.c_double * 8)()
    for bad in (dict(n_envs=0), dict(n_agents=0), dict(n_agents=2000), dict(k_closest=3, n_agents=3),
                dict(real_bytes=2)):
        kw = dict(n_envs=1, n_agents=4, k_closest=2, simplify_zstate=1, real_bytes=8, device=0)
        kw.update(bad)
        cfg = _lib.ds_config(kw["n_envs"], kw["n_agents"], kw["k_closest"], kw["simplify_zstate"],
                             kw["real_bytes"], kw["device"], ctypes.addressof(arr), ctypes.addressof(arr),
                             ctypes.addressof(arr), ctypes.addressof(arr))
        assert lib.ds_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.DS_ERR_ARG, bad
        assert lib.ds_last_error()


CTOR_CASES = [(5, [5, 5], 1.0), (10, [5, 5], 1.0), (32, [32, 32], 2.5), (128, [64, 64], 1.0),
              (32, [5, 5], 2.5), (128, [5, 5], 1.0), (7, [4, 9], 3.0), (5, [5, 5], None)]


@pytest.mark.parametrize("n,grid,delta", CTOR_CASES)
def test_host_setup_matches_reference_constructor(n, grid, delta):
    """formation / d_safety / delta clip / start lattice, bit for bit against the reference's
    constructor outputs (oracle/make_golden.py: ctor_table)."""
    from scalable_collision_avoidance_rl_b200 import formation
    tab = np.load(os.path.join(ROOT, "tests", "golden", "ctor_table.npz"))
    key = f"n{n}_g{grid[0]}x{grid[1]}_d{delta}"
    pts = formation.end_formation("O", n, grid)
    assert np.array_equal(pts, tab[key + "_end_points"])
    ds = formation.safety_distances(pts, np.ones(n) * 0.1)
    assert np.array_equal(ds, tab[key + "_d_safety"])
    deltas, clipped = formation.clip_deltas(None if delta is None else np.ones(n) * delta, ds)
    assert np.array_equal(np.asarray(deltas, np.float64), tab[key + "_deltas"])
    assert clipped == (delta is not None and bool((delta > ds).any()))
    # same Python `random` stream -> same start nodes (make_golden seeds with n; no obstacles drawn)
    random.seed(n)
    start = formation.sample_start_reference_stream(n, grid)
    assert np.array_equal(start, tab[key + "_state0"][:, 0:2])
    assert list(tab[key + "_lss"]) == [formation.local_state_space(2, True), 2]


def test_known_d_safety_values():
    """SURVEY section 4 known answers."""
    from scalable_collision_avoidance_rl_b200 import formation
    for n, grid, want in [(5, [5, 5], 2.44), (10, [5, 5], 1.19), (32, [32, 32], 2.62),
                          (128, [64, 64], 1.21), (32, [5, 5], 0.24), (128, [5, 5], -0.09)]:
        ds = formation.safety_distances(formation.end_formation("O", n, grid), np.ones(n) * 0.1)
        assert np.allclose(ds, want, atol=1e-12), (n, grid, ds[:3])


def test_batched_start_sampler_distinct_nodes():
    from scalable_collision_avoidance_rl_b200 import formation
    rng = np.random.default_rng(0)
    for n, grid in [(10, [5, 5]), (128, [64, 64])]:
        p = formation.sample_start_batched(257, n, grid, rng)
        assert p.shape == (257, n, 2)
        d0, d1 = formation.lattice_shape(grid)
        idx = np.rint(p / formation.LATTICE_PITCH).astype(int)
        assert (idx[..., 0] < d0).all() and (idx[..., 1] < d1).all() and (idx >= 0).all()
        flat = idx[..., 0] * d1 + idx[..., 1]
        assert all(len(set(row)) == n for row in flat)
    with pytest.raises(ValueError):
        formation.sample_start_batched(2, 600, [5, 5], rng)   # more agents than nodes


def test_action_table_and_helpers():
    from scalable_collision_avoidance_rl_b200 import formation
    import drone_env
    t = formation.unit_action_table(16)
    assert t.shape == (16, 2) and np.allclose(np.hypot(t[:, 0], t[:, 1]), 1)
    assert (drone_env.dim, drone_env.dt, drone_env.max_time_steps) == (2, 0.05, 200)
    for name in ("drones", "running_average", "plot_rewards", "plot_grads", "num_to_rgb",
                 "gradient_control", "proportional_control"):
        assert hasattr(drone_env, name)
    x = np.arange(100.0)
    y = drone_env.running_average(x, 50)
    assert y[49] == np.mean(x[:50]) and y[10] == 10
    assert drone_env.num_to_rgb(0, 4) == (round(128) / 255, round(np.sin(2) * 127 + 128) / 255,
                                          round(np.sin(4) * 127 + 128) / 255)

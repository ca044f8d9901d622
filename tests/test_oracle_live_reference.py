"""CPU, build container only: the C oracle stepped side by side with the UNMODIFIED reference imported
from /root/reference (oracle/ref_harness.py shims matplotlib / IPython / np.infty), on random
configurations that are NOT among the committed golden vectors -- a differential check that the
fixtures did not happen to miss something.  Skipped where the reference is absent (the GPU box)."""
import contextlib
import io
import os
import random

import numpy as np
import pytest

from helpers import FP64_TOL, assert_close, compare_obs
from oracle import c_oracle

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def _tie_rows(env):
    d_ij, _, _, _ = env.distance_data(env.state, env.deltas, env.d_safety)
    s = np.sort(d_ij, axis=1)[:, : min(env.n_agents, env.k_closest + 2)]
    return (s[:, 1:] == s[:, :-1]).any(1)


def _pad(Ni, k):
    out = np.full((len(Ni), k + 1), -1, np.int32)
    for i, lst in enumerate(Ni):
        out[i, : len(lst)] = lst
    return out


CASES = [  # n, grid, k, simplify, deltas ("u" = uniform 1.0, "h" = heterogeneous, None), box (None = lattice start), cw
    (5, [5, 5], 2, True, "u", None, 0.2),
    (9, [7, 4], 3, False, "h", None, 0.35),
    (6, [5, 5], 1, True, None, None, 0.2),
    (12, [6, 6], 4, False, "h", 2.5, 1.0),       # dense random box: collisions, many in-range neighbours
    (40, [32, 32], 2, True, "u", 9.0, 0.2),
]


from test_row_logic_host import rowlib, _frame  # noqa: E402,F401  (fixture + host build of the kernels' row logic)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c[0]}_k{c[2]}_{'s' if c[3] else 'f'}_{c[4]}_{c[5]}")
@pytest.mark.parametrize("path", [0, 1, 2], ids=["row", "worklist", "inline32"])
def test_kernel_row_logic_against_the_live_reference(rowlib, case, path):
    """The row logic of the CUDA kernels (dronestep_kernels.cuh compiled for the host, as in
    test_row_logic_host.py) evaluated on states the live reference has just produced."""
    if path == 2 and case[0] > 32:
        pytest.skip("inline mode exists for n <= 32 only")
    from oracle.ref_harness import import_reference
    ref = import_reference("drone_env")
    n, grid, k, simplify, dmode, box, cw = case
    rng = np.random.default_rng(77 + n)
    deltas = None if dmode is None else (np.ones(n) if dmode == "u" else rng.uniform(0.2, 2.0, n))
    random.seed(5); np.random.seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        env = ref.drones(n_agents=n, n_obstacles=0, grid=list(grid), end_formation="O", k_closest=k,
                         deltas=None if deltas is None else deltas.copy(), simplify_zstate=simplify)
    env.collision_weight = cw
    if box is not None:
        env.state[:, 0:2] = rng.uniform(0, box, (n, 2))
    dl = np.asarray(env.deltas, np.float64).reshape(-1)
    for step in range(12):
        act = rng.uniform(-1, 1, (n, 2))
        state, z, r, ncoll, fin, tr = env.step([a_.copy() for a_ in act])
        got_r, got_tr, got_z, got_Ni, nc, _ = _frame(rowlib, path, 8, n, k, simplify, 0, state[:, 0:2].copy(),
                                                     state[:, 2:4].copy(), env.end_points.reshape(n, 2),
                                                     env.d_safety, dl, env.drone_radius, cw)
        assert_close(got_r, np.array(r), FP64_TOL, f"r step {step}")
        assert_close(got_tr, np.array(tr), FP64_TOL, f"true_r step {step}")
        assert nc == int(ncoll), step
        _compare_obs_live(env, got_z, got_Ni, np.array(z), _pad(env.Ni, k), f"obs step {step}")


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c[0]}_k{c[2]}_{'s' if c[3] else 'f'}_{c[4]}_{c[5]}")
@pytest.mark.parametrize("seed", [11, 12])
def test_c_oracle_against_the_live_reference(case, seed):
    from oracle.ref_harness import import_reference
    ref = import_reference("drone_env")
    n, grid, k, simplify, dmode, box, cw = case
    rng = np.random.default_rng(1000 * seed + n)
    deltas = None if dmode is None else (np.ones(n) if dmode == "u" else rng.uniform(0.2, 2.0, n))
    random.seed(seed); np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        env = ref.drones(n_agents=n, n_obstacles=0, grid=list(grid), end_formation="O", k_closest=k,
                         deltas=None if deltas is None else deltas.copy(), simplify_zstate=simplify)
    env.collision_weight = cw
    if box is not None:
        env.state[:, 0:2] = rng.uniform(0, box, (n, 2))
    orc = c_oracle.OracleEnv(1, n, env.end_points, env.d_safety, np.asarray(env.deltas, np.float64).reshape(-1),
                             env.drone_radius, k, simplify, c_oracle.default_params(cw))
    t0 = int(rng.integers(0, 190))
    env.internal_t = t0
    orc.set_state(env.state[None, :, 0:2].copy(), env.state[None, :, 2:4].copy(), np.array([t0], np.int32))
    for step in range(25):
        act = rng.uniform(-1, 1, (n, 2)) if step % 2 else \
            np.stack([np.cos(a := rng.integers(0, 16, n) / 16 * 2 * np.pi), np.sin(a)], 1)
        state, z, r, ncoll, fin, tr = env.step([a_.copy() for a_ in act])
        out = orc.step(act[None])
        assert np.array_equal(out.pos[0], state[:, 0:2]) and np.array_equal(out.vel[0], state[:, 2:4]), step
        assert_close(out.r[0], np.array(r), FP64_TOL, f"r step {step}")
        assert_close(out.true_r[0], np.array(tr), FP64_TOL, f"true_r step {step}")
        assert int(out.ncoll[0]) == int(ncoll) and bool(out.finished[0]) == bool(fin), step
        _compare_obs_live(env, out.z[0], out.Ni[0], np.array(z), _pad(env.Ni, k), f"obs step {step}")
        if fin:
            break


def _compare_obs_live(env, z, Ni, z_ref, Ni_ref, what):
    """Rows without an exact tie among the first k + 2 sorted distances: slot for slot (helpers.compare_obs).
    Tie rows: np.argsort is unstable (NumPy >= 1.25 uses an AVX-512 sort), so ANY agent whose clipped
    distance equals the reference's rank-kth distance is a valid reference outcome for slot kth -- typically
    several pairs clipped to d_safety[i] that the deltas[j]-broadcast mask (drone_env.py:328) still counts as
    in range.  Required there: the self slot, the number of real neighbours, for every real slot an agent
    with exactly the rank-kth clipped distance, and slot contents that are that agent's relative state."""
    tie = _tie_rows(env)
    compare_obs(z[~tie], Ni[~tie], z_ref[~tie], Ni_ref[~tie], np.zeros(int((~tie).sum()), bool), FP64_TOL, what)
    d_ij, _, _, _ = env.distance_data(env.state, env.deltas, env.d_safety)
    d_sorted = np.sort(d_ij, axis=1)
    cols = z.shape[-1]
    for i in np.nonzero(tie)[0]:
        assert Ni[i, 0] == i and np.array_equal(z[i, 0], z_ref[i, 0], equal_nan=True), f"{what}: self slot row {i}"
        assert (Ni[i] >= 0).sum() == (Ni_ref[i] >= 0).sum(), f"{what}: neighbour count row {i}"
        for kth in range(1, Ni.shape[1]):
            j = int(Ni[i, kth])
            if j < 0:                                            # ghost slot: position part is the reference's
                assert Ni_ref[i, kth] < 0
                assert np.array_equal(z[i, kth, 0:2], z_ref[i, kth, 0:2], equal_nan=True), f"{what}: ghost row {i}"
                continue
            assert d_ij[i, j] == d_sorted[i, kth], f"{what}: row {i} slot {kth}: agent {j} is not a rank-{kth} candidate"
            want = env.state[j, 0:2] - env.state[i, 0:2]
            assert np.array_equal(z[i, kth, 0:2], want), f"{what}: row {i} slot {kth} relative position"
            if cols == 5:
                assert np.array_equal(z[i, kth, 2:5], env.state[j, 2:5]), f"{what}: row {i} slot {kth} velocity / radius"
        assert len({int(j) for j in Ni[i] if j >= 0}) == (Ni[i] >= 0).sum(), f"{what}: duplicate agent row {i}"


@pytest.mark.parametrize("n,grid,box,u_max", [(5, [5, 5], 2.0, 1.0), (11, [6, 6], 1.5, 0.7), (33, [32, 32], 6.0, 1.0)])
def test_controllers_against_the_live_reference(rowlib, n, grid, box, u_max):
    """control_action of the kernels (host build) and the C oracle against the reference's own
    gradient_control / proportional_control (drone_env.py:612-679) on dense random states: bit-exact,
    NaN patterns included."""
    import ctypes
    from oracle.ref_harness import import_reference
    ref = import_reference("drone_env")
    rng = np.random.default_rng(n)
    random.seed(1); np.random.seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        env = ref.drones(n_agents=n, n_obstacles=0, grid=list(grid), end_formation="O", k_closest=2,
                         deltas=np.ones(n), simplify_zstate=True)
    xF = np.ascontiguousarray(env.end_points.reshape(-1)); dsf = np.ascontiguousarray(env.d_safety)
    rad = np.ascontiguousarray(env.drone_radius)
    for f in range(20):
        env.state[:, 0:2] = rng.uniform(0, box, (n, 2))
        if f >= 12:                                                              # the controllers read the radii from the state
            env.state[:, 4] = rng.uniform(0.05, 0.25, n)
            rad = np.ascontiguousarray(env.state[:, 4])
        if f == 3:
            env.state[1, 0:2] = env.state[0, 0:2] + [rad[0] + rad[1], 0.0]      # exact contact: division by zero
        if f == 4:
            env.state[2, 0:2] = env.end_points.reshape(n, 2)[2]                  # on the goal
        with np.errstate(all="ignore"):
            want = {2: np.array(ref.gradient_control(env.state, env, u_max=u_max)).reshape(n, 2),
                    1: np.array(ref.proportional_control(env.state, env)).reshape(n, 2)}
        pos = np.ascontiguousarray(env.state[:, 0:2])
        for mode, um in ((2, u_max), (1, 1.0)):
            act = np.zeros((n, 2))
            rowlib.rowcheck_control(mode, n, *[x.ctypes.data_as(ctypes.c_void_p) for x in (pos, xF, dsf, rad)],
                                    ctypes.c_double(um), act.ctypes.data_as(ctypes.c_void_p))
            assert np.array_equal(act, want[mode], equal_nan=True), (f, mode)
            got = c_oracle.control(mode, pos[None], env.end_points, env.d_safety, rad, um)[0]
            assert np.array_equal(got, want[mode], equal_nan=True), (f, mode, "oracle")

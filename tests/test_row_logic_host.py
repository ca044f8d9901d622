"""Row logic of the CUDA kernels (eval_row + write_obs in dronestep_kernels.cuh), compiled for the
HOST by tests/rowcheck/rowcheck.cu and compared with the C oracle.  This checks the algorithmic
part that differs from the reference's formulation -- the provably-clipped fast path, the
Delta-disk count by correction, the (distance, index) k-nearest selection -- without a GPU.  The
GPU parity tests (tests/test_gpu_parity.py) remain the proof for the compiled device code.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import FP64_TOL, assert_close, compare_obs
from oracle import c_oracle
from scalable_collision_avoidance_rl_b200 import formation

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rowcheck", "rowcheck.cu")
HDR = os.path.join(os.path.dirname(HERE), "scalable_collision_avoidance_rl_b200", "csrc", "dronestep_kernels.cuh")
OUT = os.path.join(HERE, "_build", "librowcheck.so")


@pytest.fixture(scope="module")
def rowlib():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        nvcc = "/usr/local/cuda/bin/nvcc"
        # the header's device part uses sm_100 intrinsics: compile for the product's architecture;
        # only the HOST side of this object is ever executed
        subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
                        "-fPIC,-ffp-contract=off", "-shared", "-o", OUT, SRC], check=True)
    lib = ctypes.CDLL(OUT)
    lib.rowcheck_frame.restype = ctypes.c_int
    lib.rowcheck_log.restype = ctypes.c_double
    lib.rowcheck_log.argtypes = [ctypes.c_double]
    return lib


def _frame(lib, path, real_bytes, n, k, simplify, log_mode, pos, vel, xF, ds, dl, rad, cw):
    cols = 2 if simplify else 5
    r, tr = np.zeros(n), np.zeros(n)
    z = np.zeros((n, k + 1, cols)); Ni = np.full((n, k + 1), -1, np.int32)
    nc, ng = ctypes.c_int(0), ctypes.c_int(0)
    keep = [np.ascontiguousarray(x, np.float64) for x in (pos, vel, xF, ds, dl, rad)]
    lib.rowcheck_frame(path, real_bytes, n, k, int(simplify), log_mode, *[x.ctypes.data_as(ctypes.c_void_p) for x in keep],
                       ctypes.c_double(cw), r.ctypes.data_as(ctypes.c_void_p),
                       tr.ctypes.data_as(ctypes.c_void_p), z.ctypes.data_as(ctypes.c_void_p),
                       Ni.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nc), ctypes.byref(ng))
    return r, tr, z, Ni, nc.value, ng.value


CASES = [  # n, k, simplify, grid, delta, box, hetero_delta, hetero_radius
    (5, 2, True, [5, 5], 1.0, 2.0, False, False),
    (10, 2, True, [5, 5], 1.0, 3.0, False, False),
    (10, 2, False, [5, 5], 1.0, 2.0, True, False),
    (10, 2, True, [5, 5], None, 4.0, False, False),      # deltas=None -> Delta == d_safety: all in range
    (7, 0, True, [5, 5], 1.0, 2.0, False, False),
    (9, 1, False, [5, 5], 1.0, 2.0, False, True),
    (12, 4, True, [8, 8], 1.5, 3.0, True, True),
    (20, 6, False, [16, 16], 2.0, 5.0, False, False),
    (32, 2, True, [32, 32], 2.5, 8.0, False, False),
    (33, 2, True, [32, 32], 2.5, 30.0, False, False),    # second block of 32 agents
    (70, 3, False, [32, 32], 1.0, 6.0, True, False),
    (128, 2, True, [64, 64], 1.0, 60.0, False, False),   # sparse: nearly every pair on the fast path
    (128, 2, True, [5, 5], 1.0, 5.0, False, False),      # negative d_safety (degenerate reference config)
]


@pytest.mark.parametrize("path", [0, 1, 2], ids=["row", "worklist", "inline32"])
@pytest.mark.parametrize("log_mode", [0, 1])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c[0]}_k{c[1]}_{'s' if c[2] else 'f'}_box{c[5]}")
def test_row_logic_matches_oracle(rowlib, case, log_mode, path):
    n, k, simplify, grid, delta, box, hd, hr = case
    if path == 2 and n > 32:
        pytest.skip("inline mode exists for n <= 32 only")
    rng = np.random.default_rng(100 + n + k)
    rad = rng.uniform(0.05, 0.15, n) if hr else np.full(n, 0.1)
    xF = formation.end_formation("O", n, grid)
    ds = formation.safety_distances(xF, rad)
    deltas_in = None if delta is None else (rng.uniform(0.2, delta, n) if hd else np.ones(n) * delta)
    dl, _ = formation.clip_deltas(deltas_in, ds)
    dl = np.asarray(dl, np.float64).reshape(-1)
    frames = 40
    orc = c_oracle.OracleEnv(frames, n, xF, ds, dl, rad, k, simplify, c_oracle.default_params(0.35))
    pos = rng.uniform(0, box, (frames, n, 2))
    pos[0, 1] = pos[0, 0]                                  # coincident agents (d = -2l ties with self)
    pos[1, 1] = pos[1, 0] + [rad[0] + rad[1], 0.0]         # exact contact
    pos[2] = xF.reshape(n, 2)                              # everyone on goal (ghost = 0/0)
    vel = rng.standard_normal((frames, n, 2))
    orc.set_state(pos, vel)
    ref = orc.observe()
    for f in range(frames):
        r, tr, z, Ni, nc, ng = _frame(rowlib, path, 8, n, k, simplify, log_mode, pos[f], vel[f], xF, ds, dl, rad, 0.35)
        assert_close(r, ref.r[f], FP64_TOL, f"r frame {f}")
        assert_close(tr, ref.true_r[f], FP64_TOL, f"true_r frame {f}")
        assert nc == ref.ncoll[f], f"ncoll frame {f}"
        compare_obs(z, Ni, ref.z[f], ref.Ni[f], ref.tie[f], FP64_TOL, f"obs frame {f}")
    assert ref.ncoll.sum() > 0 or box > 20 or ds.min() < 0   # dense cases must exercise collisions


def test_row_logic_float32_close(rowlib):
    n, k = 10, 2
    rng = np.random.default_rng(5)
    xF = formation.end_formation("O", n, [5, 5]); rad = np.full(n, 0.1)
    ds = formation.safety_distances(xF, rad); dl = np.ones(n)
    orc = c_oracle.OracleEnv(64, n, xF, ds, dl, rad, k, True, c_oracle.default_params(0.2))
    pos = rng.uniform(0, 5, (64, n, 2)).astype(np.float32).astype(np.float64)
    orc.set_state(pos)
    ref = orc.observe()
    for f in range(64):
        r, tr, z, Ni, nc, ng = _frame(rowlib, 1, 4, n, k, True, 0, pos[f], np.zeros((n, 2)), xF, ds, dl, rad, 0.2)
        err = np.abs(r - ref.r[f]) / np.maximum(1, np.abs(ref.r[f]))
        assert err.max() < 5e-6


def test_table_log_accuracy(rowlib):
    """log_r (dronestep_kernels.cuh) against libm over the positive normal range: within 2 ulp,
    or within 2e-18 absolute where lc + log1p(r) cancels (results below ~1e-2 in the table
    intervals next to the one that holds 1); exact special cases; log(1) = 0."""
    rng = np.random.default_rng(7)
    xs = np.concatenate([
        np.exp(rng.uniform(-700, 700, 200000)),              # whole exponent range
        rng.uniform(0.5, 2.0, 200000),                       # the range the barrier term lives in
        1.0 + rng.uniform(-1, 1, 100000) * 10.0 ** rng.uniform(-15, -2, 100000),
        np.array([1.0, 2.0, 0.5, np.sqrt(2.0), np.sqrt(0.5), np.nextafter(1.0, 2), np.nextafter(1.0, 0),
                  2.2250738585072014e-308, 1.7976931348623157e308, 1.19 / 0.3, 1.19 / 1.1899999]),
    ])
    got = np.array([rowlib.rowcheck_log(float(x)) for x in xs])
    ref = np.log(xs)
    ulp = np.spacing(np.abs(ref))
    err = np.abs(got - ref) / np.maximum(2.0 * ulp, 2e-18)
    assert err.max() <= 1.0, f"max error {np.abs(got - ref).max():.3e} at x = {xs[err.argmax()]!r}"
    near1 = np.abs(xs - 1.0) < 1e-4                           # the rc = 1 interval: relative accuracy
    assert (np.abs(got - ref)[near1] <= 2.0 * ulp[near1]).all()
    assert rowlib.rowcheck_log(1.0) == 0.0
    with np.errstate(all="ignore"):
        assert rowlib.rowcheck_log(0.0) == -np.inf and np.isnan(rowlib.rowcheck_log(-1.0))
        assert rowlib.rowcheck_log(np.inf) == np.inf and np.isnan(rowlib.rowcheck_log(np.nan))
        assert abs(rowlib.rowcheck_log(5e-324) - np.log(5e-324)) < 1e-12


@pytest.mark.parametrize("name,key,mode", [("control_gradient_n5", "action", 2), ("control_gradient_n10", "action", 2),
                                           ("control_proportional_n8", "action", 1),
                                           ("control_dense_n7", "gradient", 2), ("control_dense_n7", "proportional", 1)])
def test_control_action_matches_reference(rowlib, name, key, mode):
    """control_action (dronestep_kernels.cuh, host build) against the actions the reference's own
    gradient_control / proportional_control produced (oracle/make_golden_control.py): bit-exact,
    NaN patterns (division by zero at exact contact, agent on its goal) included."""
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    n = int(g["n"])
    xF = np.ascontiguousarray(g["end_points"].reshape(-1)); dsf = np.ascontiguousarray(g["d_safety"])
    rad = np.full(n, 0.1)
    u_max = float(g["u_max"]) if mode == 2 else 1.0
    for f in range(g["state_in"].shape[0]):
        pos = np.ascontiguousarray(g["state_in"][f, :, 0:2])
        act = np.zeros((n, 2))
        rowlib.rowcheck_control(mode, n, *[x.ctypes.data_as(ctypes.c_void_p) for x in (pos, xF, dsf, rad)],
                                ctypes.c_double(u_max), act.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(act, g[key][f], equal_nan=True), f"frame {f}"

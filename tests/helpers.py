"""Shared test helpers: golden-fixture loading and tie-aware comparison."""
from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# float contract of BASELINE.json north_star: |delta| <= 1e-5 on state/reward.
# The fp64 paths are expected to sit many orders below it; tests assert both.
CONTRACT_TOL = 1e-5
FP64_TOL = 1e-9


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not p.endswith("ctor_table.npz") and not os.path.basename(p).startswith(("returns_", "control_", "policynet_")))


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    for key in ("n", "k", "simplify", "max_time_steps"):
        g[key] = int(g[key])
    g["collision_weight"] = float(g["collision_weight"])
    g["dt"] = float(g["dt"])
    return g


def assert_close(a, b, tol, what):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), f"{what}: NaN pattern differs"
    if a.size:
        err = np.max(np.abs(np.where(nan_a, 0, a) - np.where(nan_b, 0, b)))
        assert err <= tol, f"{what}: max abs err {err:.3e} > {tol:.1e}"


def compare_obs(z, Ni, z_ref, Ni_ref, tie_ref, tol, what):
    """Compare localized observations [.., n, k+1, cols] tie-aware.

    Rows without an exact distance tie must match slot for slot.  On tie rows the
    reference itself is platform dependent (np.argsort is unstable): require the
    self slot, the neighbour COUNT and the multiset of slot distances to match.
    """
    z = np.asarray(z, np.float64); z_ref = np.asarray(z_ref, np.float64)
    Ni = np.asarray(Ni); Ni_ref = np.asarray(Ni_ref)
    tie = np.asarray(tie_ref).astype(bool)
    assert z.shape == z_ref.shape, f"{what}: z shape {z.shape} vs {z_ref.shape}"
    flat_z = z.reshape(-1, *z.shape[-2:]); flat_zr = z_ref.reshape(-1, *z.shape[-2:])
    flat_n = Ni.reshape(-1, Ni.shape[-1]); flat_nr = Ni_ref.reshape(-1, Ni.shape[-1])
    flat_t = tie.reshape(-1)
    clean = ~flat_t
    assert_close(flat_z[clean], flat_zr[clean], tol, what + " z (tie-free rows)")
    assert np.array_equal(flat_n[clean], flat_nr[clean]), what + " Ni (tie-free rows)"
    if flat_t.any():
        a, b = flat_z[flat_t], flat_zr[flat_t]
        assert_close(a[:, 0], b[:, 0], tol, what + " z self slot (tie rows)")
        assert np.array_equal(flat_n[flat_t][:, 0], flat_nr[flat_t][:, 0]), what + " Ni[0] (tie rows)"
        assert np.array_equal((flat_n[flat_t] >= 0).sum(1), (flat_nr[flat_t] >= 0).sum(1)), \
            what + " neighbour count (tie rows)"
        da = np.sort(np.hypot(a[:, 1:, 0], a[:, 1:, 1]), axis=1)
        db = np.sort(np.hypot(b[:, 1:, 0], b[:, 1:, 1]), axis=1)
        assert_close(da, db, max(tol, 1e-12), what + " slot distances (tie rows)")

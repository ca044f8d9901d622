"""Host-side environment setup: the constants the step kernel consumes.

These run once per constructor / once per reset, stay on the host in float64 and
follow the reference's own formulas so the kernel constants are bit-identical
(reference drone_env.py: generate_formation 115-153, delta clip 85-91, start
lattice 171-205).  They are plain functions so the CPU test-suite can pin them to
the golden constructor table without a GPU.
"""
from __future__ import annotations

import math
import random as _pyrandom

import numpy as np

DIM = 2
DRONE_RADIUS = 0.1                  # drone_env.py:75,174
LATTICE_PITCH = 2 * 1.1 * 0.1       # drone_env.py:193 (delta_l)


def circle_formation(n_agents: int, grid) -> np.ndarray:
    """End points of formation "O" as the reference's [2n,1] column (drone_env.py:124-131)."""
    step = 2 * np.pi / n_agents
    out = np.zeros([n_agents * DIM, 1])
    for i in range(n_agents):
        # same operation order as the reference: ((cos * 0.9) * g) / 2 + g / 2
        out[DIM * i, 0] = np.cos(i * step) * 0.9 * grid[0] / 2 + grid[0] / 2
        out[DIM * i + 1, 0] = np.sin(i * step) * 0.9 * grid[1] / 2 + grid[1] / 2
    return out


def end_formation(label: str, n_agents: int, grid) -> np.ndarray:
    if label != "O":
        # the reference only logs an error and then crashes on the undefined name (drone_env.py:133-134)
        raise ValueError(f"{label} is Not a valid end formation identifier")
    return circle_formation(n_agents, grid)


def safety_distances(end_points: np.ndarray, radius: np.ndarray) -> np.ndarray:
    """d_hat_i = floor(100 * min_{j != i}(||xF_i - xF_j|| - l_i - l_j)) / 100 (drone_env.py:136-153).

    A vectorised pass shortlists, per row, the candidates within 1e-9 of the row
    minimum; those are then re-evaluated with the reference's exact expression
    (np.linalg.norm of the 1-D difference) so the floor() sees identical bits.
    """
    n = radius.shape[0]
    xF = np.asarray(end_points, np.float64).reshape(n, DIM)
    d_safety = np.zeros(n)
    if n == 1:
        d_safety[0] = np.inf
        return np.floor(d_safety * 100) / 100
    diff = xF[:, None, :] - xF[None, :, :]
    approx = np.sqrt((diff * diff).sum(2)) - radius[:, None] - radius[None, :]
    np.fill_diagonal(approx, np.inf)
    row_min = approx.min(1)
    for i in range(n):
        best = np.inf
        for j in np.nonzero(approx[i] <= row_min[i] + 1e-9)[0]:
            d_ij = np.linalg.norm(xF[i] - xF[j]) - radius[i] - radius[j]
            best = min([best, d_ij])
        d_safety[i] = best
    return np.floor(d_safety * 100) / 100


def clip_deltas(deltas, d_safety: np.ndarray):
    """Delta-disk radii clipped to d_safety (drone_env.py:85-91). Returns (deltas, clipped?)."""
    if deltas is None:
        return d_safety, False
    deltas = np.asarray(deltas)
    return np.minimum(deltas, d_safety), (not np.all(deltas <= d_safety))


def lattice_shape(grid):
    div = np.floor(np.array(grid) / LATTICE_PITCH)          # drone_env.py:194
    return int(div[0]), int(div[1])


def lattice_coords(flat_idx, grid) -> np.ndarray:
    """Coordinates of lattice nodes given their index in the reference's i-major list
    (drone_env.py:197-200): node = [idx * pitch, jdx * pitch]."""
    _, d1 = lattice_shape(grid)
    flat_idx = np.asarray(flat_idx)
    return np.stack([(flat_idx // d1) * LATTICE_PITCH, (flat_idx % d1) * LATTICE_PITCH], -1)


def sample_start_reference_stream(n_agents: int, grid) -> np.ndarray:
    """n distinct lattice nodes drawn from Python's global `random` stream exactly as the
    reference does (random.sample of the node list, drone_env.py:204): sampling positions
    of a sequence depends only on its length, so a range of the same length consumes the
    stream identically and returns the same picks."""
    d0, d1 = lattice_shape(grid)
    picks = _pyrandom.sample(range(d0 * d1), n_agents)
    return lattice_coords(picks, grid)


def sample_start_batched(n_envs: int, n_agents: int, grid, rng: np.random.Generator) -> np.ndarray:
    """[E,n,2] distinct lattice nodes per environment from a NumPy generator (the
    throughput path; statistically equivalent to the reference's reset, not stream-equal)."""
    d0, d1 = lattice_shape(grid)
    L = d0 * d1
    if n_agents > L:
        raise ValueError("Sample larger than population or is negative")   # as random.sample
    picks = np.empty((n_envs, n_agents), np.int64)
    block = max(1, (1 << 24) // max(L, 1))                    # bound the key matrix to ~128 MB
    for lo in range(0, n_envs, block):
        m = min(block, n_envs - lo)
        if L <= 4096:
            keys = rng.random((m, L))
            sel = np.argpartition(keys, n_agents - 1, axis=1)[:, :n_agents]
            # argpartition leaves the first n in arbitrary order; shuffle for an unbiased order
            picks[lo:lo + m] = rng.permuted(sel, axis=1)
        else:
            sel = rng.integers(0, L, size=(m, n_agents))
            for _ in range(64):                               # rejection on the rare duplicates
                srt = np.sort(sel, axis=1)
                bad = (srt[:, 1:] == srt[:, :-1]).any(1)
                if not bad.any():
                    break
                sel[bad] = rng.integers(0, L, size=(int(bad.sum()), n_agents))
            else:
                raise RuntimeError("could not draw distinct lattice nodes")
            picks[lo:lo + m] = sel
    return lattice_coords(picks, grid)


def local_state_space(k_closest: int, simplify: bool) -> int:
    return (DIM if simplify else 2 * DIM + 1) * (1 + k_closest)   # drone_env.py:180-184


def unit_action_table(n_actions: int) -> np.ndarray:
    """Action set of the discrete softmax policy (reference utils.py:262-269)."""
    rows = []
    for a in range(n_actions):
        rows.append([np.cos(a / n_actions * 2 * np.pi) * 1, np.sin(a / n_actions * 2 * np.pi) * 1])
    return np.array(rows)


def num_to_rgb(val, max_val):
    """Agent colour ramp (reference drone_env.py:41-51)."""
    if val > max_val:
        raise ValueError("val must not be greater than max_val")
    if val < 0 or max_val < 0:
        raise ValueError("arguments may not be negative")
    i = val * 255 / max_val
    return tuple(round(math.sin(0.024 * i + ph) * 127 + 128) / 255 for ph in (0, 2, 4))

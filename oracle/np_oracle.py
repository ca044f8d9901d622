"""TEST INFRASTRUCTURE ONLY -- vectorised NumPy float64 restatement of the
reference's ``drones.step()`` path over E environments.

Second, independent restatement (the first is ``drone_oracle.c``); the two are
cross-checked against each other and against the golden vectors recorded from
the unmodified reference (``tests/test_oracle_golden.py``).  Parity status:
PINNED by ``tests/golden/``.  Never imported by the product package.

Reference citations (file:line in /root/reference/drone_env.py):
  integrate 227-238 | distance_data 295-334 | rewards 260-293 |
  localized_states 336-401 | termination 248-256
"""
from __future__ import annotations

import numpy as np

SENTINEL = 9.99e3        # drone_env.py:330-332
ZERO_EPS = -10 ** -6     # drone_env.py:320
GHOST_FACTOR = 1.1       # drone_env.py:386
GOAL_TOL = 0.2           # drone_env.py:251


def _norm_blas(x, y):
    """sqrt(ddot(v, v)) as OpenBLAS evaluates it for 2-vectors: fma(y, y, x*x).

    NumPy has no fma ufunc; emulate it exactly with a Dekker/Veltkamp split
    (error-free product) so the restatement is bit-identical to the C oracle.
    """
    xx = x * x
    # two-product y*y = p + e exactly
    p = y * y
    c = 134217729.0 * y  # 2**27 + 1
    hi = c - (c - y)
    lo = y - hi
    e = ((hi * hi - p) + 2.0 * hi * lo) + lo * lo
    # fma(y,y,xx) = round(p + e + xx): two-sum of p and xx, then add the errors
    s = p + xx
    bb = s - p
    err = (p - (s - bb)) + (xx - bb)
    return np.sqrt(s + (err + e))


def observe(pos, vel, radius, xF, d_safety, deltas, k, simplify, collision_weight, dt=0.05):
    """rewards() (drone_env.py:260-293) for pos[E,n,2]; returns dict of arrays."""
    E, n, _ = pos.shape
    q, b = 2 * dt, collision_weight * dt
    dx = pos[:, :, None, 0] - pos[:, None, :, 0]
    dy = pos[:, :, None, 1] - pos[:, None, :, 1]
    with np.errstate(all="ignore"):
        raw = _norm_blas(dx, dy) - radius[None, :, None] - radius[None, None, :]
        ds = np.broadcast_to(d_safety[None, :, None], raw.shape)
        d = np.where(ds < raw, ds, raw)                       # python min(raw, ds), :318
        d = np.where(d == 0, ZERO_EPS, d)                     # :319-320
        dn = ds / d                                           # :321
        eye = np.eye(n, dtype=bool)[None]
        self_d = np.where(d_safety < -2 * radius, d_safety, -radius - radius)  # :323
        d = np.where(eye, self_d[None, :, None], d)
        dn = np.where(eye, 1.0, dn)                           # :325
        coll = dn <= 0                                        # :327
        nd = d <= deltas[None, None, :]                       # :328 (column broadcast)
        logd = np.where(coll, SENTINEL, np.log(np.where(coll, SENTINEL, dn)))  # :330-332
        g = xF[None] - pos
        nrm = np.sqrt(g[..., 0] * g[..., 0] + g[..., 1] * g[..., 1])
        goal = q * (nrm * nrm)                                # :276
        r = -np.nan_to_num(goal + b * np.sum(logd * nd, 2))   # :282,287
        true_r = -np.nan_to_num(goal + b * np.sum(logd, 2))   # :283,288
        ncoll = coll.sum((1, 2)).astype(np.int32)             # :284
        order = np.argsort(d, axis=2, kind="stable")          # :338, ties -> lowest index
        in_range = nd.sum(2) - 1                              # :346
        cols = 2 if simplify else 5
        z = np.zeros((E, n, k + 1, cols))
        Ni = np.full((E, n, k + 1), -1, np.int32)
        zi = -(xF[None] - pos)                                # :357
        z[:, :, 0, 0:2] = zi
        if not simplify:
            z[:, :, 0, 2:4] = vel
            z[:, :, 0, 4] = radius[None]
        Ni[:, :, 0] = np.arange(n)[None]
        zn = _norm_blas(zi[..., 0], zi[..., 1])
        ghost = zi / zn[..., None] * deltas[None, :, None] * GHOST_FACTOR  # :386
        ee = np.arange(E)[:, None]
        for kth in range(1, k + 1):
            j = order[:, :, kth]
            inside = kth <= in_range                          # :362
            rel = pos[ee, j] - pos                            # :368
            z[:, :, kth, 0:2] = np.where(inside[..., None], rel, ghost)
            if not simplify:
                z[:, :, kth, 2:4] = vel[ee, j]
                z[:, :, kth, 4] = radius[j]
            Ni[:, :, kth] = np.where(inside, j, -1)
        # reference appends neighbours contiguously; "inside" is a prefix so -1s trail
        dsort = np.take_along_axis(d, order[:, :, : min(n, k + 2)], 2)
        tie = (dsort[:, :, 1:] == dsort[:, :, :-1]).any(2)
    return dict(r=r, true_r=true_r, z=z, Ni=Ni, ncoll=ncoll, tie=tie, d=d)


def step(pos, vel, t, act, radius, xF, d_safety, deltas, k, simplify, collision_weight,
         dt=0.05, max_time_steps=200):
    """drones.step() (drone_env.py:214-258) for E envs; pos/vel/t updated in place."""
    pos += dt * act          # A = I, B = dt*I (:78-79,235); rounding: x + round(dt*u)
    vel[...] = act           # :238
    out = observe(pos, vel, radius, xF, d_safety, deltas, k, simplify, collision_weight, dt)
    g = xF[None] - pos
    err = np.sqrt(g[..., 0] * g[..., 0] + g[..., 1] * g[..., 1])
    out["finished"] = ((err <= GOAL_TOL).all(1) | (t >= max_time_steps - 1)).astype(np.uint8)  # :251
    t += 1                   # :256
    return out


def returns(reward_tr, Ni_tr, finished_tr, discount, baseline=None):
    """NumPy restatement of the return scan and the Delta-neighbourhood advantage gather
    (reference SAC_agents.py:304-310,333-345); same contract as ``c_oracle.returns``."""
    r = np.asarray(reward_tr, np.float64)
    Ni = np.asarray(Ni_tr); fin = np.asarray(finished_tr)
    T, E, n = r.shape
    k = Ni.shape[-1] - 1
    executed = np.cumsum(fin == 2, axis=0) == 0                       # [T,E]: slice was stepped
    prev_fin = np.concatenate([np.zeros((1, E), bool), np.cumsum(fin == 1, axis=0)[:-1] > 0])
    executed &= ~prev_fin
    G = np.zeros((T, E, n))
    nxt = np.zeros((E, n))
    started = np.zeros(E, bool)
    for t in range(T - 1, -1, -1):                                    # SAC_agents.py:306-309
        ex = executed[t]
        g = np.where(started[:, None], nxt * discount + r[t], r[t])
        G[t] = np.where(ex[:, None], g, 0.0)
        nxt = np.where(ex[:, None], g, nxt)
        started |= ex
    base = np.zeros((T, E, n)) if baseline is None else np.asarray(baseline, np.float64)
    adv = np.zeros((T, E, n)); cnt = np.zeros((T, E, n), np.int32)
    for m in range(k + 1):                                            # list order (:344-345)
        j = Ni[..., m]
        valid = (j >= 0) & executed[..., None]
        gj = np.take_along_axis(G, np.where(j >= 0, j, 0), axis=2)
        adv = np.where(valid, adv + (gj - base), adv)
        cnt += valid
    return G, adv, cnt


def control(mode, pos, end_points, d_safety, radius=None, u_max=1.0):
    """NumPy restatement of proportional_control (mode 1, drone_env.py:655-679) and gradient_control
    (mode 2, drone_env.py:612-653) over E environments; same contract as ``c_oracle.control``."""
    pos = np.asarray(pos, np.float64)
    E, n, _ = pos.shape
    xF = np.asarray(end_points, np.float64).reshape(1, n, 2)
    ds = np.asarray(d_safety, np.float64).reshape(1, n, 1)
    rad = np.full(n, 0.1) if radius is None else np.asarray(radius, np.float64).reshape(n)
    with np.errstate(all="ignore"):
        if mode == 1:
            u = 1 * (xF - pos)
            nrm = _norm_blas(u[..., 0], u[..., 1])[..., None]
            return np.where(nrm > 1, u / nrm * 1, u)
        d = pos[:, :, None, :] - pos[:, None, :, :]                       # x_i - x_j
        nrm = _norm_blas(d[..., 0], d[..., 1])
        dij = nrm - rad[None, :, None] - rad[None, None, :]
        use = (dij <= ds) & ~np.eye(n, dtype=bool)[None]
        den = dij * nrm
        t2 = np.zeros((E, n, 2))
        for j in range(n):                                                # accumulation order of :640-647
            t2 = np.where(use[:, :, j, None], t2 + d[:, :, j, :] / den[:, :, j, None], t2)
        grad = 1 * (2 * (pos - xF)) - 0.1 * t2
        return np.clip(-grad, -u_max, u_max)


def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    M = 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M, p1 & M, ((p0 >> 32) ^ c3 ^ k1) & M, p0 & M
        k0, k1 = (k0 + 0x9E3779B9) & M, (k1 + 0xBB67AE85) & M
    return [c0, c1, c2, c3]


def reset_random(n_envs, n_agents, d0, d1, pitch, seed, stream):
    """Restatement of the device-side reset sampler (reset_random_kernel): Philox4x32-10 keyed by
    the seed, counter (environment, draw block, stream); Lemire's unbiased index; duplicates
    redrawn.  Returns [E,n,2] float64 lattice coordinates [idx * pitch, jdx * pitch]
    (reference drone_env.py:193-205 fixes the lattice and the distribution, not the stream)."""
    L = d0 * d1
    thresh = ((1 << 32) - L) % L
    out = np.zeros((n_envs, n_agents, 2))
    for e in range(n_envs):
        picks, block, rnd = [], 0, []
        while len(picks) < n_agents:
            if not rnd:
                rnd = _philox4x32_10(e, block, stream, 0, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
                block += 1
            m = rnd.pop(0) * L
            if (m & 0xFFFFFFFF) < thresh:
                continue
            node = m >> 32
            if node in picks:
                continue
            picks.append(node)
        nodes = np.array(picks)
        out[e, :, 0] = (nodes // d1) * pitch
        out[e, :, 1] = (nodes % d1) * pitch
    return out


def policy_probs(z, W1, b1, W2, b2, W3, b3):
    """fp32 restatement of DiscreteSoftmaxNN.forward (reference utils.py:287-302) for a batch:
    z [B, in] -> softmax(W3 relu(W2 relu(W1 z + b1) + b2) + b3) [B, A].  The reference evaluates one
    observation at a time with softmax over dim 0 of the action vector (utils.py:283)."""
    f = np.float32
    x = np.asarray(z, np.float64).astype(f)                              # torch.tensor(z, dtype=float32) (:305)
    l1 = np.maximum(x @ np.asarray(W1, f).T + np.asarray(b1, f), f(0))
    l2 = np.maximum(l1 @ np.asarray(W2, f).T + np.asarray(b2, f), f(0))
    o = l2 @ np.asarray(W3, f).T + np.asarray(b3, f)
    o = o - o.max(axis=1, keepdims=True)
    e = np.exp(o)
    return (e / e.sum(axis=1, keepdims=True)).astype(f)


def policy_sample(probs, n_envs, n_agents, seed, stream):
    """Action index drawn by the device policy (policy_kernel): u = top 24 bits of Philox4x32-10
    keyed by the seed at counter (environment, agent, stream, 1), first index whose running fp32
    sum of probabilities exceeds u (the last index if none does)."""
    probs = np.asarray(probs, np.float32).reshape(n_envs, n_agents, -1)
    idx = np.zeros((n_envs, n_agents), np.uint8)
    for e in range(n_envs):
        for i in range(n_agents):
            r = _philox4x32_10(e, i, stream, 1, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)[0]
            u = np.float32(r >> 8) * np.float32(2.0 ** -24)
            c = np.float32(0)
            pick = probs.shape[-1] - 1
            for a in range(probs.shape[-1]):
                c = np.float32(c + probs[e, i, a])
                if u < c:
                    pick = a
                    break
            idx[e, i] = pick
    return idx

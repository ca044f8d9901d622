"""BASELINE INFRASTRUCTURE ONLY -- time the UNMODIFIED reference ``drones.step()`` on host cores.

Used by ``bench.py``'s CPU-baseline legs only (``cpu_baseline_reference`` of the GPU arm and
``--impl reference``).  The reference (``/root/reference`` here, the staged copy ``oracle/_ref``
on a GPU box) is imported through ``oracle/ref_harness`` and driven exactly as its drivers drive it
(``train_problem.py:82-107``): ``env.step(actions)`` until ``finished``, then ``env.reset()``.
Actions: uniform over the 16 unit directions of ``DiscreteSoftmaxNN(n_actions=16)``
(``utils.py:262-269``) -- the same action set the GPU arm streams.  The reference is single
threaded; ``procs`` independent processes each own one environment.
"""
from __future__ import annotations

import os
import time

import numpy as np

_ENV = {}


def available() -> bool:
    from oracle import ref_harness
    return ref_harness.reference_available()


def _get_env(n, grid, delta):
    key = (n, tuple(grid), float(delta))
    env = _ENV.get(key)
    if env is None:
        import contextlib
        import io
        from oracle import ref_harness
        mod = ref_harness.import_reference("drone_env")
        with contextlib.redirect_stdout(io.StringIO()):       # the ctor prints a warning when it clips deltas
            env = mod.drones(n_agents=n, n_obstacles=0, grid=list(grid), end_formation="O", k_closest=2,
                             deltas=np.ones(n) * delta, simplify_zstate=True)
        env.collision_weight = 0.2                            # train_problem.py:31
        _ENV[key] = env
    return env


def run_episodes(args):
    """(n, grid, delta, episodes, steps_per_episode, seed) -> (agent_steps, seconds) in THIS process."""
    n, grid, delta, episodes, T, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import random
    env = _get_env(n, grid, delta)
    rng = np.random.default_rng(seed)
    random.seed(seed)
    ang = 2 * np.pi * np.arange(16) / 16
    table = np.stack([np.cos(ang), np.sin(ang)], 1)
    done_steps = 0
    t0 = time.perf_counter()
    for _ in range(episodes):
        env.reset(renew_obstacles=False)
        for _t in range(T):
            acts = [table[a] for a in rng.integers(0, 16, n)]
            _s, _z, _r, _c, finished, _tr = env.step(acts)
            done_steps += 1
            if finished:
                break
    return done_steps * n, time.perf_counter() - t0


class Pool:
    """`procs` persistent worker processes (plain subprocesses speaking one JSON line per job over
    pipes: independent of how the parent was started and of any CUDA context it holds), one
    reference environment each."""

    def __init__(self, procs):
        import subprocess
        import sys
        self.procs = int(procs)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1",
                   CUDA_VISIBLE_DEVICES="")
        self.workers = [subprocess.Popen([sys.executable, "-u", "-m", "oracle.ref_cpu", "--worker"], cwd=root, env=env,
                                         stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
                        for _ in range(self.procs)]

    def run(self, n, grid, delta, episodes_per_proc, T, seed0=0):
        """One bench 'step': every process runs `episodes_per_proc` episodes.  Returns
        (agent_steps, wall_seconds)."""
        import json
        t0 = time.perf_counter()
        for p, w in enumerate(self.workers):
            w.stdin.write(json.dumps([n, list(grid), float(delta), int(episodes_per_proc), int(T), seed0 + 1000 * p]) + "\n")
            w.stdin.flush()
        res = []
        for w in self.workers:
            line = w.stdout.readline()
            if not line:
                raise RuntimeError("reference worker died")
            res.append(json.loads(line))
        wall = time.perf_counter() - t0
        return sum(r[0] for r in res), wall

    def close(self):
        for w in self.workers:
            try:
                w.stdin.close()
                w.wait(timeout=5)
            except Exception:
                w.kill()


def _worker_main():
    import json
    import sys
    for line in sys.stdin:
        line = line.strip()
        if not line:
            continue
        steps, sec = run_episodes(tuple(json.loads(line)))
        sys.stdout.write(json.dumps([steps, sec]) + "\n")
        sys.stdout.flush()


def rate(n, grid, delta, procs, budget_s, T=200):
    """agent-steps/s of the reference on `procs` processes over roughly `budget_s` seconds (after a
    warm-up episode per process).  Returns (rate, sample description)."""
    pool = Pool(procs)
    try:
        pool.run(n, grid, delta, 1, min(T, 20))                      # import + warm-up
        steps, wall = pool.run(n, grid, delta, 1, T, seed0=1)        # calibration episode
        per_ep = max(wall, 1e-3)
        eps = max(1, int((budget_s - wall) / per_ep))
        if eps > 1 or wall < 0.5 * budget_s:
            s2, w2 = pool.run(n, grid, delta, eps, T, seed0=2)
            steps, wall = s2, w2
        return steps / wall, f"{steps // n} env-steps of the unmodified NumPy reference over {procs} process(es), {wall:.1f} s"
    finally:
        pool.close()


if __name__ == "__main__":
    import sys
    if "--worker" in sys.argv:
        _worker_main()

"""CPU, world_size 2, gloo: the N>1 host logic -- env sharding and the single all-reduce of
episode aggregates -- using the C oracle as the stand-in stepper for each rank's shard."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, E_total, n, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    from scalable_collision_avoidance_rl_b200 import dist as dsdist, formation
    lo, hi = dsdist.shard_envs(E_total, rank, world)
    rng = np.random.default_rng(7)
    xF = formation.end_formation("O", n, [5, 5])
    ds = formation.safety_distances(xF, np.ones(n) * 0.1)
    deltas, _ = formation.clip_deltas(np.ones(n), ds)
    start = formation.sample_start_batched(E_total, n, [5, 5], rng)
    act = rng.uniform(-1, 1, (T, E_total, n, 2))
    env = c_oracle.OracleEnv(hi - lo, n, xF, ds, deltas, None, 2, True)
    env.set_state(start[lo:hi])
    res = env.rollout(np.ascontiguousarray(act[:, lo:hi]))
    agg5 = torch.cat([torch.as_tensor(res["agg"].sum(0)), torch.tensor([float(hi - lo)])])
    dsdist.allreduce_episode_aggregates(agg5)
    gathered = dsdist.gather_env_returns(torch.as_tensor(res["agg"])) if E_total % world == 0 else None
    if rank == 0:
        q.put((agg5.numpy(), None if gathered is None else gathered.numpy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("E_total", [10, 7])
def test_sharded_rollout_allreduce_matches_single_process(E_total):
    from oracle import c_oracle
    from scalable_collision_avoidance_rl_b200 import dist as dsdist, formation
    n, T, world = 5, 25, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, E_total, n, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    agg5, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth over all environments
    rng = np.random.default_rng(7)
    xF = formation.end_formation("O", n, [5, 5])
    ds = formation.safety_distances(xF, np.ones(n) * 0.1)
    deltas, _ = formation.clip_deltas(np.ones(n), ds)
    start = formation.sample_start_batched(E_total, n, [5, 5], rng)
    act = rng.uniform(-1, 1, (T, E_total, n, 2))
    env = c_oracle.OracleEnv(E_total, n, xF, ds, deltas, None, 2, True)
    env.set_state(start)
    res = env.rollout(act)
    want = np.concatenate([res["agg"].sum(0), [E_total]])
    assert np.allclose(agg5, want, rtol=0, atol=1e-9)
    if gathered is not None:
        assert np.array_equal(gathered, res["agg"])
    s = dsdist.episode_summary(torch.as_tensor(agg5))
    assert s["n_envs"] == E_total and abs(s["steps"] - T) < 1e-12


def test_shard_envs_partition():
    from scalable_collision_avoidance_rl_b200 import dist as dsdist
    for E, G in [(8192, 8), (1024, 8), (7, 2), (5, 8), (4096, 1)]:
        blocks = [dsdist.shard_envs(E, r, G) for r in range(G)]
        assert blocks[0][0] == 0 and blocks[-1][1] == E
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dsdist.shard_envs(8, 2, 2)
    with pytest.raises(ValueError):
        dsdist.allreduce_episode_aggregates(torch.zeros(4, dtype=torch.float64))

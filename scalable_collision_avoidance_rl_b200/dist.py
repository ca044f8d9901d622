"""Multi-GPU plumbing: environments shard across ranks, one all-reduce per episode.

Environments are fully independent (no cross-env term anywhere on the path, reference
drone_env.py:214-401), so the data path needs NO collective: rank g of G owns a
contiguous block of environments and steps it with its own handle and stream.  The only
exchange mirrors what the reference's driver accumulates per episode
(train_problem.py:98-100,118-121): the 5-vector
    (sum_t mean_i r, sum_t mean_i true_r, sum_t n_collisions, steps, #envs)
reduced on the device by ds_reduce_aggregates and summed across ranks with a single
all-reduce (NCCL on GPUs; gloo on CPU for the host-logic tests), enqueued on the rollout
stream with no host synchronisation.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

AGG_FIELDS = ("sum_mean_reward", "sum_mean_true_reward", "sum_collisions", "steps", "n_envs")


def shard_envs(n_envs_total: int, rank: int, world_size: int):
    """Contiguous block [lo, hi) of environments owned by `rank` (sizes differ by <= 1)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(n_envs_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_episode_aggregates(agg5: torch.Tensor, group=None, async_op: bool = False):
    """In-place SUM all-reduce of the float64 [5] episode aggregate vector.  async_op=True: the
    collective runs on NCCL's own stream behind the rollout that produced agg5 and the caller's
    stream does NOT wait for it (the next episode does not depend on the aggregate); returns the
    work handle to wait() on before the vector is read or its buffer reused (None when there is
    nothing to reduce)."""
    if agg5.dtype != torch.float64 or agg5.numel() != len(AGG_FIELDS):
        raise ValueError("expected the float64 [5] vector of BatchedDrones.episode_aggregates()")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(agg5, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return work if async_op else agg5
    return None if async_op else agg5


def episode_summary(agg5: torch.Tensor) -> dict:
    """Per-environment means of the global aggregates, as the reference's progress bar reports
    them per episode (train_problem.py:135-140)."""
    v = agg5.detach().cpu().tolist()
    n = max(v[4], 1.0)
    return {"reward": v[0] / n, "true_reward": v[1] / n, "collisions": v[2] / n, "steps": v[3] / n,
            "n_envs": int(v[4])}


def gather_env_returns(agg: torch.Tensor, group=None) -> torch.Tensor:
    """Optional all-gather of the per-environment accumulators [E_local,4] -> [E_total,4]
    (equal shard sizes required)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return agg
    out = [torch.empty_like(agg) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, agg.contiguous(), group=group)
    return torch.cat(out, 0)

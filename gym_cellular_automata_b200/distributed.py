"""Multi-GPU plumbing of the vector env: one process per GPU, environments sharded across ranks.

The path shards trivially: environments never interact, shared_context is
read-only and replicated, so a step needs NO data-path collective.  The only collective is the
all-gather of per-env episode statistics once per rollout -- the reference's (disabled)
``gather_stats`` (agents/jax_ppo.py:1330-1343).  NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(num_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """Env e lives on rank e // ceil(num_envs / world): returns this rank's [lo, hi)."""
    per = (num_envs + world - 1) // world
    lo = min(rank * per, num_envs)
    return lo, min(lo + per, num_envs)


def gather_episode_stats(stats: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
    """All-gather every per-env leaf into a (world * N_local, ...) tensor ordered by rank, i.e. by global env index --
    the reference's ``gather_stats`` (agents/jax_ppo.py:1330-1343).  The 4-byte (N_local,) leaves (float32 / int32: all
    fields of ``rollout_stats.EpisodeStatistics`` and the env's info counters) travel in ONE collective, packed as bit
    patterns; anything else is gathered leaf by leaf.  No-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()):
        return dict(stats)
    world = dist.get_world_size(group)
    out = {}
    packed = [k for k, v in stats.items() if v.dim() == 1 and v.element_size() == 4]
    if packed:
        n = stats[packed[0]].shape[0]
        packed = [k for k in packed if stats[k].shape[0] == n]
        send = torch.stack([stats[k].contiguous().view(torch.int32) for k in packed], dim=0).reshape(-1)  # L x n bit patterns
        recv = torch.empty(world * send.numel(), dtype=torch.int32, device=send.device)
        dist.all_gather_into_tensor(recv, send, group=group)
        recv = recv.view(world, len(packed), n)
        for i, k in enumerate(packed):
            out[k] = recv[:, i, :].reshape(world * n).view(stats[k].dtype)
    for k, v in stats.items():
        if k in out:
            continue
        v = v.contiguous()
        buf = torch.empty((world * v.shape[0],) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        dist.all_gather_into_tensor(buf, v, group=group)
        out[k] = buf
    return out

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


class EpisodeStatistics:
    """Per-env running episode statistics with the fields of the reference's EpisodeStatistics
    (agents/jax_ppo.py:379-398): running returns / lengths and the values of the last finished
    episode.  Plain tensor ops on (N,) vectors; lives on whatever device the env uses."""

    def __init__(self, num_envs: int, device="cpu"):
        z = lambda dt: torch.zeros(num_envs, dtype=dt, device=device)  # noqa: E731
        self.episode_returns = z(torch.float32)
        self.episode_lengths = z(torch.int32)
        self.returned_episode_returns = z(torch.float32)
        self.returned_episode_lengths = z(torch.int32)

    def update(self, reward: torch.Tensor, terminated: torch.Tensor) -> None:
        done = terminated.bool()
        ret = self.episode_returns + reward
        ln = self.episode_lengths + 1
        self.returned_episode_returns = torch.where(done, ret, self.returned_episode_returns)
        self.returned_episode_lengths = torch.where(done, ln, self.returned_episode_lengths)
        self.episode_returns = torch.where(done, torch.zeros_like(ret), ret)
        self.episode_lengths = torch.where(done, torch.zeros_like(ln), ln)

    def as_dict(self) -> Dict[str, torch.Tensor]:
        return {k: getattr(self, k) for k in ("episode_returns", "episode_lengths", "returned_episode_returns",
                                              "returned_episode_lengths")}


def gather_episode_stats(stats: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
    """All-gather every (N_local,) leaf into a (world * N_local,) tensor ordered by rank, i.e. by
    global env index.  No-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()):
        return dict(stats)
    world = dist.get_world_size(group)
    out = {}
    for k, v in stats.items():
        v = v.contiguous()
        buf = torch.empty((world * v.shape[0],) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        dist.all_gather_into_tensor(buf, v, group=group)
        out[k] = buf
    return out

"""Episode statistics of the rollout loop that drives the env (the caller side of the hot path).

Mirror of ``EpisodeStatistics`` and of the statistics half of ``step_env_wrapped`` in the reference's
``agents/jax_ppo.py:380-398,486-501,504-655``: per-env return / length accumulators, the
day / night extension-correctness counters, ``returned_episode_*`` latches and the 10-entry ring
buffers of recently finished episodes.  The reference walks the envs with a serial ``lax.scan``
(``jax_ppo.py:543-621``); here one launch of ``gca_episode_stats_update`` does the same with a
block scan, on the tensors owned by this object (device-resident, updated in place).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, current_stream, load, ptr

RECENT = 10
_PER_ENV = {"episode_returns": torch.float32, "episode_lengths": torch.int32,
            "returned_episode_returns": torch.float32, "returned_episode_lengths": torch.int32,
            "current_day_correct": torch.int32, "current_night_correct": torch.int32,
            "current_day_steps": torch.int32, "current_night_steps": torch.int32}
_RING = {"recent_returns": torch.float32, "recent_lengths": torch.int32, "recent_day_correct": torch.int32,
         "recent_night_correct": torch.int32, "recent_day_steps": torch.int32, "recent_night_steps": torch.int32}
_SCALAR = {"amount_finished": torch.int32, "recent_idx": torch.int32}


class EpisodeStatistics:
    """All fields of the reference dataclass as zero-initialised CUDA tensors (jax_ppo.py:486-501)."""

    def __init__(self, num_envs: int, device=None):
        self.num_envs = int(num_envs)
        self.device = torch.device("cuda" if device is None else device)
        if self.device.type != "cuda":
            raise _lib.GcaError("EpisodeStatistics lives on a CUDA device (libgca has no CPU path)")
        for name, dt in _PER_ENV.items():
            setattr(self, name, torch.zeros(self.num_envs, dtype=dt, device=self.device))
        for name, dt in _RING.items():
            setattr(self, name, torch.zeros(RECENT, dtype=dt, device=self.device))
        for name, dt in _SCALAR.items():
            setattr(self, name, torch.zeros(1, dtype=dt, device=self.device))
        self._c = _lib.GcaEpisodeStats(**{f: getattr(self, f).data_ptr() for f in _lib._EPISODE_FIELDS})

    def update(self, actions: torch.Tensor, step_reward: torch.Tensor, terminated: torch.Tensor,
               obs_night: torch.Tensor, truncated: torch.Tensor | None = None) -> "EpisodeStatistics":
        """One rollout step: ``actions`` (N,3) int32, ``step_reward`` = info["reward"] (N,) float32,
        ``terminated`` / ``truncated`` (N,) uint8 (or bool), ``obs_night`` (N,) uint8 = is_night of the
        observation the actions were chosen on.  Call it after the env step, on the same stream.  ``actions`` may be
        the pinned host tensor handed to ``env.step_host`` (read in place, no copy)."""
        def u8(t):
            return None if t is None else (t.view(torch.uint8) if t.dtype == torch.bool else t)
        terminated, truncated, obs_night = u8(terminated), u8(truncated), u8(obs_night)
        N = self.num_envs
        check(load().gca_episode_stats_update(
            N, C.byref(self._c), ptr(step_reward, torch.float32, N, "step_reward"),
            ptr(terminated, torch.uint8, N, "terminated"), ptr(truncated, torch.uint8, N, "truncated"),
            ptr(obs_night, torch.uint8, N, "obs_night"), ptr(actions, torch.int32, 3 * N, "actions", allow_pinned=True),
            current_stream()), "gca_episode_stats_update")
        return self

    def as_dict(self):
        return {f: getattr(self, f) for f in _lib._EPISODE_FIELDS}

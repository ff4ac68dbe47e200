"""Benchmark workload preparation: bring a batch of envs to a STATIONARY, phase-desynchronised mixture.

All envs of a batch leave ``reset()`` in phase (two burning cells each), the cost of an env step follows the
size of the fire fronts through the episode, and the fused auto-reset restores the same snapshot
(reference advanced_bulldozer.py:422-518), so phases do not mix by themselves for hundreds of steps.
``stationary_preroll`` runs ``horizon`` env steps during SETUP and force-resets env group g (envs e with
e % groups == g) at pre-roll step g * horizon / groups through ``gca_conditional_reset``: afterwards the
groups' episode ages are spread evenly over (0, horizon], which with ``horizon`` = one typical episode is the
long-run mixture of a rollout loop -- whatever window a driver then times.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, current_stream, load, ptr


def random_actions(n_steps: int, num_envs: int, device, generator: torch.Generator) -> torch.Tensor:
    """(n_steps, N, 3) int32 action triples: move U{0..8}, shoot U{0,1}, extension id U{0..2}
    (total_action_space.sample(), reference scripts/run:630)."""
    return torch.stack([torch.randint(0, 9, (n_steps, num_envs), device=device, generator=generator),
                        torch.randint(0, 2, (n_steps, num_envs), device=device, generator=generator),
                        torch.randint(0, 3, (n_steps, num_envs), device=device, generator=generator)],
                       dim=-1).to(torch.int32).contiguous()


def force_reset(env, mask_u8: torch.Tensor) -> None:
    """Restore the envs with mask != 0 from the env's reset snapshot (conditional_reset on a chosen set)."""
    st = env._state
    reward = torch.zeros(env.num_envs, dtype=torch.float32, device=env.device)
    m = mask_u8.clone()  # the call clears it
    check(load().gca_conditional_reset(C.byref(env._params), C.byref(st.cstruct()), C.byref(env._snapshot.cstruct()),
                                       ptr(env._snap_reward), ptr(reward), ptr(m), current_stream()),
          "gca_conditional_reset")
    env._version += 1


def stationary_preroll(env, horizon: int, groups: int, seed: int = 0, chunk: int = 64, settle: int = 96) -> dict:
    """Pre-roll ``horizon`` env steps with random actions, force-resetting group g at step g*horizon/groups, then
    ``settle`` free-running steps (the last groups' young fires grow in, the load balancer re-deals the envs).
    Returns {"horizon", "groups", "steps"}; the env must have been reset() and use auto_reset."""
    N, dev = env.num_envs, env.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x5EED0000 + seed)
    gid = torch.arange(N, device=dev) % groups
    stride = max(1, horizon // groups)
    done = 0
    while done < horizon:
        n = min(chunk, horizon - done)
        acts = random_actions(n, N, dev, gen)
        for i in range(n):
            t = done + i
            if t > 0 and t % stride == 0 and t // stride < groups:
                force_reset(env, (gid == t // stride).to(torch.uint8))
            env.step_device(acts[i])
        done += n
    left = settle
    while left > 0:
        n = min(chunk, left)
        acts = random_actions(n, N, dev, gen)
        for i in range(n):
            env.step_device(acts[i])
        left -= n
    return {"horizon": horizon, "groups": groups, "steps": done + settle}

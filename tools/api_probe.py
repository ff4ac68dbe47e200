#!/usr/bin/env python
"""Diagnostics: wall-clock cost per env step of the reference-style rollout loop through the mirrored API
(stateless_step + conditional_reset, observations on), against the fused device calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
N, steps, warm = 4096, 128, 100
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(0)
acts = torch.stack([torch.randint(0, 9, (warm + steps, N), device=dev, generator=gen), torch.randint(0, 2, (warm + steps, N), device=dev, generator=gen),
                    torch.randint(0, 3, (warm + steps, N), device=dev, generator=gen)], -1).to(torch.int32).contiguous()
def make(obs_mode, ext, auto_reset):
    env = AdvancedForestFireBulldozerEnv(64, 64, key=1, num_envs=N, speed_move=0.48, speed_act=0.12, use_hidden=True, substeps=4,
                                         rng_mode="legacy", seed=0, hidden="random", obs_mode=obs_mode, auto_reset=auto_reset,
                                         collect_stats=True, device=dev, balance_every=8, enable_extensions=ext)
    obs, info = env.reset()
    for i in range(warm): env.step_device(acts[i])
    torch.cuda.synchronize()
    return env, obs, info
def timed(name, f):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(steps): f(i)
    torch.cuda.synchronize(); print("%-64s: %.1f us per step" % (name, (time.perf_counter() - t0) / steps * 1e6), flush=True)
for obs_mode, ext in (("rgb_f32", True), ("rgb_f32", False), ("rgb_u8", False), ("none", False)):
    env, obs, info = make(obs_mode, ext, False)
    def loop(i, env=env):
        a = acts[warm + i]
        st = env.stateless_step(a)
        env.conditional_reset(st, a)
    timed(f"stateless_step + conditional_reset, obs {obs_mode}, ext {ext}", loop)
    env, obs, info = make(obs_mode, ext, True)
    timed(f"stateless_step with fused auto-reset,  obs {obs_mode}, ext {ext}", lambda i, env=env: env.stateless_step(acts[warm + i]))
    timed(f"step_device + observe_device,           obs {obs_mode}, ext {ext}", lambda i, env=env: (env.step_device(acts[warm + i]), env.observe_device(acts[warm + i])))

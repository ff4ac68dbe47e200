#!/usr/bin/env python
"""Diagnostics: a short run of step_device + observe_device (for an ncu launch list of the observation path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
N = 4096
mode = sys.argv[1] if len(sys.argv) > 1 else "rgb_u8"
ext = len(sys.argv) > 2 and sys.argv[2] == "ext"
dev = torch.device("cuda", 0)
env = AdvancedForestFireBulldozerEnv(64, 64, key=1, num_envs=N, speed_move=0.48, speed_act=0.12, use_hidden=True, substeps=4,
                                     rng_mode="legacy", seed=0, hidden="random", obs_mode=mode, auto_reset=True,
                                     collect_stats=True, device=dev, balance_every=8, enable_extensions=ext)
env.reset()
gen = torch.Generator(device=dev); gen.manual_seed(0)
acts = torch.stack([torch.randint(0, 9, (40, N), device=dev, generator=gen), torch.randint(0, 2, (40, N), device=dev, generator=gen),
                    torch.randint(0, 3, (40, N), device=dev, generator=gen)], -1).to(torch.int32).contiguous()
for i in range(40):
    env.step_device(acts[i])
    rgb = env.observe_device(acts[i])
torch.cuda.synchronize()
print("ok", rgb.shape, rgb.dtype)

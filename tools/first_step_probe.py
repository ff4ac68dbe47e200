"""Diagnostics: the first kernel after an idle gap runs ~50 us slower -- why bench.py primes the timed region."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
from gym_cellular_automata_b200.workload import random_actions, stationary_preroll
dev = torch.device("cuda", 0)
env = AdvancedForestFireBulldozerEnv(64, 64, key=1, num_envs=4096, speed_move=0.48, speed_act=0.12, use_hidden=True, substeps=4,
                                     seed=0, hidden="random", obs_mode="none", auto_reset=True, collect_stats=True, device=dev, balance_every=8)
env.reset(); stationary_preroll(env, 512, 32)
gen = torch.Generator(device=dev); gen.manual_seed(0)
acts = random_actions(400, 4096, dev, gen)
flush = (torch.empty(256 << 20, dtype=torch.uint8, device=dev), torch.zeros(256 << 20, dtype=torch.uint8, device=dev), torch.zeros(1, dtype=torch.int64, device=dev))
for i in range(5): env.step_device(acts[i])
torch.cuda.synchronize()
for rep in range(4):
    us = bench.timed_steps(env, acts, 5 + 20 * rep, 20, flush, torch)
    print("rep", rep, [round(float(x), 1) for x in us[:6]], "mean %.1f" % us.mean())
    if rep == 1:
        s = env.stats()
    if rep == 2:
        import time; time.sleep(0.5)

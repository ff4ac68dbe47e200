#!/usr/bin/env python
"""tools/phase_series.py OUT.json [--steps S] [--preroll H --groups G]
Per-step series of the 64x64 bench workload (4096 envs, K = 4, L2 flushed): CUDA-event time of each env step,
front cells / draws per env sub-step and terminations per step -- from reset with all envs in phase (default), or
after workload.stationary_preroll.  Diagnostics, not part of the product path."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
from gym_cellular_automata_b200.workload import random_actions, stationary_preroll

ap = argparse.ArgumentParser()
ap.add_argument("out")
ap.add_argument("--steps", type=int, default=1200)
ap.add_argument("--preroll", type=int, default=0)
ap.add_argument("--groups", type=int, default=16)
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--substeps", type=int, default=4)
ap.add_argument("--stats-every", type=int, default=8)
a = ap.parse_args()
N, K = a.envs, a.substeps
dev = torch.device("cuda", 0)
env = AdvancedForestFireBulldozerEnv(64, 64, key=1, num_envs=N, speed_move=0.48, speed_act=0.12, use_hidden=True,
                                     substeps=K, rng_mode="legacy", seed=0, hidden="random", obs_mode="none",
                                     auto_reset=True, collect_stats=True, device=dev, balance_every=8)
env.reset()
if a.preroll:
    stationary_preroll(env, a.preroll, a.groups)
gen = torch.Generator(device=dev)
gen.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
times, terms, fronts, draws = [], [], [], []
term_acc = torch.zeros((), dtype=torch.int64, device=dev)
done = 0
prev = env.stats()
while done < a.steps:
    n = min(a.stats_every, a.steps - done)
    acts = random_actions(n, N, dev, gen)
    evs = []
    term_acc.zero_()
    for i in range(n):
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = env.step_device(acts[i])
        e1.record()
        term_acc += out.terminated.sum()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    times += [e0.elapsed_time(e1) * 1e3 for e0, e1 in evs]
    cur = env.stats()
    d = (cur - prev).astype(np.float64)
    prev = cur
    sub = max(d[5] * K, 1.0)
    fronts.append(d[0] / sub)
    draws.append(d[1] / sub)
    terms.append(int(term_acc.item()))
    done += n
t = np.array(times)
res = {"envs": N, "K": K, "preroll": a.preroll, "groups": a.groups, "steps": a.steps, "stats_every": a.stats_every,
       "us_mean": float(t.mean()), "us_per_block": [float(x) for x in t.reshape(-1, a.stats_every).mean(1)],
       "front_per_env_substep": fronts, "draws_per_env_substep": draws, "terminated_per_block": terms}
json.dump(res, open(a.out, "w"))
print("mean us/step %.1f ; first/last 100: %.1f / %.1f ; max block %.1f min block %.1f ; terminations %d" % (
    t.mean(), t[:100].mean(), t[-100:].mean(), max(res["us_per_block"]), min(res["us_per_block"]), sum(terms)))

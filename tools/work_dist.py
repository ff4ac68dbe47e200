#!/usr/bin/env python
"""Diagnostics: distribution of per-env step cost (SM cycles per warp, GCA_FLAG_WORK_CYCLES) of the
64x64 kernel on the bench workload, next to the kernel duration.  Not part of the product path."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_cellular_automata_b200 import _lib
from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv

N = int(os.environ.get("N", 4096)); K = 4; warm = int(os.environ.get("WARM", 100))
dev = torch.device("cuda", 0)
env = AdvancedForestFireBulldozerEnv(64, 64, key=1, num_envs=N, speed_move=0.48, speed_act=0.12, use_hidden=True,
                                     substeps=K, rng_mode="legacy", seed=0, hidden="random", obs_mode="none",
                                     auto_reset=True, collect_stats=True, device=dev, balance_every=0)
env.reset()
gen = torch.Generator(device=dev); gen.manual_seed(0)
def act():
    return torch.stack([torch.randint(0, 9, (N,), device=dev, generator=gen), torch.randint(0, 2, (N,), device=dev, generator=gen),
                        torch.randint(0, 3, (N,), device=dev, generator=gen)], -1).to(torch.int32).contiguous()
for _ in range(warm): env.step_device(act())
torch.cuda.synchronize()
out = {}
for rep in range(3):
    a = act()
    env._flags |= _lib.FLAG_WORK_CYCLES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.step_device(a); e1.record(); torch.cuda.synchronize()
    cyc = env._state.work.cpu().numpy().astype(np.int64)
    env._flags &= ~_lib.FLAG_WORK_CYCLES
    env.step_device(act()); torch.cuda.synchronize()
    est = env._state.work.cpu().numpy().astype(np.int64)
    us = e0.elapsed_time(e1) * 1e3
    q = lambda p: float(np.percentile(cyc, p))
    out[rep] = {"kernel_us": us, "kernel_cycles_at_1965MHz": us * 1965, "cyc_mean": float(cyc.mean()), "cyc_p50": q(50),
                "cyc_p90": q(90), "cyc_p99": q(99), "cyc_max": int(cyc.max()), "est_mean": float(est.mean()),
                "est_max": int(est.max()), "est_p99": float(np.percentile(est, 99))}
    print(json.dumps(out[rep]))
    for _ in range(20): env.step_device(act())

#!/usr/bin/env python
"""Diagnostics: where the end-to-end step time goes (host call overhead, copies, synchronisation)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
N, K, steps = 4096, 4, 256
dev = torch.device("cuda", 0)
env = AdvancedForestFireBulldozerEnv(64, 64, key=1, num_envs=N, speed_move=0.48, speed_act=0.12, use_hidden=True, substeps=K,
                                     rng_mode="legacy", seed=0, hidden="random", obs_mode="none", auto_reset=True,
                                     collect_stats=True, device=dev, balance_every=8)
env.reset()
gen = torch.Generator(device=dev); gen.manual_seed(0)
acts = torch.stack([torch.randint(0, 9, (100 + 4 * steps, N), device=dev, generator=gen), torch.randint(0, 2, (100 + 4 * steps, N), device=dev, generator=gen),
                    torch.randint(0, 3, (100 + 4 * steps, N), device=dev, generator=gen)], -1).to(torch.int32).contiguous()
for i in range(100): env.step_device(acts[i])
torch.cuda.synchronize()
h_act = acts[100:].cpu().pin_memory()
h_rew, h_term = env.host_result_buffers()
def timed(f, n=steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): f(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
st = torch.cuda.current_stream()
print("step_host zero-copy (mapped pinned buffers): %.1f us" % timed(lambda i: env.step_host(h_act[i], h_rew, h_term)))
print("step_host staged (H2D + step + D2H + sync) : %.1f us" % timed(lambda i: env.step_host(h_act[i], h_rew, h_term, staged=True)))
print("step_host zero-copy again                  : %.1f us" % timed(lambda i: env.step_host(h_act[i], h_rew, h_term)))
print("step_device + stream sync (no copies)      : %.1f us" % timed(lambda i: (env.step_device(acts[100 + steps + i]), st.synchronize())))
print("step_device back to back (no sync)         : %.1f us" % timed(lambda i: env.step_device(acts[100 + 2 * steps + i])))
d_act = torch.empty((N, 3), dtype=torch.int32, device=dev)
print("torch copies + step_device + sync          : %.1f us" % timed(lambda i: (d_act.copy_(h_act[3 * steps + i], non_blocking=True), env.step_device(d_act), h_rew.copy_(env._out.reward, non_blocking=True), st.synchronize())))
print("python-only overhead of h_act[i] indexing  : %.2f us" % timed(lambda i: h_act[i]))

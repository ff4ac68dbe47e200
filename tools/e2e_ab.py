#!/usr/bin/env python
"""Diagnostics: A/B of the transports of gca_env_step_host at the SAME episode phase (a fresh env of the same seed
per variant: 100 warm-up steps, then 256 timed steps)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_cellular_automata_b200 import _lib
from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
N, K, steps, warm = 4096, 4, 256, 100
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(0)
acts = torch.stack([torch.randint(0, 9, (warm + steps, N), device=dev, generator=gen), torch.randint(0, 2, (warm + steps, N), device=dev, generator=gen),
                    torch.randint(0, 3, (warm + steps, N), device=dev, generator=gen)], -1).to(torch.int32).contiguous()
h_act = acts[warm:].cpu().pin_memory()
st = torch.cuda.current_stream()
def run(name, f, reps=2):
    best = []
    for _ in range(reps):
        env = AdvancedForestFireBulldozerEnv(64, 64, key=1, num_envs=N, speed_move=0.48, speed_act=0.12, use_hidden=True, substeps=K,
                                             rng_mode="legacy", seed=0, hidden="random", obs_mode="none", auto_reset=True,
                                             collect_stats=True, device=dev, balance_every=8)
        env.reset()
        if os.environ.get("PREROLL"):
            from gym_cellular_automata_b200.workload import stationary_preroll
            stationary_preroll(env, int(os.environ["PREROLL"]), 32)
        for i in range(warm): env.step_device(acts[i])
        h_rew, h_term = env.host_result_buffers()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(steps): f(env, i, h_rew, h_term)
        torch.cuda.synchronize(); best.append((time.perf_counter() - t0) / steps * 1e6)
    print("%-52s: %s us" % (name, " / ".join("%.1f" % b for b in best)), flush=True)
if os.environ.get("E2E_AB_QUICK"):
    run("step_host zero-copy in + out", lambda e, i, r, t: e.step_host(h_act[i], r, t), reps=3)
    sys.exit(0)
run("step_device back to back (no sync, no copies)", lambda e, i, r, t: e.step_device(acts[warm + i]))
run("step_device + stream sync (no copies)", lambda e, i, r, t: (e.step_device(acts[warm + i]), st.synchronize()))
run("step_host zero-copy in + out", lambda e, i, r, t: e.step_host(h_act[i], r, t))
run("step_host zero-copy in, staged out", lambda e, i, r, t: e.step_host(h_act[i], r, t, staged=_lib.FLAG_HOST_COPY_OUT))
run("step_host staged in, zero-copy out", lambda e, i, r, t: e.step_host(h_act[i], r, t, staged=_lib.FLAG_HOST_COPY_IN))
run("step_host staged in + out", lambda e, i, r, t: e.step_host(h_act[i], r, t, staged=True))

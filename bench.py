#!/usr/bin/env python
"""bench.py -- throughput of the fused environment step (BASELINE.json metric: CA cell-updates/s
and env-steps/s, 64x64, N envs) on 1..8 B200, with the roofline, the end-to-end figure through
the host API and the CPU baseline beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one env step (clock + K_sub CA sub-steps + move + douse + reward + done +
auto-reset) of every env of the batch = ONE launch of env_step64_kernel per GPU.
Workload = BASELINE.json configs[1]: 64x64 Alexandridis env, 4096 envs per GPU, speed-multiplier 4
(K_sub = 4 CA sub-steps per env step), hidden layers on, synthetic random hidden layers, random
actions.  Envs are independent -> sharded across ranks with no data-path collective (weak
scaling); NCCL only all-gathers the per-env episode statistics.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALGO_BYTES_PER_CELL = 7  # read cell u8 + age u16 + hidden u8, write cell u8 + age u16
ALGO_BYTES_PER_ENV = 64
# Workload phase (SURVEY 8d asks for a developed fire front): every env made by this file is first stepped
# PREROLL_HORIZON times in SETUP while env group g (e % PREROLL_GROUPS == g) is force-reset at pre-roll step
# g * PREROLL_HORIZON / PREROLL_GROUPS (gym_cellular_automata_b200/workload.py), so the batch is a stationary
# mixture of episode phases whatever --steps / --warmup a driver passes.
PREROLL_HORIZON = 512
PREROLL_GROUPS = 32


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--substeps", type=int, default=4)
    ap.add_argument("--rng-mode", default="legacy")
    ap.add_argument("--no-hidden", action="store_true")
    ap.add_argument("--hidden", default="random", choices=["random", "device", "reference"],
                    help="hidden layers: iid synthetic (host NumPy), generated on the device (large grids), reference-like")
    ap.add_argument("--generic-tiles", action="store_true", help="grids other than 64x64: force the generic tiled kernels")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--flush-mode", default="write+read", choices=["write", "write+read"],
                    help="L2 flush between timed steps: a 256 MiB write, or the write followed by a 256 MiB read sweep (clean L2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the BASELINE config 3 / 4 lines (other_configs)")
    ap.add_argument("--no-obs-leg", action="store_true", help="skip the extra device-timed leg with RGB observations")
    ap.add_argument("--cpu-envs", type=int, default=256, help="envs of the CPU sample (cpu_baseline leg and --impl reference)")
    ap.add_argument("--cpu-steps", type=int, default=1024, help="env steps of the cpu_baseline sample (about 10-20 s of host work)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--ruleset", default="alexandridis", choices=["alexandridis", "v3"],
                    help="v3 = the registered ForestFireBulldozer256x256-v3 rule set (WindyForestFire), an extra line")
    ap.add_argument("--preroll", type=int, default=PREROLL_HORIZON,
                    help="setup: env steps of the phase-desynchronising pre-roll (0 = all envs start in phase from reset)")
    ap.add_argument("--preroll-groups", type=int, default=PREROLL_GROUPS)
    ap.add_argument("--two-group", action="store_true",
                    help="also time the end-to-end loop with the batch split into two asynchronous env groups (EnvPool style)")
    ap.add_argument("--config5", action="store_true",
                    help="also time BASELINE config 5's per-GPU share (8192 envs per GPU); on by default at 8 GPUs")
    ap.add_argument("--long-run", type=int, default=384,
                    help="extra device-timed steps after the timed window that state the long-run mean (0 = off)")
    ap.add_argument("--balance-every", type=int, default=8,
                    help="re-deal envs to warps every this many steps by last step's cost (0 = off)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU side (the oracle port): cpu_baseline leg and --impl reference.  The only place bench.py runs
# anything from oracle/.
# ------------------------------------------------------------------------------------------------
def cpu_port_throughput(size, K, n_envs, steps, warmup, mode="legacy", use_hidden=True, seed=0):
    from oracle import alexandridis as ax
    from oracle import init_state as oinit
    from oracle import prng
    from oracle.c_oracle import COracle
    m = prng.LEGACY if mode == "legacy" else prng.PARTITIONABLE
    state, _ = oinit.initial_state(size, size, n_envs, seed=seed, use_hidden=use_hidden, mode=m, hidden="random")
    E = ax.EnvConstants(size, size, speed_move=0.12 * 4, speed_act=0.03 * 4)
    co = COracle(E, oinit.get_winds(), K=K, mode=m)
    rng = np.random.default_rng(seed)
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which is not what this arm measures)
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    threads = max(threads, COracle.max_threads())

    def act():
        return np.stack([rng.integers(0, 9, n_envs), rng.integers(0, 2, n_envs), rng.integers(0, 3, n_envs)],
                        1).astype(np.int32)

    for _ in range(warmup):
        co.step(state, act(), nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        co.step(state, act(), nthreads=threads)
    dt = time.perf_counter() - t0
    env_steps = n_envs * steps / dt
    return {"cell_updates_per_s": env_steps * size * size * K, "env_steps_per_s": env_steps,
            "seconds": dt, "threads": threads}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # every step of this arm is a bounded sample of the workload: cpu_envs envs instead of envs_per_gpu,
    # the requested --steps / --warmup are honoured as given
    n_envs = args.cpu_envs
    steps = max(1, args.steps)
    warmup = max(0, args.warmup)
    r = cpu_port_throughput(args.size, args.substeps, n_envs, steps, warmup, args.rng_mode, not args.no_hidden,
                            args.seed)
    sample = (f"each step = {n_envs} envs (of the workload's {args.envs_per_gpu}) x 1 env step ({args.substeps} CA sub-steps) "
              f"of the {args.size}x{args.size} workload; dense C port of the reference step (oracle/gca_oracle.c), "
              f"OpenMP over envs on {r['threads']} host threads")
    line = {
        "impl": "reference", "metric": "cell_updates_per_s", "value": r["cell_updates_per_s"],
        "unit": "cell-updates/s", "env_steps_per_s": r["env_steps_per_s"], "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * r["seconds"] / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
        "config": dict(workload_config(args), cpu_sample_envs_per_step=n_envs,
                       cpu_sample=f"this arm steps {n_envs} envs per step, not {args.envs_per_gpu}: a bounded sample of the "
                                  "workload (the dense CPU step costs the same for every env and phase)"),
        "cpu_baseline": {"value": r["cell_updates_per_s"], "unit": "cell-updates/s", "cores": r["threads"],
                         "kind": "port", "sample": sample},
        "e2e": {"value": r["cell_updates_per_s"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "the reference is JAX-on-CPU and cannot be installed in this image (no jax/jaxlib wheel, no "
                "network); this arm times the repo's dense C restatement of the same step on all host cores",
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"advanced-bulldozer Alexandridis CA {args.size}x{args.size}, {args.envs_per_gpu} envs/GPU, "
                        f"K={args.substeps} CA sub-steps per env step (speed-multiplier 4), hidden "
                        f"{'off' if args.no_hidden else 'on (synthetic random layers)'}, random actions, auto-reset on",
            "envs_per_gpu": args.envs_per_gpu, "grid": [args.size, args.size], "substeps": args.substeps,
            "phase": (f"stationary mixture: {args.preroll}-step pre-roll in setup, {args.preroll_groups} env groups force-reset "
                      f"{args.preroll // max(args.preroll_groups, 1)} steps apart (episode ages spread over one episode)"
                      if args.preroll > 0 else "all envs in phase from reset"),
            "rng_mode": args.rng_mode, "l2": (f"flushed between timed steps (256 MiB write{', then a 256 MiB read sweep: cold and clean' if args.flush_mode == 'write+read' else ''})") if not args.no_flush
            else "not flushed", "parallelism": f"env-sharded x{args.gpus}, no data-path collective"}


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU, sampled in a side thread while the timed loops run: NVML
    (nvidia_ml_py, one query every 2 ms) when available, else `nvidia-smi` (one query per ~50 ms)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index, self.uuid, self._halt = index, uuid, threading.Event()
        self.sm, self.mx, self.reasons, self.how = [], None, set(), "nvidia-smi"
        self._nvml = self._handle = None
        self.period = 0.002  # seconds between NVML queries
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                for cand in (f"GPU-{uuid}", str(uuid)):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand if isinstance(cand, bytes) else cand.encode())
                        break
                    except Exception:
                        h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml, self._handle, self.how = pynvml, h, "nvml"
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n, h = self._nvml, self._handle
        self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        for name, bit in self.REASONS:
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        if len(f) >= 7:
            self.sm.append(float(f[0]))
            self.mx = float(f[1])
            for (name, _), v in zip(self.REASONS, f[3:7]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._halt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                if self._nvml is not None:
                    self._nvml = None  # fall back to nvidia-smi
            self._halt.wait(self.period if self._nvml is not None else 0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": self.how}


def pin_rank_to_cores(local, world):
    """Give every rank of a node its own slice of the host cores (eight synchronous host loops otherwise share
    whatever cores the scheduler picks).  Returns (previous affinity, cores now used) or (None, None)."""
    try:
        prev = sorted(os.sched_getaffinity(0))
        per = len(prev) // world
        if world <= 1 or per < 1:
            return None, None
        mine = prev[local * per:(local + 1) * per]
        os.sched_setaffinity(0, mine)
        return prev, mine
    except (AttributeError, OSError):
        return None, None


FLUSH_MODE = "write+read"


def flush_l2(flush, i, torch):
    """Evict everything from the 126 MB L2 between timed steps: write a 256 MiB buffer (the env state is displaced, its
    dirty lines written back), then -- mode "write+read" -- sweep a second 256 MiB buffer with reads so that the L2 ends up
    holding CLEAN lines: the timed kernel then misses in L2 on every access (cold), but does not also pay for the
    write-back of the flush buffer's own 126 MB of dirty lines, which is traffic of the measurement, not of the path."""
    flush[0].fill_(i & 0xFF)
    if FLUSH_MODE == "write+read":
        flush[2][0] = flush[1].view(torch.int64).max()  # full read sweep (no L2-sized write)


def timed_steps(env, acts, first, n, flush, torch):
    """n env steps (acts[first + i]) with one CUDA-event pair each, the L2 flushed before every step (outside the
    events).  Returns the per-step times in microseconds after a synchronize."""
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    if flush is not None:
        # the synchronize() in front of the timed region leaves the GPU idle; the first kernel after an idle gap of a
        # few milliseconds runs ~50 us slower (measured: 140 us vs 85, tools/first_step_probe.py).  A millisecond of
        # un-timed flush sweeps (no env step) brings it back to its loaded state and loads torch's kernels.
        for i in range(16):
            flush_l2(flush, i, torch)
    torch.cuda.nvtx.range_push("timed_steps")  # (ncu --nvtx --nvtx-include "timed_steps/" profiles exactly these launches)
    for i in range(n):
        if flush is not None:
            flush_l2(flush, i, torch)  # evict the env state from the 126 MB L2 (outside the timed events)
        starts[i].record()
        env.step_device(acts[first + i])
        ends[i].record()
    torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
    return np.array([s.elapsed_time(e) * 1e3 for s, e in zip(starts, ends)])


def other_config_line(torch, dev, peak, flush, size, n_envs, use_hidden, K, steps, warmup, label):
    """One of the BASELINE configurations that are not the headline (config 3: 1024 envs of 256x256, hidden on / off;
    config 4: one 4096x4096 grid): device-timed env steps with the L2 flushed, from reset + `warmup` steps (young
    fires), device-generated hidden layers.  Same metric, same 7 B/cell/env-step roofline accounting."""
    from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
    from gym_cellular_automata_b200.workload import random_actions
    env = AdvancedForestFireBulldozerEnv(size, size, key=7, num_envs=n_envs, speed_move=0.12 * 4, speed_act=0.03 * 4,
                                         use_hidden=use_hidden, substeps=K, seed=3, hidden="device", obs_mode="none",
                                         auto_reset=True, collect_stats=True, device=dev)
    env.reset()
    gen = torch.Generator(device=dev)
    gen.manual_seed(11)
    acts = random_actions(warmup + steps, n_envs, dev, gen)
    for i in range(warmup):
        env.step_device(acts[i])
    torch.cuda.synchronize()
    st0, l0 = env.stats(), env.kernel_launches
    us = timed_steps(env, acts, warmup, steps, flush, torch)
    d = (env.stats() - st0).astype(np.float64)
    sub = max(d[5] * K, 1.0)
    sec = float(us.mean()) * 1e-6
    algo = ALGO_BYTES_PER_CELL * n_envs * size * size + ALGO_BYTES_PER_ENV * n_envs
    return {"workload": label, "grid": [size, size], "envs": n_envs, "substeps": K, "hidden": "on (device-generated layers)" if use_hidden else "off",
            "value": n_envs * size * size * K / sec, "unit": "cell-updates/s", "env_steps_per_s": n_envs / sec,
            "us_per_step": sec * 1e6, "steps": steps, "warmup": warmup, "launches_per_step": (env.kernel_launches - l0) / steps,
            "roofline": {"bound": "hbm", "achieved": algo / sec / 1e9, "peak": peak, "unit": "GB/s", "frac": algo / sec / 1e9 / peak,
                         "algorithmic_bytes_per_launch": algo},
            "front_cells_per_env_substep": d[0] / sub, "draws_per_env_substep": d[1] / sub, "threshold_cells": int(d[4]),
            "phase": f"steps {warmup}..{warmup + steps} after reset (young fires), L2 flushed between steps"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
    from gym_cellular_automata_b200.workload import random_actions, stationary_preroll

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    prev_affinity, my_cores = pin_rank_to_cores(local, world)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    K, size = args.substeps, args.size
    ROLLOUT = 128  # steps per rollout of the reference's trainer (agents/args.py:59): cadence of the statistics all-gather

    def make_env(n_envs, obs_mode="none", env_offset=0, total=None):
        e = AdvancedForestFireBulldozerEnv(
            size, size, key=1 + rank, num_envs=n_envs, speed_move=0.12 * 4, speed_act=0.03 * 4, use_hidden=not args.no_hidden,
            substeps=K, rng_mode=args.rng_mode, seed=args.seed + rank + 1000 * env_offset, hidden=args.hidden, obs_mode=obs_mode,
            auto_reset=True, collect_stats=True, device=dev, balance_every=args.balance_every, env_offset=env_offset,
            total_envs=total, generic_tiles=args.generic_tiles)
        e.reset()
        if args.preroll > 0:  # SETUP: stationary, phase-desynchronised mixture (not part of --warmup)
            stationary_preroll(e, args.preroll, args.preroll_groups, seed=args.seed + rank)
        return e

    def gather_episode_info(e):
        """The path's only collective (reference agents/jax_ppo.py:1330-1343): all-gather of the per-env episode
        counters the step kernel maintains (info["steps_elapsed"], info["reward_accumulated"])."""
        ep = torch.stack([e._state.steps_elapsed, e._state.reward_accumulated], dim=1).contiguous()
        if world == 1:
            return ep
        out = torch.empty((world * ep.shape[0], 2), dtype=ep.dtype, device=dev)
        dist.all_gather_into_tensor(out, ep)
        return out

    def measure(n_envs, with_long_run):
        """Device-timed loop, L2-warm loop and the end-to-end loop for one batch size; the same env steps in all three."""
        env = make_env(n_envs)
        gen = torch.Generator(device=dev)
        gen.manual_seed(args.seed + rank)
        n_long = args.long_run if with_long_run else 0
        total = args.warmup + args.steps + n_long
        acts = random_actions(total, n_envs, dev, gen)  # resident in HBM before the timed region
        for i in range(args.warmup):
            env.step_device(acts[i])
        torch.cuda.synchronize()
        stats0 = env.stats()
        launches0 = env.kernel_launches
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        us = timed_steps(env, acts, args.warmup, args.steps, flush, torch)
        if world > 1:
            dist.barrier()
        stats1 = env.stats()
        launches = env.kernel_launches - launches0  # env_step64_kernel per step + the re-balancing sort every few steps
        us_long = timed_steps(env, acts, args.warmup + args.steps, n_long, flush, torch) if n_long else None
        stats2 = env.stats()
        # L2-warm variant: back-to-back launches, one event pair (the steps after the ones above)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            env.step_device(acts[args.warmup + i])
        e1.record()
        torch.cuda.synchronize()
        ms_warm = e0.elapsed_time(e1)
        del env
        # ---- end to end through the host API: pinned host actions in, reward/terminated out, per step.
        # (1) synchronous: ONE C call per step for the whole batch, results valid on return.  The SAME steps as the
        # device-timed loop: a second env built from the same seeds, same pre-roll, same warm-up steps
        env_host = make_env(n_envs)
        for i in range(args.warmup):
            env_host.step_device(acts[i])
        h_act = acts[args.warmup:args.warmup + args.steps].cpu().pin_memory()
        h_rew, h_term = env_host.host_result_buffers()  # pinned
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if flush is not None:  # wake the GPU after the idle gap of the set-up (see timed_steps); no env step, not timed
            for i in range(16):
                flush_l2(flush, i, torch)
            torch.cuda.synchronize()
        n_coll = 0
        t0 = time.perf_counter()
        for i in range(args.steps):
            # the kernel reads this step's actions from the pinned host buffer and stores reward + terminated to the
            # pinned host buffers itself (zero-copy transport); the host polls the completion word of the launch
            env_host.step_host(h_act[i], h_rew, h_term)
            if (i + 1) % ROLLOUT == 0 or i + 1 == args.steps:  # end of a rollout: the statistics all-gather
                gathered = gather_episode_info(env_host)
                n_coll += 1
        torch.cuda.synchronize()
        e2e_sync_s = time.perf_counter() - t0
        del env_host
        # (2) two env groups (EnvPool style): the batch is two half-size envs on two CUDA streams; the host handles the
        # results of one group while the other group steps (step_host(wait=False) / step_host_wait()).  Every step of
        # every group still takes its actions from pinned host memory and delivers reward + terminated to the host.
        if not args.two_group:
            return {"us": us, "us_long": us_long, "ms_warm": ms_warm, "e2e_s": float("nan"), "e2e_sync_s": e2e_sync_s,
                    "launches": launches, "collectives2": 0,
                    "d_stats": (stats1 - stats0).astype(np.float64),
                    "d_stats_long": (stats2 - stats1).astype(np.float64) if n_long else None,
                    "gathered": gathered, "collectives": n_coll}
        half = n_envs // 2
        groups = [make_env(half, env_offset=0, total=n_envs), make_env(n_envs - half, env_offset=half, total=n_envs)]
        streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        g_act = [h_act[:, :half].contiguous().pin_memory(), h_act[:, half:].contiguous().pin_memory()]
        g_buf = [g.host_result_buffers() for g in groups]
        for g in range(2):
            for i in range(args.warmup):
                groups[g].step_device(acts[i][:half] if g == 0 else acts[i][half:])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        n_coll2 = 0
        sink = 0.0
        t0 = time.perf_counter()
        for g in range(2):
            with torch.cuda.stream(streams[g]):
                groups[g].step_host(g_act[g][0], g_buf[g][0], g_buf[g][1], wait=False)
        for i in range(args.steps):
            for g in range(2):
                groups[g].step_host_wait()          # this group's reward / terminated are in host memory now
                sink += float(g_buf[g][0][0])       # (the host reads them)
                if i + 1 < args.steps:
                    with torch.cuda.stream(streams[g]):
                        groups[g].step_host(g_act[g][i + 1], g_buf[g][0], g_buf[g][1], wait=False)
            if (i + 1) % ROLLOUT == 0 or i + 1 == args.steps:
                for g in range(2):
                    with torch.cuda.stream(streams[g]):
                        gather_episode_info(groups[g])
                n_coll2 += 1
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        del groups
        return {"us": us, "us_long": us_long, "ms_warm": ms_warm, "e2e_s": e2e_s, "e2e_sync_s": e2e_sync_s,
                "launches": launches, "collectives2": n_coll2,
                "d_stats": (stats1 - stats0).astype(np.float64),
                "d_stats_long": (stats2 - stats1).astype(np.float64) if n_long else None,
                "gathered": gathered, "collectives": n_coll}

    N = args.envs_per_gpu
    global FLUSH_MODE
    FLUSH_MODE = args.flush_mode
    flush = None if args.no_flush else (torch.empty(256 << 20, dtype=torch.uint8, device=dev),
                                        torch.zeros(256 << 20, dtype=torch.uint8, device=dev),
                                        torch.zeros(1, dtype=torch.int64, device=dev))
    sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(local), "uuid", None)) if rank == 0 else None
    if os.environ.get("GCA_BENCH_NO_SAMPLER"):  # diagnostics only: the JSON line then has no clocks
        sampler = None
    if sampler:
        sampler.start()
    m = measure(N, True)
    # BASELINE config 5 (65 536 envs of 64x64 over 8 GPUs = 8192 per GPU): timed beside the headline at 8 GPUs
    m5 = measure(8192, False) if (args.config5 or (world == 8 and N != 8192 and size == 64)) else None
    clocks = sampler.stop() if sampler else None  # sampled over the timed loops above (device-timed, L2-warm, end-to-end)

    # ---- reported separately: the same env steps with the RGB observation of each step
    with_obs = {}
    if not args.no_obs_leg:
        n_obs = min(args.steps, 64)
        gen = torch.Generator(device=dev)
        gen.manual_seed(args.seed + rank)
        acts = random_actions(args.warmup + n_obs, N, dev, gen)
        try:
            for mode in ("rgb_u8", "rgb_f32"):
                env_obs = make_env(N, mode)
                for i in range(args.warmup):
                    env_obs.step_device(acts[i])
                torch.cuda.synchronize()
                # back to back, one event pair (like value_l2_warm): per-step event pairs would count the host's enqueue
                # latency between the two launches of a step whenever the device runs dry
                o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                o0.record()
                for i in range(n_obs):
                    # ONE launch: the step kernel's epilogue draws the frame (GCA_FLAG_RENDER)
                    _, rgb = env_obs.step_observe_device(acts[args.warmup + i])
                o1.record()
                torch.cuda.synchronize()
                ms_obs = o0.elapsed_time(o1)
                with_obs[mode] = {"us_per_step": ms_obs / n_obs * 1e3, "steps": n_obs, "l2": "warm (back-to-back steps)",
                                  "launches_per_step": 1 if env_obs._can_fuse_render() else 2,
                                  "cell_updates_per_s_per_gpu": N * n_obs / (ms_obs * 1e-3) * size * size * K,
                                  "obs_bytes_per_step": int(rgb.numel() * rgb.element_size())}
                del env_obs, rgb
        except Exception as exc:  # a reported-separately figure must never take the bench line down
            with_obs = {"error": f"{type(exc).__name__}: {exc}"}

    def reduce_ranks(mm):
        """max over ranks of the three timed regions (+ the per-rank values, for attribution)"""
        mine = torch.tensor([float(mm["us"].sum()) * 1e-3, mm["ms_warm"], mm["e2e_s"], mm["e2e_sync_s"]], dtype=torch.float64,
                            device=dev)
        if world == 1:
            return [float(x) for x in mine.tolist()], None
        allr = torch.empty((world, 4), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        a = allr.cpu().numpy()
        per_rank = {"device_ms_per_step": [float(x) / args.steps for x in a[:, 0]],
                    "e2e_us_per_step": [float(x) / args.steps * 1e6 for x in a[:, 3]]}
        if args.two_group:
            per_rank["e2e_two_group_us_per_step"] = [float(x) / args.steps * 1e6 for x in a[:, 2]]
        return [float(x) for x in a.max(0)], per_rank

    (ms, ms_warm, e2e_s, e2e_sync_s), per_rank = reduce_ranks(m)
    red5 = reduce_ranks(m5) if m5 is not None else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    def throughput(n_envs, seconds):
        return n_envs * world * args.steps / seconds * size * size * K

    cells = N * size * size
    total_envs = N * world
    env_steps_s = total_envs * args.steps / (ms * 1e-3)
    cu_s = env_steps_s * size * size * K
    algo_bytes = ALGO_BYTES_PER_CELL * cells + ALGO_BYTES_PER_ENV * N  # per launch, per GPU
    launch_s = ms * 1e-3 / args.steps
    achieved = algo_bytes / launch_s / 1e9
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
            traffic, traffic_src = tj.get("env_step64_kernel_bytes_per_launch"), tj.get("source")
    except Exception:
        pass

    def wl(d):
        sub = max(d[5] * K, 1.0)
        return {"front_cells_per_env_substep": d[0] / sub, "draws_per_env_substep": d[1] / sub,
                "ignitions_per_env_substep": d[2] / sub, "burnouts_per_env_substep": d[3] / sub,
                "threshold_cells": int(d[4])}

    us = m["us"]
    step_us = {"mean": float(us.mean()), "min": float(us.min()), "median": float(np.median(us)), "max": float(us.max()),
               "series": [round(float(x), 2) for x in us[:64]]}
    workload_stats = wl(m["d_stats"])
    if m["us_long"] is not None and len(m["us_long"]):
        lr = float(m["us_long"].mean())
        workload_stats["long_run"] = dict(wl(m["d_stats_long"]), steps=int(len(m["us_long"])), us_per_step=lr,
                                          timed_window_over_long_run=float(us.mean()) / lr,
                                          what="the steps right after the timed window, same timing method: the "
                                               "stationary mean the timed window must agree with (+-10 %)")
    gathered = m["gathered"]
    line = {
        "metric": "cell_updates_per_s", "value": cu_s, "unit": "cell-updates/s", "env_steps_per_s": env_steps_s,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 bit-boards + f32 + u32 threefry",
        "data": "synthetic", "config": workload_config(args),
        "value_l2_warm": throughput(N, ms_warm * 1e-3),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                     "algorithmic_bytes_per_launch": algo_bytes, "launch_us": launch_s * 1e6},
        "e2e": {"value": throughput(N, e2e_sync_s), "unit": "cell-updates/s",
                "env_steps_per_s": total_envs * args.steps / e2e_sync_s,
                "h2d_bytes_per_step": int(N * 3 * 4), "d2h_bytes_per_step": int(N * 5),
                "collectives_in_timed_region": m["collectives"], "us_per_step": e2e_sync_s / args.steps * 1e6,
                "what": "one synchronous gca_env_step_host call per step for the whole batch, host buffers in and out: the fused step kernel reads the pinned host actions (H2D over the bus, zero-copy), the warp whose env ends last copies reward + terminated of all envs to pinned host memory in one burst (D2H) and stores the completion word the host polls -- results valid on return; the all-gather of the per-env episode counters runs inside this loop every 128 steps (one rollout) and at its end; same env steps as the device-timed loop (second env, same seeds, pre-roll and warm-up; L2 not flushed in this loop)",
                "two_group_async": None if not args.two_group else {"value": throughput(N, e2e_s), "us_per_step": e2e_s / args.steps * 1e6,
                                    "collectives_in_timed_region": m["collectives2"],
                                    "what": "the same traffic with the batch split into two env groups of N/2 on two CUDA streams (EnvPool style: GCA_FLAG_HOST_ASYNC + gca_host_wait, the host handles one group's results while the other steps); not the headline: a half-size launch lasts almost as long as a full one (the step is latency-bound per CTA), so splitting the batch does not pay on one GPU"}},
        "gpu_launches": m["launches"], "clocks": clocks, "step_us": step_us, "with_observation": with_obs,
        "workload_stats": workload_stats,
        "episode_stats": {"envs": int(gathered.shape[0]), "mean_steps_elapsed": float(gathered[:, 0].mean()),
                          "mean_reward_accumulated": float(gathered[:, 1].mean())},
    }
    if per_rank is not None:
        line["per_rank"] = dict(per_rank, host_cores_per_rank=len(my_cores) if my_cores else None)
    if m5 is not None:
        (ms5, ms5_warm, e2e5_s, e2e5_sync_s), per_rank5 = red5
        line["config5"] = {"workload": f"BASELINE config 5: {8192 * world} envs of {size}x{size}, 8192 per GPU, same pre-roll",
                           "value": throughput(8192, ms5 * 1e-3), "ms_per_step": ms5 / args.steps,
                           "env_steps_per_s": 8192 * world * args.steps / (ms5 * 1e-3),
                           "e2e": {"value": throughput(8192, e2e5_sync_s), "env_steps_per_s": 8192 * world * args.steps / e2e5_sync_s,
                                   "collectives_in_timed_region": m5["collectives"],
                                   "two_group_async_value": throughput(8192, e2e5_s) if args.two_group else None},
                           "roofline_frac": (ALGO_BYTES_PER_CELL * 8192 * size * size + ALGO_BYTES_PER_ENV * 8192)
                           / (ms5 * 1e-3 / args.steps) / 1e9 / peak,
                           "workload_stats": wl(m5["d_stats"]), "per_rank": per_rank5}
    if not args.no_other_configs and world == 1 and size == 64:
        # reported beside the headline, not part of it: BASELINE configs 3 and 4 (parity at full size: tests/test_gpu_parity.py)
        try:
            line["other_configs"] = [
                other_config_line(torch, dev, peak, flush, 256, 1024, True, 4, 24, 40,
                                  "BASELINE config 3: 1024 envs of 256x256 (R = 6), hidden layers on; whole-grid bit-board kernel, one launch per env step"),
                other_config_line(torch, dev, peak, flush, 256, 1024, False, 4, 24, 40,
                                  "BASELINE config 3: 1024 envs of 256x256 (R = 6), hidden layers off"),
                other_config_line(torch, dev, peak, flush, 4096, 1, True, 4, 16, 24,
                                  "BASELINE config 4: one 4096x4096 grid (R = 10); generic tiled kernels (active-tile lists, TMA-staged bit-row tiles), one CUDA graph launch per env step"),
                other_config_line(torch, dev, peak, flush, 64, 4096, True, 1, 24, 200,
                                  "config 2 at K = 1: ONE CA update per env step, the reference's own RepeatCAJax semantics (repeat_ca_jax.py:61-63)"),
                v3_line(256, 1024, 64, 16)]
        except Exception as exc:
            line["other_configs"] = {"error": f"{type(exc).__name__}: {exc}"}
    if not args.no_cpu_baseline and world == 1:
        if prev_affinity:
            os.sched_setaffinity(0, prev_affinity)
        r = cpu_port_throughput(size, K, args.cpu_envs, args.cpu_steps, 2, args.rng_mode, not args.no_hidden, args.seed)
        line["cpu_baseline"] = {
            "value": r["cell_updates_per_s"], "unit": "cell-updates/s", "cores": r["threads"], "kind": "port",
            "sample": f"{args.cpu_envs} envs x {args.cpu_steps} env steps of the same workload, dense C port of the "
                      f"reference step (oracle/gca_oracle.c), OpenMP over envs, {r['seconds']:.1f} s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def v3_line(size, N, steps, warmup, seed=0):
    """The v3 rule set (WindyForestFire + Move / Modify(cut) / RepeatCA of ForestFireBulldozer256x256-v3), batched: env
    steps per second with the reference's clock (most steps run 0 CA updates) -- a line beside the headline."""
    import torch
    from gym_cellular_automata_b200.forest_fire.bulldozer import ForestFireBulldozerEnv
    env = ForestFireBulldozerEnv(size, size, num_envs=N, seed=seed, max_repeats=2)
    env.reset()
    dev = env.device
    total = warmup + steps
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    acts = torch.stack([torch.randint(0, 9, (total, N), device=dev, generator=gen),
                        torch.randint(0, 2, (total, N), device=dev, generator=gen)], -1).to(torch.int32)
    rolls = torch.rand((N, 2, 9), dtype=torch.float64, device=dev, generator=gen)
    from gym_cellular_automata_b200._lib import check, current_stream, load, ptr
    st = env._state

    def step(i):
        check(load().gca_windy_env_step(N, size, size, ptr(st.tree), ptr(st.fire), ptr(st.position), ptr(st.time),
                                        ptr(acts[i]), ptr(env._wind_dev), ptr(rolls), 2, float(env._t_act_move),
                                        float(env._t_act_shoot), float(env._t_env_any), ptr(env._reward),
                                        ptr(env._term), ptr(env._counts), ptr(env._repeats), current_stream()))
    reps = 0
    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i)
        reps += env._repeats
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ca_updates = int(torch.as_tensor(reps).sum())
    return {"metric": "env_steps_per_s", "value": N * steps / (ms * 1e-3), "unit": "env-steps/s",
            "ruleset": "v3 WindyForestFire", "n_gpus": 1, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "us_per_step": ms / steps * 1e3, "ca_updates": ca_updates,
            "cell_updates_per_s": ca_updates * size * size / (ms * 1e-3), "data": "synthetic",
            "workload": f"BASELINE config 3, the registered v3 rule set: {N} envs of {size}x{size}, reference clock (most steps run 0 CA updates)",
            "config": {"workload": f"v3 rule set {size}x{size}, {N} envs, reference clock (most steps run 0 CA updates)"}}


def run_v3(args):
    """Extra (non-headline) line: the v3 rule set, 256x256 x 1024 envs by default (BASELINE config 3)."""
    size = args.size if args.size != 64 else 256
    N = args.envs_per_gpu if args.envs_per_gpu != 4096 else 1024
    print(json.dumps(v3_line(size, N, args.steps, args.warmup, args.seed)))


def main():
    args = parse()
    if args.size != 64 and args.preroll == PREROLL_HORIZON:
        args.preroll = 0  # the pre-roll horizon is one 64x64 episode; other grids start from reset unless told otherwise
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.ruleset == "v3":
        run_v3(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()

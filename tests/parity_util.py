"""Lock-step comparison of the CUDA env (through the C ABI) with the CPU oracle.

Both sides start from the SAME injected state (grid, ages, hidden layers, slope-factor table,
keys, wind, position) and receive the SAME actions; after every env step all state components
are compared bit for bit.  On a mismatch the report says where, and the CUDA state is re-synced
from the oracle so later steps stay comparable."""
from __future__ import annotations

import copy

import numpy as np
import torch

from oracle import alexandridis as ax
from oracle import init_state as oinit
from oracle import prng
from oracle.c_oracle import COracle

from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv

MODES = {"legacy": prng.LEGACY, "partitionable": prng.PARTITIONABLE}


def make_pair(N=8, size=64, K=1, mode="legacy", use_hidden=True, seed=0, hidden="reference", p_tree=0.0,
              speed_mult=4.0, jax_seed=1, collect_stats=True, obs_mode="none", enable_extensions=False,
              ncols=None, scatter_fire=0.0, use_tma=True, fast_slope=False, generic_tiles=False):
    nrows, ncols = size, (ncols or size)
    slope_fn = None
    if fast_slope:
        from gym_cellular_automata_b200.forest_fire.bulldozer.utils.init_utils import get_slope as slope_fn
    state, info = oinit.initial_state(nrows, ncols, N, seed=seed, jax_seed=jax_seed, use_hidden=use_hidden,
                                      mode=MODES[mode], hidden=hidden, slope_fn=slope_fn)
    if scatter_fire > 0:
        # richer start than the two seed cells: random burning cells with random remaining ages
        rs = np.random.default_rng(seed + 1000)
        ctx = state["per_env_context"]
        m = rs.random(ctx["true_grid"].shape) < scatter_fire
        ctx["true_grid"][m] = 2.0
        ctx["fire_age"][m] = rs.integers(1, 40, size=int(m.sum())).astype(np.float32)
    E = ax.EnvConstants(nrows, ncols, speed_move=0.12 * speed_mult, speed_act=0.03 * speed_mult, p_tree_ca=p_tree)
    winds = oinit.get_winds()
    state["shared_context"] = E.shared_context(winds)
    co = COracle(E, winds, K=K, mode=MODES[mode])
    env = AdvancedForestFireBulldozerEnv(nrows, ncols, key=jax_seed, num_envs=N, speed_move=0.12 * speed_mult,
                                         speed_act=0.03 * speed_mult, use_hidden=use_hidden, substeps=K,
                                         rng_mode=mode, seed=seed, hidden="random" if use_hidden else "reference",
                                         obs_mode=obs_mode, ca_p_tree=p_tree, collect_stats=collect_stats,
                                         enable_extensions=enable_extensions, use_tma=use_tma, generic_tiles=generic_tiles)
    sync(env, state, as_snapshot=True)
    return env, co, E, state, info


def sync(env, state, as_snapshot=False, info=None):
    env.set_state(state["per_env_context"], state["position"], state["time"], as_snapshot=as_snapshot, info=info)


def random_actions(rng, N, shoot_p=0.5):
    return np.stack([rng.integers(0, 9, N), (rng.random(N) < shoot_p).astype(np.int64), rng.integers(0, 3, N)],
                    axis=1).astype(np.int32)


def read_cuda_state(env):
    st = env._state
    ref = st.unpack_to_reference(env._params)
    out = {k: v.cpu().numpy() for k, v in ref.items()}
    out["wind_index"] = st.wind_index.cpu().numpy()
    out["key"] = st.key.cpu().numpy()
    out["is_night"] = st.is_night.cpu().numpy()
    out["time_step"] = st.time_step.cpu().numpy()
    out["position"] = st.position.cpu().numpy()
    out["time"] = st.time.cpu().numpy()
    return out


def compare(env, state, reward=None, term=None, counts=None, max_report=5):
    """Returns a list of mismatch descriptions (empty = bit-exact)."""
    got = read_cuda_state(env)
    ctx = state["per_env_context"]
    bad = []
    for k in ("true_grid", "fire_age", "dousing_count", "wind_index", "key", "is_night", "time_step"):
        a, b = np.asarray(ctx[k]), got[k]
        if not np.array_equal(a, b):
            idx = np.argwhere(a != b)
            ex = [f"{tuple(i)}: oracle {a[tuple(i)]} cuda {b[tuple(i)]}" for i in idx[:max_report]]
            bad.append(f"{k}: {len(idx)} mismatches, e.g. " + "; ".join(ex))
    for k in ("position", "time"):
        if not np.array_equal(np.asarray(state[k]), got[k]):
            bad.append(f"{k}: oracle {np.asarray(state[k]).ravel()[:8]} cuda {got[k].ravel()[:8]}")
    out = env._out
    if reward is not None and not np.array_equal(reward, out.step_reward.cpu().numpy()):
        bad.append(f"reward: oracle {reward[:4]} cuda {out.step_reward.cpu().numpy()[:4]}")
    if term is not None and not np.array_equal(term, out.terminated.cpu().numpy().astype(bool)):
        bad.append("terminated differs")
    if counts is not None and not np.array_equal(counts, out.counts.cpu().numpy()):
        bad.append(f"counts: oracle {counts[:3].tolist()} cuda {out.counts.cpu().numpy()[:3].tolist()}")
    return bad


def lockstep(env, co, state, steps, rng, inject_fn=None, shoot_p=0.5, resync=True, verbose=False):
    """Runs `steps` env steps on both sides.  Returns (n_bad_steps, reports, stats)."""
    N = env.num_envs
    reports = []
    for s in range(steps):
        act = random_actions(rng, N, shoot_p)
        inject = inject_fn(s) if inject_fn else None
        reward, term, counts = co.step(state, act, inject=inject)
        adev = torch.as_tensor(act, device=env.device)
        if inject is not None:
            env._actions.copy_(adev)
            env._launch_step(env._actions, inject)
        else:
            env.step_device(adev)
        bad = compare(env, state, reward, term, counts)
        if bad:
            reports.append((s, bad))
            if verbose:
                print(f"step {s}: " + " | ".join(bad))
            if resync:
                sync(env, state)
    return len(reports), reports, env.stats()

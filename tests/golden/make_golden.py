"""Generates tests/golden/env_step_golden.npz from the NumPy oracle (oracle/alexandridis.py).

These are ORACLE-generated vectors (K = 1, 2, 4 sub-steps, long runs): they pin the oracle (NumPy and C)
and the CUDA path against regressions.  The vectors recorded from the reference's own source are
tests/golden/reference_shim_golden.npz (make_reference_golden.py).
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import alexandridis as ax  # noqa: E402
from oracle import init_state as oinit  # noqa: E402
from oracle import prng  # noqa: E402

CASES = {
    "legacy_k1": dict(mode=prng.LEGACY, K=1, N=3, steps=60, seed=1, use_hidden=True),
    "part_k2": dict(mode=prng.PARTITIONABLE, K=2, N=2, steps=40, seed=2, use_hidden=True),
    "legacy_k4_nohidden": dict(mode=prng.LEGACY, K=4, N=2, steps=30, seed=3, use_hidden=False),
}


def actions_for(case, step):
    rng = np.random.default_rng(1000 * case["seed"] + step)
    N = case["N"]
    return np.stack([rng.integers(0, 9, N), rng.integers(0, 2, N), rng.integers(0, 3, N)], 1).astype(np.int32)


def run_case(case):
    state, info = oinit.initial_state(64, 64, case["N"], seed=case["seed"], jax_seed=1,
                                      use_hidden=case["use_hidden"], mode=case["mode"])
    E = ax.EnvConstants(64, 64, speed_move=0.48, speed_act=0.12)
    state["shared_context"] = E.shared_context(oinit.get_winds())
    rewards = []
    for s in range(case["steps"]):
        _, state, reward, term, _, info = ax.stateless_step(E, actions_for(case, s), state, info, K=case["K"],
                                                            mode=case["mode"], render_obs=False)
        rewards.append(reward)
    ctx = state["per_env_context"]
    return {"grid": ctx["true_grid"].astype(np.uint8), "fire_age": ctx["fire_age"].astype(np.uint16),
            "dousing": np.packbits(ctx["dousing_count"].astype(np.uint8), axis=-1), "key": ctx["key"],
            "wind_index": ctx["wind_index"], "position": state["position"], "time": state["time"],
            "rewards": np.stack(rewards), "reward_accumulated": info["reward_accumulated"]}


if __name__ == "__main__":
    out = {}
    for name, case in CASES.items():
        for k, v in run_case(case).items():
            out[f"{name}/{k}"] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "env_step_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

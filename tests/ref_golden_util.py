"""Replay helpers for tests/golden/reference_shim_golden.npz -- vectors produced by the REFERENCE'S OWN SOURCE run
under oracle/ref_shim (tests/golden/make_reference_golden.py).  Everything here is test infrastructure."""
from __future__ import annotations

import hashlib
import os

import numpy as np

from oracle import alexandridis as ax
from oracle import init_state as oinit

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_shim_golden.npz")
CASES = {"ref64_legacy_ext": dict(mode=0, ext=True, use_hidden=True),
         "ref32_nohidden_regrow": dict(mode=0, ext=False, use_hidden=False),
         "ref64_partitionable": dict(mode=1, ext=False, use_hidden=True)}
STATE_KEYS = ("grid", "fire_age", "dousing", "key", "wind_index", "time_step", "is_night", "position", "time")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_case(name):
    z = np.load(PATH)
    pre = name + "/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


def _ctx(fx, which, slope, W):
    g = fx[which + "/grid"]
    return {
        "wind_index": fx[which + "/wind_index"].astype(np.int32),
        "density": fx["density"].astype(np.int32),
        "vegetation": fx["vegetation"].astype(np.int32),
        "altitude": fx["altitude"].astype(np.float32),
        "slope": slope,
        "pslope": ax.p_slope_table(slope),
        "fire_age": fx[which + "/fire_age"].astype(np.float32),
        "key": fx[which + "/key"].astype(np.uint32),
        "is_night": fx[which + "/is_night"].astype(np.int32),
        "true_grid": g.astype(np.float32),
        "time_step": fx[which + "/time_step"].astype(np.int32),
        "dousing_count": np.unpackbits(fx[which + "/dousing"], axis=-1)[..., :W].astype(np.int32),
    }


def oracle_states(fx):
    """(E, start state, start info, conditional_reset snapshot) in the oracle's layout; the slope table is
    recomputed from the recorded float64 altitude with the ORACLE's get_slope and must hash to the reference's."""
    N, H, W = fx["start/grid"].shape
    slope = oinit.get_slope(fx["altitude64"].astype(np.float64)).astype(np.float32)
    p_fire, p_tree, p_wind, day = [float(x) for x in fx["shared_scalars"]]
    E = ax.EnvConstants(H, W, speed_move=0.48, speed_act=0.12, p_tree_ca=p_tree, p_wind_change=p_wind, p_fire=p_fire)
    assert E.day_length == int(day)
    shared = E.shared_context(fx["winds"])
    start = {"per_env_context": _ctx(fx, "start", slope, W), "shared_context": shared,
             "position": fx["start/position"].astype(np.int32), "time": fx["start/time"].astype(np.float32)}
    snap = {"per_env_context": _ctx(fx, "snapshot", slope, W), "shared_context": shared,
            "position": fx["snapshot/position"].astype(np.int32), "time": fx["snapshot/time"].astype(np.float32)}
    info = {"TimeLimit.truncated": np.zeros(N, dtype=bool), "terminated": np.zeros(N, dtype=bool),
            "steps_elapsed": np.zeros(N, dtype=np.float32), "reward_accumulated": np.zeros(N, dtype=np.float32),
            "reward": np.zeros(N, dtype=np.float32)}
    return E, start, info, snap, slope


def state_record(state):
    """The recorded components of a state, in the fixture's storage types."""
    ctx = state["per_env_context"]
    return {"grid": ctx["true_grid"].astype(np.uint8), "fire_age": ctx["fire_age"].astype(np.uint16),
            "dousing": np.packbits(ctx["dousing_count"].astype(np.uint8), axis=-1), "key": ctx["key"].astype(np.uint32),
            "wind_index": ctx["wind_index"].astype(np.int32), "time_step": ctx["time_step"].astype(np.int32),
            "is_night": ctx["is_night"].astype(np.int32), "position": state["position"].astype(np.int32),
            "time": state["time"].astype(np.float32)}


def compare_step(fx, s, got: dict, where: str):
    """got: name -> array for step s; returns a list of mismatch descriptions."""
    bad = []
    for k, v in got.items():
        want = fx["steps/" + k][s]
        v = np.asarray(v)
        if v.shape != want.shape or not np.array_equal(v, want.astype(v.dtype) if want.dtype != v.dtype else want):
            n = int(np.sum(v != want)) if v.shape == want.shape else -1
            bad.append(f"{where} step {s}: {k} differs ({n} elements)")
    return bad


def replay_oracle(fx, mode: int, ext: bool, where: str = "oracle"):
    """Runs the NumPy oracle through the recorded rollout (stateless_step then conditional_reset per step) and
    compares every recorded component, incl. the float32 RGB observations (by hash).  Returns mismatches."""
    E, state, info, snap, slope = oracle_states(fx)
    bad = []
    if sha(slope) != str(fx["slope_sha256"]):
        bad.append(f"{where}: get_slope differs from the reference's slope table")
    if not np.array_equal(oinit.get_winds().astype(np.float32), fx["winds"]):
        bad.append(f"{where}: get_winds differs from the reference's wind matrices")
    for s in range(fx["actions"].shape[0]):
        a = fx["actions"][s]
        rgb, st2, rew, term, _, info2 = ax.stateless_step(E, a, state, info, K=1, mode=mode, enable_extensions=ext,
                                                          render_obs=True)
        pre_sha = sha(rgb.astype(np.float32))
        rgb2, state, reward, term2, info = ax.conditional_reset(E, rgb, st2, rew, term, info2, a, snap,
                                                                enable_extensions=ext, render_obs=True)
        got = state_record(state)
        got.update(step_reward=rew.astype(np.float32), terminated=term.astype(np.uint8),
                   reward=reward.astype(np.float32), terminated_after_reset=term2.astype(np.uint8),
                   steps_elapsed=info["steps_elapsed"], reward_accumulated=info["reward_accumulated"])
        bad += compare_step(fx, s, got, where)
        if pre_sha != str(fx["steps/pre_rgb_sha256"][s]):
            bad.append(f"{where} step {s}: observation before conditional_reset differs")
        if sha(rgb2.astype(np.float32)) != str(fx["steps/rgb_sha256"][s]):
            bad.append(f"{where} step {s}: observation differs")
    if not np.array_equal(rgb2.astype(np.float32), fx["last_rgb"]):
        bad.append(f"{where}: last observation differs")
    return bad

"""TEST INFRASTRUCTURE (CPU oracle) -- literal NumPy restatement of the statistics half of
``step_env_wrapped`` in /root/reference/gym_cellular_automata/agents/jax_ppo.py:504-655, including
the serial per-env scan of ``update_recent_stats`` (:543-621).  Only tests / smoke / bench may import it.
Parity status: pinned to the reference's own source -- ``step_env_wrapped`` and ``EpisodeStatistics`` are cut out of
jax_ppo.py by ``ast`` and executed under oracle/ref_shim (tests/golden/make_reference_golden.py run_rollout_stats);
tests/test_oracle.py::test_rollout_stats_oracle_reproduces_reference_source_golden replays the recorded steps."""
import numpy as np

RECENT = 10
PER_ENV = {"episode_returns": np.float32, "episode_lengths": np.int32, "returned_episode_returns": np.float32,
           "returned_episode_lengths": np.int32, "current_day_correct": np.int32, "current_night_correct": np.int32,
           "current_day_steps": np.int32, "current_night_steps": np.int32}
RING = {"recent_returns": np.float32, "recent_lengths": np.int32, "recent_day_correct": np.int32,
        "recent_night_correct": np.int32, "recent_day_steps": np.int32, "recent_night_steps": np.int32}


def new_stats(num_envs):
    """jax_ppo.py:486-501"""
    st = {k: np.zeros(num_envs, dt) for k, dt in PER_ENV.items()}
    st.update({k: np.zeros(RECENT, dt) for k, dt in RING.items()})
    st["amount_finished"] = np.int32(0)
    st["recent_idx"] = np.int32(0)
    return st


def update(st, actions, info_reward, terminated, truncated, is_night):
    """jax_ppo.py:520-655 (statistics only).  Returns a new dict."""
    st = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in st.items()}
    terminated = np.asarray(terminated).astype(np.int32)
    truncated = np.asarray(truncated).astype(np.int32)
    is_night = np.asarray(is_night).astype(np.int32)
    new_ret = (st["episode_returns"] + np.asarray(info_reward, np.float32)).astype(np.float32)   # :520
    new_len = st["episode_lengths"] + 1                                                          # :521
    ext = np.asarray(actions)[:, -1]                                                             # :525
    st["current_day_correct"] += (1 - is_night) * (ext == 2)                                     # :528-541
    st["current_night_correct"] += is_night * (ext == 1)
    st["current_day_steps"] += 1 - is_night
    st["current_night_steps"] += is_night
    mask = (terminated + truncated) != 0                                                         # :623
    idx = int(st["recent_idx"])
    for e in range(len(mask)):                                                                   # :547-612 (lax.scan)
        if mask[e]:
            st["recent_returns"][idx] = new_ret[e]
            st["recent_lengths"][idx] = new_len[e]
            st["recent_day_correct"][idx] = st["current_day_correct"][e]
            st["recent_night_correct"][idx] = st["current_night_correct"][e]
            st["recent_day_steps"][idx] = st["current_day_steps"][e]
            st["recent_night_steps"][idx] = st["current_night_steps"][e]
            idx = (idx + 1) % RECENT
    st["recent_idx"] = np.int32((int(st["recent_idx"]) + int(mask.sum())) % RECENT)              # :614-615
    st["amount_finished"] = np.int32(int(st["amount_finished"]) + int(terminated.sum()))         # :627-629
    keep = ((1 - terminated) * (1 - truncated))
    st["episode_returns"] = (new_ret * keep.astype(np.float32)).astype(np.float32)               # :630-635
    st["episode_lengths"] = (new_len * keep).astype(np.int32)
    st["returned_episode_returns"] = np.where(mask, new_ret, st["returned_episode_returns"]).astype(np.float32)
    st["returned_episode_lengths"] = np.where(mask, new_len, st["returned_episode_lengths"]).astype(np.int32)
    return st

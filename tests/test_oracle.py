"""CPU tests of the ORACLE (test infrastructure): PRNG known-answer vectors, NumPy vs C restatement,
rule-level checks mirroring the reference's own operator tests, golden fixtures."""
import copy
import importlib.util
import os
import sys

import numpy as np
import pytest

from oracle import alexandridis as ax
from oracle import init_state as oinit
from oracle import prng

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- PRNG: Random123 KATs and public JAX values ------------------------------
def test_threefry_random123_known_answers():
    kat = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for k, x, y in kat:
        a, b = prng.threefry2x32(np.uint32(k[0]), np.uint32(k[1]), np.uint32(x[0]), np.uint32(x[1]))
        assert (int(a), int(b)) == y


def test_jax_documented_values_both_layouts():
    k0 = prng.key_from_seed(0)
    assert prng.split(k0, 2, prng.LEGACY).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert int(prng.random_bits(k0, 1, prng.LEGACY)[0]) == 1797259609
    assert prng.uniform(k0, 1, prng.LEGACY)[0] == np.float32(0.41845703)
    assert prng.split(k0, 2, prng.PARTITIONABLE).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]


@pytest.mark.parametrize("mode", [prng.LEGACY, prng.PARTITIONABLE])
def test_lazy_element_equals_dense_stream(mode):
    key = np.array([123, 456], np.uint32)
    for n in (1, 2, 3, 9, 36864):
        dense = prng.random_bits(key, n, mode)
        assert np.array_equal(dense, prng.random_bits_at(key, np.arange(n), n, mode))


def test_uniform_and_randint_construction():
    bits = np.array([0, 511, 512, 0xFFFFFFFF], np.uint32)
    u = prng.bits_to_uniform(bits)
    assert u.tolist() == [0.0, 0.0, 2.0 ** -23, 1.0 - 2.0 ** -23]
    assert prng.randint_params(144.0, 168.0) == (144, 24, 16)      # 64x64
    assert prng.randint_params(576.0, 672.0) == (576, 96, 64)      # 256x256
    assert prng.randint_params(9216.0, 10752.0) == (9216, 1536, 1024)
    assert prng.randint_params(1, 8) == (1, 7, 4)
    r = prng.randint(np.array([1, 2], np.uint32), 1000, 144.0, 168.0)
    assert r.min() >= 144 and r.max() < 168 and r.dtype == np.int32


# ---- constants (A0) ------------------------------------------------------------------------------------
def test_ca_constants_table():
    c = ax.CAConstants(64)
    assert (c.initial_spread_time, c.radius, c.fire_age_min, c.fire_age_max) == (96, 4, 144.0, 168.0)
    assert np.allclose(c.ring_weights, [4.3333e-3, 9.75e-4, 2.6e-4, 1.3e-4], rtol=1e-4)
    assert abs(float(c.burn_kernel.sum(dtype=np.float64)) - 0.065) < 1e-7
    assert np.isclose(c.dousing_weights[0, 0], 0.0588) and np.isclose(c.dousing_weights[2, 2], 0.504)
    assert [ax.CAConstants(s).radius for s in (32, 256, 4096)] == [3, 6, 10]
    E = ax.EnvConstants(64, 64, 0.48, 0.12)
    t = (E.movement_timings[0] + E.shooting_timings[1]) + E.t_any_f32
    assert abs(float(t) - 0.13120833) < 1e-7


# ---- rule-level checks (the reference's test style: deterministic corner cases) -------------------------
def _tiny_state(N=2, size=16, seed=0):
    state, info = oinit.initial_state(size, size, N, seed=seed, use_hidden=True, hidden="random")
    E = ax.EnvConstants(size, size)
    state["shared_context"] = E.shared_context(oinit.get_winds())
    return E, state, info


def test_rule_injected_uniforms_extremes():
    E, state, info = _tiny_state()
    ctx = state["per_env_context"]
    N, H, W = ctx["true_grid"].shape
    zeros = {"u_burn": np.zeros((N, H, W, 3, 3), np.float32), "u_grow": np.ones((N, H, W), np.float32),
             "age_new": np.full((N, H, W), 7, np.int32), "u_wind": np.ones(N, np.float32),
             "wind_step": np.ones(N, np.int32)}
    g, c, dbg = ax.ca_update(E.ca, ctx["true_grid"], ctx, state["shared_context"], inject=zeros, want_debug=True)
    old = ctx["true_grid"]
    pad = np.pad(old, ((0, 0), (1, 1), (1, 1)))
    nb = np.zeros_like(old, dtype=bool)
    pos_p = np.zeros_like(old, dtype=bool)
    for i in range(3):
        for j in range(3):
            f = pad[:, i:i + H, j:j + W] == 2
            nb |= f
            pos_p |= f & (dbg["p"][..., i, j] > 0)
    # u = 0: every tree with a burning neighbour whose probability is positive ignites, nothing else
    assert np.array_equal(g == 2, ((old == 1) & pos_p) | ((old == 2) & (ctx["fire_age"] > 1)))
    assert np.all(c["fire_age"][(g == 2) & (old == 1)] == 7)
    assert np.array_equal(c["wind_index"], ctx["wind_index"])  # u_wind = 1 -> no wind change
    ones = dict(zeros, u_burn=np.ones((N, H, W, 3, 3), np.float32))
    g1, c1 = ax.ca_update(E.ca, old, ctx, state["shared_context"], inject=ones)
    assert not ((g1 == 2) & (old == 1)).any()                     # u = 1: no ignition at all
    assert np.array_equal(c1["fire_age"][old == 2], ctx["fire_age"][old == 2] - 1)


def test_fire_burns_out_when_age_reaches_one():
    E, state, info = _tiny_state()
    ctx = state["per_env_context"]
    ctx["fire_age"][ctx["true_grid"] == 2] = 1.0
    inj = {"u_burn": np.ones(ctx["true_grid"].shape + (3, 3), np.float32)}
    g, c = ax.ca_update(E.ca, ctx["true_grid"], ctx, state["shared_context"], inject=inj)
    assert not (g == 2).any() and np.all(c["fire_age"][ctx["true_grid"] == 2] == 0)
    assert ax.is_done(g).all() and np.all(ax.award(g) == 0)


def _independent_new_position(action, position, nrows, ncols):
    # independent boundary oracle, same idea as the reference's test_move_modify.py:132-170
    r, c = position
    dr = {0: -1, 1: -1, 2: -1, 3: 0, 4: 0, 5: 0, 6: 1, 7: 1, 8: 1}[action]
    dc = {0: -1, 1: 0, 2: 1, 3: -1, 4: 0, 5: 1, 6: -1, 7: 0, 8: 1}[action]
    nr, nc = r + dr, c + dc
    if nr < 0 or nr >= nrows:
        nr = r
    if nc < 0 or nc >= ncols:
        nc = c
    return nr, nc


def test_move_against_independent_oracle():
    rng = np.random.default_rng(0)
    for _ in range(300):
        H, W = rng.integers(2, 9, 2)
        pos = np.array([[rng.integers(0, H), rng.integers(0, W)]], dtype=np.int32)
        a = int(rng.integers(0, 9))
        got = ax.move(pos, np.array([a]), H, W)[0]
        assert tuple(got) == _independent_new_position(a, tuple(pos[0]), H, W)


def test_modify_clock_reward_daynight():
    dc = np.zeros((2, 8, 8), np.int32)
    out = ax.modify(dc, np.array([1, 0]), np.array([[3, 4], [1, 1]]))
    assert out[0, 3, 4] == 1 and out.sum() == 1 and dc.sum() == 0
    grid = np.zeros((1, 4, 4), np.float32)
    grid[0, 0, :] = 1
    grid[0, 1, 0] = 2
    assert ax.award(grid)[0] == np.float32(-(np.float32(1) / (np.float32(5) + np.float32(1e-8))))
    assert not ax.is_done(grid)[0] and ax.is_done(np.zeros((1, 4, 4), np.float32))[0]
    E, state, info = _tiny_state(N=2)
    state["per_env_context"]["time_step"][:] = [399, 7]
    act = np.array([[4, 0, 0], [4, 1, 0]])
    _, ns, _, _, _, ninfo = ax.stateless_step(E, act, state, info, render_obs=False)
    assert ns["per_env_context"]["is_night"].tolist() == [1, 0]
    assert ns["per_env_context"]["time_step"].tolist() == [400, 8]
    assert ns["per_env_context"]["dousing_count"][1].sum() == 1 and ns["per_env_context"]["dousing_count"][0].sum() == 0
    assert ninfo["steps_elapsed"].tolist() == [1.0, 1.0]
    assert np.all((ns["time"] >= 0) & (ns["time"] < 1))


def test_k_substeps_equals_k_single_updates():
    # repeats semantics the JAX operator dropped but the NumPy RepeatCA has (reference test_repeat_ca.py:68-90)
    E, state, info = _tiny_state(N=2, size=16, seed=3)
    ctx = state["per_env_context"]
    g, c = ctx["true_grid"], ctx
    for _ in range(3):
        g, c = ax.ca_update(E.ca, g, c, state["shared_context"])
    act = np.array([[4, 0, 0], [4, 0, 0]])
    _, ns, *_ = ax.stateless_step(E, act, state, info, K=3, render_obs=False)
    assert np.array_equal(ns["per_env_context"]["true_grid"], g)
    assert np.array_equal(ns["per_env_context"]["key"], c["key"])


def test_observation_quirks():
    g = np.ones((1, 6, 6), np.float32)
    g[0, 0, :] = 0
    g[0, 3, 3] = 2
    pos = np.array([[5, 5]], np.int32)
    dc = np.zeros((1, 6, 6), np.int32)
    dc[0, 2, 2] = 1
    base = ax.build_observation(g, pos, np.array([[4, 0, 0, 0]]), dc, np.array([0]), False)
    assert base.dtype == np.float32 and base.shape == (1, 6, 6, 3)
    assert base[0, 5, 5].tolist() == [0, 0, 0]                       # bulldozer pixel is black
    assert base[0, 1, 1].tolist() == [0xA9, 0xC4, 0x99] and base[0, 3, 3].tolist() == [0xE6, 0x81, 0x81]
    assert base[0, 2, 2].tolist() == [0xA9 * 0.25, 0xC4 * 0.25, 0x99 * 0.25 + 150.0]  # doused tree, day tint blue
    night = ax.build_observation(g, pos, np.array([[4, 0, 0, 0]]), dc, np.array([1]), False)
    assert night[0, 1, 1].tolist() == [0x2F, 0x4F, 0x4F]
    # extension bit 0 (raw grid): the first row with a positive entry is row 1 -> channel index 1 (zeros)
    ext = ax.build_observation(g, pos, np.array([[4, 0, 1, 0]]), dc, np.array([0]), True)
    assert ext[0, 1, 1].tolist() == [0xDD, 0xD1, 0xD3]
    g2 = g.copy()
    g2[0, 0, 0] = 1                                                # now row 0 is positive -> channel 0 = raw grid
    ext2 = ax.build_observation(g2, pos, np.array([[4, 0, 1, 0]]), dc, np.array([0]), True)
    assert ext2[0, 1, 1].tolist() == [0xA9, 0xC4, 0x99]
    assert ax.apply_blur(np.full((1, 4, 4), 2.0, np.float32)).tolist() == np.full((1, 4, 4), 2).tolist()


# ---- C restatement == NumPy restatement -------------------------------------------------------------------
def _c_oracle():
    from oracle.c_oracle import COracle
    return COracle


@pytest.mark.parametrize("mode,K,size,p_tree", [(prng.LEGACY, 1, 64, 0.0), (prng.PARTITIONABLE, 2, 64, 0.0),
                                               (prng.LEGACY, 2, 32, 0.01), (prng.LEGACY, 1, 40, 0.0)])
def test_c_oracle_equals_numpy_oracle(mode, K, size, p_tree):
    COracle = _c_oracle()
    N = 2
    E = ax.EnvConstants(size, size, 0.48, 0.12, p_tree_ca=p_tree)
    winds = oinit.get_winds()
    state, info = oinit.initial_state(size, size, N, seed=1, mode=mode, hidden="random")
    state["shared_context"] = E.shared_context(winds)
    cst = copy.deepcopy({k: v for k, v in state.items() if k != "shared_context"})
    co = COracle(E, winds, K=K, mode=mode)
    rng = np.random.default_rng(0)
    for step in range(12):
        act = np.stack([rng.integers(0, 9, N), rng.integers(0, 2, N), rng.integers(0, 3, N)], 1)
        _, state, reward, term, _, info = ax.stateless_step(E, act, state, info, K=K, mode=mode, render_obs=False)
        r2, t2, cnt = co.step(cst, act)
        for k in ("true_grid", "fire_age", "dousing_count", "wind_index", "key", "is_night", "time_step"):
            assert np.array_equal(state["per_env_context"][k], cst["per_env_context"][k]), (step, k)
        assert np.array_equal(state["position"], cst["position"]) and np.array_equal(state["time"], cst["time"])
        assert np.array_equal(reward, r2) and np.array_equal(term, t2)


def test_c_oracle_injected_fields():
    COracle = _c_oracle()
    N, K, size = 2, 2, 32
    E = ax.EnvConstants(size, size, 0.48, 0.12, p_tree_ca=0.05)
    winds = oinit.get_winds()
    state, info = oinit.initial_state(size, size, N, seed=2, hidden="random")
    state["shared_context"] = E.shared_context(winds)
    cst = copy.deepcopy({k: v for k, v in state.items() if k != "shared_context"})
    co = COracle(E, winds, K=K)
    rng = np.random.default_rng(1)
    for step in range(6):
        inj = {"u_burn": (rng.random((K, N, size, size, 9)) * 0.3).astype(np.float32),
               "u_grow": rng.random((K, N, size, size)).astype(np.float32),
               "age_new": rng.integers(3, 9, (K, N, size, size)).astype(np.int32),
               "u_wind": rng.random((K, N)).astype(np.float32), "wind_step": rng.integers(1, 8, (K, N)).astype(np.int32)}
        act = np.stack([rng.integers(0, 9, N), rng.integers(0, 2, N), rng.integers(0, 3, N)], 1)
        inj_list = [{k: (v[j].reshape(N, size, size, 3, 3) if k == "u_burn" else v[j]) for k, v in inj.items()}
                    for j in range(K)]
        _, state, reward, term, _, info = ax.stateless_step(E, act, state, info, K=K, inject=inj_list, render_obs=False)
        r2, t2, cnt = co.step(cst, act, inject=inj)
        assert np.array_equal(state["per_env_context"]["true_grid"], cst["per_env_context"]["true_grid"]), step
        assert np.array_equal(state["per_env_context"]["fire_age"], cst["per_env_context"]["fire_age"]), step
        assert np.array_equal(state["per_env_context"]["wind_index"], cst["per_env_context"]["wind_index"]), step


# ---- golden fixtures -------------------------------------------------------------------------------------
def _golden_module():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_oracle_reproduces_golden_fixtures():
    mg = _golden_module()
    gold = np.load(os.path.join(HERE, "golden", "env_step_golden.npz"))
    for name, case in mg.CASES.items():
        out = mg.run_case(case)
        for k, v in out.items():
            assert np.array_equal(v, gold[f"{name}/{k}"]), (name, k)


def test_c_oracle_reproduces_golden_fixtures():
    COracle = _c_oracle()
    mg = _golden_module()
    gold = np.load(os.path.join(HERE, "golden", "env_step_golden.npz"))
    for name, case in mg.CASES.items():
        state, info = oinit.initial_state(64, 64, case["N"], seed=case["seed"], jax_seed=1,
                                          use_hidden=case["use_hidden"], mode=case["mode"])
        E = ax.EnvConstants(64, 64, speed_move=0.48, speed_act=0.12)
        co = COracle(E, oinit.get_winds(), K=case["K"], mode=case["mode"])
        rewards = [co.step(state, mg.actions_for(case, s))[0] for s in range(case["steps"])]
        ctx = state["per_env_context"]
        assert np.array_equal(ctx["true_grid"].astype(np.uint8), gold[f"{name}/grid"]), name
        assert np.array_equal(ctx["fire_age"].astype(np.uint16), gold[f"{name}/fire_age"]), name
        assert np.array_equal(ctx["key"], gold[f"{name}/key"]) and np.array_equal(np.stack(rewards), gold[f"{name}/rewards"])


# ---- v3 rule set (WindyForestFire), reference operators/tests/test_ca_windy.py:55-102 ----------------------
def test_windy_oracle_rules_with_certain_wind():
    from oracle import windy
    rng = np.random.default_rng(0)
    for _ in range(40):
        g = rng.choice([0, 3, 25], size=(5, 6), p=[0.3, 0.5, 0.2]).astype(np.int64)
        ng = windy.windy_update(g, np.ones((3, 3)), rng.random((3, 3)))
        pad = np.pad(g, 1)
        for r in range(5):
            for c in range(6):
                nb = pad[r:r + 3, c:c + 3]
                if g[r, c] == 3:
                    assert ng[r, c] == (25 if (nb == 25).any() else 3)
                else:
                    assert ng[r, c] == 0
    # a failed direction does not propagate: fire BELOW a tree spreads up with the kernel's "up" entry
    g = np.zeros((3, 3), np.int64)
    g[1, 1], g[2, 1] = 3, 25
    wind = windy.DEFAULT_WIND
    roll = np.zeros((3, 3))
    roll[0, 1] = 0.99  # wind[0,1] = 0.64 <= 0.99 -> "up" fails
    assert windy.windy_update(g, wind, roll)[1, 1] == 3
    roll[0, 1] = 0.10
    assert windy.windy_update(g, wind, roll)[1, 1] == 25


def test_v3_env_step_oracle_clock_and_cut():
    from oracle import windy
    C = windy.V3Constants(8, 8, t_move=0.6, t_shoot=0.7)
    g = np.full((8, 8), 3, np.int64)
    g[4, 4] = 25
    rolls = np.zeros((3, 3, 3))
    # move + shoot: 0.6 + 0.7 + 0.001 -> one CA update, fraction kept; the tree under the bulldozer is cut
    ng, pos, t, rew, done, rep = windy.v3_env_step(C, g, np.array([0, 0]), 0.0, (8, 1), windy.DEFAULT_WIND, rolls)
    assert rep == 1 and abs(t - 0.301) < 1e-12 and tuple(pos) == (1, 1) and ng[1, 1] == 0
    assert ng[4, 4] == 0 and (ng == 25).sum() == 8 and not done
    # not moving and not shooting costs only t_any (bulldozer.py:285-286)
    ng, pos, t, rew, done, rep = windy.v3_env_step(C, g, np.array([0, 0]), 0.0, (4, 0), windy.DEFAULT_WIND, rolls)
    assert rep == 0 and t == 0.001 and np.array_equal(ng, g)


# ---- rollout statistics (agents/jax_ppo.py:504-655) ---------------------------------------------------
def test_rollout_stats_ring_buffer_semantics():
    """Hand-worked case of update_recent_stats: finished envs enter the ring in env order from recent_idx,
    wrap modulo 10 (later ones overwrite), counters are latched and the running totals are cleared."""
    from oracle import rollout
    N = 14
    st = rollout.new_stats(N)
    st["recent_idx"] = np.int32(7)
    st["episode_returns"][:] = np.arange(N, dtype=np.float32)
    st["episode_lengths"][:] = 5
    actions = np.zeros((N, 3), np.int32)
    actions[:, 2] = 2
    term = np.ones(N, np.uint8)
    term[3] = 0
    trunc = np.zeros(N, np.uint8)
    trunc[3] = 1  # truncated counts as finished, not as terminated
    out = rollout.update(st, actions, np.full(N, -0.5, np.float32), term, trunc, np.zeros(N, np.int32))
    assert int(out["recent_idx"]) == (7 + N) % 10 and int(out["amount_finished"]) == N - 1
    # env e has rank e; slot (7 + e) % 10 keeps the LAST writer: ranks 4..13 survive
    for e in range(4, N):
        assert out["recent_returns"][(7 + e) % 10] == np.float32(e - 0.5)
        assert out["recent_lengths"][(7 + e) % 10] == 6
        assert out["recent_day_correct"][(7 + e) % 10] == 1 and out["recent_day_steps"][(7 + e) % 10] == 1
    assert not out["episode_returns"].any() and not out["episode_lengths"].any()
    assert np.array_equal(out["returned_episode_returns"], np.arange(N, dtype=np.float32) - np.float32(0.5))
    # nothing finished: accumulators run on, the ring stays
    out2 = rollout.update(out, actions, np.ones(N, np.float32), np.zeros(N, np.uint8), trunc * 0, np.ones(N, np.int32))
    assert np.array_equal(out2["recent_returns"], out["recent_returns"]) and int(out2["recent_idx"]) == int(out["recent_idx"])
    assert np.all(out2["episode_returns"] == 1) and np.all(out2["episode_lengths"] == 1)
    assert np.all(out2["current_night_steps"] == 1) and np.all(out2["current_night_correct"] == 0)


# ---- device-side hidden-layer generator: the NumPy mirror keeps the reference's layer models ------------------
def test_hidden_device_mirror_layer_models():
    """oracle/hidden_device.py (what gca_generate_hidden is checked against): value ranges and structure of
    init_utils.py:10-116, and its slope equals the package's get_slope on the same altitude."""
    from oracle import hidden_device as hd
    from gym_cellular_automata_b200.forest_fire.bulldozer.utils.init_utils import get_slope
    H, W, seed = 64, 48, 0x1234567890AB
    for env in (0, 5, 70000):
        veg = hd.patches(H, W, seed, env, hd.VEG_RECT, hd.VEG_FILL)
        den = hd.patches(H, W, seed, env, hd.DEN_RECT, hd.DEN_FILL)
        assert veg.min() >= 1 and veg.max() <= 5 and den.min() >= 1 and den.max() <= 5
        assert not np.array_equal(veg, den)
        # types 4 and 5 only come from rectangles: they form at most 7 axis-aligned blocks, so few rows change
        assert (veg >= 4).sum() == 0 or np.unique(np.nonzero(veg >= 4)[0]).size <= H
        alt = hd.altitude(H, W, seed, env)
        assert alt.min() >= 0.0 and alt.max() <= (5 + 9 * 6 + 7 * 4) / 10
        s = hd.slope(alt)
        assert np.array_equal(s, get_slope(alt[None])[0].astype(np.float32))
        assert not s[0].any() and not s[-1].any() and not s[:, 0].any() and not s[:, -1].any() and not s[:, :, 1, 1].any()
    # an env's layers depend on (seed, env) only
    assert np.array_equal(hd.patches(H, W, seed, 5, 1, 2), hd.patches(H, W, seed, 5, 1, 2))
    assert not np.array_equal(hd.patches(H, W, seed, 5, 1, 2), hd.patches(H, W, seed + 1, 5, 1, 2))


# ----------------------------------------------------------------------------------------------------------
# the reference's OWN source, executed under oracle/ref_shim (NumPy stand-ins for jax / flax / gymnasium)
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["ref64_legacy_ext", "ref32_nohidden_regrow", "ref64_partitionable"])
def test_oracle_reproduces_reference_source_golden(name):
    """tests/golden/reference_shim_golden.npz was recorded from the reference's advanced_bulldozer.py / operator
    files themselves (tests/golden/make_reference_golden.py): rollouts of stateless_step + conditional_reset with
    burn-outs, dousing, wind changes, a day/night flip, regrowth, a termination and its reset, extensions on / off,
    both jax.random stream layouts.  The oracle must reproduce every state component and every observation pixel
    of every step; get_slope and get_winds must reproduce the reference's tables."""
    import ref_golden_util as R
    fx = R.load_case(name)
    c = R.CASES[name]
    bad = R.replay_oracle(fx, c["mode"], c["ext"], name)
    assert not bad, "\n".join(bad[:20])
    if name != "ref64_partitionable":
        assert fx["steps/terminated"].sum() >= 1, "the fixture should hold a termination"
    assert (fx["steps/is_night"] == 1).any() and fx["steps/dousing"].any()


def test_v3_oracle_reproduces_reference_source_golden():
    """v3 rule set: the reference's own WindyForestFire / Move / Modify / RepeatCA / MDP operators (NumPy + scipy,
    run through the gymnasium stand-in) stepped for 40 steps on 4 grids of 32x48 with recorded wind rolls; the
    oracle's v3_env_step must give the same grid, position, clock, reward, done flag and CA-update count."""
    import ref_golden_util as R
    from oracle import windy
    fx = R.load_case("v3_32x48")
    N, H, W = fx["grid0"].shape
    tm, ts, ta = fx["t_move_shoot_any"]
    C = windy.V3Constants(H, W, t_any=float(ta), t_move=float(tm), t_shoot=float(ts))
    assert np.array_equal(fx["wind"], windy.DEFAULT_WIND)
    grid, pos, time = fx["grid0"].astype(np.int64), fx["position0"].copy(), fx["time0"].copy()
    for s in range(fx["actions"].shape[0]):
        for e in range(N):
            if fx["frozen"][s, e]:
                continue
            g, p, t, r, d, rep = windy.v3_env_step(C, grid[e], pos[e], time[e], fx["actions"][s, e], fx["wind"],
                                                   fx["rolls"][s, e])
            grid[e], pos[e], time[e] = g, p, t
            assert np.array_equal(g, fx["steps/grid"][s, e]), (s, e)
            assert np.array_equal(p, fx["steps/position"][s, e]) and t == fx["steps/time"][s, e], (s, e)
            want = fx["steps/reward"][s, e]
            assert r == want or (np.isnan(r) and np.isnan(want)), (s, e)
            assert d == bool(fx["steps/terminated"][s, e]) and rep == int(fx["steps/repeats"][s, e]), (s, e)
    assert {0, 1, 2} <= set(fx["steps/repeats"].ravel().tolist())


def test_move_douse_clock_oracle_reproduces_reference_source_golden():
    """Operator level, exhaustive: the reference env's own MoveModifyJax for every move x shoot action at every cell of
    the two outer rings of a 16x16 grid (2052 cases), and its RepeatCAJax clock for every action pair at accumulated
    times on both sides of the wrap (0.8687916 + 0.13120833 is the float32 boundary at speed-multiplier 4)."""
    import ref_golden_util as R
    fx = R.load_case("operator_edges")
    S = int(fx["size"])
    a = fx["actions"]
    new = ax.move(fx["pos_in"], a[:, 0], S, S)
    assert np.array_equal(new, fx["pos_out"])
    n = len(a)
    dc = ax.modify(np.zeros((n, S, S), np.int32), a[:, 1], new)
    want = np.zeros((n, S, S), np.int32)
    hit = fx["doused"][:, 0] >= 0
    want[np.nonzero(hit)[0], fx["doused"][hit, 0], fx["doused"][hit, 1]] = 1
    assert np.array_equal(dc, want) and np.array_equal(hit, a[:, 1] == 1)
    E = ax.EnvConstants(S, S, speed_move=0.48, speed_act=0.12)
    ca = fx["clock_actions"]
    t = (fx["clock_time_in"] + ((E.movement_timings[ca[:, 0]] + E.shooting_timings[ca[:, 1]]) + E.t_any_f32)).astype(np.float32)
    assert np.array_equal(np.modf(t)[0].astype(np.float32), fx["clock_frac"])
    assert (np.modf(t)[1] == 0).any() and (np.modf(t)[1] == 1).any(), "both sides of the wrap"


def test_ca_constants_reproduce_reference_source_golden():
    """A0 for every BASELINE grid size (32 ... 4096, R = 3 ... 10): the oracle's CAConstants against what the reference's
    own constructor derived (heat kernel, dousing weights, fire-age range), bit for bit; gca_params_init is checked
    against CAConstants in tests/test_host_api.py."""
    import ref_golden_util as R
    fx = R.load_case("constants")
    for s in fx["sizes"].tolist():
        c = ax.CAConstants(s)
        assert np.array_equal(c.burn_kernel, fx[f"{s}/burn_kernel"]), s
        assert np.array_equal(c.dousing_weights, fx[f"{s}/dousing_weights"]), s
        assert [c.initial_spread_time, c.fire_age_min, c.fire_age_max, c.radius] == fx[f"{s}/scalars"].tolist(), s
    assert ax.CAConstants(4096).burn_kernel.shape == (21, 21)


def test_rollout_stats_oracle_reproduces_reference_source_golden():
    """Episode statistics: vectors recorded from the reference's own ``step_env_wrapped`` / ``EpisodeStatistics``
    source (cut out of agents/jax_ppo.py by ast and run under the shim, make_reference_golden.run_rollout_stats);
    every field, value and dtype, after every step."""
    import ref_golden_util as R
    from oracle import rollout
    fx = R.load_case("rollout_stats")
    steps, N = fx["actions"].shape[:2]
    st = rollout.new_stats(N)
    for s in range(steps):
        st = rollout.update(st, fx["actions"][s], fx["reward"][s], fx["terminated"][s], fx["truncated"][s],
                            fx["is_night"][s])
        for k, v in st.items():
            want = fx["out/" + k][s]
            assert np.array_equal(np.asarray(v), want) and np.asarray(v).dtype == want.dtype, (s, k)
    assert int(fx["out/amount_finished"][-1]) > 100 and fx["truncated"].any()


@pytest.mark.skipif(not os.environ.get("GCA_SHIM_FUZZ"), reason="opt-in: GCA_SHIM_FUZZ=<number of random cases>")
def test_reference_source_fuzz():
    """Opt-in sweep (minutes of CPU, needs /root/reference): random configurations -- grid sizes incl. non-powers of
    two, 1-3 envs, both stream layouts, hidden layers / extensions / regrowth on and off -- each rolled out with the
    reference's own source under the shim and replayed by the oracle."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("no reference tree here")
    import contextlib
    import io
    from oracle.ref_shim import jax_shim
    spec = importlib.util.spec_from_file_location("make_reference_golden",
                                                  os.path.join(HERE, "golden", "make_reference_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    import ref_golden_util as R
    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k.split(".")[0] in ("jax", "flax", "gymnasium", "gym_cellular_automata")}
    rng = np.random.default_rng(int(os.environ.get("GCA_SHIM_FUZZ_SEED", "0")))
    try:
        jax = ref_shim.install(prng.LEGACY)
        ab = ref_shim.load("forest_fire.bulldozer.advanced_bulldozer")
        for i in range(int(os.environ["GCA_SHIM_FUZZ"])):
            n = int(rng.integers(1, 4))
            case = dict(size=int(rng.choice([16, 20, 24, 32, 40])), N=n, steps=int(rng.integers(4, 10)),
                        mode=int(rng.integers(0, 2)), use_hidden=bool(rng.integers(0, 2)), ext=bool(rng.integers(0, 2)),
                        seed=1000 + i, scatter=float(rng.choice([0.02, 0.05, 0.1])),
                        dying_env=(int(rng.integers(0, n)) if rng.random() < 0.5 else None),
                        p_tree_ca=float(rng.choice([0.0, 0.0, 0.03])))
            with contextlib.redirect_stdout(io.StringIO()):
                fx = mg.run_case(f"fuzz{i}", case, ab, jax.numpy)
            bad = R.replay_oracle(fx, case["mode"], case["ext"], f"fuzz{i} {case}")
            assert not bad, "\n".join(bad[:20])
    finally:
        jax_shim.set_rng_mode(prng.LEGACY)
        for k in [k for k in sys.modules if k.split(".")[0] in ("jax", "flax", "gymnasium", "gym_cellular_automata")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


def test_reference_source_runs_live_under_the_shim():
    """Where the reference tree exists (this container, not the GPU box): import its env through the shim, run a
    short 16x16 rollout and replay the oracle on it -- the generator of the fixtures above, exercised end to end."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("no reference tree here")
    import contextlib
    import importlib.util
    import io
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_reference_golden",
                                                  os.path.join(here, "golden", "make_reference_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    import ref_golden_util as R
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.split(".")[0] in ("jax", "flax", "gymnasium", "gym_cellular_automata")}
    try:
        jax = ref_shim.install(prng.LEGACY)
        ab = ref_shim.load("forest_fire.bulldozer.advanced_bulldozer")
        case = dict(size=16, N=2, steps=6, mode=prng.LEGACY, use_hidden=True, ext=True, seed=21, scatter=0.04,
                    dying_env=1, p_tree_ca=0.0)
        with contextlib.redirect_stdout(io.StringIO()):
            fx = mg.run_case("live16", case, ab, jax.numpy)
        bad = R.replay_oracle(fx, prng.LEGACY, True, "live16")
        assert not bad, "\n".join(bad[:20])
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in ("jax", "flax", "gymnasium", "gym_cellular_automata")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


def test_reference_golden_is_reproducible():
    """tests/golden/make_reference_golden.py is deterministic: regenerating sections from the reference's source (every
    generator the reference leaves unseeded is seeded from the case seed) reproduces the committed arrays exactly.
    Needs /root/reference (skipped on the GPU box)."""
    import importlib.util
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("the reference tree is not mounted here")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_reference_golden.py")
    spec = importlib.util.spec_from_file_location("make_reference_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    committed = np.load(mod.GOLDEN_PATH)
    out = mod.generate(only={"ref64_partitionable", "v3_32x48", "rollout_stats", "constants", "reset_frame"})
    assert len(out) > 50
    for k, v in out.items():
        v = np.asarray(v)
        assert k in committed.files, k
        assert v.shape == committed[k].shape and np.array_equal(v, committed[k]), k


@pytest.mark.parametrize("tag", ["ext", "plain"])
def test_reset_frame_rule_reproduces_reference_source_golden(tag):
    """The frame the reference's reset() returns (advanced_bulldozer.py:405-409) is grid_to_rgb of the display grid that
    grid_to_rgb_with_extensions picks out of the raw multi-channel sample -- NOT of channel 0.  The host rule the env
    uses for reset_obs="reference" plus the oracle's grid_to_rgb reproduce the recorded frame; channel 0 does not."""
    from gym_cellular_automata_b200.forest_fire.bulldozer.advanced_bulldozer import reference_reset_display
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_shim_golden.npz"))
    g = z[f"reset_frame/{tag}/sample"].astype(np.float32)
    pos, rgb = z[f"reset_frame/{tag}/position"], z[f"reset_frame/{tag}/rgb"]
    N, H, W, _ = g.shape
    night, dous = np.zeros(N, np.int32), np.zeros((N, H, W), np.int32)
    assert np.array_equal(ax.grid_to_rgb(reference_reset_display(g), night, dous, pos), rgb)
    assert not np.array_equal(ax.grid_to_rgb(g[..., 0], night, dous, pos), rgb)

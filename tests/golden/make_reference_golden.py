"""Generates tests/golden/reference_shim_golden.npz by executing the REFERENCE'S OWN SOURCE.

The reference env (``/root/reference/gym_cellular_automata/forest_fire/bulldozer/advanced_bulldozer.py`` and the
operator files it composes) is imported where it lies and run on CPU under ``oracle/ref_shim`` -- NumPy stand-ins
for the jax / flax / gymnasium names it touches (none of them is installable in this image).  The loop below is the
reference's rollout loop (``agents/jax_ppo.py``: ``stateless_step`` then ``conditional_reset``); every state
component after every step is recorded.  ``tests/test_oracle.py`` replays the oracle, ``tests/test_gpu_parity.py``
the CUDA path, on the recorded start states and actions and demand identical results.

What the vectors pin and what they cannot: see ``oracle/ref_shim/__init__.py`` (all reference Python on the path;
not jax.random's bit stream -- delegated to oracle/prng.py, pinned by known answers -- nor XLA's float32 summation
order).  jit semantics: ``initial_state`` is evaluated once per env instance (a jitted function bakes the sample it
saw at trace time); the subclass below does just that and nothing else.

Run (only here, /root/reference is needed):  python tests/golden/make_reference_golden.py
"""
import hashlib
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import prng, ref_shim  # noqa: E402
from oracle.ref_shim import gymnasium_shim, jax_shim  # noqa: E402

CASES = {
    # name: grid size, envs, steps, stream layout, hidden layers, extensions, seed, start-state recipe
    "ref64_legacy_ext": dict(size=64, N=3, steps=36, mode=prng.LEGACY, use_hidden=True, ext=True, seed=11,
                             scatter=0.02, dying_env=0, p_tree_ca=0.0),
    "ref32_nohidden_regrow": dict(size=32, N=4, steps=48, mode=prng.LEGACY, use_hidden=False, ext=False, seed=12,
                                  scatter=0.03, dying_env=1, p_tree_ca=0.01),
    "ref64_partitionable": dict(size=64, N=2, steps=12, mode=prng.PARTITIONABLE, use_hidden=True, ext=False, seed=13,
                                scatter=0.02, dying_env=None, p_tree_ca=0.0),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def actions_for(case, step):
    rng = np.random.default_rng(7000 * case["seed"] + step)
    N = case["N"]
    return np.stack([rng.integers(0, 9, N), rng.integers(0, 2, N), rng.integers(0, 3, N)], 1).astype(np.int32)


def u16(a, name):
    a = np.asarray(a)
    assert np.all(a == np.round(a)) and a.min() >= 0 and a.max() < 65536, name
    return a.astype(np.uint16)


def record(ctx_all):
    ctx = ctx_all["per_env_context"]
    return {
        "grid": np.asarray(ctx["true_grid"]).astype(np.uint8),
        "fire_age": u16(ctx["fire_age"], "fire_age"),
        "dousing": np.packbits(np.asarray(ctx["dousing_count"]).astype(np.uint8), axis=-1),
        "key": np.asarray(ctx["key"]).astype(np.uint32),
        "wind_index": np.asarray(ctx["wind_index"]).astype(np.int32),
        "time_step": np.asarray(ctx["time_step"]).astype(np.int32),
        "is_night": np.asarray(ctx["is_night"]).astype(np.int32),
        "position": np.asarray(ctx_all["position"]).astype(np.int32),
        "time": np.asarray(ctx_all["time"]).astype(np.float32),
    }


def run_case(name, case, ab, jnp):
    jax_shim.set_rng_mode(case["mode"])
    np.random.seed(case["seed"])
    random.seed(case["seed"])
    gymnasium_shim.seed_all(case["seed"])  # the spaces' / env's own generators (the reference leaves them unseeded)
    jax = sys.modules["jax"]

    class TraceOnce(ab.AdvancedForestFireBulldozerEnv):
        """initial_state evaluated once, as under jax.jit (the traced sample becomes a constant)."""
        _baked = None

        @property
        def initial_state(self):
            if self._baked is None:
                self._baked = ab.AdvancedForestFireBulldozerEnv.initial_state.fget(self)
            grid, c = self._baked
            return grid, {"per_env_context": dict(c["per_env_context"]), "shared_context": dict(c["shared_context"]),
                          "position": c["position"], "time": c["time"]}

    captured = []  # the float64 altitude get_slope is called with (the env keeps only a float32 copy)
    orig_get_slope = ab.get_slope
    ab.get_slope = lambda alt, *a, **k: (captured.append(np.array(alt, dtype=np.float64)), orig_get_slope(alt, *a, **k))[1]
    key = jax.random.split(jax.random.PRNGKey(1))[0]  # scripts/run:554-555
    S, N = case["size"], case["N"]
    env = TraceOnce(S, S, key=key, num_envs=N, speed_move=0.48, speed_act=0.12, use_hidden=case["use_hidden"],
                    enable_extensions=case["ext"])
    ab.get_slope = orig_get_slope
    obs, info = env.reset()
    rgb, ctx_all = obs
    out = {}
    snap = record(ctx_all)
    pec = ctx_all["per_env_context"]
    out["altitude"] = np.asarray(pec["altitude"]).astype(np.float32)
    out["altitude64"] = captured[0]
    assert np.array_equal(captured[0].astype(np.float32), out["altitude"])
    out["vegetation"] = np.asarray(pec["vegetation"]).astype(np.uint8)
    out["density"] = np.asarray(pec["density"]).astype(np.uint8)
    out["slope_sha256"] = np.array(sha(np.asarray(pec["slope"]).astype(np.float32)))
    out["slope_sample"] = np.asarray(pec["slope"]).astype(np.float32)[0, 5:9, 5:9]
    out["winds"] = np.asarray(ctx_all["shared_context"]["winds"]).astype(np.float32)
    sc = ctx_all["shared_context"]
    if case["p_tree_ca"]:
        sc["p_tree"] = jnp.array(case["p_tree_ca"], dtype=jnp.float32)  # the CA's regrowth probability (default 0)
    out["shared_scalars"] = np.array([float(sc["p_fire"]), float(sc["p_tree"]), float(sc["p_wind_change"]),
                                      float(sc["day_length"])], dtype=np.float64)
    for k, v in snap.items():
        out["snapshot/" + k] = v
    # ---- a richer start than two seed cells: scattered fires with short remaining ages; one env about to end
    rs = np.random.default_rng(case["seed"] + 500)
    grid = np.asarray(pec["true_grid"]).copy()
    age = np.asarray(pec["fire_age"]).copy()
    m = rs.random(grid.shape) < case["scatter"]
    grid[m] = 2.0
    age[m] = rs.integers(1, 30, size=int(m.sum())).astype(np.float32)
    d = case["dying_env"]
    if d is not None:  # a lone pair of cells that burns out at once: terminated -> conditional_reset
        grid[d][grid[d] == 2.0] = 1.0
        r, c = 3 * S // 4, S // 4
        grid[d, r - 2:r + 3, c - 3:c + 3] = 0.0
        grid[d, r, c] = grid[d, r, c - 1] = 2.0
        age[d, r, c], age[d, r, c - 1] = 2.0, 3.0
    pec["true_grid"] = jnp.asarray(grid.astype(np.float32))
    pec["fire_age"] = jnp.asarray(age.astype(np.float32))
    pec["time_step"] = jnp.asarray(np.full(N, 396, dtype=np.int32))  # day/night flips at time_step 400
    for k, v in record(ctx_all).items():
        out["start/" + k] = v
    steps = case["steps"]
    acts = np.stack([actions_for(case, s) for s in range(steps)])
    out["actions"] = acts
    rec = {}
    n_term = 0
    t0 = time.time()
    for s in range(steps):
        a = jnp.asarray(acts[s])
        step_tuple = env.stateless_step(a, obs, info)
        pre_rgb = np.asarray(step_tuple[0][0]).astype(np.float32)
        pre = {"step_reward": np.asarray(step_tuple[1]).astype(np.float32),
               "terminated": np.asarray(step_tuple[2]).astype(np.uint8),
               "pre_rgb_sha256": np.array(sha(pre_rgb))}
        n_term += int(pre["terminated"].sum())
        obs, reward, terminated, truncated, info = env.conditional_reset(step_tuple, a)
        post = record(obs[1])
        post["reward"] = np.asarray(reward).astype(np.float32)
        post["terminated_after_reset"] = np.asarray(terminated).astype(np.uint8)
        post["steps_elapsed"] = np.asarray(info["steps_elapsed"]).astype(np.float32)
        post["reward_accumulated"] = np.asarray(info["reward_accumulated"]).astype(np.float32)
        post["rgb_sha256"] = np.array(sha(np.asarray(obs[0]).astype(np.float32)))
        for k, v in {**pre, **post}.items():
            rec.setdefault(k, []).append(v)
    out["last_rgb"] = np.asarray(obs[0]).astype(np.float32)
    for k, v in rec.items():
        out["steps/" + k] = np.stack(v)
    g = out["steps/grid"]
    print(f"{name}: {steps} steps x {N} envs in {time.time() - t0:.1f} s; terminations {n_term}; "
          f"cells changed per step {np.mean((g[1:] != g[:-1]).sum(axis=(1, 2, 3))):.1f}; "
          f"burning at end {int((g[-1] == 2).sum())}; doused {int(np.unpackbits(out['steps/dousing'][-1]).sum())}")
    return out


def run_v3(ref_shim_mod, N=4, H=32, W=48, steps=40, seed=31):
    """The registered v3 rule set: the reference's ForestFireBulldozerEnv (bulldozer/bulldozer.py, NumPy
    WindyForestFire / Move / Modify / RepeatCA) is single-env; N independent instances are stepped with recorded
    actions.  Its only random draw per CA update, ``spaces.Box(0, 1, (3, 3)).sample()`` (ca_windy.py:53-60), is fed
    from a seeded stream and recorded."""
    import types
    from oracle.ref_shim import gymnasium_shim
    ops = sys.modules["gym_cellular_automata.forest_fire.operators"]
    for mod, names in (("ca_windy", ("WindyForestFire",)), ("move_modify", ("Move", "Modify", "MoveModify")),
                       ("repeat_ca", ("RepeatCA",))):
        m = ref_shim_mod.load("forest_fire.operators." + mod)
        for n in names:
            setattr(ops, n, getattr(m, n))
    rname = "gym_cellular_automata.forest_fire.bulldozer.utils.render"
    if rname not in sys.modules:
        r = types.ModuleType(rname)
        r.render = None
        sys.modules[rname] = r
    bd = ref_shim_mod.load("forest_fire.bulldozer.bulldozer")
    rng = np.random.default_rng(seed)
    consumed = []
    orig_sample = gymnasium_shim.Box.sample

    def sample(self):
        if self.shape == (3, 3):
            roll = rng.random((3, 3))
            consumed.append(roll)
            return roll
        return orig_sample(self)
    gymnasium_shim.Box.sample = sample
    try:
        envs = [bd.ForestFireBulldozerEnv(nrows=H, ncols=W, t_move=0.45, t_shoot=0.8) for _ in range(N)]
        first = [e.reset()[0] for e in envs]
        consumed.clear()
        out = {"grid0": np.stack([np.asarray(o[0]) for o in first]).astype(np.uint8),
               "position0": np.stack([np.asarray(o[1][1]) for o in first]).astype(np.int32),
               "time0": np.array([float(o[1][2]) for o in first], dtype=np.float64),
               "wind": np.asarray(first[0][1][0]["wind"], dtype=np.float64),
               "t_move_shoot_any": np.array([0.45, 0.8, 0.001])}
        acts = np.stack([rng.integers(0, 9, (steps, N)), rng.integers(0, 2, (steps, N))], -1).astype(np.int32)
        out["actions"] = acts
        rolls = np.zeros((steps, N, 3, 3, 3))
        rec = {k: [] for k in ("grid", "position", "time", "reward", "terminated", "repeats")}
        for s in range(steps):
            row = {k: [] for k in rec}
            for e, env in enumerate(envs):
                consumed.clear()
                if env.done:  # the reference refuses to step a finished env; keep it frozen
                    g, (cp, pos, t) = env.state
                    rew, term = np.nan, True
                else:
                    # CAEnv.step's sequence (ca_env.py:27-48) -- MDP transition, done check, reward -- with ONE
                    # repair: the env's context carries {"wind": array} where WindyForestFire.update compares the
                    # array itself (bulldozer.py:270 vs ca_windy.py:65: TypeError as shipped), so the
                    # MDP operator is handed the array
                    g0, (cp, pos0, t0) = env.state
                    wind = cp["wind"] if isinstance(cp, dict) else cp
                    g, (cp, pos, t) = env.MDP(g0, acts[s, e], (wind, pos0, t0))
                    env.state = env.grid, env.context = g, (cp, pos, t)
                    env._is_done()
                    try:
                        rew = env._award()
                    except ZeroDivisionError:
                        rew = np.nan
                    term = env.done
                assert len(consumed) <= 3
                for k, roll in enumerate(consumed):
                    rolls[s, e, k] = roll
                row["grid"].append(np.asarray(g).astype(np.uint8))
                row["position"].append(np.asarray(pos).astype(np.int32))
                row["time"].append(float(t))
                row["reward"].append(float(rew))
                row["terminated"].append(bool(term))
                row["repeats"].append(len(consumed))
            for k in rec:
                rec[k].append(np.stack([np.asarray(v) for v in row[k]]))
        out["rolls"] = rolls
        out["frozen"] = np.stack([np.array([False] * N)] + [r.astype(bool) for r in rec["terminated"][:-1]])
        for k, v in rec.items():
            out["steps/" + k] = np.stack(v)
        print(f"v3: {steps} steps x {N} envs of {H}x{W}; repeats seen {sorted(set(out['steps/repeats'].ravel().tolist()))}; "
              f"burning at end {int((out['steps/grid'][-1] == 25).sum())}")
        return out
    finally:
        gymnasium_shim.Box.sample = orig_sample


def run_operator_edges(ab, jnp, S=16):
    """Operator level, exhaustive where it is cheap: the reference env's own ``move_modify`` (MoveModifyJax) for
    every move x shoot action at every cell of the border ring (+ interior cells) of a 16x16 grid, and its
    ``time_per_action`` / RepeatCAJax clock arithmetic for every action pair at accumulated times around 1."""
    jax = sys.modules["jax"]
    key = jax.random.split(jax.random.PRNGKey(1))[0]
    env = ab.AdvancedForestFireBulldozerEnv(S, S, key=key, num_envs=1, speed_move=0.48, speed_act=0.12, use_hidden=False)
    cells = [(r, c) for r in range(S) for c in range(S) if r in (0, 1, S - 2, S - 1) or c in (0, 1, S - 2, S - 1)]
    cells += [(5, 7), (8, 8)]
    pos_in, act, pos_out, doused = [], [], [], []
    grid = jnp.zeros((S, S))
    for (r, c) in cells:
        for a0 in range(9):
            for a1 in range(2):
                pec = {"dousing_count": jnp.zeros((S, S), dtype=jnp.int32)}
                _, p, pec2 = env.move_modify(grid, (jnp.array(a0), jnp.array(a1)), jnp.array([r, c]), pec)
                pos_in.append((r, c)); act.append((a0, a1)); pos_out.append(np.asarray(p))
                d = np.argwhere(np.asarray(pec2["dousing_count"]) != 0)
                assert len(d) <= 1
                doused.append(d[0] if len(d) else np.array([-1, -1]))
    times_in, times_act, frac = [], [], []
    obs, _ = env.reset()
    pec0 = {k: v[0] for k, v in obs[1]["per_env_context"].items()}
    shared = obs[1]["shared_context"]
    for t_in in (0.0, 0.5, 0.86, 0.8687916, 0.8687917, 0.87, 0.95, 0.999):
        for a0 in range(9):
            for a1 in range(2):
                _, (_, f) = env.repeater(pec0["true_grid"], (jnp.array(a0), jnp.array(a1)), dict(pec0), shared,
                                         jnp.asarray(np.float32(t_in)))
                times_in.append(t_in); times_act.append((a0, a1)); frac.append(np.asarray(f, dtype=np.float32))
    print(f"operator edges: {len(pos_in)} move/douse cases, {len(frac)} clock cases")
    return {"size": np.array(S), "pos_in": np.array(pos_in, np.int32), "actions": np.array(act, np.int32),
            "pos_out": np.stack(pos_out).astype(np.int32), "doused": np.stack(doused).astype(np.int32),
            "clock_time_in": np.array(times_in, np.float32), "clock_actions": np.array(times_act, np.int32),
            "clock_frac": np.array(frac, np.float32)}


def run_reset_frame(ab, jnp, S=32, N=3, seed=17):
    """The frame reset() itself returns (advanced_bulldozer.py:401-420): grid_to_rgb_with_extensions applied to the raw
    multi-channel initial sample -- with extensions enabled its channels 3.. are independent random grids, and the
    row-index-as-channel-index quirk of :1028-1032 picks which one is shown.  Recorded: the whole sample, the position,
    the frame, with and without extension channels."""
    jax = sys.modules["jax"]
    out = {}
    for tag, ext in (("ext", True), ("plain", False)):
        np.random.seed(seed); random.seed(seed); gymnasium_shim.seed_all(seed)
        key = jax.random.split(jax.random.PRNGKey(1))[0]
        env = ab.AdvancedForestFireBulldozerEnv(S, S, key=key, num_envs=N, speed_move=0.48, speed_act=0.12,
                                                use_hidden=False, enable_extensions=ext)
        grid, ctx = env.initial_state
        env.__class__.initial_state = property(lambda self, _v=(grid, ctx): _v)  # jit bakes the traced sample
        try:
            (rgb, ctx_all), _ = env.reset()
        finally:
            env.__class__.initial_state = ab.AdvancedForestFireBulldozerEnv.__dict__["initial_state"]
        g = np.asarray(grid)
        assert g.ndim == 4 and np.all(g == np.round(g))
        out[tag + "/sample"] = g.astype(np.uint8)
        out[tag + "/position"] = np.asarray(ctx_all["position"]).astype(np.int32)
        out[tag + "/rgb"] = np.asarray(rgb).astype(np.float32)
        assert np.array_equal(np.asarray(ctx_all["per_env_context"]["true_grid"]), g[..., 0])
    print(f"reset frame: {S}x{S}, {N} envs, sample channels {out['ext/sample'].shape[-1]} / {out['plain/sample'].shape[-1]}")
    return out


CONSTANT_SIZES = (32, 64, 100, 128, 200, 256, 1024, 4096)


def run_constants(ref_shim_mod):
    """A0: what the reference's PartiallyObservableForestFireJax constructor derives from the grid size (heat kernel
    up to 21x21, dousing weights, fire-age range) for every BASELINE grid size, and the env's clock constants."""
    m = ref_shim_mod.load("forest_fire.operators.ca_alexandridis_jax")
    out = {"sizes": np.array(CONSTANT_SIZES, np.int32)}
    for s in CONSTANT_SIZES:
        op = m.PartiallyObservableForestFireJax(s, 0, 1, 2)
        out[f"{s}/burn_kernel"] = np.asarray(op.burn_kernel, dtype=np.float32)[0, 0]
        out[f"{s}/dousing_weights"] = np.asarray(op.dousing_weights, dtype=np.float32)
        out[f"{s}/scalars"] = np.array([op.initial_spread_time, op.fire_age_min, op.fire_age_max, op.burn_kernel_radius],
                                       dtype=np.float64)
    print(f"constants: sizes {CONSTANT_SIZES}")
    return out


def run_rollout_stats(N=37, steps=40, seed=41):
    """The statistics half of the PPO rollout step: ``step_env_wrapped`` is a closure inside
    agents/jax_ppo.py:run_rollout_loop (the module itself needs flax.linen / optax / orbax / tensorboard), so its
    source and the EpisodeStatistics dataclass are cut out of the reference file by ``ast`` at run time and executed
    under the shim against a stand-in env whose stateless_step / conditional_reset hand back prepared tuples
    (bursts of more than 10 simultaneous finishes, truncations, day and night)."""
    import ast
    import types
    path = os.path.join(ref_shim.REFERENCE_ROOT, "gym_cellular_automata", "agents", "jax_ppo.py")
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "EpisodeStatistics")
    loop = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_rollout_loop")
    fn = next(n for n in ast.walk(loop) if isinstance(n, ast.FunctionDef) and n.name == "step_env_wrapped")
    jax = sys.modules["jax"]
    jnp = jax.numpy
    env = types.SimpleNamespace()
    ns = {"jax": jax, "jnp": jnp, "flax": sys.modules["flax"], "env": env, "IS_FIRE_FIGHTER": True}
    exec(compile(ast.Module(body=[cls, fn], type_ignores=[]), path, "exec"), ns)
    Stats, step_env_wrapped = ns["EpisodeStatistics"], ns["step_env_wrapped"]
    z = lambda n, dt: jnp.zeros(n, dtype=dt)  # noqa: E731  (initial values: jax_ppo.py:486-501)
    st = Stats(episode_returns=z(N, jnp.float32), episode_lengths=z(N, jnp.int32),
               returned_episode_returns=z(N, jnp.float32), returned_episode_lengths=z(N, jnp.int32),
               recent_returns=z(10, jnp.float32), recent_lengths=z(10, jnp.int32),
               recent_idx=jnp.array(0, dtype=jnp.int32),
               current_day_correct=z(N, jnp.int32), current_night_correct=z(N, jnp.int32),
               current_day_steps=z(N, jnp.int32), current_night_steps=z(N, jnp.int32),
               recent_day_correct=z(10, jnp.int32), recent_night_correct=z(10, jnp.int32),
               recent_day_steps=z(10, jnp.int32), recent_night_steps=z(10, jnp.int32))
    rng = np.random.default_rng(seed)
    fields = [f for f in Stats.__dataclass_fields__]
    rec = {k: [] for k in ["actions", "reward", "terminated", "truncated", "is_night"] + ["out/" + f for f in fields]}
    for s in range(steps):
        p_fin = [0.0, 0.02, 0.5, 1.0][s % 4]
        actions = rng.integers(0, 3, (N, 3)).astype(np.int32)
        reward = (-rng.random(N)).astype(np.float32)
        term = rng.random(N) < p_fin
        trunc = (rng.random(N) < 0.02) & ~term
        night = rng.integers(0, 2, N).astype(np.int32)
        obs = (None, {"per_env_context": {"is_night": jnp.asarray(night)}})
        info = {"reward": jnp.asarray(reward), "terminated": jnp.asarray(term), "TimeLimit.truncated": jnp.asarray(trunc)}
        step_tuple = (obs, jnp.asarray(reward), jnp.asarray(term), jnp.asarray(trunc), info)
        env.stateless_step = lambda a, o, i, _t=step_tuple: _t
        env.conditional_reset = lambda t, a: t
        st, _ = step_env_wrapped(st, jnp.asarray(actions), obs, info)
        for k, v in (("actions", actions), ("reward", reward), ("terminated", term.astype(np.uint8)),
                     ("truncated", trunc.astype(np.uint8)), ("is_night", night.astype(np.uint8))):
            rec[k].append(v)
        for f in fields:
            rec["out/" + f].append(np.asarray(getattr(st, f)))
    out = {k: np.stack(v) for k, v in rec.items()}
    print(f"rollout stats: {steps} steps x {N} envs; finished {int(out['out/amount_finished'][-1])}")
    return out


GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_shim_golden.npz")


def generate(only=None, keep=None):
    """All sections (or the names in ``only``) as a dict of arrays; ``keep`` = arrays of an existing file whose
    sections are not regenerated.  Deterministic: every generator the reference leaves unseeded is seeded from the case
    seed (np.random / random / the shim's spaces), so a rerun reproduces the committed file array by array."""
    assert ref_shim.available(), "the reference tree is needed to generate these vectors"
    import contextlib
    import io
    out = {}
    if only is not None and keep is not None:
        out = {k: keep[k] for k in keep.files if k.split("/")[0] not in only}
    jax = ref_shim.install(prng.LEGACY)
    ab = ref_shim.load("forest_fire.bulldozer.advanced_bulldozer")
    for name, case in CASES.items():
        if only is not None and name not in only:
            continue
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):  # the reference constructor prints its slope histogram
            res = run_case(name, case, ab, jax.numpy)
        print(buf.getvalue().strip().splitlines()[-1])
        for k, v in res.items():
            out[f"{name}/{k}"] = v
    jax_shim.set_rng_mode(prng.LEGACY)
    gymnasium_shim.seed_all(1)
    if only is None or "v3_32x48" in only:
        for k, v in run_v3(ref_shim).items():
            out[f"v3_32x48/{k}"] = v
    if only is None or "operator_edges" in only:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            res = run_operator_edges(ab, jax.numpy)
        print(buf.getvalue().strip().splitlines()[-1])
        for k, v in res.items():
            out[f"operator_edges/{k}"] = v
    if only is None or "reset_frame" in only:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            res = run_reset_frame(ab, jax.numpy)
        print(buf.getvalue().strip().splitlines()[-1])
        for k, v in res.items():
            out[f"reset_frame/{k}"] = v
    if only is None or "constants" in only:
        for k, v in run_constants(ref_shim).items():
            out[f"constants/{k}"] = v
    if only is None or "rollout_stats" in only:
        for k, v in run_rollout_stats().items():
            out[f"rollout_stats/{k}"] = v
    gymnasium_shim.seed_all(None)
    return out


if __name__ == "__main__":
    # usage: make_reference_golden.py [--only name[,name...]]   (names: the CASES keys, v3_32x48, operator_edges, reset_frame, constants, rollout_stats);
    # with --only the other sections of the existing file are kept as they are
    only = set(sys.argv[sys.argv.index("--only") + 1].split(",")) if "--only" in sys.argv else None
    out = generate(only, np.load(GOLDEN_PATH) if only is not None else None)
    np.savez_compressed(GOLDEN_PATH, **out)
    print("wrote", GOLDEN_PATH, os.path.getsize(GOLDEN_PATH), "bytes")

"""ORACLE (test infrastructure, not product code) -- dense NumPy restatement of the
advanced-bulldozer environment step of frasermince/gym-cellular-automata.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  It is deliberately DENSE and literal (every cell draws its 12
random words, every neighbourhood is materialised), i.e. it follows the reference's data
flow, not the lazy/bit-board formulation of the CUDA kernels it checks.

Parity status: **pinned to the reference's own Python source** -- advanced_bulldozer.py and the operator
files are imported from /root/reference and executed under ``oracle/ref_shim`` (NumPy stand-ins for the
jax / flax / gymnasium names they touch; jax itself cannot be installed here); the rollouts recorded that
way (tests/golden/make_reference_golden.py -> tests/golden/reference_shim_golden.npz: stateless_step +
conditional_reset, observations included) are reproduced bit for bit by this module
(tests/test_oracle.py::test_oracle_reproduces_reference_source_golden).  **Still unpinned** (no jax / XLA
here, and the reference's tests hold no golden values): the bit stream of
``jax.random`` -- pinned only to public known-answer vectors (oracle/prng.py) -- and XLA's float32
evaluation details; reduction order (row-major, accumulator from +0) and ``exp`` (NumPy float32) are
stated assumptions, shared by the shim.

All paths below are relative to /root/reference/gym_cellular_automata/.

Batched convention: the reference vmaps single-env operators over axis 0; here every
array simply carries the leading env axis N.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from . import prng

EMPTY, TREE, FIRE = 0, 1, 2  # forest_fire/bulldozer/advanced_bulldozer.py:164-166
F32 = np.float32

# forest_fire/operators/ca_alexandridis_jax.py:170-173 (index 0 is an unreachable sentinel)
VEG_PROBS = np.array([-999, -0.1, 0.2, 0.5, 0.8, 1.2], dtype=F32)
DEN_PROBS = np.array([-999, -0.2, 0.2, 0.5, 0.8, 1.2], dtype=F32)
SLOPE_A = F32(0.078)  # ca_alexandridis_jax.py:199


class CAConstants:
    """Constants of PartiallyObservableForestFireJax.__init__
    (forest_fire/operators/ca_alexandridis_jax.py:54-160)."""

    def __init__(self, grid_size: int):
        self.grid_size = int(grid_size)
        self.initial_spread_time = self.grid_size + self.grid_size // 2  # :59
        self.fire_age_min = self.initial_spread_time * 1.5  # :60 (python float)
        self.fire_age_max = self.initial_spread_time * 1.75  # :61
        self.radius = math.ceil(math.log2(self.grid_size)) - 2  # :62
        border = 0.0007 * self.fire_age_max * 0.50  # :64
        inner = 0.006 * self.fire_age_max * 0.50  # :65
        dw = np.full((5, 5), border, dtype=np.float64)
        dw[1:4, 1:4] = inner
        self.dousing_weights = dw.astype(F32)  # :67-105 (python doubles -> f32 array)
        self.burn_kernel = self._burn_kernel(self.radius)  # :108-153

    @staticmethod
    def _burn_kernel(radius: int) -> np.ndarray:
        # :108-151 -- 0.065 total; ring i gets 60 % of what is left, shared by its cells
        # (ring 0 counts the centre too); the last ring takes the remainder.
        remaining = 0.065
        ring_w = []
        for i in range(radius):
            cells = (2 * i + 3) ** 2 - (2 * i + 1) ** 2 + (1 if i == 0 else 0)
            if i == radius - 1:
                ring_w.append(remaining / cells)
            else:
                ring_w.append(remaining * 0.60 / cells)
                remaining = remaining * 0.40
        size = 2 * radius + 1
        k = np.zeros((size, size), dtype=F32)
        c = radius
        if radius > 0:
            k[c, c] = ring_w[0]
        for i in range(radius):
            ring = i + 1
            for a in range(size):
                for b in range(size):
                    if max(abs(a - c), abs(b - c)) == ring:
                        k[a, b] = ring_w[i]
        return k

    @property
    def ring_weights(self) -> np.ndarray:
        """f32 weight of ring 1..R (centre shares ring 1's weight)."""
        c = self.radius
        return np.array([self.burn_kernel[c, c + r] for r in range(1, c + 1)], dtype=F32)


# ----------------------------------------------------------------------------------------
# batched random helpers (one key per env)
# ----------------------------------------------------------------------------------------

def _bits_batch(keys: np.ndarray, n: int, mode: int) -> np.ndarray:
    """(N,2) keys -> (N,n) uint32, each row = prng.random_bits(key, n)."""
    keys = np.asarray(keys, dtype=np.uint32)
    k0 = keys[:, 0:1]
    k1 = keys[:, 1:2]
    if mode == prng.LEGACY:
        m = n + (n & 1)
        h = m // 2
        cnt = np.arange(m, dtype=np.uint32)
        if n & 1:
            cnt[-1] = 0
        a, b = prng.threefry2x32(k0, k1, cnt[None, :h], cnt[None, h:])
        return np.concatenate([a, b], axis=1)[:, :n]
    idx = np.arange(n, dtype=np.uint64)
    hi = (idx >> np.uint64(32)).astype(np.uint32)[None]
    lo = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)[None]
    a, b = prng.threefry2x32(k0, k1, hi, lo)
    return (a ^ b).astype(np.uint32)


def split_batch(keys: np.ndarray, mode: int):
    """jax.random.split(key) per env -> (new_key (N,2), subkey (N,2))."""
    keys = np.asarray(keys, dtype=np.uint32)
    if mode == prng.LEGACY:
        out = _bits_batch(keys, 4, mode).reshape(-1, 2, 2)
    else:
        z = np.zeros((keys.shape[0], 1), np.uint32)
        a0, b0 = prng.threefry2x32(keys[:, 0:1], keys[:, 1:2], z, z)
        a1, b1 = prng.threefry2x32(keys[:, 0:1], keys[:, 1:2], z, z + np.uint32(1))
        out = np.stack([np.concatenate([a0, b0], 1), np.concatenate([a1, b1], 1)], axis=1)
    return out[:, 0].copy(), out[:, 1].copy()


def uniform_batch(keys, n, mode):
    return prng.bits_to_uniform(_bits_batch(keys, n, mode))


def randint_batch(keys, n, minval, maxval, mode):
    k1, k2 = split_batch(keys, mode)
    return prng.randint_from_bits(_bits_batch(k1, n, mode), _bits_batch(k2, n, mode), minval, maxval)


# ----------------------------------------------------------------------------------------
# A2/A3: burn probability + synchronous grid update
# ----------------------------------------------------------------------------------------

def p_slope_table(slope: np.ndarray) -> np.ndarray:
    """exp(f32(0.078) * slope) in float32 (ca_alexandridis_jax.py:199-200).  The same NumPy
    call produces the table handed to the CUDA kernels, so the two sides agree by
    construction (XLA's exp may differ from NumPy's in the last ulp: stated assumption)."""
    return np.exp(SLOPE_A * np.asarray(slope, dtype=F32)).astype(F32)


def _window_sum(padded_mask_f32: np.ndarray, weights: np.ndarray, H: int, W: int) -> np.ndarray:
    """Row-major sequential float32 sum of weights[i,j] * mask[r-n+i, c-n+j], acc from +0."""
    size = weights.shape[0]
    acc = np.zeros((padded_mask_f32.shape[0], H, W), dtype=F32)
    for i in range(size):
        for j in range(size):
            acc = (acc + padded_mask_f32[:, i:i + H, j:j + W] * weights[i, j]).astype(F32)
    return acc


def burn_probability(C: CAConstants, grid, dousing_count, vegetation, density, wind_matrix, pslope):
    """_compute_burn_probability (ca_alexandridis_jax.py:164-206) on a batch.

    grid (N,H,W) f32; dousing_count/vegetation/density (N,H,W) int; wind_matrix (N,3,3) f32;
    pslope (N,H,W,3,3) f32 -> (N,H,W,3,3) f32.
    """
    N, H, W = grid.shape
    R = C.radius
    fire_pad = np.pad((grid == FIRE).astype(F32), ((0, 0), (R, R), (R, R)))
    heat = _window_sum(fire_pad, C.burn_kernel, H, W)  # :348-349
    dous_pad = np.pad(dousing_count.astype(F32), ((0, 0), (2, 2), (2, 2)))
    dousing = _window_sum(dous_pad, C.dousing_weights, H, W)  # :345-346
    p_veg = VEG_PROBS[np.clip(vegetation, 1, 5)]  # :176-184
    p_den = DEN_PROBS[np.clip(density, 1, 5)]
    p_h = (heat - dousing).astype(F32)  # :198
    base = (p_h * (F32(1) + p_veg)).astype(F32)
    base = (base * (F32(1) + p_den)).astype(F32)
    p = (base[..., None, None] * wind_matrix[:, None, None, :, :]).astype(F32)
    p = (p * pslope).astype(F32)  # :206, evaluated left to right
    return p


def update_grid(C, grid, fire_age, dousing_count, vegetation, density, wind_matrix, pslope,
                p_tree, u_burn, u_grow, age_new):
    """_update_grid rule (ca_alexandridis_jax.py:379-398,423) given the random fields."""
    N, H, W = grid.shape
    p = burn_probability(C, grid, dousing_count, vegetation, density, wind_matrix, pslope)
    pad1 = np.pad(grid, ((0, 0), (1, 1), (1, 1)))
    hit = np.zeros((N, H, W), dtype=bool)
    for i in range(3):
        for j in range(3):
            nb_fire = pad1[:, i:i + H, j:j + W] == FIRE
            hit |= nb_fire & (u_burn[..., i, j] < p[..., i, j])
    tree = grid == TREE
    fire = grid == FIRE
    empty = grid == EMPTY
    new_grid = np.where(
        tree & hit, F32(FIRE),
        np.where(empty & (u_grow < F32(p_tree)), F32(TREE),
                 np.where(fire & (fire_age <= 1), F32(EMPTY), grid))).astype(F32)
    new_age = np.where((new_grid == FIRE) & (grid != FIRE), age_new.astype(F32), fire_age).astype(F32)
    new_age = np.where(fire, new_age - F32(1), new_age).astype(F32)
    return new_grid, new_age, p


def ca_update(C, grid, ctx, shared, mode=prng.LEGACY, inject: Optional[dict] = None, want_debug=False):
    """PartiallyObservableForestFireJax.update (ca_alexandridis_jax.py:426-460), batched.

    ctx keys used: fire_age, dousing_count, vegetation, density, pslope, wind_index, key.
    ``inject`` may hold u_burn (N,H,W,3,3), u_grow (N,H,W), age_new (N,H,W), u_wind (N,),
    wind_step (N,) to replace the corresponding draws (rule-parity mode); the key chain is
    advanced regardless.  Returns (new_grid, new_ctx[, debug]).
    """
    N, H, W = grid.shape
    winds = shared["winds"]  # (8,2,3,3) f32
    wind_matrix = winds[ctx["wind_index"], 0]  # :428
    key = ctx["key"]
    key, sub = split_batch(key, mode)  # :437  K1, S1
    # _update_grid key schedule :352-368
    ka, s_burn = split_batch(sub, mode)
    kb, s_grow = split_batch(ka, mode)
    _kc, s_age = split_batch(kb, mode)
    inject = inject or {}
    n_cells = H * W
    u_burn = inject.get("u_burn")
    if u_burn is None:
        u_burn = uniform_batch(s_burn, 9 * n_cells, mode).reshape(N, H, W, 3, 3)
    u_grow = inject.get("u_grow")
    if u_grow is None:
        u_grow = uniform_batch(s_grow, n_cells, mode).reshape(N, H, W)
    age_new = inject.get("age_new")
    if age_new is None:
        age_new = randint_batch(s_age, n_cells, C.fire_age_min, C.fire_age_max, mode).reshape(N, H, W)
    new_grid, new_age, p = update_grid(
        C, grid, ctx["fire_age"], ctx["dousing_count"], ctx["vegetation"], ctx["density"],
        wind_matrix, ctx["pslope"], shared["p_tree"], u_burn, u_grow, age_new)
    # wind random walk :443-451
    key, s_wind = split_batch(key, mode)  # K2
    u_wind = inject.get("u_wind")
    if u_wind is None:
        u_wind = uniform_batch(s_wind, 1, mode)[:, 0]
    key, s_idx = split_batch(key, mode)  # K3
    wind_step = inject.get("wind_step")
    if wind_step is None:
        wind_step = randint_batch(s_idx, 1, 1, 8, mode)[:, 0]
    change = u_wind < F32(shared["p_wind_change"])
    new_wind = np.where(change, (ctx["wind_index"] + wind_step) % 8, ctx["wind_index"]).astype(np.int32)
    new_ctx = dict(ctx)
    new_ctx["fire_age"] = new_age
    new_ctx["wind_index"] = new_wind
    new_ctx["key"] = key
    if want_debug:
        return new_grid, new_ctx, {"p": p, "u_burn": u_burn, "age_new": age_new}
    return new_grid, new_ctx


# ----------------------------------------------------------------------------------------
# A6-A11: clock, move, douse, MDP, reward, done
# ----------------------------------------------------------------------------------------

class EnvConstants:
    """Scalars of AdvancedForestFireBulldozerEnv.__init__/_init_time_mappings
    (forest_fire/bulldozer/advanced_bulldozer.py:81-303,745-777)."""

    def __init__(self, nrows, ncols, speed_move=0.12, speed_act=0.03, t_any=0.001,
                 t_move=None, t_shoot=None, p_tree_ca=0.0, p_wind_change=0.06, p_fire=0.00033):
        self.nrows, self.ncols = int(nrows), int(ncols)
        scale = (self.nrows + self.ncols) // 2  # :238
        self.t_env_any = t_any
        self.t_act_move = (1 / (speed_move * scale)) - t_any if t_move is None else t_move  # :240-242
        self.t_act_shoot = (1 / (speed_act * scale)) - self.t_act_move if t_shoot is None else t_shoot  # :244-246
        # every move (incl. not_move) and both shoot values cost the same (:746-754)
        self.movement_timings = np.full(9, self.t_act_move, dtype=np.float64).astype(F32)
        self.shooting_timings = np.full(2, self.t_act_shoot, dtype=np.float64).astype(F32)
        self.t_any_f32 = F32(t_any)
        self.p_tree = F32(p_tree_ca)  # :209 (0: no regrowth)
        self.p_wind_change = F32(p_wind_change)  # :210
        self.p_fire = F32(p_fire)
        self.day_length = 400  # :733
        self.ca = CAConstants(self.nrows)  # :276-282 uses nrows as grid_size

    def shared_context(self, winds):
        return {"winds": np.asarray(winds, dtype=F32), "p_fire": self.p_fire, "p_tree": self.p_tree,
                "p_wind_change": self.p_wind_change, "day_length": self.day_length}


UP_SET, DOWN_SET = (0, 1, 2), (6, 7, 8)  # advanced_bulldozer.py:262-268
LEFT_SET, RIGHT_SET = (0, 3, 6), (2, 5, 8)


def move(position, a0, nrows, ncols):
    """MoveJax.update (forest_fire/operators/move_modify_jax.py:39-62), batched."""
    row = position[:, 0].astype(np.int32).copy()
    col = position[:, 1].astype(np.int32).copy()
    valid_up, valid_down = row > 0, row < nrows - 1
    valid_left, valid_right = col > 0, col < ncols - 1
    row = np.where(np.isin(a0, UP_SET) & valid_up, row - 1, row)
    row = np.where(np.isin(a0, DOWN_SET) & valid_down, row + 1, row)
    col = np.where(np.isin(a0, LEFT_SET) & valid_left, col - 1, col)
    col = np.where(np.isin(a0, RIGHT_SET) & valid_right, col + 1, col)
    return np.stack([row, col], axis=1).astype(np.int32)


def modify(dousing_count, a1, position):
    """ModifyJax.update (move_modify_jax.py:102-114): shoot -> dousing_count[row,col] = 1."""
    d = dousing_count.copy()
    sel = np.nonzero(np.asarray(a1) == 1)[0]
    d[sel, position[sel, 0], position[sel, 1]] = 1
    return d


def count_cells(grid):
    """count_cells (advanced_bulldozer.py:941-953): int32 counts per env."""
    flat = grid.reshape(grid.shape[0], -1)
    return ((flat == TREE).sum(1).astype(np.int32), (flat == FIRE).sum(1).astype(np.int32))


def award(grid):
    """_award (advanced_bulldozer.py:597-630): -(f / (t + f + 1e-8)) in float32."""
    t, f = count_cells(grid)
    denom = ((t + f).astype(F32) + F32(1e-8)).astype(F32)
    return (-(f.astype(F32) / denom)).astype(F32)


def is_done(grid):
    """_is_done (advanced_bulldozer.py:632-633)."""
    return ~np.any(grid.reshape(grid.shape[0], -1) == FIRE, axis=1)


EXT_LOOKUP = np.array([[0, 0], [1, 0], [0, 1]], dtype=np.int32)  # create_up_to_k_mappings(2,1)


def full_actions(action):
    """_create_full_actions (advanced_bulldozer.py:308-330): (N,3) -> (N,4)."""
    action = np.asarray(action)
    return np.concatenate([action[:, :2], EXT_LOOKUP[action[:, 2]]], axis=1).astype(np.int32)


def mdp_update(E: EnvConstants, grid, action4, ctx, shared, position, time, K=1,
               mode=prng.LEGACY, inject=None, enable_extensions=False, render_obs=True):
    """MDP.update (advanced_bulldozer.py:1103-1133) + RepeatCAJax.update
    (forest_fire/operators/repeat_ca_jax.py:34-71), batched.

    K = number of CA sub-steps per env step.  K = 1 is the reference (repeat_ca_jax.py:61-63
    runs exactly one CA update whatever ``repeats`` is); K > 1 threads the context through
    K successive ca updates, which is what the commented fori_loop (:64-69) and the NumPy
    RepeatCA (forest_fire/operators/repeat_ca.py:42-43) do.  ``inject`` is a list of K dicts.
    """
    a0, a1 = action4[:, 0], action4[:, 1]
    # clock: repeat_ca_jax.py:35-41
    t_action = (E.movement_timings[a0] + E.shooting_timings[a1]).astype(F32)
    t_taken = (t_action + E.t_any_f32).astype(F32)
    new_time = (time.astype(F32) + t_taken).astype(F32)
    frac, _repeats = np.modf(new_time)
    cur_grid, cur_ctx = grid, ctx
    for k in range(K):
        inj = inject[k] if inject is not None else None
        cur_grid, cur_ctx = ca_update(E.ca, cur_grid, cur_ctx, shared, mode, inj)
    # move + douse (move_modify_jax.py:148-157)
    new_pos = move(position, a0, E.nrows, E.ncols)
    next_ctx = dict(cur_ctx)
    next_ctx["dousing_count"] = modify(cur_ctx["dousing_count"], a1, new_pos)
    next_ctx["true_grid"] = cur_grid  # :1117
    next_ctx["time_step"] = (cur_ctx["time_step"] + 1).astype(np.int32)  # :1118
    rgb = None
    if render_obs:
        # observation from the NEW grid/position but the INPUT context (:1120-1122)
        rgb = build_observation(cur_grid, new_pos, action4, ctx["dousing_count"], ctx["is_night"],
                                enable_extensions)
    flip = (next_ctx["time_step"] % shared["day_length"]) == 0  # :1123-1127
    next_ctx["is_night"] = np.where(flip, 1 - cur_ctx["is_night"], cur_ctx["is_night"]).astype(np.int32)
    return (rgb, cur_grid), (next_ctx, new_pos, frac.astype(F32))


def stateless_step(E, action, state, info, K=1, mode=prng.LEGACY, inject=None,
                   enable_extensions=False, render_obs=True):
    """stateless_step (advanced_bulldozer.py:332-399).  ``state`` = dict(per_env_context,
    shared_context, position, time); returns (rgb, state', reward, terminated, truncated, info')."""
    a4 = full_actions(action)
    ctx = state["per_env_context"]
    (rgb, grid), (nctx, npos, ntime) = mdp_update(
        E, ctx["true_grid"], a4, ctx, state["shared_context"], state["position"], state["time"],
        K, mode, inject, enable_extensions, render_obs)
    reward = award(grid)
    terminated = is_done(grid)
    truncated = np.zeros(grid.shape[0], dtype=bool)
    ninfo = dict(info)
    ninfo["reward"] = reward
    ninfo["terminated"] = terminated
    ninfo["TimeLimit.truncated"] = truncated
    ninfo["steps_elapsed"] = (info["steps_elapsed"] + F32(1)).astype(F32)  # :390
    ninfo["reward_accumulated"] = (info["reward_accumulated"] + reward).astype(F32)  # :391
    nstate = {"per_env_context": nctx, "shared_context": state["shared_context"],
              "position": npos, "time": ntime}
    return rgb, nstate, reward, terminated, truncated, ninfo


RESET_KEYS = ("wind_index", "density", "vegetation", "altitude", "slope", "pslope", "fire_age",
              "key", "true_grid", "dousing_count")  # every per-env key but time_step/is_night (:489-499)


def conditional_reset(E, rgb, state, reward, terminated, info, action, initial_state,
                      enable_extensions=False, render_obs=True):
    """conditional_reset (advanced_bulldozer.py:422-518).  ``initial_state`` is the snapshot
    the reference bakes in at trace time (F10)."""
    if not terminated.any():  # lax.cond(step_tuple[2].any(), ...) :513-518
        return rgb, state, reward, terminated, info
    N = terminated.shape[0]
    ctx = dict(state["per_env_context"])
    ictx = initial_state["per_env_context"]
    t3 = terminated[:, None, None]
    grid_new = np.where(t3, ictx["true_grid"], ctx["true_grid"]).astype(F32)  # :438-442
    pos_new = np.where(terminated[:, None], initial_state["position"], state["position"]).astype(np.int32)
    time_new = np.where(terminated, initial_state["time"], state["time"]).astype(F32)
    new_rgb = rgb
    if render_obs:
        # re-render terminated envs from the restored grid/position with the NOT yet restored
        # per-env context (post-step dousing_count / is_night) (:462-487)
        a4 = full_actions(action)
        fresh = build_observation(grid_new, pos_new, a4, ctx["dousing_count"], ctx["is_night"],
                                  enable_extensions)
        new_rgb = np.where(terminated[:, None, None, None], fresh, rgb).astype(F32)
    for k in RESET_KEYS:
        if k not in ctx:
            continue
        sel = terminated.reshape((N,) + (1,) * (ctx[k].ndim - 1))
        ctx[k] = np.where(sel, ictx[k], ctx[k]).astype(ctx[k].dtype)
    ctx["true_grid"] = grid_new  # :501
    ninfo = dict(info)
    ninfo["steps_elapsed"] = np.where(terminated, F32(0), info["steps_elapsed"]).astype(F32)
    ninfo["reward_accumulated"] = np.where(terminated, F32(0), info["reward_accumulated"]).astype(F32)
    new_reward = award(grid_new)  # :508
    nstate = {"per_env_context": ctx, "shared_context": state["shared_context"],
              "position": pos_new, "time": time_new}
    return new_rgb, nstate, new_reward, np.zeros_like(terminated), ninfo


# ----------------------------------------------------------------------------------------
# A13: observation
# ----------------------------------------------------------------------------------------

COLORS_DAY = np.array([[0xDD, 0xD1, 0xD3], [0xA9, 0xC4, 0x99], [0xE6, 0x81, 0x81]], dtype=np.int32)  # :41-44
COLORS_NIGHT = np.array([[0x69, 0x69, 0x69], [0x2F, 0x4F, 0x4F], [0x8B, 0x00, 0x00]], dtype=np.int32)  # :47-49
TINT_DAY = np.array([0, 0, 200], dtype=np.int32)  # :1082-1084
TINT_NIGHT = np.array([255, 165, 0], dtype=np.int32)


def apply_blur(grid):
    """apply_blur (forest_fire/bulldozer/utils/extension_utils.py:99-116): float32 op order kept."""
    N, H, W = grid.shape
    normalized = (grid.astype(F32) / F32(3.0)).astype(F32)
    k = F32(1.0) / F32(9.0)
    padded = np.pad(normalized, ((0, 0), (1, 1), (1, 1)), mode="edge")
    blurred = np.zeros_like(normalized)
    for i in range(3):
        for j in range(3):
            blurred = (blurred + k * padded[:, i:i + H, j:j + W]).astype(F32)
    return np.round(blurred * F32(3)).astype(np.int32)


def grid_to_rgb(display, is_night, dousing_count, position):
    """MDP.grid_to_rgb (advanced_bulldozer.py:1035-1101) -> (N,H,W,3) float32."""
    N, H, W = display.shape
    night = (np.asarray(is_night) != 0)
    pal = np.where(night[:, None, None], COLORS_NIGHT[None], COLORS_DAY[None])  # (N,3,3)
    rgb = np.broadcast_to(pal[:, 0][:, None, None, :], (N, H, W, 3)).astype(F32).copy()
    for v in (TREE, FIRE):  # later where() wins, same as the reference's nesting order
        m = (display == v)[..., None]
        rgb = np.where(m, pal[:, v][:, None, None, :].astype(F32), rgb)
    strength = np.where(dousing_count == 1, F32(0.75), F32(0)).astype(F32)[..., None]
    tint = np.where(night[:, None], TINT_NIGHT[None], TINT_DAY[None]).astype(F32)[:, None, None, :]
    blended = (rgb * (F32(1) - strength) + tint * strength).astype(F32)
    rgb = np.where((dousing_count > 0)[..., None], blended, rgb).astype(F32)
    e = np.arange(N)
    rgb[e, position[:, 0], position[:, 1]] = 0.0  # position colour is black day and night
    return rgb


def build_observation(grid, position, action4, dousing_count, is_night, enable_extensions):
    """build_observation_on_extensions + grid_to_rgb_with_extensions
    (advanced_bulldozer.py:988-1033; extension_utils.py:119-195), batched.

    Channel 0 = blurred grid when extensions are enabled else the raw grid; extension channel
    0 = raw grid if action bit 0, channel 1 = blurred grid if action bit 1.  The display
    channel quirk (:1028-1032): has_extension is evaluated PER ROW of the (H,W,2) extension
    block, first_valid = first row with a positive entry, and that ROW index is used
    (clamped to 1) as the CHANNEL index.
    """
    N, H, W = grid.shape
    g = grid.astype(F32)
    if enable_extensions:
        blur = apply_blur(g).astype(F32)
        base = blur
        ext0 = np.where((action4[:, 2] != 0)[:, None, None], g, F32(0))
        ext1 = np.where((action4[:, 3] != 0)[:, None, None], blur, F32(0))
    else:
        base = g
        ext0 = np.zeros_like(g)
        ext1 = np.zeros_like(g)
    ext = np.stack([ext0, ext1], axis=-1)  # (N,H,W,2)
    has_row = (ext > 0).any(axis=(2, 3))  # (N,H)
    any_ext = has_row.any(axis=1)
    first_valid = np.argmax(has_row, axis=1)
    ch = np.minimum(first_valid, 1)
    chosen = ext[np.arange(N), :, :, ch]
    display = np.where(any_ext[:, None, None], chosen, base)
    return grid_to_rgb(display, is_night, dousing_count, position)

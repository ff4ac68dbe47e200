"""AdvancedForestFireBulldozerEnv -- batched vector env on B200, drop-in for the reference's
functional API (reference forest_fire/bulldozer/advanced_bulldozer.py:63-953):

    obs, info = env.reset()
    obs, reward, terminated, truncated, info = env.stateless_step(action, obs, info)
    step_tuple = env.conditional_reset(step_tuple, action)

``obs = (rgb, context)``: ``rgb`` is the float32 (N,H,W,3) image of MDP.grid_to_rgb; ``context``
has the reference's keys (per_env_context / shared_context / position / time) as torch CUDA
tensors, unpacked lazily from the packed device state.  All state lives in HBM between calls;
one fused kernel launch performs clock + K CA sub-steps + move + douse + reward + done
(csrc/gca_step64.cu).  The host code here only marshals pointers through the C ABI.

Differences that are deliberate and documented:
 * arrays carry the env axis themselves (the reference vmaps a single-env MDP);
 * state is owned by the env and updated in place -- ``obs``/``info`` passed back into
   ``stateless_step`` must be the ones returned by the previous call (as in the reference's
   rollout loop); ``set_state`` injects arbitrary states;
 * ``substeps`` = K CA updates per env step (1 = the reference, repeat_ca_jax.py:61-63);
 * the initial grid is sampled once per env instance (the reference bakes one sample into its
   jitted reset/conditional_reset);
 * a working stateful ``step(action)`` (the reference's CAEnv.step is broken for this env, F9).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import numpy as np
import torch

from ... import _lib
from ..._config import TYPE_BOX, TYPE_INT
from ..._lib import check, current_stream, load, ptr
from ...ca_env import CAEnv
from ...grid_space import GridSpace
from ...operator import Operator
from ...packed import PackedState, StepOutputs, make_inject, make_params
from ... import spaces
from ..operators import (ModifyCUDA, MoveCUDA, MoveModifyCUDA, PartiallyObservableForestFireCUDA, RepeatCACUDA)
from .utils.init_utils import (create_up_to_k_mappings, get_slope, get_winds, init_altitude, init_altitude_same,
                               init_density, init_density_same, init_vegetation, init_vegetation_same,
                               p_slope_table)

EXTENSION_CHOICES = [(2, 1)]  # one registry of two extensions, at most one active (extension_utils.py:241-258)


def key_data(key) -> np.ndarray:
    """Raw (2,) uint32 threefry key from an int seed (jax.random.key(seed)) or raw key data."""
    if isinstance(key, (int, np.integer)):
        return np.array([(int(key) >> 32) & 0xFFFFFFFF, int(key) & 0xFFFFFFFF], dtype=np.uint32)
    k = key.detach().cpu().numpy() if torch.is_tensor(key) else np.asarray(key)
    k = k.astype(np.uint32).reshape(-1)
    assert k.size == 2, "key must be an int seed or 2 uint32 words"
    return k


def split_keys(key2: np.ndarray, num: int, rng_mode: str) -> np.ndarray:
    """jax.random.split(key, num) evaluated on the device through the C ABI test hook."""
    k = torch.as_tensor(np.asarray(key2, dtype=np.uint32)).cuda()
    out = torch.empty((num, 2), dtype=torch.uint32, device="cuda")
    mode = 0 if rng_mode in ("legacy", 0) else 1
    check(load().gca_threefry_split(ptr(k), num, mode, ptr(out), current_stream()), "gca_threefry_split")
    return out.cpu().numpy()


class _LazyPerEnv(dict):
    """per_env_context whose grid-sized float32/int32 arrays are unpacked on first access."""

    _LAZY = ("true_grid", "fire_age", "dousing_count")

    def __init__(self, env, static):
        super().__init__(static)
        self._env = env
        self._version = env._version

    def __missing__(self, k):
        if k in self._LAZY:
            if self._version != self._env._version:
                raise RuntimeError("stale context: the env has stepped since this observation was returned")
            out = self._env._state.unpack_to_reference(self._env._params, want=(k,))
            self[k] = out[k]
            return out[k]
        raise KeyError(k)

    def keys(self):
        return list(super().keys()) + [k for k in self._LAZY if not super().__contains__(k)]

    def __contains__(self, k):
        return super().__contains__(k) or k in self._LAZY

    # the rest of the mapping protocol goes through keys() / __getitem__, so the lazy arrays are part of every
    # iteration and copy.  (CPython's dict(d) / {**d} read the underlying table directly and still miss them: code of
    # this package that copies a context uses materialise_context(), below.)
    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def get(self, k, default=None):
        return self[k] if k in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]

    def copy(self):
        return {k: self[k] for k in self.keys()}


def materialise_context(per_env_context) -> dict:
    """A plain dict of a per_env_context, lazily unpacked arrays included (dict(ctx) would drop them) and host-lazy
    static layers resolved to their arrays."""
    out = {}
    for k in per_env_context.keys():
        v = per_env_context[k]
        out[k] = v.tensor() if isinstance(v, _HostLazy) else v
    return out


def reference_reset_display(grid_sample):
    """Display grid of the reference's reset() frame (advanced_bulldozer.py:405-409 -> :1021-1033) for the raw
    (N, H, W, 3 + extensions) initial sample: channel 0 when there are no extension channels; otherwise the
    reference tests "any entry > 0" per ROW of the (H, W, E) extension block and uses the first such row's index,
    clamped to E - 1, as the CHANNEL index (the same quirk every step's frame has)."""
    g = np.asarray(grid_sample)
    base, ext = g[..., 0], g[..., 3:]
    if ext.shape[-1] == 0:
        return base
    has_row = (ext > 0).any(axis=(2, 3))
    ch = np.minimum(np.argmax(has_row, axis=1), ext.shape[-1] - 1)
    chosen = ext[np.arange(g.shape[0]), :, :, ch]
    return np.where(has_row.any(axis=1)[:, None, None], chosen, base)


class AdvancedForestFireBulldozerEnv(CAEnv):
    metadata = {"render_modes": ["human"], "render_mode": "rgb_array"}

    @property
    def MDP(self):
        return self._MDP

    @property
    def initial_state(self):
        """(grid (N,H,W,5) float32, context) as NumPy, like the reference property (:70-79)."""
        self._ensure_initial()
        return self._init_grid5, self._init_context

    def __init__(self, nrows, ncols, key=0, num_envs=8, speed_move=0.12, speed_act=0.03, speed_multiplier=1.0,
                 pos_bull: Optional[Tuple] = None, pos_fire: Optional[Tuple] = None, t_move: Optional[float] = None,
                 t_shoot: Optional[float] = None, t_any=0.001, p_tree=0.90, p_empty=0.10, use_hidden: bool = True,
                 middle_fire: bool = False, enable_extensions: bool = False, *, device="cuda", substeps: int = 1,
                 rng_mode: str = "legacy", seed: Optional[int] = None, hidden: str = "reference",
                 obs_mode: str = "rgb_f32", auto_reset: bool = False, ca_p_tree: float = 0.0,
                 p_wind_change: float = 0.06, collect_stats: bool = False, use_tma: bool = True,
                 env_offset: int = 0, total_envs: Optional[int] = None, balance_every: int = 0,
                 generic_tiles: bool = False, reset_obs: str = "true_grid", **kwargs):
        super().__init__(nrows, ncols, **kwargs)
        if not torch.cuda.is_available():
            raise _lib.GcaError("AdvancedForestFireBulldozerEnv needs a CUDA device (sm_100a); there is no CPU path")
        load()
        self.device = torch.device(device)
        self.speed_multiplier = speed_multiplier
        self.middle_fire = middle_fire
        self.use_hidden = use_hidden
        self.enable_extensions = bool(enable_extensions)
        self.starting_key = key_data(key)
        self.rng_mode = rng_mode
        self.substeps = int(substeps)
        self.obs_mode = obs_mode
        self.auto_reset = bool(auto_reset)
        # observation returned by reset(): "true_grid" (channel 0 of the initial sample: what every later step shows) or
        # "reference" (the reference's own reset frame: grid_to_rgb_with_extensions on the raw multi-channel sample,
        # advanced_bulldozer.py:405-409 -- with extensions its channels 3.. are independent random grids)
        if reset_obs not in ("true_grid", "reference"):
            raise ValueError("reset_obs must be 'true_grid' or 'reference'")
        self.reset_obs = reset_obs
        # 64x64 kernel: re-deal envs to warps every `balance_every` steps by last step's cost (0 = off)
        self.balance_every = int(balance_every)
        self._steps_since_balance = 0
        self._host_act_dev = None
        self._host_args = None
        self.kernel_launches = 0  # kernels of libgca launched by the step path (step + re-balancing)
        self.num_envs = int(num_envs)
        # multi-GPU sharding: this instance holds envs [env_offset, env_offset + num_envs) of a batch of
        # total_envs; per-env keys come from ONE split over the whole batch, so they do not depend on
        # how many ranks share it
        self.env_offset = int(env_offset)
        self.total_envs = int(total_envs) if total_envs is not None else self.env_offset + self.num_envs
        self.title = "ForestFireBulldozer" + str(nrows) + "x" + str(ncols)
        self._host_rng = np.random.RandomState(seed) if seed is not None else np.random
        self.np_random = np.random.default_rng(seed)
        self.shared_context_keys = {"winds", "p_fire", "p_tree", "p_wind_change", "day_length"}
        self.per_env_context_keys = {"wind_index", "density", "vegetation", "altitude", "slope", "fire_age", "key",
                                     "is_night", "true_grid", "time_step", "dousing_count"}
        self._empty, self._tree, self._fire = 0, 1, 2
        self._shoots = {"shoot": 1, "none": 0}
        self._moves = {n: i for i, n in enumerate(("up_left", "up", "up_right", "left", "not_move", "right",
                                                   "down_left", "down", "down_right"))}
        self._action_sets = {"up": {0, 1, 2}, "down": {6, 7, 8}, "left": {0, 3, 6}, "right": {2, 5, 8},
                             "not_move": {4}}
        self._pos_bull = pos_bull
        self._pos_fire = [pos_fire] * num_envs if pos_fire is not None else None
        self._p_tree_init, self._p_empty_init = p_tree, p_empty
        self._p_fire, self._p_tree, self._p_wind_change = 0.00033, float(ca_p_tree), float(p_wind_change)
        self._t_env_any = t_any

        # static inputs (host, A14)
        winds = get_winds(use_hidden)
        self._winds = winds.astype(np.float32)
        N = self.num_envs
        self._pslope_dev = None
        if use_hidden and hidden == "device":
            # layers generated on the GPU (gca_generate_hidden): same layer models, counter-based random numbers that
            # depend only on (seed, global env index); everything stays a device tensor
            density, vegetation, altitude, self._slope, self._pslope_dev = self._generate_hidden_device(
                nrows, ncols, N, int(seed or 0), int(env_offset))
        elif use_hidden and hidden == "reference":
            density = init_density(nrows, ncols, N, self._host_rng)
            vegetation = init_vegetation(nrows, ncols, N, self._host_rng)
            altitude = init_altitude(nrows, ncols, N, self._host_rng)
        elif use_hidden:  # "random": iid synthetic layers (benchmarks)
            density = self.np_random.integers(1, 6, size=(N, nrows, ncols))
            vegetation = self.np_random.integers(1, 6, size=(N, nrows, ncols))
            altitude = self.np_random.uniform(0, 1.1, size=(N, nrows, ncols))
        else:
            density = init_density_same(nrows, ncols, N)
            vegetation = init_vegetation_same(nrows, ncols, N)
            altitude = init_altitude_same(nrows, ncols, N)
        if self._pslope_dev is None:
            self._density = np.asarray(density).astype(np.int32)
            self._vegitation = np.asarray(vegetation).astype(np.int32)
            self._altitude = np.asarray(altitude).astype(np.float32)
            self._slope = get_slope(np.asarray(altitude, dtype=np.float64)).astype(np.float32) if use_hidden else None
        else:
            self._density, self._vegitation, self._altitude = density, vegetation, altitude

        self._params = make_params(nrows, ncols, self.substeps, speed_move, speed_act, t_any, t_move, t_shoot,
                                   self._p_tree, self._p_wind_change, rng_mode, self._winds[:, 0])
        self._t_act_move = float(self._params.t_move[0])
        self._t_act_shoot = float(self._params.t_shoot[0])
        self._flags = 0 if use_hidden else _lib.FLAG_NO_HIDDEN
        if not use_tma:
            self._flags |= _lib.FLAG_NO_TMA
        if generic_tiles:  # the tiled kernels even where the whole-grid bit-board kernel applies (tests, A/B runs)
            self._flags |= _lib.FLAG_GENERIC_TILES

        self._set_spaces()
        self.ca = PartiallyObservableForestFireCUDA(nrows, self._empty, self._tree, self._fire, params=self._params,
                                                    use_hidden=use_hidden, **self.ca_space)
        self.move = MoveCUDA(self._action_sets, params=self._params, **self.move_space)
        self.modify = ModifyCUDA({}, params=self._params, **self.modify_space)
        self.move_modify = MoveModifyCUDA(self.move, self.modify, **self.move_modify_space)
        self.repeater = RepeatCACUDA(self.ca, self.time_per_action, self.time_per_state, **self.repeater_space)
        self._MDP = MDP(self, **self.MDP_space)

        self._state = None
        self._snapshot = None
        self._init_grid5 = None
        self._version = 0
        self._version_structs = 0  # bumped whenever a state/snapshot buffer is re-allocated
        self._fast_args = None
        self._out = StepOutputs(N, self.device, with_stats=collect_stats)
        self._rgb = None
        self._scratch = torch.zeros(N, dtype=torch.int32, device=self.device)
        self._actions = torch.zeros((N, 3), dtype=torch.int32, device=self.device)
        self._shared_ctx = None
        self._false = None
        self._host_lazy = None
        self._slope_packed = False

    # ------------------------------------------------------------------------------------------
    # clock helpers (advanced_bulldozer.py:745-777)
    # ------------------------------------------------------------------------------------------
    def time_per_action(self, action):
        a0, a1 = action
        tm = torch.tensor(list(self._params.t_move), dtype=torch.float32, device=self.device)
        ts = torch.tensor(list(self._params.t_shoot), dtype=torch.float32, device=self.device)
        return tm[a0.long()] + ts[a1.long()]

    def time_per_state(self, s):
        return torch.tensor(float(self._params.t_any), dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------------------------------
    # spaces (advanced_bulldozer.py:779-939)
    # ------------------------------------------------------------------------------------------
    def _set_spaces(self):
        N, H, W = self.num_envs, self.nrows, self.ncols
        inf = float("inf")
        self.per_env_context_space = {
            "wind_index": spaces.Box(0, 7, shape=(N,), dtype=TYPE_INT),
            "density": spaces.Box(0, 5, shape=(N, H, W), dtype=TYPE_INT),
            "vegetation": spaces.Box(0, 5, shape=(N, H, W), dtype=TYPE_INT),
            "altitude": spaces.Box(0.0, inf, shape=(N, H, W), dtype=TYPE_BOX),
            "slope": spaces.Box(-90.0, 90.0, shape=(N, H, W, 3, 3), dtype=TYPE_BOX),
            "fire_age": spaces.Box(0, inf, shape=(N, H, W), dtype=TYPE_BOX),
            "is_night": spaces.Box(0, 1, shape=(N,), dtype=TYPE_BOX),
            "time_step": spaces.Box(0, inf, shape=(N,), dtype=TYPE_BOX),
            "true_grid": spaces.Box(0, 2, shape=(N, H, W), dtype=TYPE_BOX),
        }
        self.shared_context_space = {
            "winds": spaces.Box(0.0, 1.0, shape=(8, 3, 3), dtype=TYPE_BOX),
            "p_fire": spaces.Box(0.0, 1.0, shape=(), dtype=TYPE_BOX),
            "p_tree": spaces.Box(0.0, 1.0, shape=(), dtype=TYPE_BOX),
            "p_wind_change": spaces.Box(0.0, 1.0, shape=(), dtype=TYPE_BOX),
            "day_length": spaces.Box(0.0, inf, shape=(), dtype=TYPE_BOX),
        }
        self.position_space = spaces.Box(low=np.zeros((N, 2), dtype=TYPE_INT),
                                         high=np.tile(np.array([H, W]), (N, 1)), shape=(N, 2), dtype=TYPE_INT)
        self.time_space = spaces.Box(0.0, inf, shape=(N,), dtype=TYPE_BOX)
        self.context_space = spaces.Dict({
            "per_env_context": spaces.Dict(self.per_env_context_space),
            "shared_context": spaces.Dict(self.shared_context_space),
            "position": self.position_space,
            "time": self.time_space,
        })
        m, n = len(self._moves), len(self._shoots)
        self.action_space = spaces.MultiDiscrete(nvec=np.array([[m, n]] * N), dtype=TYPE_INT)
        self.extension_choices = list(EXTENSION_CHOICES)
        ext_nvec = np.array([sum(math.comb(a, i) for i in range(k + 1)) for a, k in self.extension_choices])
        self.extension_space = spaces.MultiDiscrete(
            nvec=np.array([math.comb(a, k) for a, k in self.extension_choices]), dtype=TYPE_INT)
        self.total_action_space = spaces.MultiDiscrete(
            nvec=np.array([np.concatenate([np.array([m, n]), ext_nvec])] * N), dtype=TYPE_INT)
        self.grid_space = GridSpace(values=[self._empty, self._tree, self._fire], shape=(N, H, W, 3))
        self._extension_lookups = [create_up_to_k_mappings(a, k)[0] for a, k in self.extension_choices]
        self.observation_space = spaces.Tuple((self.grid_space, self.context_space))
        base = {"grid_space": self.grid_space, "context_space": self.context_space}
        self.ca_space = dict(base, action_space=self.action_space)
        self.move_space = {"grid_space": self.grid_space, "action_space": spaces.Discrete(m),
                           "context_space": self.position_space}
        self.modify_space = {"grid_space": self.grid_space, "action_space": spaces.Discrete(n),
                             "context_space": self.position_space}
        self.move_modify_space = {"grid_space": self.grid_space, "action_space": self.action_space,
                                  "context_space": self.position_space}
        self.repeater_space = dict(base, action_space=self.action_space)
        self.MDP_space = dict(base, action_space=self.action_space)

    # ------------------------------------------------------------------------------------------
    # initial distributions (advanced_bulldozer.py:650-743), host NumPy
    # ------------------------------------------------------------------------------------------
    def _initial_grid_distribution(self):
        N, H, W = self.num_envs, self.nrows, self.ncols
        total_ext = sum(n for n, _ in self.extension_choices)
        gs = GridSpace(values=[self._empty, self._tree, self._fire],
                       probs=[self._p_empty_init, self._p_tree_init, 0.0], shape=(N, H, W, total_ext + 3))
        gs._np_random = self.np_random
        grid = gs.sample().astype(np.float32)
        if self._pos_fire is None:
            if self.middle_fire:
                r, c = H // 2, W // 2
            else:
                r, c = 3 * H // 4, W // 4
            self._pos_fire = [[(r, c), (r, c - 1)] for _ in range(N)]
        init_age = (H + H // 2) * 2
        fire_age = np.zeros((N, H, W), dtype=np.float32)
        for e in range(N):
            cells = self._pos_fire[e]
            if isinstance(cells[0], (int, np.integer)):
                cells = [tuple(cells)]
            for r, c in cells:
                grid[e, r, c] = self._fire
                fire_age[e, r, c] = init_age
        return grid, fire_age

    def _initial_context_distribution(self, fire_age, grid):
        N, H, W = self.num_envs, self.nrows, self.ncols
        if self._pos_bull is None:
            self._pos_bull = [(int(H * 0.15), int(W * 0.85)) for _ in range(N)]
        elif isinstance(self._pos_bull[0], (int, np.integer)):
            self._pos_bull = [tuple(self._pos_bull)] * N
        keys = split_keys(self.starting_key, self.total_envs, self.rng_mode)[self.env_offset:self.env_offset + N]
        wind_index = (self.np_random.integers(0, 8, size=N) if self.use_hidden else np.zeros(N)).astype(np.int32)
        per_env = {
            "wind_index": wind_index,
            "density": self._density,
            "vegetation": self._vegitation,
            "altitude": self._altitude,
            "slope": self._slope if self._slope is not None else np.zeros((N, H, W, 3, 3), dtype=np.float32),
            **({"pslope": self._pslope_dev} if self._pslope_dev is not None else {}),
            "fire_age": fire_age,
            "key": keys,
            "is_night": np.zeros(N, dtype=np.int32),
            "true_grid": np.ascontiguousarray(grid[..., 0]),
            "time_step": np.ones(N, dtype=np.int32),
            "dousing_count": np.zeros((N, H, W), dtype=np.int32),
        }
        shared = {"winds": self._winds, "p_fire": np.float32(self._p_fire), "p_tree": np.float32(self._p_tree),
                  "p_wind_change": np.float32(self._p_wind_change), "day_length": 400}
        return {"per_env_context": per_env, "shared_context": shared,
                "position": np.array(self._pos_bull, dtype=np.int32), "time": np.zeros(N, dtype=np.float32)}

    def _generate_hidden_device(self, H, W, N, seed, env_offset):
        d = self.device
        veg = torch.empty((N, H, W), dtype=torch.int32, device=d)
        den = torch.empty((N, H, W), dtype=torch.int32, device=d)
        alt = torch.empty((N, H, W), dtype=torch.float32, device=d)
        alt64 = torch.empty((N, H, W), dtype=torch.float64, device=d)
        slope = torch.empty((N, H, W, 3, 3), dtype=torch.float32, device=d)
        pslope = torch.empty((N, H, W, 3, 3), dtype=torch.float32, device=d)
        check(load().gca_generate_hidden(N, H, W, seed, env_offset, ptr(veg), ptr(den), ptr(alt), ptr(alt64), ptr(slope),
                                         ptr(pslope), current_stream()), "gca_generate_hidden")
        return den, veg, alt, slope, pslope

    def _ensure_initial(self):
        if self._init_grid5 is None:
            grid5, fire_age = self._initial_grid_distribution()
            self._init_grid5 = grid5
            self._init_context = self._initial_context_distribution(fire_age, grid5)

    # ------------------------------------------------------------------------------------------
    # device state
    # ------------------------------------------------------------------------------------------
    def set_state(self, per_env_context, position, time, *, as_snapshot: bool = False, info=None):
        """Load an arbitrary reference-layout state (NumPy or torch arrays) into the packed device
        state.  ``pslope`` (N,H,W,3,3) may be given instead of ``slope``."""
        N, H, W = self.num_envs, self.nrows, self.ncols
        if self._state is None:
            self._version_structs += 1
            self._state = PackedState(N, H, W, self.device, use_hidden=self.use_hidden)
            self._slope_packed = False
        ctx = materialise_context(per_env_context)
        if self.use_hidden:
            own = self._host_lazy.get("slope") if self._host_lazy else None
            own_t = own._t if isinstance(own, _HostLazy) else own
            if "pslope" not in ctx and "slope" in ctx and self._slope_packed and own_t is not None and ctx["slope"] is own_t:
                del ctx["slope"]  # the env's own slope tensor (from a previous observation): its factor table is packed already
            elif "pslope" not in ctx and "slope" not in ctx and not self._slope_packed:
                if self._pslope_dev is not None:
                    ctx["pslope"] = self._pslope_dev
                else:
                    ctx["slope"] = self._slope
            self._slope_packed = True
        self._state.tick.zero_()  # ages are stored as burn-out ticks relative to tick 0
        self._state.pack_from_reference(self._params, ctx, position, time)
        if info is not None:
            self._state.steps_elapsed.copy_(torch.as_tensor(np.asarray(info["steps_elapsed"], dtype=np.float32)))
            self._state.reward_accumulated.copy_(
                torch.as_tensor(np.asarray(info["reward_accumulated"], dtype=np.float32)))
        else:
            self._state.steps_elapsed.zero_()
            self._state.reward_accumulated.zero_()
        if as_snapshot or self._snapshot is None:
            self._version_structs += 1
            self._snapshot = self._state.clone()
            self._snap_reward = torch.zeros(N, dtype=torch.float32, device=self.device)
            check(load().gca_reward_done(C.byref(self._params), C.byref(self._snapshot.cstruct()),
                                         ptr(self._snap_reward), None, None, current_stream()), "gca_reward_done")
        self._version += 1

    def _static_context(self, sc=None):
        d = self.device
        if sc is None:
            sc = PackedState.carve_scalars(self._state._scalars.clone(), self.num_envs)
        if self._host_lazy is None:  # static host arrays, uploaded on first use
            def lazy(a):  # device tensors (hidden="device") are handed out as they are
                return a if torch.is_tensor(a) and a.is_cuda else _HostLazy(a, d)
            self._host_lazy = {
                "density": lazy(self._density), "vegetation": lazy(self._vegitation), "altitude": lazy(self._altitude),
                "slope": lazy(self._slope if self._slope is not None
                              else np.zeros((self.num_envs, self.nrows, self.ncols, 3, 3), np.float32))}
        static = {"wind_index": sc["wind_index"], "key": sc["key"], "is_night": sc["is_night"],
                  "time_step": sc["time_step"]}
        static.update(self._host_lazy)
        return static

    def _shared_context(self):
        """shared_context (advanced_bulldozer.py:724-731): constants, uploaded once (a host-to-device copy per step
        would synchronise the rollout loop with the device) and handed out like the reference hands its dict through."""
        if self._shared_ctx is None:
            self._shared_ctx = {
                "winds": torch.as_tensor(self._winds, device=self.device),
                "p_fire": torch.tensor(self._p_fire, dtype=torch.float32, device=self.device),
                "p_tree": torch.tensor(self._p_tree, dtype=torch.float32, device=self.device),
                "p_wind_change": torch.tensor(self._p_wind_change, dtype=torch.float32, device=self.device),
                "day_length": 400}
        return dict(self._shared_ctx)

    def _all_false(self):
        """The (N,) all-False tensor of ``truncated`` (never written: shared between steps)."""
        if self._false is None:
            self._false = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        return self._false

    def _context_view(self, sc=None):
        """The context pytree of the current state.  ``sc``: views of ONE snapshot copy of the per-env scalar block
        (PackedState.carve_scalars) -- what used to be eight small clones per step."""
        if sc is None:
            sc = PackedState.carve_scalars(self._state._scalars.clone(), self.num_envs)
        per_env = _LazyPerEnv(self, self._static_context(sc))
        shared = self._shared_context()
        return {"per_env_context": per_env, "shared_context": shared, "position": sc["position"], "time": sc["time"]}

    def _render(self, cell, doused, position, night_u8, ext_action, env_mask=None, actions=None):
        """``ext_action``: (N,) int32 extension ids, or ``actions``: the step's (N,3) int32 CUDA action triples (the
        kernel then reads the third column itself, gca_render_rgb_actions)."""
        if self.obs_mode == "none":
            return None
        N, H, W = self.num_envs, self.nrows, self.ncols
        u8 = self.obs_mode == "rgb_u8"
        if self._rgb is None or env_mask is None:
            # a fresh buffer per step keeps earlier observations valid (rollout storage keeps them)
            self._rgb = torch.empty((N, H, W, 3), dtype=torch.uint8 if u8 else torch.float32, device=self.device)
        if actions is not None:
            check(load().gca_render_rgb_actions(C.byref(self._params), N, ptr(cell), ptr(doused), ptr(position),
                                                ptr(night_u8), ptr(actions, torch.int32, 3 * N, "actions"), ptr(env_mask),
                                                int(self.enable_extensions), int(u8), ptr(self._scratch),
                                                ptr(self._rgb), current_stream()), "gca_render_rgb_actions")
        else:
            check(load().gca_render_rgb(C.byref(self._params), N, ptr(cell), ptr(doused), ptr(position), ptr(night_u8),
                                        ptr(ext_action), ptr(env_mask), int(self.enable_extensions), int(u8),
                                        ptr(self._scratch), ptr(self._rgb), current_stream()), "gca_render_rgb")
        return self._rgb

    def _info(self, terminated=None, sc=None, oc=None):
        st, out = self._state, self._out
        return {"reward": out.step_reward.clone() if oc is None else oc["step_reward"],
                "terminated": out.terminated.bool() if terminated is None else terminated,
                "TimeLimit.truncated": self._all_false(),
                "steps_elapsed": st.steps_elapsed.clone() if sc is None else sc["steps_elapsed"],
                "reward_accumulated": st.reward_accumulated.clone() if sc is None else sc["reward_accumulated"]}

    # ------------------------------------------------------------------------------------------
    # functional API
    # ------------------------------------------------------------------------------------------
    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        """reset (advanced_bulldozer.py:401-420): initial grid/context, RGB rendered from channel 0."""
        self._ensure_initial()
        ic = self._init_context
        self._state = None
        self.set_state(ic["per_env_context"], ic["position"], ic["time"], as_snapshot=True)
        N = self.num_envs
        st = self._state
        zeros_u8 = torch.zeros(N, dtype=torch.uint8, device=self.device)
        ext0 = torch.zeros(N, dtype=torch.int32, device=self.device)
        self._rgb = None
        # the reference renders reset observations with grid_to_rgb_with_extensions on the raw
        # 5-channel sample (:405-409): channels 3,4 are independent random grids there.  By default the
        # observation is drawn from the true grid (channel 0), which is what every later step shows;
        # reset_obs="reference" reproduces the reference's frame (display grid chosen on the host: reset is
        # not on the hot path).
        shown = st.cell
        if self.reset_obs == "reference":
            disp = reference_reset_display(self._init_grid5)
            shown = torch.as_tensor(np.ascontiguousarray(disp.astype(np.uint8)), device=self.device)
        enable = self.enable_extensions
        self.enable_extensions = False
        rgb = self._render(shown, st.doused, st.position, zeros_u8, ext0)
        self.enable_extensions = enable
        info = {"TimeLimit.truncated": torch.zeros(N, dtype=torch.bool, device=self.device),
                "terminated": torch.zeros(N, dtype=torch.bool, device=self.device),
                "steps_elapsed": torch.zeros(N, dtype=torch.float32, device=self.device),
                "reward_accumulated": torch.zeros(N, dtype=torch.float32, device=self.device),
                "reward": torch.zeros(N, dtype=torch.float32, device=self.device)}
        self._out.terminated.zero_()
        return (rgb, self._context_view()), info

    def _kernels_per_step(self, auto_reset: bool) -> int:
        """libgca kernels one env step launches: 64x64 -- the fused kernel; whole-grid bit-board grids (W % 64 == 0, up to
        256x256) -- one kernel (+ the reset kernel); anything else -- the dense tile count, the key schedules + the
        step's tile list, one tile kernel per CA sub-step (+ a copy-back when their number is odd) and the epilogue,
        replayed as one CUDA graph (+ the reset kernel)."""
        if self._state.work is not None:
            return 1
        H, W, R = self.nrows, self.ncols, int(self._params.R)
        bb = (W % 64 == 0 and W <= 256 and H <= 256 and H * W <= 65536 and 4 <= R <= 6
              and not (self._flags & _lib.FLAG_GENERIC_TILES))
        return (1 if bb else self.substeps + 3 + (self.substeps & 1)) + (1 if auto_reset else 0)

    def _can_fuse_render(self) -> bool:
        """The step kernel draws the observation itself (GCA_FLAG_RENDER): 64x64 grids without extensions."""
        return (self.obs_mode != "none" and not self.enable_extensions and self._state is not None
                and self._state.work is not None)

    def _new_rgb(self):
        u8 = self.obs_mode == "rgb_u8"
        # a fresh buffer per step keeps earlier observations valid (rollout storage keeps them)
        self._rgb = torch.empty((self.num_envs, self.nrows, self.ncols, 3), dtype=torch.uint8 if u8 else torch.float32,
                                device=self.device)
        return self._rgb

    def _launch_step(self, actions_dev, inject=None, auto_reset=None, render=False):
        flags = self._flags
        if self.auto_reset if auto_reset is None else auto_reset:
            flags |= _lib.FLAG_AUTO_RESET
        if render:
            flags |= _lib.FLAG_RENDER
            c = self._out.cstruct()
            c.rgb = self._new_rgb().data_ptr()
            c.rgb_u8 = 1 if self.obs_mode == "rgb_u8" else 0
        if self.balance_every and self._state.work is not None:
            if self._state.order is None:
                self._state.enable_balancing()
                self._version_structs += 1
            self._steps_since_balance += 1
            if self._steps_since_balance >= self.balance_every:
                self._state.rebalance()
                self._steps_since_balance = 0
                self.kernel_launches += 1
        self.kernel_launches += self._kernels_per_step(bool(flags & _lib.FLAG_AUTO_RESET))
        if inject is None:
            # hot path: every struct pointer is cached; only the action pointer and the stream vary
            fa = self._fast_args
            if fa is None or fa[0] != self._version_structs:
                fa = self._fast_args = (self._version_structs, load().gca_env_step, C.byref(self._params),
                                        C.byref(self._state.cstruct()), C.byref(self._out.cstruct()),
                                        C.byref(self._snapshot.cstruct()), ptr(self._snap_reward))
            if not (actions_dev.is_cuda and actions_dev.dtype == torch.int32 and actions_dev.is_contiguous()):
                raise _lib.GcaError("actions must be a contiguous int32 CUDA tensor of shape (N, 3)")
            rc = fa[1](fa[2], fa[3], actions_dev.data_ptr(), fa[4], None, fa[5], fa[6], flags,
                       torch.cuda.current_stream().cuda_stream)
            if rc:
                check(rc, "gca_env_step")
            self._version += 1
            return None
        inj, keep = make_inject(inject, self.device)
        check(load().gca_env_step(C.byref(self._params), C.byref(self._state.cstruct()), ptr(actions_dev),
                                  C.byref(self._out.cstruct()), None if inj is None else C.byref(inj),
                                  C.byref(self._snapshot.cstruct()), ptr(self._snap_reward), flags,
                                  current_stream()), "gca_env_step")
        self._version += 1
        return keep

    def stateless_step(self, action, obs=None, info=None, *, inject=None):
        """stateless_step (advanced_bulldozer.py:332-399).  ``action``: (N,3) [move, shoot, extension id]."""
        if self._state is None:
            raise RuntimeError("call reset() first")
        a = action if torch.is_tensor(action) else torch.as_tensor(np.asarray(action))
        if (a.is_cuda and a.dtype == torch.int32 and a.dim() == 2 and a.shape[0] == self.num_envs and a.shape[1] == 3
                and a.is_contiguous() and a.device == self._actions.device):
            acts = a  # already what the kernel reads: no staging copy
        else:
            self._actions.copy_(a.to(self.device, non_blocking=True).reshape(self.num_envs, -1)[:, :3])
            acts = self._actions
        fused = self._can_fuse_render()
        self._launch_step(acts, inject, render=fused)
        st, out = self._state, self._out
        rgb = self._rgb if fused else None
        if self.obs_mode != "none" and not fused:
            rgb = self._render(st.cell, st.doused, st.position, out.obs_night, None, actions=acts)
        # two snapshot copies (per-env scalars, step outputs); everything handed out is a view of them
        sc = PackedState.carve_scalars(st._scalars.clone(), self.num_envs)
        oc = StepOutputs.carve(out._base.clone(), self.num_envs)
        reward = oc["reward"]
        terminated = oc["terminated"].bool()
        truncated = self._all_false()
        terminated_out = truncated if self.auto_reset else terminated
        return (rgb, self._context_view(sc)), reward, terminated_out, truncated, self._info(terminated, sc, oc)

    def conditional_reset(self, step_tuple, action, *, seed=None, options=None):
        """conditional_reset (advanced_bulldozer.py:422-518): restore terminated envs from the
        initial snapshot, re-render their observation, zero their info counters, recompute reward and
        clear ``terminated`` -- all on the device, no host round trip."""
        obs, reward, terminated, truncated, info = step_tuple
        st = self._state
        mask = terminated.to(self.device).to(torch.uint8).contiguous()
        rgb = obs[0]
        if self.obs_mode != "none" and rgb is not None:
            a = action if torch.is_tensor(action) else torch.as_tensor(np.asarray(action))
            ext = a.to(self.device).reshape(self.num_envs, -1)[:, 2].to(torch.int32).contiguous()
            night = st.is_night.to(torch.uint8)
            self._rgb = rgb
            # restored grid/position, but the not-yet-restored dousing marks and day/night (:462-487)
            rgb = self._render(self._snapshot.cell, st.doused, self._snapshot.position, night, ext, env_mask=mask)
        new_reward = reward.to(self.device).clone().contiguous()
        check(load().gca_conditional_reset(C.byref(self._params), C.byref(st.cstruct()),
                                           C.byref(self._snapshot.cstruct()), ptr(self._snap_reward),
                                           ptr(new_reward), ptr(mask), current_stream()), "gca_conditional_reset")
        self._version += 1
        sc = PackedState.carve_scalars(st._scalars.clone(), self.num_envs)
        ninfo = dict(info)
        ninfo["steps_elapsed"] = sc["steps_elapsed"]
        ninfo["reward_accumulated"] = sc["reward_accumulated"]
        return (rgb, self._context_view(sc)), new_reward, self._all_false(), truncated, ninfo

    # ------------------------------------------------------------------------------------------
    # stateful / fast paths
    # ------------------------------------------------------------------------------------------
    def step(self, action):
        """Stateful gymnasium-style step over the whole batch (no implicit reset)."""
        return self.stateless_step(action)

    def step_device(self, actions_dev: torch.Tensor, auto_reset: Optional[bool] = None) -> StepOutputs:
        """Hot path: one fused launch on the current stream; actions and results stay in HBM."""
        self._launch_step(actions_dev, None, auto_reset)
        return self._out

    def step_observe_device(self, actions_dev: torch.Tensor, auto_reset: Optional[bool] = None):
        """``step_device`` + the step's RGB observation.  64x64 grids without extensions: ONE launch (the step kernel's
        epilogue draws the frame from the bit-boards it holds, GCA_FLAG_RENDER); otherwise the step followed by
        ``gca_render_rgb``.  Returns (outputs, rgb)."""
        if self._can_fuse_render():
            self._launch_step(actions_dev, None, auto_reset, render=True)
            return self._out, self._rgb
        self._launch_step(actions_dev, None, auto_reset)
        return self._out, self.observe_device(actions_dev)

    def observe_device(self, actions_dev: torch.Tensor) -> Optional[torch.Tensor]:
        """The RGB observation ``stateless_step`` would return for the step just made with ``step_device(actions_dev)``
        (new grid / position, pre-step dousing marks and day/night, extension channel of the actions); one launch
        of ``gca_render_rgb`` into a fresh (N,H,W,3) buffer of the env's ``obs_mode`` (None for ``"none"``)."""
        st, out = self._state, self._out
        return self._render(st.cell, st.doused, st.position, out.obs_night, None, actions=actions_dev)

    def host_result_buffers(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Pinned host (reward, terminated) tensors laid out like the device outputs, for ``step_host``."""
        N = self.num_envs
        buf = torch.empty(5 * N, dtype=torch.uint8).pin_memory()
        return buf[:4 * N].view(torch.float32), buf[4 * N:]

    def step_host(self, actions_host: torch.Tensor, reward_host: torch.Tensor, terminated_host: torch.Tensor,
                  staged: bool = False, wait: bool = True) -> None:
        """The step for a CPU-side rollout loop, one C call (``gca_env_step_host``): ``actions_host`` (N,3)
        int32, ``reward_host`` (N,) float32 and ``terminated_host`` (N,) uint8 are HOST tensors.  When all
        three are pinned (``pin_memory()``) and the grid is 64x64 the step kernel reads the actions and stores
        reward / terminated over the bus itself (zero-copy: one launch, no copy engine) and the env that finishes last
        stores a completion word the host polls; otherwise -- or with ``staged=True`` -- actions are copied in, the
        fused step runs, reward / terminated are copied out and the stream is synchronised (``staged`` may also be
        ``FLAG_HOST_COPY_IN`` or ``FLAG_HOST_COPY_OUT`` to stage one direction only).
        ``wait=True``: the results are valid on return.  ``wait=False`` returns right after the launch; call
        ``step_host_wait()`` before reading the buffers -- a rollout loop that owns two env groups (two env objects on
        two CUDA streams) handles the results of one while the other steps, EnvPool style.  State stays on the device."""
        N = self.num_envs
        if (actions_host.is_cuda or actions_host.dtype != torch.int32 or actions_host.numel() != 3 * N
                or not actions_host.is_contiguous()):
            raise _lib.GcaError(f"actions_host: expected a contiguous host tensor of {3 * N} x int32")
        ha = self._host_args
        if ha is None or ha[0] != (reward_host.data_ptr(), terminated_host.data_ptr(), self._version_structs):
            for t, dt, name in ((reward_host, torch.float32, "reward_host"), (terminated_host, torch.uint8, "terminated_host")):
                if t.is_cuda or t.dtype != dt or t.numel() != N or not t.is_contiguous():
                    raise _lib.GcaError(f"{name}: expected a contiguous host tensor of {N} x {dt}")
            if self._host_act_dev is None:
                self._host_act_dev = torch.empty((N, 3), dtype=torch.int32, device=self.device)
            # every pointer that does not change from step to step, bound once
            ha = self._host_args = (
                (reward_host.data_ptr(), terminated_host.data_ptr(), self._version_structs), load().gca_env_step_host,
                C.byref(self._params), C.byref(self._state.cstruct()), self._host_act_dev.data_ptr(),
                C.byref(self._out.cstruct()), C.byref(self._snapshot.cstruct()), ptr(self._snap_reward),
                reward_host.data_ptr(), terminated_host.data_ptr(),
                reward_host.is_pinned() and terminated_host.is_pinned() and self._state.work is not None)
        flags = self._flags | (_lib.FLAG_AUTO_RESET if self.auto_reset else 0) | (_lib.FLAG_HOST_COPY if staged is True else int(staged))
        if not wait:
            flags |= _lib.FLAG_HOST_ASYNC
        if ha[10] and actions_host.is_pinned():
            flags |= _lib.FLAG_HOST_MAPPED  # torch pins with cudaHostAlloc: no per-call pointer queries in the C layer
        if self.balance_every and self._state.work is not None:
            if self._state.order is None:
                self._state.enable_balancing()
                self._version_structs += 1
                return self.step_host(actions_host, reward_host, terminated_host, staged, wait)  # re-bind the cached pointers
            self._steps_since_balance += 1
        # the re-dealing of envs to CTA slots is launched AFTER the step (it reads the costs the step leaves and only the
        # next step needs its result): the host already has this step's results while it runs
        rebalance = (self.balance_every and self._state.work is not None
                     and self._steps_since_balance >= self.balance_every)
        self.kernel_launches += self._kernels_per_step(self.auto_reset)
        self._out.next_token()
        stream = torch.cuda.current_stream().cuda_stream
        rc = ha[1](ha[2], ha[3], actions_host.data_ptr(), ha[4], ha[5], ha[6], ha[7],
                   flags | (_lib.FLAG_HOST_ASYNC if rebalance else 0), ha[8], ha[9], stream)
        if rc:
            check(rc, "gca_env_step_host")
        self._version += 1
        # what a wait has to wait for: the completion word (zero-copy out on the 64x64 path) or the stream
        word = (self._state.work is not None and not (flags & _lib.FLAG_HOST_COPY_OUT) and reward_host.is_pinned()
                and terminated_host.is_pinned())
        self._host_pending = (self._out.host_done.data_ptr() if word else None, self._out.token, stream)
        if rebalance:
            self._state.rebalance()
            self._steps_since_balance = 0
            self.kernel_launches += 1
            if wait:
                self.step_host_wait()
        if wait:
            self._host_pending = None

    def step_host_wait(self, timeout_s: float = 30.0) -> None:
        """Block until the ``step_host(..., wait=False)`` in flight has delivered its results to the host buffers."""
        pend = getattr(self, "_host_pending", None)
        if pend is None:
            return
        self._host_pending = None
        check(load().gca_host_wait(pend[0], pend[1], float(timeout_s), pend[2]), "gca_host_wait")

    # ------------------------------------------------------------------------------------------
    # reference helpers
    # ------------------------------------------------------------------------------------------
    def count_cells(self, grid=None):
        g = self._state.cell if grid is None else torch.as_tensor(grid)
        return {v: (g == v).sum() for v in (self._empty, self._tree, self._fire)}

    def _award(self, grid=None):
        c = self.count_cells(grid)
        t, f = c[self._tree], c[self._fire]
        return -(f.float() / ((t + f).float() + torch.tensor(1e-8, dtype=torch.float32, device=f.device)))

    def _is_done(self, grid=None):
        g = self._state.cell if grid is None else torch.as_tensor(grid)
        return ~(g == self._fire).any()

    def _report(self):
        return {"hit": False}

    def stats(self):
        return None if self._out.stats is None else self._out.stats.cpu().numpy().copy()


class _HostLazy:
    """Static host array uploaded to the device on first use (slope alone is 36 B/cell)."""

    def __init__(self, arr, device):
        self._arr, self._device, self._t = arr, device, None

    def tensor(self):
        if self._t is None:
            self._t = torch.as_tensor(self._arr, device=self._device)
        return self._t

    def __array__(self, dtype=None, copy=None):
        a = self._arr.detach().cpu().numpy() if torch.is_tensor(self._arr) else self._arr
        return np.asarray(a, dtype=dtype)

    def to(self, *args, **kwargs):
        return self.tensor().to(*args, **kwargs)

    def cpu(self):
        return self.tensor().cpu()

    @property
    def dtype(self):
        return self.tensor().dtype

    @property
    def shape(self):
        return self._arr.shape

    def __getitem__(self, i):
        return self.tensor()[i]


class MDP(Operator):
    """Top operation (reference advanced_bulldozer.py:956-1133): RepeatCA -> MoveModify -> observation,
    batched.  ``update(grid, action, per_env_context, shared_context, position, time)`` returns
    ``((rgb, grid, None), (per_env_context, position, time))`` from ONE fused kernel launch."""
    grid_dependant = True
    action_dependant = True
    context_dependant = True
    deterministic = False

    def __init__(self, env, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.env = env
        self.repeat_ca = env.repeater
        self.move_modify = env.move_modify
        self.suboperators = (self.repeat_ca, self.move_modify)
        self.should_transform_grid = env.enable_extensions
        self.enable_extensions = env.enable_extensions
        self.tree, self.fire, self.empty = env._tree, env._fire, env._empty

    def update(self, grid, action, per_env_context, shared_context, position, time):
        env = self.env
        ctx = materialise_context(per_env_context)
        ctx["true_grid"] = grid
        env.set_state(ctx, position, time)
        a = action if torch.is_tensor(action) else torch.as_tensor(np.asarray(action))
        a = a.to(env.device).reshape(env.num_envs, -1)
        if a.shape[1] == 4:  # full action with the 2-bit extension vector -> extension id
            ext_id = a[:, 2] * 1 + a[:, 3] * 2
            a = torch.stack([a[:, 0], a[:, 1], ext_id], dim=1)
        obs, _reward, _term, _trunc, _info = env.stateless_step(a)
        rgb, context = obs
        pe = context["per_env_context"]
        new_grid = pe["true_grid"]
        _ = pe["fire_age"], pe["dousing_count"]
        return (rgb, new_grid, None), (pe, context["position"], context["time"])

    def grid_to_rgb(self, display_grid, per_env_context, position):
        """MDP.grid_to_rgb (:1035-1101) for a batch: display grid (N,H,W) values 0/1/2."""
        env = self.env
        d = env.device
        g = torch.as_tensor(np.asarray(display_grid) if not torch.is_tensor(display_grid) else display_grid)
        cell = g.to(d).to(torch.uint8).contiguous()
        N = cell.shape[0]
        dc = per_env_context["dousing_count"]
        dc = torch.as_tensor(np.asarray(dc) if not torch.is_tensor(dc) else dc).to(d)
        H, W = cell.shape[1], cell.shape[2]
        WW = (W + 63) // 64
        bits = torch.zeros((N, H, WW * 64), dtype=torch.int64, device=d)
        bits[:, :, :W] = (dc != 0).long()
        sh = torch.arange(64, device=d, dtype=torch.int64)
        doused = (bits.reshape(N, H, WW, 64) << sh).sum(-1).contiguous()
        night = torch.as_tensor(np.asarray(per_env_context["is_night"]) if not torch.is_tensor(
            per_env_context["is_night"]) else per_env_context["is_night"]).to(d).to(torch.uint8).contiguous()
        pos = torch.as_tensor(np.asarray(position) if not torch.is_tensor(position) else position).to(d).to(
            torch.int32).contiguous()
        out = torch.empty((N, H, W, 3), dtype=torch.float32, device=d)
        check(load().gca_render_rgb(C.byref(env._params), N, ptr(cell), ptr(doused), ptr(pos), ptr(night), None, None,
                                    0, 0, None, ptr(out), current_stream()), "gca_render_rgb")
        return out


BatchedAdvancedBulldozerEnv = AdvancedForestFireBulldozerEnv

"""Device-resident packed state of a batch of environments + the constant block, as the C ABI
(include/gca.h) sees them.  Torch tensors own the memory; libgca only gets raw pointers."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import GcaInject, GcaParams, GcaState, GcaStepOut, check, current_stream, load, ptr

RNG_MODES = {"legacy": _lib.RNG_LEGACY, "partitionable": _lib.RNG_PARTITIONABLE, 0: 0, 1: 1}


def make_params(nrows: int, ncols: int, K: int = 1, speed_move: float = 0.12, speed_act: float = 0.03,
                t_any: float = 0.001, t_move: Optional[float] = None, t_shoot: Optional[float] = None,
                p_tree: float = 0.0, p_wind_change: float = 0.06, rng_mode="legacy",
                winds: Optional[np.ndarray] = None) -> GcaParams:
    """gca_params_init: constants of PartiallyObservableForestFireJax.__init__
    (reference forest_fire/operators/ca_alexandridis_jax.py:54-160) and the clock mapping
    (forest_fire/bulldozer/advanced_bulldozer.py:238-246)."""
    p = GcaParams()
    w = None
    if winds is not None:
        w = np.ascontiguousarray(np.asarray(winds, dtype=np.float32).reshape(-1))
        assert w.size == 72, "winds must be 8 wind matrices of 3x3"
    rc = load().gca_params_init(
        C.byref(p), nrows, ncols, K, float(speed_move), float(speed_act), float(t_any),
        -1.0 if t_move is None else float(t_move), -1.0 if t_shoot is None else float(t_shoot),
        float(p_tree), float(p_wind_change), RNG_MODES[rng_mode], None if w is None else w.ctypes.data)
    check(rc, "gca_params_init")
    return p


class StepOutputs:
    """Per-step device outputs (gca_step_out)."""

    def __init__(self, N: int, device, with_stats: bool = False):
        # reward and terminated share one allocation ([N] f32 followed by [N] u8) so that a host caller
        # gets both with a single device-to-host copy (gca_env_step_host)
        # ... and step_reward sits in front of them, so that one copy snapshots all three (carve())
        self.N = N
        self._base = torch.zeros(9 * N, dtype=torch.uint8, device=device)
        v = self.carve(self._base, N)
        self.step_reward, self.reward, self.terminated = v["step_reward"], v["reward"], v["terminated"]
        self.counts = torch.zeros((N, 2), dtype=torch.int32, device=device)
        self.obs_night = torch.zeros(N, dtype=torch.uint8, device=device)
        self.stats = torch.zeros(8, dtype=torch.int64, device=device) if with_stats else None
        # completion word of the host step (gca_env_step_host / gca_host_wait): one word of pinned (mapped) host memory the
        # step kernel stores the step's token to, and the device word the launch counts its envs in
        self.done_counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.host_done = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.token = 0
        self._c = GcaStepOut(ptr(self.reward).value, ptr(self.step_reward).value, ptr(self.terminated).value,
                             ptr(self.counts).value, ptr(self.obs_night).value,
                             None if self.stats is None else ptr(self.stats).value, None, None,
                             self.host_done.data_ptr(), ptr(self.done_counter).value, 0, 0, None)

    def next_token(self) -> int:
        """A fresh non-zero token for the next host step; written into the struct the C call reads."""
        self.token = (self.token % 0x7FFFFFFF) + 1
        self._c.done_token = self.token
        return self.token

    def cstruct(self) -> GcaStepOut:
        return self._c

    @staticmethod
    def carve(base: torch.Tensor, N: int) -> dict:
        """Views of a (9 N,) uint8 buffer: step_reward (N,) float32 | reward (N,) float32 | terminated (N,) uint8."""
        f = base.narrow(0, 0, 8 * N).view(torch.float32)
        return {"step_reward": f.narrow(0, 0, N), "reward": f.narrow(0, N, N), "terminated": base.narrow(0, 8 * N, N)}


class PackedState:
    """cell u8 / death u16 / hidden u8 / doused bit-board / slope-factor table + per-env scalars."""

    def __init__(self, N: int, H: int, W: int, device, use_hidden: bool = True, with_pslope: bool = True):
        self.N, self.H, self.W = N, H, W
        self.WW = (W + 63) // 64
        self.device = torch.device(device)
        d = self.device
        self.cell = torch.zeros((N, H, W), dtype=torch.uint8, device=d)
        self.death = torch.zeros((N, H, W), dtype=torch.uint16, device=d)
        self.hidden = torch.zeros((N, H, W), dtype=torch.uint8, device=d) if use_hidden else None
        self.doused = torch.zeros((N, H, self.WW), dtype=torch.int64, device=d)
        self.pslope = (torch.ones((N, H, W, 8), dtype=torch.float32, device=d)
                       if (use_hidden and with_pslope) else None)
        self.row_min = torch.full((N, H), -1, dtype=torch.int32, device=d)  # 0xFFFFFFFF
        self.tick = torch.zeros(N, dtype=torch.int32, device=d)
        # the per-env scalars live in ONE buffer, so that a rollout loop can snapshot them with one copy
        self._carve_scalars(torch.zeros(self.scalar_words(N), dtype=torch.int32, device=d))
        self.time_step.fill_(1)
        tiled = not (H == 64 and W == 64)
        self.scratch_cell = torch.zeros((N, H, W), dtype=torch.uint8, device=d) if tiled else None
        # per env the key schedule of up to 8 sub-steps (96 words) + tree / fire counts (2), 16 list counters, then per
        # 32x64 tile its burning-cell count and two slots of the active-tile lists (include/gca.h: scratch_u32)
        n_tiles = N * ((H + 31) // 32) * ((W + 63) // 64)
        self.scratch_u32 = torch.zeros(N * 98 + 16 + 3 * n_tiles, dtype=torch.int32, device=d) if tiled else None
        self.work = None if tiled else torch.zeros(N, dtype=torch.int32, device=d)
        self.order = None   # set by enable_balancing()
        # tree / fire bit-boards, the grid representation the 64x64 kernel reads (kept in step with `cell`)
        self.bb = None if tiled else torch.zeros((N, H, self.WW, 2), dtype=torch.int64, device=d)
        self._c = None

    SCALARS = ("wind_index", "is_night", "time_step", "position", "key", "time", "steps_elapsed", "reward_accumulated")

    @staticmethod
    def scalar_words(N: int) -> int:
        return 10 * ((N + 3) // 4 * 4)

    @staticmethod
    def carve_scalars(base: torch.Tensor, N: int) -> dict:
        """Views of a (scalar_words(N),) int32 buffer: wind_index, is_night, time_step (N,) int32, position (N,2) int32,
        key (N,2) uint32, time, steps_elapsed, reward_accumulated (N,) float32 -- each block 16-byte aligned."""
        M = (N + 3) // 4 * 4
        f = base.view(torch.float32)
        return {"wind_index": base.narrow(0, 0, N), "is_night": base.narrow(0, M, N), "time_step": base.narrow(0, 2 * M, N),
                "position": base.narrow(0, 3 * M, 2 * N).view(N, 2),
                "key": base.narrow(0, 5 * M, 2 * N).view(torch.uint32).view(N, 2),
                "time": f.narrow(0, 7 * M, N), "steps_elapsed": f.narrow(0, 8 * M, N),
                "reward_accumulated": f.narrow(0, 9 * M, N)}

    def _carve_scalars(self, base: torch.Tensor) -> None:
        self._scalars = base
        for k, v in self.carve_scalars(base, self.N).items():
            setattr(self, k, v)

    _FIELDS = ("cell", "death", "hidden", "doused", "pslope", "row_min", "tick", "key", "wind_index", "position",
               "time", "time_step", "is_night", "steps_elapsed", "reward_accumulated", "scratch_cell", "scratch_u32", "work", "order", "bb")

    def cstruct(self) -> GcaState:
        if self._c is None:
            s = GcaState()
            s.N = self.N
            for f in self._FIELDS:
                t = getattr(self, f)
                setattr(s, f, None if t is None else t.data_ptr())
            self._c = s
        return self._c

    def clone(self, share_static: bool = True) -> "PackedState":
        """Deep copy of the dynamic arrays; hidden / pslope are static and shared by default."""
        o = PackedState.__new__(PackedState)
        o.N, o.H, o.W, o.WW, o.device = self.N, self.H, self.W, self.WW, self.device
        o._carve_scalars(self._scalars.clone())
        for f in self._FIELDS:
            t = getattr(self, f)
            if f in self.SCALARS:
                continue
            if t is None:
                setattr(o, f, None)
            elif share_static and f in ("hidden", "pslope", "scratch_cell", "scratch_u32", "work", "order"):
                setattr(o, f, t)
            else:
                setattr(o, f, t.clone())
        o._c = None
        return o

    def copy_from(self, other: "PackedState") -> None:
        for f in self._FIELDS:
            t, s = getattr(self, f), getattr(other, f)
            if t is not None and s is not None and t.data_ptr() != s.data_ptr() and not f.startswith("scratch") and f not in ("work", "order"):
                t.copy_(s)

    def enable_balancing(self) -> None:
        """Allocate the warp-slot -> env permutation (identity until rebalance() is called)."""
        if self.order is None and self.work is not None:
            self.order = torch.arange(self.N, dtype=torch.int32, device=self.device)
            self._c = None

    def rebalance(self) -> None:
        if self.order is not None:
            check(load().gca_balance_order(self.N, ptr(self.work), ptr(self.order), current_stream()),
                  "gca_balance_order")

    # ---- reference layout <-> packed ------------------------------------------------------------
    def pack_from_reference(self, params: GcaParams, ctx: Dict[str, torch.Tensor], position=None, time=None,
                            check_inputs: bool = True) -> None:
        """ctx: reference per_env_context arrays (true_grid f32, fire_age f32, dousing_count i32,
        vegetation/density i32, [pslope f32 (N,H,W,3,3) or slope], wind_index, key, is_night, time_step)."""
        d = self.device

        def dev(x, dt):
            t = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x)
            return t.to(device=d, dtype=dt).contiguous()

        grid = dev(ctx["true_grid"], torch.float32)
        age = dev(ctx["fire_age"], torch.float32)
        dous = dev(ctx["dousing_count"], torch.int32)
        veg = den = None
        if self.hidden is not None:
            veg = dev(ctx["vegetation"], torch.int32)
            den = dev(ctx["density"], torch.int32)
            if self.pslope is not None:
                ps = None
                if "pslope" in ctx:
                    ps = dev(ctx["pslope"], torch.float32)
                elif "slope" in ctx:
                    # exp(f32(0.078) slope) is tabulated with NumPy on the host (the oracle's bits, whatever the libm)
                    from .forest_fire.bulldozer.utils.init_utils import p_slope_table
                    sl = ctx["slope"]
                    sl = sl.detach().cpu().numpy() if torch.is_tensor(sl) else np.asarray(sl)
                    ps = dev(p_slope_table(sl), torch.float32)
                # neither given: the table already packed stays (set_state of the env's own, unchanged slope)
                if ps is not None:
                    ps = ps.reshape(self.N, self.H, self.W, 9)
                    self.pslope.copy_(ps[..., [0, 1, 2, 3, 5, 6, 7, 8]])
        if "tick" in ctx:
            self.tick.copy_(dev(ctx["tick"], torch.int32))
        err = torch.zeros(1, dtype=torch.int32, device=d)
        check(load().gca_pack_state(C.byref(params), C.byref(self.cstruct()), ptr(grid), ptr(age), ptr(dous),
                                    ptr(veg), ptr(den), ptr(self.hidden), ptr(err), current_stream()),
              "gca_pack_state")
        if check_inputs:
            e = int(err.item())
            if e:
                raise _lib.GcaError(f"gca_pack_state: input outside the packed domain (flags {e:#x}): "
                                    "1 cell not in {0,1,2}; 2 fire cell age not an integer in [1,32767]; "
                                    "4 non-fire age not an integer in [0,65535]; 8 veg/den not in [0,7]; "
                                    "16 dousing_count not in {0,1}")
        self.wind_index.copy_(dev(ctx["wind_index"], torch.int32))
        k = ctx["key"]
        k = torch.as_tensor(np.asarray(k).astype(np.uint32)) if not torch.is_tensor(k) else k
        self.key.copy_(k.to(d).reshape(self.N, 2))
        self.is_night.copy_(dev(ctx["is_night"], torch.int32))
        self.time_step.copy_(dev(ctx["time_step"], torch.int32))
        if position is not None:
            self.position.copy_(dev(position, torch.int32))
        if time is not None:
            self.time.copy_(dev(time, torch.float32))

    def unpack_to_reference(self, params: GcaParams, want=("true_grid", "fire_age", "dousing_count")):
        """float32 / int32 arrays in the reference's context layout (device tensors)."""
        d = self.device
        shape = (self.N, self.H, self.W)
        grid = torch.empty(shape, dtype=torch.float32, device=d) if "true_grid" in want else None
        age = torch.empty(shape, dtype=torch.float32, device=d) if "fire_age" in want else None
        dous = torch.empty(shape, dtype=torch.int32, device=d) if "dousing_count" in want else None
        check(load().gca_unpack_state(C.byref(params), C.byref(self.cstruct()), ptr(grid), ptr(age), ptr(dous),
                                      current_stream()), "gca_unpack_state")
        out = {}
        if grid is not None:
            out["true_grid"] = grid
        if age is not None:
            out["fire_age"] = age
        if dous is not None:
            out["dousing_count"] = dous
        return out


def make_inject(inject: Optional[dict], device) -> (Optional[GcaInject], list):
    """dict of arrays (leading K axis) -> gca_inject; returns the struct and the tensors to keep alive."""
    if not inject:
        return None, []
    keep = []
    s = GcaInject()
    for name, dt in (("u_burn", torch.float32), ("u_grow", torch.float32), ("age_new", torch.int32),
                     ("u_wind", torch.float32), ("wind_step", torch.int32)):
        a = inject.get(name)
        if a is not None:
            t = (a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a))).to(device=device, dtype=dt).contiguous()
            keep.append(t)
            setattr(s, name, t.data_ptr())
    return s, keep

"""MoveCUDA / ModifyCUDA / MoveModifyCUDA -- bulldozer movement and dousing
(reference forest_fire/operators/move_modify_jax.py:11-157), batched.

Move: 9-way move clamped at the borders.  Modify: ``shoot`` marks ``dousing_count[row, col] = 1``
at the (already moved) position; the grid itself is never modified in the advanced env
(SURVEY.md F4).  Both go through gca_move_modify on the packed dousing bit-board."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Set

import numpy as np
import torch

from ... import _lib
from ..._lib import check, current_stream, load, ptr
from ...operator import Operator
from ...packed import make_params
from ... import spaces


def _dev(x, device, dtype):
    t = x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))
    return t.to(device=device, dtype=dtype).contiguous()


def _run_move_modify(params, H, W, position, a0, a1, dousing=None, device="cuda"):
    """positions (N,2), actions (N,), optional dousing_count (N,H,W) -> new positions, dousing."""
    d = torch.device(device)
    pos = _dev(position, d, torch.int32).reshape(-1, 2).clone()
    N = pos.shape[0]
    WW = (W + 63) // 64
    if dousing is None:
        doused = torch.zeros((N, H, WW), dtype=torch.int64, device=d)
    else:
        dc = _dev(dousing, d, torch.int64).reshape(N, H, W)
        bits = torch.zeros((N, H, WW * 64), dtype=torch.int64, device=d)
        bits[:, :, :W] = (dc != 0).long()
        doused = (bits.reshape(N, H, WW, 64) << torch.arange(64, device=d)).sum(-1).contiguous()
    act = torch.zeros((N, 3), dtype=torch.int32, device=d)
    act[:, 0] = _dev(a0, d, torch.int32).reshape(-1)
    act[:, 1] = _dev(a1, d, torch.int32).reshape(-1)
    st = _lib.GcaState()
    st.N = N
    st.position = pos.data_ptr()
    st.doused = doused.data_ptr()
    check(load().gca_move_modify(C.byref(params), C.byref(st), ptr(act), current_stream()), "gca_move_modify")
    new_dousing = None
    if dousing is not None:
        sh = torch.arange(64, device=d)
        new_dousing = ((doused[..., None] >> sh) & 1).reshape(N, H, WW * 64)[:, :, :W].to(torch.int32)
    return pos, new_dousing


class MoveCUDA(Operator):
    grid_dependant = False
    action_dependant = True
    context_dependant = True
    deterministic = True

    def __init__(self, directions_sets: Dict[str, Set], *args, params=None, device="cuda", **kwargs):
        super().__init__(*args, **kwargs)
        self.up_set, self.down_set = directions_sets["up"], directions_sets["down"]
        self.left_set, self.right_set = directions_sets["left"], directions_sets["right"]
        self.not_move_set = directions_sets["not_move"]
        self.movement_set = self.up_set | self.down_set | self.left_set | self.right_set | self.not_move_set
        expected = {"up": {0, 1, 2}, "down": {6, 7, 8}, "left": {0, 3, 6}, "right": {2, 5, 8}}
        for k, v in expected.items():
            if set(directions_sets[k]) != v:
                raise _lib.GcaError("MoveCUDA implements the reference's 3x3 keypad action layout only")
        self._params = params
        self.device = device

    def update(self, grid, action, context):
        H, W = grid.shape[-2], grid.shape[-1]
        P = self._params if self._params is not None else make_params(max(H, 8), max(W, 8))
        if P.H != H or P.W != W:
            P = _lib.GcaParams.from_buffer_copy(P)
            P.H, P.W = H, W
        pos = torch.as_tensor(np.asarray(context) if not torch.is_tensor(context) else context)
        single = pos.dim() == 1
        a0 = torch.as_tensor(np.asarray(action) if not torch.is_tensor(action) else action).reshape(-1)
        new_pos, _ = _run_move_modify(P, H, W, pos.reshape(-1, 2), a0, torch.zeros_like(a0), None, self.device)
        return grid, (new_pos[0] if single else new_pos)


class ModifyCUDA(Operator):
    hit = False
    grid_dependant = True
    action_dependant = True
    context_dependant = True
    deterministic = True

    def __init__(self, effects: dict, *args, params=None, device="cuda", **kwargs):
        super().__init__(*args, **kwargs)
        self.effects = effects
        self._params = params
        self.device = device

    def update(self, grid, action, context, per_env_context):
        H, W = grid.shape[-2], grid.shape[-1]
        P = self._params if self._params is not None else make_params(max(H, 8), max(W, 8))
        if P.H != H or P.W != W:
            P = _lib.GcaParams.from_buffer_copy(P)
            P.H, P.W = H, W
        pos = torch.as_tensor(np.asarray(context) if not torch.is_tensor(context) else context)
        single = pos.dim() == 1
        a1 = torch.as_tensor(np.asarray(action) if not torch.is_tensor(action) else action).reshape(-1)
        dc = per_env_context["dousing_count"]
        dc = torch.as_tensor(np.asarray(dc) if not torch.is_tensor(dc) else dc)
        stay = torch.full_like(a1, 4)
        _, new_dc = _run_move_modify(P, H, W, pos.reshape(-1, 2), stay, a1, dc.reshape(-1, H, W), self.device)
        per_env_context["dousing_count"] = new_dc[0] if single else new_dc  # mutated in place, as the reference does
        return grid, context, per_env_context


class MoveModifyCUDA(Operator):
    grid_dependant = True
    action_dependant = True
    context_dependant = True
    deterministic = True

    def __init__(self, move, modify, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.suboperators = move, modify
        self.move, self.modify = move, modify
        if self.action_space is None and self.move.action_space is not None:
            self.action_space = spaces.Tuple((self.move.action_space, self.move.action_space))
        if self.context_space is None and self.move.context_space is not None \
                and self.modify.context_space is not None:
            self.context_space = self.move.context_space

    def update(self, grid, subactions, position, per_env_context):
        grid, position = self.move(grid, subactions[0], position)
        grid, position, per_env_context = self.modify(grid, subactions[1], position, per_env_context)
        return grid, position, per_env_context

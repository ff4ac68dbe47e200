"""WindyForestFireCUDA -- the v3 fire CA (reference forest_fire/operators/ca_windy.py:11-173), batched.

One 3x3 uniform roll per CA update decides which of the 8 directions fail (``wind <= roll``); a
tree with a burning neighbour in a direction that did not fail ignites, fire burns out after one
update, empty stays empty.  ``update(grid, action, wind, roll=None)``: grid (N,H,W) or (H,W) with
the reference's cell values 0 / 3 / 25; ``roll`` (N,3,3) may be injected (rule parity), otherwise
it is drawn from this operator's seeded ``np_random`` (the reference draws from an unseeded Box).
"""
from __future__ import annotations

import numpy as np
import torch

from ... import spaces
from ..._config import TYPE_BOX
from ..._lib import check, current_stream, load, ptr
from ...operator import Operator

VALUES = (0, 3, 25)  # empty, tree, fire (reference bulldozer/bulldozer.py:85-87)


def to_codes(grid, device, empty=0, tree=3, fire=25):
    g = grid if torch.is_tensor(grid) else torch.as_tensor(np.asarray(grid))
    g = g.to(device)
    return ((g == tree).to(torch.uint8) + 2 * (g == fire).to(torch.uint8)).contiguous()


def from_codes(codes, empty=0, tree=3, fire=25):
    lut = torch.tensor([empty, tree, fire], dtype=torch.int64, device=codes.device)
    return lut[codes.long()]


class WindyState:
    """tree / fire bit-boards + position + float64 clock of N envs (device)."""

    def __init__(self, N, H, W, device):
        self.N, self.H, self.W, self.WW = N, H, W, (W + 63) // 64
        self.device = torch.device(device)
        self.tree = torch.zeros((N, H, self.WW), dtype=torch.int64, device=self.device)
        self.fire = torch.zeros((N, H, self.WW), dtype=torch.int64, device=self.device)
        self.position = torch.zeros((N, 2), dtype=torch.int32, device=self.device)
        self.time = torch.zeros(N, dtype=torch.float64, device=self.device)

    def pack(self, codes_u8):
        check(load().gca_windy_pack(self.N, self.H, self.W, ptr(codes_u8), ptr(self.tree), ptr(self.fire),
                                    current_stream()), "gca_windy_pack")

    def unpack(self):
        out = torch.empty((self.N, self.H, self.W), dtype=torch.uint8, device=self.device)
        check(load().gca_windy_unpack(self.N, self.H, self.W, ptr(self.tree), ptr(self.fire), ptr(out),
                                      current_stream()), "gca_windy_unpack")
        return out


class WindyForestFireCUDA(Operator):
    grid_dependant = True
    action_dependant = False
    context_dependant = True
    deterministic = False
    _identity = 2 ** 11
    _propagation = 2 ** 3

    def __init__(self, empty=0, tree=3, fire=25, *args, device="cuda", **kwargs):
        super().__init__(*args, **kwargs)
        self._empty, self._tree, self._fire = empty, tree, fire
        assert empty < tree < fire
        self.device = torch.device(device)
        if self.context_space is None:
            self.context_space = spaces.Box(0.0, 1.0, shape=(3, 3), dtype=TYPE_BOX)

    def update(self, grid, action, wind, roll=None):
        g = grid if torch.is_tensor(grid) else torch.as_tensor(np.asarray(grid))
        single = g.dim() == 2
        if single:
            g = g[None]
        N, H, W = g.shape
        st = WindyState(N, H, W, self.device)
        st.pack(to_codes(g, self.device, self._empty, self._tree, self._fire))
        if roll is None:
            roll = self.np_random.random((N, 3, 3))
        rolls = torch.as_tensor(np.asarray(roll, dtype=np.float64).reshape(N, 1, 9)).to(self.device).contiguous()
        w = torch.as_tensor(np.asarray(wind, dtype=np.float64).reshape(9)).to(self.device)
        # exactly one CA update: clock pre-loaded with 1.0 and zero action cost, "stay / no shoot" action
        st.time.fill_(1.0)
        st.position.zero_()
        act = torch.tensor([[4, 0]] * N, dtype=torch.int32, device=self.device)
        check(load().gca_windy_env_step(N, H, W, ptr(st.tree), ptr(st.fire), ptr(st.position), ptr(st.time), ptr(act),
                                        ptr(w), ptr(rolls), 1, 0.0, 0.0, 0.0, None, None, None, None,
                                        current_stream()), "gca_windy_env_step")
        out = from_codes(st.unpack(), self._empty, self._tree, self._fire)
        return (out[0] if single else out), wind

"""ForestFireBulldozerEnv -- the v3 env (registered ForestFireBulldozer256x256-v3; reference
forest_fire/bulldozer/bulldozer.py:21-400), batched over ``num_envs`` independent grids on one GPU.

Same gym surface as the reference's single env, with a leading env axis: ``reset() -> (obs, info)``,
``step(action (N,2)) -> (obs, reward, terminated, truncated, info)``; ``obs = (grid, context)`` with
grid values 0 / 3 / 25 and ``context = (wind (3,3), position (N,2), time (N,))``.  One kernel launch
per step (clock + int(repeats) WindyForestFire updates + move + cut + reward + done).  Rolls come from
the env's seeded generator (or are injected through ``step(action, rolls=...)``)."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from ... import _lib, spaces
from ..._config import TYPE_BOX, TYPE_INT
from ..._lib import check, current_stream, load, ptr
from ...ca_env import CAEnv
from ...grid_space import GridSpace
from ..operators.ca_windy_cuda import WindyForestFireCUDA, WindyState, from_codes, to_codes

DEFAULT_WIND = {"up_left": 0.48, "up": 0.64, "up_right": 0.98, "left": 0.12, "right": 0.64, "down_left": 0.06,
                "down": 0.12, "down_right": 0.48}


class ForestFireBulldozerEnv(CAEnv):
    metadata = {"render_modes": ["human"], "render_mode": "rgb_array"}

    @property
    def MDP(self):
        return self._mdp

    @property
    def initial_state(self):
        if self._initial is None:
            self._initial = self._sample_initial()
        return self._initial

    def __init__(self, nrows, ncols, speed_move=0.12, speed_act=0.03, pos_bull=None, pos_fire=None, t_move=None,
                 t_shoot=None, t_any=0.001, p_tree=0.90, p_empty=0.10, wind=DEFAULT_WIND, *, num_envs=1, device="cuda",
                 seed: Optional[int] = None, max_repeats=4, **kwargs):
        super().__init__(nrows, ncols, **kwargs)
        if not torch.cuda.is_available():
            raise _lib.GcaError("ForestFireBulldozerEnv needs a CUDA device; there is no CPU path")
        load()
        self.device = torch.device(device)
        self.num_envs = int(num_envs)
        self.np_random = np.random.default_rng(seed)
        self.title = "ForestFireBulldozer" + str(nrows) + "x" + str(ncols)
        self._empty, self._tree, self._fire = 0, 3, 25
        self._pos_bull, self._pos_fire = pos_bull, pos_fire
        self._p_tree, self._p_empty = p_tree, p_empty
        self._wind = np.array([[wind["up_left"], wind["up"], wind["up_right"]], [wind["left"], 0.0, wind["right"]],
                               [wind["down_left"], wind["down"], wind["down_right"]]], dtype=TYPE_BOX)
        assert ((self._wind >= 0) & (self._wind <= 1)).all(), "Bad Wind Data, check ranges [0.0, 1.0]"
        scale = (nrows + ncols) // 2
        self._t_env_any = t_any
        self._t_act_move = (1 / (speed_move * scale)) - t_any if t_move is None else t_move
        self._t_act_shoot = (1 / (speed_act * scale)) - self._t_act_move if t_shoot is None else t_shoot
        self.max_repeats = int(max_repeats)
        N = self.num_envs
        self.grid_space = GridSpace(values=[self._empty, self._tree, self._fire], shape=(N, nrows, ncols))
        self.action_space = spaces.MultiDiscrete(np.array([[9, 2]] * N), dtype=TYPE_INT)
        self.ca = WindyForestFireCUDA(self._empty, self._tree, self._fire, device=self.device)
        self._mdp = self.ca
        self._initial = None
        self._state = WindyState(N, nrows, ncols, self.device)
        self._wind_dev = torch.as_tensor(self._wind.reshape(9)).to(self.device)
        self._reward = torch.zeros(N, dtype=torch.float64, device=self.device)
        self._term = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self._counts = torch.zeros((N, 2), dtype=torch.int32, device=self.device)
        self._repeats = torch.zeros(N, dtype=torch.int32, device=self.device)

    def _noise(self, ax_len):
        upper = int(ax_len / 12)
        return int(self.np_random.integers(0, upper)) if upper > 0 else 0

    def _sample_initial(self):
        N, H, W = self.num_envs, self.nrows, self.ncols
        gs = GridSpace(values=[self._empty, self._tree, self._fire], probs=[self._p_empty, self._p_tree, 0.0],
                       shape=(N, H, W))
        gs._np_random = self.np_random
        grid = gs.sample()
        pos = np.zeros((N, 2), dtype=np.int32)
        for e in range(N):
            if self._pos_fire is None:
                r, c = 3 * H // 4 + self._noise(H), W // 4 + self._noise(W)
            else:
                r, c = self._pos_fire
            grid[e, r, c] = self._fire
            if self._pos_bull is None:
                pos[e] = (H // 4 + self._noise(H), 3 * W // 4 + self._noise(W))
            else:
                pos[e] = self._pos_bull
        return grid, (self._wind, pos, np.zeros(N, dtype=TYPE_BOX))

    def set_state(self, grid, position, time):
        self._state.pack(to_codes(grid, self.device, self._empty, self._tree, self._fire))
        self._state.position.copy_(torch.as_tensor(np.asarray(position, dtype=np.int32)).reshape(self.num_envs, 2))
        self._state.time.copy_(torch.as_tensor(np.asarray(time, dtype=np.float64)).reshape(self.num_envs))

    def _obs(self):
        grid = from_codes(self._state.unpack(), self._empty, self._tree, self._fire)
        return grid, (self._wind, self._state.position.clone(), self._state.time.clone())

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
            self._initial = None
        grid, (wind, pos, time) = self.initial_state
        self.set_state(grid, pos, time)
        self.steps_elapsed = 0
        return self._obs(), self._report()

    def step(self, action, rolls=None):
        N = self.num_envs
        a = action if torch.is_tensor(action) else torch.as_tensor(np.asarray(action))
        a = a.to(self.device).to(torch.int32).reshape(N, -1)[:, :2].contiguous()
        if rolls is None:
            rolls = self.np_random.random((N, self.max_repeats, 9))
        r = torch.as_tensor(np.asarray(rolls, dtype=np.float64).reshape(N, -1, 9)).to(self.device).contiguous()
        st = self._state
        check(load().gca_windy_env_step(N, self.nrows, self.ncols, ptr(st.tree), ptr(st.fire), ptr(st.position),
                                        ptr(st.time), ptr(a), ptr(self._wind_dev), ptr(r), r.shape[1],
                                        float(self._t_act_move), float(self._t_act_shoot), float(self._t_env_any),
                                        ptr(self._reward), ptr(self._term), ptr(self._counts), ptr(self._repeats),
                                        current_stream()), "gca_windy_env_step")
        self.steps_elapsed += 1
        truncated = torch.zeros(N, dtype=torch.bool, device=self.device)
        return self._obs(), self._reward.clone(), self._term.bool(), truncated, self._report()

    def _award(self):
        return self._reward

    def _is_done(self):
        return self._term.bool()

    def _report(self):
        return {"repeats": self._repeats.clone()}

"""ORACLE (test infrastructure, not product code) -- the v3 rule set of the registered
ForestFireBulldozer256x256-v3 env: WindyForestFire + NumPy Move / Modify / RepeatCA + MDP.

Follows /root/reference/gym_cellular_automata/forest_fire/operators/ca_windy.py:11-173,
operators/move_modify.py:9-134, operators/repeat_ca.py:10-45 and bulldozer/bulldozer.py:21-400.
Literal: the CA goes through scipy.signal.convolve2d and the three break points exactly like
the reference.  The reference draws its 3x3 roll from a freshly constructed gymnasium Box
(unseeded), so only RULE parity is possible: rolls are always passed in.
Parity status: pinned to the reference's own v3 operator source, executed through oracle/ref_shim's gymnasium
stand-in with recorded rolls (tests/golden/make_reference_golden.py run_v3,
tests/test_oracle.py::test_v3_oracle_reproduces_reference_source_golden).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.signal import convolve2d

EMPTY, TREE, FIRE = 0, 3, 25  # bulldozer.py:85-87
IDENTITY, PROPAGATION = 2 ** 11, 2 ** 3  # ca_windy.py:19-20
KEEP = IDENTITY * TREE  # ca_windy.py:84-100
PROPAGATE = IDENTITY * TREE + PROPAGATION * FIRE
CONSUME = IDENTITY * FIRE

DEFAULT_WIND = np.array([[0.48, 0.64, 0.98], [0.12, 0.0, 0.64], [0.06, 0.12, 0.48]], dtype=np.float64)  # bulldozer.py:46-55


def windy_update(grid: np.ndarray, wind: np.ndarray, roll: np.ndarray) -> np.ndarray:
    """WindyForestFire.update (ca_windy.py:41-51) with the uniform roll given."""
    failed = wind <= roll  # :64-65
    kernel = np.full((3, 3), PROPAGATION, dtype=np.int64)
    kernel[failed] = EMPTY
    kernel[1, 1] = IDENTITY  # :69-77
    signal = convolve2d(grid, kernel, mode="same", boundary="fill", fillvalue=EMPTY)  # :79-82
    new = np.full(grid.shape, EMPTY, dtype=grid.dtype)  # :102-139
    new[(signal >= KEEP) & (signal < PROPAGATE)] = TREE
    new[(signal >= PROPAGATE) & (signal < CONSUME)] = FIRE
    new[signal >= CONSUME] = EMPTY
    return new


class V3Constants:
    """Clock costs of ForestFireBulldozerEnv (bulldozer.py:126-136,277-298): moving costs t_move,
    not moving 0; shooting costs t_shoot, not shooting 0; every step t_any."""

    def __init__(self, nrows, ncols, speed_move=0.12, speed_act=0.03, t_any=0.001, t_move=None, t_shoot=None):
        scale = (nrows + ncols) // 2
        self.t_any = t_any
        self.t_move = (1 / (speed_move * scale)) - t_any if t_move is None else t_move
        self.t_shoot = (1 / (speed_act * scale)) - self.t_move if t_shoot is None else t_shoot
        self.nrows, self.ncols = nrows, ncols

    def time_per_action(self, move, shoot):
        return (0.0 if int(move) == 4 else self.t_move) + (0.0 if int(shoot) == 0 else self.t_shoot)


UP, DOWN, LEFT, RIGHT = {0, 1, 2}, {6, 7, 8}, {0, 3, 6}, {2, 5, 8}


def v3_env_step(C: V3Constants, grid, position, time, action, wind, rolls):
    """One CAEnv.step of the v3 env for ONE env (ca_env.py:27-48, bulldozer.py:393-400).
    rolls: (R, 3, 3) uniform rolls, consumed one per CA update.  Returns
    (grid, position, time, reward, terminated, repeats)."""
    move, shoot = int(action[0]), int(action[1])
    # RepeatCA.update (repeat_ca.py:32-45), float64 clock
    accu = float(time) + (C.time_per_action(move, shoot) + C.t_any)
    accu, repeats = math.modf(accu)
    g = grid
    for k in range(int(repeats)):
        g = windy_update(g, wind, rolls[k])
    # Move (move_modify.py:39-66)
    row, col = int(position[0]), int(position[1])
    H, W = g.shape
    if move in UP and row > 0:
        row -= 1
    if move in DOWN and row < H - 1:
        row += 1
    if move in LEFT and col > 0:
        col -= 1
    if move in RIGHT and col < W - 1:
        col += 1
    # Modify (move_modify.py:81-94): shoot cuts a tree (effects {tree: empty}, bulldozer.py:105)
    g = g.copy()
    if shoot and g[row, col] == TREE:
        g[row, col] = EMPTY
    t = int((g == TREE).sum())
    f = int((g == FIRE).sum())
    # _award (bulldozer.py:196-199) raises ZeroDivisionError when nothing is left; NaN here
    reward = -(f / (t + f)) if (t + f) > 0 else float("nan")
    return g, np.array([row, col]), accu, reward, f == 0, int(repeats)

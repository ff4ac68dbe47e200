"""ORACLE (test infrastructure, not product code) -- literal restatement of the host-side
initial-state generators of the advanced bulldozer env.

Follows /root/reference/gym_cellular_automata/forest_fire/bulldozer/utils/init_utils.py
(:10-116 hidden layers, :166-200 slope, :203-245 winds) and
forest_fire/bulldozer/advanced_bulldozer.py (:650-743 initial grid / context).  The reference
draws from the GLOBAL unseeded ``np.random``; here the same call sequence is made on an
explicit ``np.random.RandomState`` so a seeded run is reproducible (seeding the reference's
global generator with the same seed yields the same layers).  Written with plain loops on
purpose: the product's vectorised generators are checked against it.
"""
from __future__ import annotations

import math

import numpy as np

from . import prng
from .alexandridis import EMPTY, TREE, FIRE, F32, p_slope_table

WIND_THETAS = np.array([
    [[45, 0, 45], [90, 0, 90], [135, 180, 135]],      # N
    [[90, 45, 0], [135, 0, 45], [180, 135, 90]],      # NE
    [[135, 90, 45], [180, 0, 0], [135, 90, 45]],      # E
    [[180, 135, 90], [135, 0, 45], [90, 45, 0]],      # SE
    [[135, 180, 135], [90, 0, 90], [45, 0, 45]],      # S
    [[90, 135, 180], [45, 0, 135], [0, 45, 90]],      # SW
    [[45, 90, 135], [0, 0, 180], [45, 90, 135]],      # W
    [[0, 45, 90], [45, 0, 135], [90, 135, 180]],      # NW
], dtype=np.float64)  # init_utils.py:203-220


def get_winds() -> np.ndarray:
    """(8,2,3,3) float64: [k,0] = wind matrix with centre 0, [k,1] = ft (init_utils.py:225-245).
    The ``use_hidden`` branch of the reference is dead: the loop always walks wind_thetas."""
    out = np.zeros((8, 2, 3, 3))
    for k in range(8):
        t = np.radians(WIND_THETAS[k])
        ft = np.exp(10 * 0.131 * (np.cos(t) - 1))
        w = np.exp(0.045 * 10) * ft
        w[1, 1] = 0
        out[k, 0], out[k, 1] = w, ft
    return out


def _patch_layer(rng, rows, cols, num_envs):
    """init_vegetation / init_density share one recipe (init_utils.py:10-73)."""
    m = np.zeros((num_envs, rows, cols), dtype=int)
    for e in range(num_envs):
        for _ in range(rng.randint(4, 8)):
            cr = rng.randint(0, rows)
            cc = rng.randint(0, cols)
            ph = rng.randint(3, rows // 2)
            pw = rng.randint(3, cols // 2)
            kind = rng.randint(1, 6)
            m[e, max(0, cr - ph // 2):min(rows, cr + ph // 2),
              max(0, cc - pw // 2):min(cols, cc + pw // 2)] = kind
        zero = m[e] == 0
        m[e][zero] = rng.randint(1, 4, size=int(zero.sum()))
    return m


def init_vegetation(rng, rows, cols, num_envs):
    return _patch_layer(rng, rows, cols, num_envs)


def init_density(rng, rows, cols, num_envs):
    return _patch_layer(rng, rows, cols, num_envs)


def init_altitude(rng, rows, cols, num_envs):
    """init_utils.py:76-116: U(0,5) noise + 6-9 cosine hills + 4-7 ramps, then /10."""
    alt = np.zeros((num_envs, rows, cols))
    for e in range(num_envs):
        alt[e] = rng.uniform(0, 5, (rows, cols))
        for _ in range(rng.randint(6, 10)):
            cr = rng.randint(0, rows)
            cc = rng.randint(0, cols)
            radius = rng.randint(2, min(rows, cols) // 4)
            height = rng.uniform(2, 6)
            for i in range(rows):
                for j in range(cols):
                    d = np.sqrt((i - cr) ** 2 + (j - cc) ** 2)
                    if d < radius:
                        alt[e, i, j] += height * np.cos(d / radius * np.pi / 2)
        for _ in range(rng.randint(4, 8)):
            sr = rng.randint(0, rows - 4)
            sc = rng.randint(0, cols - 4)
            width = rng.randint(3, cols // 4)
            height = rng.randint(3, rows // 4)
            diff = rng.uniform(1, 4)
            for i in range(sr, min(sr + height, rows)):
                for j in range(sc, min(sc + width, cols)):
                    alt[e, i, j] += diff * ((i - sr) / height)
    return alt / 10


def get_slope(altitude):
    """init_utils.py:166-200: degrees(atan(alt[r,c] - alt[nbr])), diagonals / 1.414, centre 0,
    border cells all 0.  (N,H,W) -> (N,H,W,3,3) float64."""
    N, H, W = altitude.shape
    s = np.zeros((N, H, W, 3, 3))
    for e in range(N):
        for r in range(1, H - 1):
            for c in range(1, W - 1):
                d = altitude[e, r, c] - altitude[e, r - 1:r + 2, c - 1:c + 2]
                d = d.astype(np.float64)
                for (i, j) in ((0, 0), (0, 2), (2, 0), (2, 2)):
                    d[i, j] /= 1.414
                s[e, r, c] = np.degrees(np.arctan(d))
                s[e, r, c, 1, 1] = 0
    return s


def initial_state(nrows, ncols, num_envs, seed=0, jax_seed=1, use_hidden=True, mode=prng.LEGACY,
                  p_tree=0.90, p_empty=0.10, middle_fire=False, hidden="reference", slope_fn=None):
    """Initial (state, info) in the reference's pytree layout (advanced_bulldozer.py:650-743,
    401-420), NumPy arrays with the reference's dtypes.

    hidden: "reference" = restated init_utils patches/hills; "random" = iid synthetic layers
    (SURVEY.md section 8d); use_hidden False -> veg = den = 3, altitude 0, wind_index 0.
    Keys: key(jax_seed) -> split -> first -> split(num_envs) (scripts/run:554-555,
    advanced_bulldozer.py:705).
    """
    rs = np.random.RandomState(seed)
    gen = np.random.default_rng(seed)
    if use_hidden and hidden == "reference":
        density = init_density(rs, nrows, ncols, num_envs)
        vegetation = init_vegetation(rs, nrows, ncols, num_envs)
        altitude = init_altitude(rs, nrows, ncols, num_envs)
    elif use_hidden:
        density = gen.integers(1, 6, size=(num_envs, nrows, ncols))
        vegetation = gen.integers(1, 6, size=(num_envs, nrows, ncols))
        altitude = gen.uniform(0, 1.1, size=(num_envs, nrows, ncols))
    else:
        density = np.full((num_envs, nrows, ncols), 3, dtype=int)
        vegetation = np.full((num_envs, nrows, ncols), 3, dtype=int)
        altitude = np.zeros((num_envs, nrows, ncols))
    # slope_fn: tests of very large grids pass a vectorised equivalent of get_slope (checked against
    # this literal one on small grids) because the per-cell Python loop is O(N H W)
    slope = (slope_fn or get_slope)(np.asarray(altitude, dtype=np.float64)).astype(F32)
    # grid: iid empty/tree, two fire seeds (:650-688)
    grid = gen.choice(np.array([EMPTY, TREE, FIRE]), size=(num_envs, nrows, ncols),
                      p=[p_empty, p_tree, 0.0]).astype(F32)
    if middle_fire:
        r, c = nrows // 2, ncols // 2
    else:
        r, c = 3 * nrows // 4, ncols // 4
    fire_age = np.zeros((num_envs, nrows, ncols), dtype=F32)
    init_age = (nrows + nrows // 2) * 2
    for (fr, fc) in ((r, c), (r, c - 1)):
        grid[:, fr, fc] = FIRE
        fire_age[:, fr, fc] = init_age
    wind_index = (gen.integers(0, 8, size=num_envs) if use_hidden
                  else np.zeros(num_envs)).astype(np.int32)
    start = prng.split(prng.key_from_seed(jax_seed), 2, mode)[0]
    keys = prng.split(start, num_envs, mode)
    ctx = {
        "wind_index": wind_index,
        "density": density.astype(np.int32),
        "vegetation": vegetation.astype(np.int32),
        "altitude": np.asarray(altitude).astype(F32),
        "slope": slope,
        "pslope": p_slope_table(slope),
        "fire_age": fire_age,
        "key": keys.astype(np.uint32),
        "is_night": np.zeros(num_envs, dtype=np.int32),
        "true_grid": grid,
        "time_step": np.ones(num_envs, dtype=np.int32),
        "dousing_count": np.zeros((num_envs, nrows, ncols), dtype=np.int32),
    }
    pos = np.tile(np.array([[int(nrows * 0.15), int(ncols * 0.85)]], dtype=np.int32), (num_envs, 1))
    state = {"per_env_context": ctx, "position": pos, "time": np.zeros(num_envs, dtype=F32)}
    info = {
        "TimeLimit.truncated": np.zeros(num_envs, dtype=bool),
        "terminated": np.zeros(num_envs, dtype=bool),
        "steps_elapsed": np.zeros(num_envs, dtype=F32),
        "reward_accumulated": np.zeros(num_envs, dtype=F32),
        "reward": np.zeros(num_envs, dtype=F32),
    }
    return state, info

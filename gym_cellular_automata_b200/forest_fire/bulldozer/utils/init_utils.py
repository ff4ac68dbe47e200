"""Host-side generators of the static inputs of the advanced bulldozer env: hidden layers
(vegetation, density, altitude -> slope), wind tables, extension-action lookup.

Same distributions and -- given the same ``np.random.RandomState`` -- the same random call
sequence as the reference (forest_fire/bulldozer/utils/init_utils.py:10-116,166-245), but
vectorised: the reference's per-cell Python loops are O(N H W hills) and take minutes at 4096^2
or 65536 envs.  These are inputs of the hot path, not part of it; the slope
factor ``exp(f32(0.078) * slope)`` is tabulated here once because slope is static."""
from __future__ import annotations

import itertools

import numpy as np

WIND_THETAS = np.array([
    [[45, 0, 45], [90, 0, 90], [135, 180, 135]],
    [[90, 45, 0], [135, 0, 45], [180, 135, 90]],
    [[135, 90, 45], [180, 0, 0], [135, 90, 45]],
    [[180, 135, 90], [135, 0, 45], [90, 45, 0]],
    [[135, 180, 135], [90, 0, 90], [45, 0, 45]],
    [[90, 135, 180], [45, 0, 135], [0, 45, 90]],
    [[45, 90, 135], [0, 0, 180], [45, 90, 135]],
    [[0, 45, 90], [45, 0, 135], [90, 135, 180]],
], dtype=np.float64)


def get_winds(use_hidden: bool = True) -> np.ndarray:
    """(8, 2, 3, 3) float64: per wind direction the (wind matrix with centre 0, ft) pair.
    ``use_hidden`` is accepted for signature parity; it has no effect in the reference either."""
    t = np.radians(WIND_THETAS)
    ft = np.exp(10 * 0.131 * (np.cos(t) - 1))
    w = np.exp(0.045 * 10) * ft
    w[:, 1, 1] = 0
    return np.stack([w, ft], axis=1)


def _patches(rng, rows, cols, num_envs):
    m = np.zeros((num_envs, rows, cols), dtype=int)
    for e in range(num_envs):
        for _ in range(rng.randint(4, 8)):
            cr, cc = rng.randint(0, rows), rng.randint(0, cols)
            ph, pw = rng.randint(3, rows // 2), rng.randint(3, cols // 2)
            kind = rng.randint(1, 6)
            m[e, max(0, cr - ph // 2):min(rows, cr + ph // 2), max(0, cc - pw // 2):min(cols, cc + pw // 2)] = kind
        zero = m[e] == 0
        m[e][zero] = rng.randint(1, 4, size=int(zero.sum()))
    return m


def init_vegetation(row_count, column_count, num_envs, rng=None):
    return _patches(rng or np.random, row_count, column_count, num_envs)


def init_density(row_count, column_count, num_envs, rng=None):
    return _patches(rng or np.random, row_count, column_count, num_envs)


def init_altitude(row_count, column_count, num_envs, rng=None):
    rng = rng or np.random
    ii, jj = np.meshgrid(np.arange(row_count), np.arange(column_count), indexing="ij")
    alt = np.zeros((num_envs, row_count, column_count))
    for e in range(num_envs):
        alt[e] = rng.uniform(0, 5, (row_count, column_count))
        for _ in range(rng.randint(6, 10)):
            cr, cc = rng.randint(0, row_count), rng.randint(0, column_count)
            radius = rng.randint(2, min(row_count, column_count) // 4)
            height = rng.uniform(2, 6)
            dist = np.sqrt((ii - cr) ** 2 + (jj - cc) ** 2)
            alt[e] += np.where(dist < radius, height * np.cos(dist / radius * np.pi / 2), 0.0)
        for _ in range(rng.randint(4, 8)):
            sr, sc = rng.randint(0, row_count - 4), rng.randint(0, column_count - 4)
            width, height = rng.randint(3, column_count // 4), rng.randint(3, row_count // 4)
            diff = rng.uniform(1, 4)
            r1, c1 = min(sr + height, row_count), min(sc + width, column_count)
            alt[e, sr:r1, sc:c1] += (diff * ((np.arange(sr, r1) - sr) / height))[:, None]
    return alt / 10


def init_density_same(row_count, column_count, num_envs):
    return np.full((num_envs, row_count, column_count), 3, dtype=int)


def init_vegetation_same(row_count, column_count, num_envs):
    return np.full((num_envs, row_count, column_count), 3, dtype=int)


def init_altitude_same(row_count, column_count, num_envs):
    return np.zeros((num_envs, row_count, column_count), dtype=int)


def get_slope(altitude, row_count=None, column_count=None, num_envs=None):
    """(N,H,W) altitude -> (N,H,W,3,3) slope in degrees: atan(alt[r,c] - alt[neighbour]), diagonal
    differences / 1.414, centre 0, border cells flat."""
    alt = np.asarray(altitude, dtype=np.float64)
    N, H, W = alt.shape
    s = np.zeros((N, H, W, 3, 3))
    cur = alt[:, 1:H - 1, 1:W - 1]
    for i in range(3):
        for j in range(3):
            if i == 1 and j == 1:
                continue
            d = cur - alt[:, i:i + H - 2, j:j + W - 2]
            if i != 1 and j != 1:
                d = d / 1.414
            s[:, 1:H - 1, 1:W - 1, i, j] = np.degrees(np.arctan(d))
    return s


def p_slope_table(slope) -> np.ndarray:
    """exp(f32(0.078) * slope) in float32 (reference ca_alexandridis_jax.py:199-200), evaluated once
    on the host with NumPy; the CPU oracle uses the identical call, so both sides see the same
    bits whatever the libm."""
    return np.exp(np.float32(0.078) * np.asarray(slope, dtype=np.float32)).astype(np.float32)


def create_up_to_k_mappings(n, k):
    """ids of all subsets of size <= k of n items <-> their indicator vectors
    (for (2,1): 0 -> (0,0), 1 -> (1,0), 2 -> (0,1))."""
    vectors, to_id = [], {}
    for size in range(k + 1):
        for combo in itertools.combinations(range(n), size):
            v = tuple(1 if i in combo else 0 for i in range(n))
            to_id[v] = len(vectors)
            vectors.append(v)
    return np.array(vectors, dtype=np.int32), to_id

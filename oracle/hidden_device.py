"""TEST INFRASTRUCTURE (CPU oracle) -- NumPy restatement of the device-side hidden-layer generator
(gym_cellular_automata_b200/csrc/gca_hidden.cu): the layer models of the reference's
/root/reference/gym_cellular_automata/forest_fire/bulldozer/utils/init_utils.py:10-73 (patches), :76-116
(altitude), :166-200 (get_slope), with every random number taken from one threefry2x32 block addressed by
(stream, env, index).  Only tests / smoke / bench may import it.  Parity unpinned against the reference: its
generators draw from Python's unseeded RNG, so only the layer model (not the values) can be compared."""
import numpy as np

from . import prng

VEG_RECT, VEG_FILL, DEN_RECT, DEN_FILL, ALT_NOISE, ALT_HILL, ALT_RAMP = 1, 2, 3, 4, 5, 6, 7


def _hrand(seed, stream, env, idx):
    k0 = np.uint32((seed & 0xFFFFFFFF) ^ stream)
    k1 = np.uint32(seed >> 32)
    a, b = prng.threefry2x32(k0, k1, np.asarray(env, np.uint32), np.asarray(idx, np.uint32))
    return np.asarray(a, np.uint64), np.asarray(b, np.uint64)


def _unif(a, b):
    return (((a << np.uint64(32)) | b) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def patches(H, W, seed, env, rect_stream, fill_stream):
    """vegetation / density of global env `env`: (H, W) int32."""
    m = np.zeros((H, W), np.int32)
    n = 4 + int(_hrand(seed, rect_stream, env, 100)[0] % 4)
    for k in range(n):
        a, b = _hrand(seed, rect_stream, env, 3 * k)
        c, d = _hrand(seed, rect_stream, env, 3 * k + 1)
        kk, _ = _hrand(seed, rect_stream, env, 3 * k + 2)
        cr, cc = int(a % H), int(b % W)
        ph, pw = 3 + int(c % (H // 2 - 3)), 3 + int(d % (W // 2 - 3))
        m[max(0, cr - ph // 2):min(H, cr + ph // 2), max(0, cc - pw // 2):min(W, cc + pw // 2)] = 1 + int(kk % 5)
    fill, _ = _hrand(seed, fill_stream, np.full(H * W, env), np.arange(H * W))
    fill = (1 + (fill % 3)).astype(np.int32).reshape(H, W)
    return np.where(m == 0, fill, m)


def altitude(H, W, seed, env):
    """(H, W) float64."""
    a, b = _hrand(seed, ALT_NOISE, np.full(H * W, env), np.arange(H * W))
    alt = (5.0 * _unif(a, b)).reshape(H, W)
    ii, jj = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    nh = 6 + int(_hrand(seed, ALT_HILL, env, 100)[0] % 4)
    nr = 4 + int(_hrand(seed, ALT_RAMP, env, 100)[0] % 4)
    for k in range(nh):
        a, b = _hrand(seed, ALT_HILL, env, 3 * k)
        c, _ = _hrand(seed, ALT_HILL, env, 3 * k + 1)
        x, y = _hrand(seed, ALT_HILL, env, 3 * k + 2)
        cr, cc, rad = int(a % H), int(b % W), 2 + int(c % (min(H, W) // 4 - 2))
        hh = 2.0 + 4.0 * float(_unif(x, y))
        dist = np.sqrt((ii - cr).astype(np.float64) ** 2 + (jj - cc).astype(np.float64) ** 2)
        alt = alt + np.where(dist < rad, hh * np.cos(dist / rad * 3.141592653589793 / 2.0), 0.0)
    for k in range(nr):
        a, b = _hrand(seed, ALT_RAMP, env, 3 * k)
        c, d = _hrand(seed, ALT_RAMP, env, 3 * k + 1)
        x, y = _hrand(seed, ALT_RAMP, env, 3 * k + 2)
        sr, sc = int(a % (H - 4)), int(b % (W - 4))
        w, h = 3 + int(c % (W // 4 - 3)), 3 + int(d % (H // 4 - 3))
        diff = 1.0 + 3.0 * float(_unif(x, y))
        r1, c1 = min(sr + h, H), min(sc + w, W)
        alt[sr:r1, sc:c1] += (diff * ((np.arange(sr, r1) - sr) / h))[:, None]
    return alt / 10.0


def slope(alt):
    """(H, W) float64 altitude -> (H, W, 3, 3) float32 degrees (init_utils.py:166-200)."""
    H, W = alt.shape
    s = np.zeros((H, W, 3, 3))
    cur = alt[1:H - 1, 1:W - 1]
    for i in range(3):
        for j in range(3):
            if i == 1 and j == 1:
                continue
            d = cur - alt[i:i + H - 2, j:j + W - 2]
            if i != 1 and j != 1:
                d = d / 1.414
            s[1:H - 1, 1:W - 1, i, j] = np.arctan(d) * (180.0 / 3.141592653589793)
    return s.astype(np.float32)

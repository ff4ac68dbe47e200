"""ORACLE (test infrastructure, not product code) -- CPU restatement of jax.random.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (gym_cellular_automata_b200) never does.

What is restated
----------------
The reference draws every random number through ``jax.random`` (call sites:
gym_cellular_automata/forest_fire/operators/ca_alexandridis_jax.py:352,362-364,367-370,
437,443-448 and forest_fire/bulldozer/advanced_bulldozer.py:705).  ``jax``/``jaxlib`` are a
third-party dependency that is NOT vendored under /root/reference and is unpinned
(setup.py:25-26 lists bare "jax", "jaxlib"); it cannot be installed in this image.  This
file restates the published algorithm of ``jax._src.prng`` / ``jax._src.random``:

* threefry2x32 (Random123, 20 rounds) -- pinned by the Random123 known-answer vectors.
* ``split``, ``random_bits`` in BOTH stream layouts:
    - "legacy"  = jax_threefry_partitionable False (default before JAX 0.5.0): counters
      0..n-1 (odd n padded with one 0) are cut into halves (lo, hi), block b =
      threefry(key; lo[b], hi[b]) and the output is concat(word0s, word1s)[:n].
    - "partitionable" = default from JAX 0.5.0: element idx = word0 ^ word1 of block
      (hi32(idx), lo32(idx)); split(key, m)[i] = both words of block (0, i).
  Pinned by the public JAX documentation values (see tests/test_oracle_prng.py).
* ``uniform`` float32 in [0,1): (bits >> 9) * 2**-23 exactly.
* ``randint`` int32: two draws and the multiply-mod construction of ``_randint``.

Parity status: PRNG layer pinned to public known-answer vectors only (jax cannot run here and the
reference's own tests hold no random golden values).  The operators that consume
these numbers are pinned to the reference's own source through oracle/ref_shim, which delegates
``jax.random`` to this module.
"""
from __future__ import annotations

import numpy as np

LEGACY = 0
PARTITIONABLE = 1

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M32 = np.uint64(0xFFFFFFFF)


def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def threefry2x32(k0, k1, x0, x1):
    """threefry2x32-20.  All arguments broadcastable uint32 arrays; returns (y0, y1)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=np.uint32)
        k1 = np.asarray(k1, dtype=np.uint32)
        x0 = np.asarray(x0, dtype=np.uint32).copy()
        x1 = np.asarray(x1, dtype=np.uint32).copy()
        ks = (k0, k1, (k0 ^ k1 ^ np.uint32(0x1BD11BDA)).astype(np.uint32))
        x0 = (x0 + ks[0]).astype(np.uint32)
        x1 = (x1 + ks[1]).astype(np.uint32)
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = _rotl(x1, r)
                x1 = (x1 ^ x0).astype(np.uint32)
            x0 = (x0 + ks[(g + 1) % 3]).astype(np.uint32)
            x1 = (x1 + ks[(g + 2) % 3] + np.uint32(g + 1)).astype(np.uint32)
        return x0, x1


def key_from_seed(seed: int) -> np.ndarray:
    """jax.random.key(seed) / PRNGKey(seed) raw data for 32-bit non-negative seeds."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def random_bits(key, n: int, mode: int = LEGACY) -> np.ndarray:
    """``jax.random.bits(key, (n,), uint32)`` for one key (2,) -> (n,) uint32.

    Multi-dimensional shapes flatten row-major, so callers pass n = prod(shape).
    """
    key = np.asarray(key, dtype=np.uint32)
    n = int(n)
    if mode == LEGACY:
        m = n + (n & 1)
        h = m // 2
        cnt = np.arange(m, dtype=np.uint32)
        if n & 1:
            cnt[-1] = 0
        a, b = threefry2x32(key[0], key[1], cnt[:h], cnt[h:])
        return np.concatenate([a, b])[:n]
    elif mode == PARTITIONABLE:
        idx = np.arange(n, dtype=np.uint64)
        hi = (idx >> np.uint64(32)).astype(np.uint32)
        lo = (idx & _M32).astype(np.uint32)
        a, b = threefry2x32(key[0], key[1], hi, lo)
        return (a ^ b).astype(np.uint32)
    raise ValueError("mode")


def random_bits_at(key, idx, n: int, mode: int = LEGACY) -> np.ndarray:
    """Element(s) ``idx`` of random_bits(key, n) without generating the rest (lazy form)."""
    key = np.asarray(key, dtype=np.uint32)
    idx = np.asarray(idx, dtype=np.int64)
    if mode == LEGACY:
        m = n + (n & 1)
        h = m // 2
        first = idx < h
        lo = np.where(first, idx, idx - h).astype(np.uint32)
        hi_c = np.where(first, idx + h, idx)
        # the padded odd counter is 0, not m-1
        if n & 1:
            hi_c = np.where(hi_c == m - 1, 0, hi_c)
        a, b = threefry2x32(key[0], key[1], lo, hi_c.astype(np.uint32))
        return np.where(first, a, b).astype(np.uint32)
    elif mode == PARTITIONABLE:
        a, b = threefry2x32(key[0], key[1], (idx >> 32).astype(np.uint32), (idx & 0xFFFFFFFF).astype(np.uint32))
        return (a ^ b).astype(np.uint32)
    raise ValueError("mode")


def split(key, num: int = 2, mode: int = LEGACY) -> np.ndarray:
    """``jax.random.split(key, num)`` -> (num, 2) uint32."""
    key = np.asarray(key, dtype=np.uint32)
    if mode == LEGACY:
        return random_bits(key, 2 * num, LEGACY).reshape(num, 2)
    elif mode == PARTITIONABLE:
        i = np.arange(num, dtype=np.uint32)
        a, b = threefry2x32(key[0], key[1], np.zeros(num, np.uint32), i)
        return np.stack([a, b], axis=1).astype(np.uint32)
    raise ValueError("mode")


def bits_to_uniform(bits) -> np.ndarray:
    """float32 in [0,1): bitcast((bits >> 9) | 0x3F800000) - 1.0 == (bits >> 9) * 2**-23."""
    bits = np.asarray(bits, dtype=np.uint32)
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).astype(np.uint32).view(np.float32)
    return (f - np.float32(1.0)).astype(np.float32)


def uniform(key, n: int, mode: int = LEGACY) -> np.ndarray:
    return bits_to_uniform(random_bits(key, n, mode))


def randint_params(minval, maxval):
    """(lo, span, mult) of jax.random.randint for int32; non-integer bounds are truncated
    (``minval.astype(int)``), e.g. fire_age_min = 144.0 -> 144."""
    lo = int(minval)
    hi = int(maxval)
    span = (hi - lo) & 0xFFFFFFFF
    if hi <= lo:
        span = 1
    mult = (2 ** 16) % span
    mult = (mult * mult) % span
    return lo, span, mult


def randint_from_bits(hb, lb, minval, maxval) -> np.ndarray:
    lo, span, mult = randint_params(minval, maxval)
    hb = np.asarray(hb, dtype=np.uint32)
    lb = np.asarray(lb, dtype=np.uint32)
    with np.errstate(over="ignore"):
        s = np.uint32(span)
        off = ((hb % s) * np.uint32(mult) + (lb % s)).astype(np.uint32)
        off = off % s
    return (np.int64(lo) + off.astype(np.int64)).astype(np.int32)


def randint(key, n: int, minval, maxval, mode: int = LEGACY) -> np.ndarray:
    """``jax.random.randint(key, (n,), minval, maxval)`` int32."""
    k = split(key, 2, mode)
    hb = random_bits(k[0], n, mode)
    lb = random_bits(k[1], n, mode)
    return randint_from_bits(hb, lb, minval, maxval)

"""ORACLE SUPPORT (test infrastructure) -- the few ``gymnasium`` names the reference's operator / env source imports
(space containers with ``sample`` / ``contains``, ``Env``, ``logger``, ``utils.seeding``).  No behaviour of the
hot path depends on them; they exist so the reference files import."""
from __future__ import annotations

import types

import numpy as np


# Reproducible vector generation: the reference never seeds its spaces / env generator, so a stand-in seeded from the
# OS would make every run of tests/golden/make_reference_golden.py different.  seed_all(s) makes every generator that is
# created WITHOUT a seed afterwards draw its seed from one stream started at s (creation order is deterministic).
_SEED_SOURCE = None


def seed_all(seed=None):
    global _SEED_SOURCE
    _SEED_SOURCE = None if seed is None else np.random.default_rng(seed)


def _fresh_rng(seed=None):
    if seed is None and _SEED_SOURCE is not None:
        seed = int(_SEED_SOURCE.integers(0, 2 ** 62))
    return np.random.default_rng(seed)


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._np_random = _fresh_rng(seed)

    @property
    def shape(self):
        return self._shape

    @property
    def np_random(self):
        return self._np_random

    def seed(self, seed=None):
        self._np_random = _fresh_rng(seed)
        return [seed]

    def contains(self, x):
        return True

    def __contains__(self, x):
        return self.contains(x)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        super().__init__(shape, dtype, seed)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

    def sample(self):
        if self.dtype.kind == "f":
            return self._np_random.uniform(self.low, self.high, self.shape).astype(self.dtype)
        return self._np_random.integers(self.low, self.high, self.shape, endpoint=True).astype(self.dtype)

    def __eq__(self, o):
        return isinstance(o, Box) and self.shape == o.shape and np.array_equal(self.low, o.low) and np.array_equal(self.high, o.high)

    __hash__ = None


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        super().__init__((), np.int64, seed)
        self.n, self.start = int(n), int(start)

    def sample(self):
        return int(self.start + self._np_random.integers(self.n))


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec = np.asarray(nvec, dtype=dtype)
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self):
        return (self._np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)


class MultiBinary(Space):
    def __init__(self, n, seed=None):
        super().__init__((n,) if np.isscalar(n) else tuple(n), np.int8, seed)

    def sample(self):
        return self._np_random.integers(0, 2, self.shape).astype(np.int8)


class Tuple(Space):
    def __init__(self, spaces, seed=None):
        super().__init__(None, None, seed)
        self.spaces = tuple(spaces)

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def __iter__(self):
        return iter(self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **kw):
        super().__init__(None, None, seed)
        self.spaces = dict(spaces or {}, **kw)

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()


class Env:
    metadata: dict = {}
    render_mode = None

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.default_rng(seed)

    @property
    def np_random(self):
        if not hasattr(self, "_np_random"):
            self._np_random = _fresh_rng()
        return self._np_random


def build_modules():
    g = types.ModuleType("gymnasium")
    g._GCA_SHIM = True
    spaces = types.ModuleType("gymnasium.spaces")
    for c in (Space, Box, Discrete, MultiDiscrete, MultiBinary, Tuple, Dict):
        setattr(spaces, c.__name__, c)
    g.spaces = spaces
    g.Env = Env
    g.Space = Space
    logger = types.ModuleType("gymnasium.logger")
    for n in ("warn", "info", "error", "debug"):
        setattr(logger, n, lambda *a, **k: None)
    g.logger = logger
    utils = types.ModuleType("gymnasium.utils")
    seeding = types.ModuleType("gymnasium.utils.seeding")
    seeding.np_random = lambda seed=None: (_fresh_rng(seed), seed)
    utils.seeding = seeding
    g.utils = utils
    error = types.ModuleType("gymnasium.error")
    error.Error = type("Error", (Exception,), {})
    g.error = error
    return {"gymnasium": g, "gymnasium.spaces": spaces, "gymnasium.logger": logger, "gymnasium.utils": utils,
            "gymnasium.utils.seeding": seeding, "gymnasium.error": error}

"""Minimal gymnasium.spaces stand-ins (gymnasium is not installed in the target image).

If gymnasium is importable its classes are used unchanged; otherwise these cover what the
reference's env and its PPO caller touch: ``sample()``, ``contains()``, iteration, ``[]``,
``.shape``, ``.nvec`` (reference agents/jax_ppo.py:708-735,790-791,857-861)."""
from __future__ import annotations

from typing import Optional

import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium.spaces import Box, Dict, Discrete, MultiDiscrete, Space, Tuple  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # gymnasium absent
    HAVE_GYMNASIUM = False

    class Space:
        def __init__(self, shape=None, dtype=None, seed: Optional[int] = None):
            self._shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)
            self._np_random = None
            self._seed = seed

        @property
        def shape(self):
            return self._shape

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random = np.random.default_rng(self._seed)
            return self._np_random

        def seed(self, seed=None):
            self._seed = seed
            self._np_random = np.random.default_rng(seed)
            return [seed]

        def sample(self):
            raise NotImplementedError

        def contains(self, x) -> bool:
            raise NotImplementedError

        def __contains__(self, x):
            return self.contains(x)

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            super().__init__(shape, dtype, seed)
            is_int = np.issubdtype(self.dtype, np.integer)

            def bound(v, inf_value):
                v = np.asarray(v, dtype=np.float64)
                if is_int:
                    v = np.where(np.isinf(v), inf_value, v)
                return np.broadcast_to(v, self.shape).astype(self.dtype)

            self.low = bound(low, np.iinfo(self.dtype).min if is_int else 0)
            self.high = bound(high, np.iinfo(self.dtype).max if is_int else 0)

        def sample(self):
            lo = self.low.astype(np.float64)
            hi = self.high.astype(np.float64)
            if np.issubdtype(self.dtype, np.integer):
                return self.np_random.integers(self.low, self.high.astype(np.int64) + 1, size=self.shape).astype(self.dtype)
            unb = np.isinf(hi)
            out = np.empty(self.shape, dtype=np.float64)
            out[unb] = lo[unb] + self.np_random.exponential(size=int(unb.sum()))
            out[~unb] = self.np_random.uniform(lo[~unb], hi[~unb])
            return out.astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __eq__(self, other):
            return isinstance(other, Box) and self.shape == other.shape and np.array_equal(self.low, other.low) \
                and np.array_equal(self.high, other.high)

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Discrete(Space):
        def __init__(self, n: int, seed=None, start: int = 0):
            super().__init__((), np.int64, seed)
            self.n, self.start = int(n), int(start)

        def sample(self):
            return int(self.start + self.np_random.integers(self.n))

        def contains(self, x) -> bool:
            try:
                xi = int(x)
            except Exception:
                return False
            return self.start <= xi < self.start + self.n

        def __eq__(self, other):
            return isinstance(other, Discrete) and self.n == other.n and self.start == other.start

        def __repr__(self):
            return f"Discrete({self.n})"

    class MultiDiscrete(Space):
        def __init__(self, nvec, dtype=np.int64, seed=None):
            self.nvec = np.array(nvec, dtype=dtype, copy=True)
            super().__init__(self.nvec.shape, dtype, seed)

        def sample(self):
            return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

        def __eq__(self, other):
            return isinstance(other, MultiDiscrete) and np.array_equal(self.nvec, other.nvec)

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"

    class Tuple(Space):
        def __init__(self, spaces, seed=None):
            self.spaces = tuple(spaces)
            super().__init__(None, None, seed)

        def sample(self):
            return tuple(s.sample() for s in self.spaces)

        def contains(self, x) -> bool:
            return isinstance(x, (tuple, list)) and len(x) == len(self.spaces) and \
                all(s.contains(v) for s, v in zip(self.spaces, x))

        def __getitem__(self, i):
            return self.spaces[i]

        def __len__(self):
            return len(self.spaces)

        def __iter__(self):
            return iter(self.spaces)

        def __eq__(self, other):
            return isinstance(other, Tuple) and self.spaces == other.spaces

    class Dict(Space):
        def __init__(self, spaces=None, seed=None, **kw):
            self.spaces = dict(spaces or {})
            self.spaces.update(kw)
            super().__init__(None, None, seed)

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x) -> bool:
            return isinstance(x, dict) and x.keys() == self.spaces.keys() and \
                all(self.spaces[k].contains(v) for k, v in x.items())

        def __getitem__(self, k):
            return self.spaces[k]

        def __iter__(self):
            return iter(self.spaces)

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def __len__(self):
            return len(self.spaces)

        def __eq__(self, other):
            return isinstance(other, Dict) and self.spaces == other.spaces

"""GridSpace -- the space of integer lattices the envs declare for their grids.

Same public surface as the reference type (gym_cellular_automata/grid_space.py:11-90): built from a number of cell
states ``n`` (states 0..n-1) or from explicit ``values``, optionally with one sampling probability per state;
``sample()`` draws an iid lattice of ``shape``, ``contains(x)`` checks shape and state set, two spaces are equal when
shape and states agree."""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

from ._config import TYPE_INT
from .spaces import Space


def _cell_states(n, values, dtype) -> np.ndarray:
    """Sorted unique cell states from either constructor form."""
    if values is not None:
        return np.unique(np.asarray(list(values), dtype=dtype))
    if n is None:
        raise ValueError("'n' or 'values' must be provided.")
    if not n > 0:
        raise AssertionError("'n' must be a positive integer.")
    return np.arange(int(n), dtype=dtype)


class GridSpace(Space):
    def __init__(self, n: Optional[int] = None, values: Optional[Sequence[int]] = None, shape: tuple = tuple(),
                 probs: Optional[Sequence[float]] = None, dtype=TYPE_INT, seed: Optional[int] = None):
        if not shape:
            raise AssertionError("Shape must be a non-empty tuple.")
        super().__init__(shape, dtype, seed)
        self._from_values = values is not None
        self.values = _cell_states(n, values, dtype)
        self.n = int(self.values.size)
        self.probs = np.full(self.n, 1.0 / self.n) if probs is None else probs
        if len(self.probs) != self.n:
            raise AssertionError("Unique values do NOT MATCH with assigned probabilities.")
        self.size = math.prod(self.shape)

    # -- gymnasium.Space protocol ---------------------------------------------------------------------
    def sample(self) -> np.ndarray:
        flat = self.np_random.choice(a=self.values, size=self.size, p=self.probs)
        return flat.reshape(self.shape)

    def contains(self, x) -> bool:
        lattice = np.asarray(x, dtype=self.dtype) if isinstance(x, list) else np.asarray(x)
        if lattice.shape != self.shape:
            return False
        return bool(np.isin(lattice, self.values).all())

    @property
    def is_np_flattenable(self):
        return True

    # -- value semantics -------------------------------------------------------------------------------
    def __eq__(self, other):
        if not isinstance(other, GridSpace) or self.shape != other.shape:
            return False
        return np.array_equal(self.values, other.values)

    __hash__ = None

    def __repr__(self):
        what = f"values={self.values}" if self._from_values else f"n={self.n}"
        return f"GridSpace({what}, shape={self.shape})"

"""Operator: the reference's abstract operator contract, unchanged in meaning
(reference gym_cellular_automata/operator.py:10-75).

``update(grid, action, context) -> (new_grid, new_context)``; ``__call__`` forwards to
``update``; ``seed`` installs a NumPy generator; the four class flags and ``suboperators``
describe the operator.  Concrete CUDA operators carry a leading env axis on every array (the
reference vmaps single-env operators from stateless_step; here the batch is inside)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from copy import copy
from typing import Any, Optional, Tuple

import numpy as np

from .spaces import Space


class Operator(ABC):
    suboperators: Tuple = tuple()

    grid_dependant: Optional[bool] = None
    action_dependant: Optional[bool] = None
    context_dependant: Optional[bool] = None

    deterministic: Optional[bool] = None

    @abstractmethod
    def __init__(self, grid_space: Optional[Space] = None, action_space: Optional[Space] = None,
                 context_space: Optional[Space] = None) -> None:
        self.grid_space = grid_space
        self.action_space = action_space
        self.context_space = context_space
        self.seed()

    @abstractmethod
    def update(self, grid, action: Any, context: Any):
        """Returns (new_grid, new_context); the base implementation is the identity on copies."""
        return copy(grid), copy(context)

    def __call__(self, *args, **kwargs):
        return self.update(*args, **kwargs)

    def seed(self, seed=None):
        self._seed = seed
        self.np_random = np.random.default_rng(seed)
        return [seed]

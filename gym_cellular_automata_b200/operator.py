"""Operator -- the abstract building block every CA / move / modify / clock / MDP class derives from.

Contract of the reference type (gym_cellular_automata/operator.py:10-75), unchanged in meaning:

* construction takes the three optional spaces (grid, action, context) and seeds a NumPy generator;
* ``update(grid, action, context) -> (new_grid, new_context)`` is the one abstract method, ``op(...)`` calls it;
* four tri-state class flags say what the result depends on and whether it is deterministic, ``suboperators`` lists
  the operators a composite one is made of;
* ``seed(s)`` re-seeds and returns ``[s]``.

The CUDA operators of this package keep the contract with a leading env axis on every array (the reference vmaps
single-env operators from ``stateless_step``; here the batch is inside)."""
from __future__ import annotations

import copy as _copy
from abc import ABC, abstractmethod
from typing import Any, Optional, Tuple

import numpy as np

from .spaces import Space


class Operator(ABC):
    # description of the operator, to be set by every subclass
    suboperators: Tuple = ()
    grid_dependant: Optional[bool] = None
    action_dependant: Optional[bool] = None
    context_dependant: Optional[bool] = None
    deterministic: Optional[bool] = None

    @abstractmethod
    def __init__(self, grid_space: Optional[Space] = None, action_space: Optional[Space] = None,
                 context_space: Optional[Space] = None) -> None:
        self.grid_space, self.action_space, self.context_space = grid_space, action_space, context_space
        self.seed()

    @abstractmethod
    def update(self, grid, action: Any, context: Any):
        """(new_grid, new_context).  Subclasses that call ``super().update`` get shallow copies of their inputs."""
        return _copy.copy(grid), _copy.copy(context)

    def __call__(self, *args, **kwargs):
        return self.update(*args, **kwargs)

    def seed(self, seed=None):
        self._seed, self.np_random = seed, np.random.default_rng(seed)
        return [seed]

"""CAEnv -- abstract base of the cellular-automaton environments (reference gym_cellular_automata/ca_env.py:9-99).

A concrete env provides ``MDP`` (the top operator), ``initial_state`` (grid, context) and the three hooks ``_award``,
``_is_done`` (sets ``self.done``), ``_report``; this base supplies the stateful gym loop around them:

* ``reset()`` -> ``(obs, info)`` with ``obs = (grid, context)`` from ``initial_state`` and the episode counters zeroed;
* ``step(action)`` -> ``(obs, reward, terminated, truncated, info)``: one ``MDP`` transition, then the hooks; a step
  after termination warns once and returns the frozen state with reward 0 (reference :50-62);
* ``status()`` / ``count_cells()``.

gymnasium is optional: when it is importable ``CAEnv`` is a ``gymnasium.Env``, otherwise a plain class with the same
``reset(seed=...)`` seeding behaviour."""
from __future__ import annotations

import collections
import warnings
from abc import ABC, abstractmethod
from typing import Optional

import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium import Env as _EnvBase
except Exception:  # gymnasium absent
    class _EnvBase:
        metadata: dict = {}

        def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
            if seed is not None or not hasattr(self, "np_random"):
                self.np_random = np.random.default_rng(seed)

_STEP_AFTER_DONE = ("You are calling 'step()' even though this environment has already returned done = True. "
                    "You should always call 'reset()' once you receive 'done = True' -- any further steps are "
                    "undefined behavior.")


class CAEnv(ABC, _EnvBase):
    def __init__(self, nrows, ncols, debug=False, **kwargs):
        self.nrows, self.ncols = nrows, ncols
        self._debug = debug
        if not hasattr(self, "np_random"):
            self.np_random = np.random.default_rng()

    # -- what a concrete env supplies ---------------------------------------------------------------------
    @property
    @abstractmethod
    def MDP(self):
        raise NotImplementedError

    @property
    @abstractmethod
    def initial_state(self):
        self._resample_initial = False

    @abstractmethod
    def _award(self):
        raise NotImplementedError

    @abstractmethod
    def _is_done(self):
        raise NotImplementedError

    @abstractmethod
    def _report(self):
        raise NotImplementedError

    # -- the stateful loop -----------------------------------------------------------------------------------
    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        _EnvBase.reset(self, seed=seed)
        self.done, self.steps_beyond_done = False, 0
        self.steps_elapsed, self.reward_accumulated = 0, 0.0
        self._resample_initial = True
        self.grid, self.context = self.state = self.initial_state
        return self.state, self._report()

    def step(self, action):
        if self.done:
            if self.steps_beyond_done == 0:
                warnings.warn(_STEP_AFTER_DONE)
            self.steps_beyond_done += 1
            return self.state, 0.0, True, False, self._report()
        self.grid, self.context = self.state = self.MDP(self.grid, action, self.context)
        self._is_done()
        reward = self._award()
        self.steps_elapsed += 1
        self.reward_accumulated += reward
        return self.state, reward, self.done, False, self._report()

    def status(self):
        return {"steps_elapsed": self.steps_elapsed, "reward_accumulated": self.reward_accumulated}

    def count_cells(self, grid=None):
        cells = np.asarray(self.grid if grid is None else grid)
        return collections.Counter(cells.ravel().tolist())

"""RepeatCACUDA -- the clock operator (reference forest_fire/operators/repeat_ca_jax.py:12-71):
adds the action and state times to the accumulator, keeps the fractional part and applies the CA.
The reference applies exactly ONE CA update whatever ``repeats`` is (:61-63); ``substeps`` of the
wrapped CA's params generalises that to K updates."""
from __future__ import annotations

from typing import Callable

import torch

from ...operator import Operator


class RepeatCACUDA(Operator):
    grid_dependant = True
    action_dependant = True
    context_dependant = True

    def __init__(self, cellular_automaton, t_acting: Callable, t_perception: Callable, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.t_acting = t_acting
        self.t_perception = t_perception
        self.ca = cellular_automaton
        self.suboperators = (self.ca,)
        self.deterministic = self.ca.deterministic

    def update(self, grid, action, per_env_context, shared_context, accu_time):
        time_action = self.t_acting(action)
        time_state = self.t_perception((grid, per_env_context, shared_context))
        time_taken = time_action + time_state  # float32, same association as the reference
        new_accu = torch.as_tensor(accu_time, dtype=torch.float32, device=time_taken.device) + time_taken
        frac = new_accu - torch.trunc(new_accu)
        grid, new_per_env, _ = self.ca(grid, action, per_env_context, shared_context)
        return grid, (new_per_env, frac)

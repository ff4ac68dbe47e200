"""PartiallyObservableForestFireCUDA -- the stochastic fire-spread CA operator
(reference forest_fire/operators/ca_alexandridis_jax.py:47-460), batched, on B200.

``update(grid, action, per_env_context, shared_context) -> (grid, per_env_context,
shared_context)`` with the reference's argument meaning; every array carries a leading env axis.
The operator-level call marshals the reference float32/int32 layout into the packed device state
(gca_pack_state), runs gca_alexandridis_step and unpacks; the env keeps the packed state
resident and never pays that conversion on its hot path."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ... import _lib
from ..._lib import check, current_stream, load, ptr
from ...operator import Operator
from ...packed import PackedState, StepOutputs, make_inject, make_params
from ... import spaces
from ..._config import TYPE_BOX


def _as_dev(x, device, dtype=None):
    t = x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))
    return t.to(device=device, dtype=dtype) if dtype is not None else t.to(device)


class PartiallyObservableForestFireCUDA(Operator):
    grid_dependant = True
    action_dependant = False
    context_dependant = True
    deterministic = False

    def __init__(self, grid_size, empty, tree, fire, *args, params=None, use_hidden=True, rng_mode="legacy",
                 device="cuda", **kwargs):
        super().__init__(*args, **kwargs)
        self.grid_size = grid_size
        self.empty, self.tree, self.fire = empty, tree, fire
        self.device = torch.device(device)
        self.use_hidden = use_hidden
        self._params = params if params is not None else make_params(grid_size, grid_size, 1, rng_mode=rng_mode)
        # constants of __init__ (:57-65), exposed under the reference's attribute names
        self.initial_spread_time = grid_size + grid_size // 2
        self.fire_age_min = self.initial_spread_time * 1.5
        self.fire_age_max = self.initial_spread_time * 1.75
        self.burn_kernel_radius = int(self._params.R)
        if self.context_space is None:
            self.context_space = spaces.Box(0.0, 1.0, shape=(2,), dtype=TYPE_BOX)

    @property
    def params(self):
        return self._params

    def update(self, grid, action, per_env_context, shared_context, *, inject=None, substeps=None):
        d = self.device
        g = _as_dev(grid, d, torch.float32)
        single = g.dim() == 2
        # a plain copy that keeps lazily unpacked / host-lazy entries of the env's own context (dict(ctx) drops them)
        ctx = {k: (per_env_context[k].tensor() if hasattr(per_env_context[k], "tensor") else per_env_context[k])
               for k in per_env_context.keys()}
        if single:  # the reference's single-env signature: add the env axis
            g = g[None]
            ctx = {k: (_as_dev(v, d)[None] if k != "key" else _as_dev(v, d).reshape(1, 2)) for k, v in ctx.items()}
        N, H, W = g.shape
        ctx["true_grid"] = g
        P = self._params
        if substeps is not None and substeps != P.K:
            P = _lib.GcaParams.from_buffer_copy(P)
            P.K = int(substeps)
        if shared_context is not None and "p_tree" in shared_context:
            pt = float(shared_context["p_tree"])
            pw = float(shared_context["p_wind_change"])
            if pt != P.p_tree or pw != P.p_wind_change:
                P = _lib.GcaParams.from_buffer_copy(P)
                P.p_tree, P.p_wind_change = pt, pw
        st = PackedState(N, H, W, d, use_hidden=self.use_hidden)
        st.pack_from_reference(P, ctx)
        out = StepOutputs(N, d)
        inj, keep = make_inject(inject, d)
        flags = 0 if self.use_hidden else _lib.FLAG_NO_HIDDEN
        check(load().gca_alexandridis_step(C.byref(P), C.byref(st.cstruct()), C.byref(out.cstruct()),
                                           None if inj is None else C.byref(inj), flags, current_stream()),
              "gca_alexandridis_step")
        res = st.unpack_to_reference(P, want=("true_grid", "fire_age"))
        new_ctx = {k: per_env_context[k] for k in per_env_context.keys()}
        new_grid = res["true_grid"]
        new_ctx["fire_age"] = res["fire_age"][0] if single else res["fire_age"]
        new_ctx["wind_index"] = st.wind_index[0] if single else st.wind_index
        new_ctx["key"] = st.key[0] if single else st.key
        return (new_grid[0] if single else new_grid), new_ctx, shared_context

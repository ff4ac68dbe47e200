"""ORACLE SUPPORT (test infrastructure, not product code) -- run the reference's OWN Python source without jax.

The reference's hot path is written against ``jax`` / ``gymnasium``, neither of which can be installed in this
image.  This package provides NumPy-backed stand-ins for exactly the API surface those files
touch, so that ``tests/golden/make_reference_golden.py`` can import the operator and env source files where they
lie under ``/root/reference`` (nothing is copied) and execute them, producing golden input/output vectors that
the oracle -- and through it the CUDA path -- is then checked against.

What this pins: every line of reference *Python* on the path (rules, masks, ordering of operators, key schedule,
constants, clipping, dtype conversions, the ``.at[].set`` / ``where`` data flow).
What it cannot pin (no jax here): the bits ``jax.random`` produces for a key -- the shim delegates ``jax.random``
to ``oracle/prng.py``, which is pinned separately by the Random123 / JAX documentation known answers -- and XLA's
float32 evaluation details: reductions over window axes are done here in row-major sequential order from +0 (the
oracle's stated assumption), ``exp`` is NumPy's float32 ``exp``.

``install()`` puts the stand-ins into ``sys.modules`` (jax, jax.numpy, jax.random, jax.lax, gymnasium, ...) and
registers empty parent packages for ``gym_cellular_automata`` so that sub-modules import from the reference tree
without running the package ``__init__`` (which pulls in matplotlib and every other env).  Only the golden
generator and the optional live test call it; it refuses to run when a real ``jax`` is importable.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

from . import jnp_shim, jax_shim, gymnasium_shim

REFERENCE_ROOT = os.environ.get("GCA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_cellular_automata"))


def install(rng_mode: int = 0):
    """Register the stand-ins.  ``rng_mode``: oracle.prng.LEGACY (0) or PARTITIONABLE (1) stream layout."""
    if importlib.util.find_spec("jax") is not None and not getattr(sys.modules.get("jax"), "_GCA_SHIM", False):
        raise RuntimeError("a real jax is importable: run the reference itself instead of the shim")
    jax_shim.set_rng_mode(rng_mode)
    jax = jax_shim.build_module()
    sys.modules["jax"] = jax
    sys.modules["jax.numpy"] = jax.numpy
    sys.modules["jax.random"] = jax.random
    sys.modules["jax.lax"] = jax.lax
    sys.modules["jax.debug"] = jax.debug
    for name, mod in gymnasium_shim.build_modules().items():
        sys.modules[name] = mod
    if "flax" not in sys.modules:  # flax.struct.dataclass: a frozen dataclass with .replace
        import dataclasses
        flax = types.ModuleType("flax")
        flax._GCA_SHIM = True
        struct = types.ModuleType("flax.struct")

        def _dataclass(cls):
            d = dataclasses.dataclass(frozen=True)(cls)
            d.replace = lambda self, **kw: dataclasses.replace(self, **kw)
            return d
        struct.dataclass = _dataclass
        flax.struct = struct
        sys.modules["flax"], sys.modules["flax.struct"] = flax, struct
    pkg_root = os.path.join(REFERENCE_ROOT, "gym_cellular_automata")
    for name, sub in (("gym_cellular_automata", ""), ("gym_cellular_automata.forest_fire", "forest_fire"),
                      ("gym_cellular_automata.forest_fire.operators", "forest_fire/operators"),
                      ("gym_cellular_automata.forest_fire.bulldozer", "forest_fire/bulldozer"),
                      ("gym_cellular_automata.forest_fire.bulldozer.utils", "forest_fire/bulldozer/utils")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(pkg_root, sub)]
            m._GCA_SHIM = True
            sys.modules[name] = m
    # the matplotlib renderer is not on the path
    rname = "gym_cellular_automata.forest_fire.bulldozer.utils.advanced_bulldozer_render"
    if rname not in sys.modules:
        r = types.ModuleType(rname)
        r.render = r.plot_grid_attribute = None
        sys.modules[rname] = r
    # what the operators package __init__ would export to advanced_bulldozer.py
    ops = sys.modules["gym_cellular_automata.forest_fire.operators"]
    for mod, names in (("ca_alexandridis_jax", ("PartiallyObservableForestFireJax",)),
                       ("move_modify_jax", ("MoveJax", "ModifyJax", "MoveModifyJax")),
                       ("repeat_ca_jax", ("RepeatCAJax",))):
        m = load("forest_fire.operators." + mod)
        for n in names:
            setattr(ops, n, getattr(m, n))
    return jax


def load(module: str):
    """Import ``gym_cellular_automata.<module>`` from the reference tree (after ``install()``)."""
    return importlib.import_module("gym_cellular_automata." + module)

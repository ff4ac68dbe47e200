"""ORACLE SUPPORT (test infrastructure) -- stand-ins for ``jax`` (jit, vmap, lax, random, debug) on NumPy.

``jax.random`` delegates to ``oracle/prng.py`` (threefry2x32, both stream layouts).  ``vmap`` is a plain Python
loop over axis 0 that maps pytrees (tuples / lists / dicts of arrays) in and stacks them out; ``jit`` is the
identity; ``lax.cond`` / ``lax.switch`` evaluate only the branch taken (results are the same as tracing both)."""
from __future__ import annotations

import types

import numpy as np

from .. import prng
from . import jnp_shim as jnp

_MODE = prng.LEGACY


def set_rng_mode(mode: int):
    global _MODE
    _MODE = mode


# ---------------------------------------------------------------- pytrees
def _tree_map(f, t):
    if isinstance(t, dict):
        return {k: _tree_map(f, v) for k, v in t.items()}
    if isinstance(t, tuple):
        return tuple(_tree_map(f, v) for v in t)
    if isinstance(t, list):
        return [_tree_map(f, v) for v in t]
    return f(t)


def _first_leaf(t):
    if isinstance(t, dict):
        for v in t.values():
            r = _first_leaf(v)
            if r is not None:
                return r
        return None
    if isinstance(t, (tuple, list)):
        for v in t:
            r = _first_leaf(v)
            if r is not None:
                return r
        return None
    return t


def _tree_stack(items):
    h = items[0]
    if isinstance(h, dict):
        return {k: _tree_stack([it[k] for it in items]) for k in h}
    if isinstance(h, tuple):
        return tuple(_tree_stack([it[i] for it in items]) for i in range(len(h)))
    if isinstance(h, list):
        return [_tree_stack([it[i] for it in items]) for i in range(len(h))]
    return jnp.asarray(np.stack([np.asarray(it) for it in items], axis=0))


def _slice_arg(a, ax, i):
    """in_axes entry ``ax`` may be None, 0, or a pytree prefix (dict / tuple) of such entries."""
    if ax is None:
        return a
    if isinstance(ax, dict):
        return {k: _slice_arg(a[k], ax.get(k, 0), i) for k in a}
    if isinstance(ax, (tuple, list)) and isinstance(a, (tuple, list)):
        return type(a)(_slice_arg(v, x, i) for v, x in zip(a, ax))
    if ax != 0:
        raise NotImplementedError("vmap shim: only axis 0")
    return _tree_map(lambda x: jnp.asarray(x)[i], a)


def _batch_size(a, ax):
    if ax is None:
        return None
    if isinstance(ax, dict):
        for k in a:
            n = _batch_size(a[k], ax.get(k, 0))
            if n is not None:
                return n
        return None
    if isinstance(ax, (tuple, list)) and isinstance(a, (tuple, list)):
        for v, x in zip(a, ax):
            n = _batch_size(v, x)
            if n is not None:
                return n
        return None
    leaf = _first_leaf(a)
    return None if leaf is None else np.asarray(leaf).shape[0]


def vmap(f, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            n = _batch_size(a, ax)
            if n is not None:
                break
        if n is None:
            raise ValueError("vmap shim: nothing to map over")
        outs = [f(*[_slice_arg(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)]
        return _tree_stack(outs)
    return mapped


def jit(f=None, **kwargs):
    if f is None:
        return lambda g: g
    return f


# ---------------------------------------------------------------- lax
def _dynamic_slice(operand, start_indices, slice_sizes):
    a = np.asarray(operand)
    idx = []
    for s, size, dim in zip(start_indices, slice_sizes, a.shape):
        s = int(np.asarray(s))
        s = min(max(s, 0), dim - size)  # lax.dynamic_slice clamps the start so the slice stays in bounds
        idx.append(slice(s, s + size))
    return jnp.asarray(a[tuple(idx)])


def _fori_loop(lo, hi, body, init):
    v = init
    for i in range(int(np.asarray(lo)), int(np.asarray(hi))):
        v = body(i, v)
    return v


def _scan(f, init, xs=None, length=None):
    n = int(length) if xs is None else int(np.asarray(_first_leaf(xs)).shape[0])
    carry, ys = init, []
    for i in range(n):
        x = None if xs is None else _tree_map(lambda a: jnp.asarray(a)[i], xs)
        carry, y = f(carry, x)
        ys.append(y)
    return carry, (None if not ys or ys[0] is None else _tree_stack(ys))


def _cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(np.asarray(pred)) else false_fun(*operands)


def _switch(index, branches, *operands):
    i = min(max(int(np.asarray(index)), 0), len(branches) - 1)
    return branches[i](*operands)


# ---------------------------------------------------------------- random
def _n(shape):
    if shape is None:
        return 1, ()
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    shape = tuple(int(s) for s in shape)
    n = 1
    for s in shape:
        n *= s
    return n, shape


def _key(key):
    return np.asarray(key, dtype=np.uint32).reshape(2)


def _split(key, num=2):
    return jnp.asarray(prng.split(_key(key), int(num), _MODE))


def _uniform(key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0):
    n, shape = _n(shape)
    u = prng.uniform(_key(key), n, _MODE).reshape(shape)
    if minval != 0.0 or maxval != 1.0:
        u = (u * np.float32(maxval - minval) + np.float32(minval)).astype(np.float32)
    return jnp.asarray(u)


def _randint(key, shape, minval, maxval, dtype=np.int32):
    n, shape = _n(shape)
    return jnp.asarray(prng.randint(_key(key), n, minval, maxval, _MODE).reshape(shape))


def _PRNGKey(seed):
    return jnp.asarray(prng.key_from_seed(int(seed)))


def _unsupported(name):
    def f(*a, **k):
        raise NotImplementedError(f"jax shim: {name} is not on the hot path and is not provided")
    return f


def build_module():
    jax = types.ModuleType("jax")
    jax._GCA_SHIM = True
    jax.numpy = jnp
    jax.jit = jit
    jax.vmap = vmap
    lax = types.ModuleType("jax.lax")
    lax.dynamic_slice = _dynamic_slice
    lax.fori_loop = _fori_loop
    lax.cond = _cond
    lax.scan = _scan
    lax.switch = _switch
    jax.lax = lax
    random = types.ModuleType("jax.random")
    random.split = _split
    random.uniform = _uniform
    random.randint = _randint
    random.PRNGKey = _PRNGKey
    random.key = _PRNGKey
    random.poisson = _unsupported("random.poisson")
    random.normal = _unsupported("random.normal")
    random.choice = _unsupported("random.choice")
    jax.random = random
    debug = types.ModuleType("jax.debug")
    debug.callback = lambda *a, **k: None
    debug.print = lambda *a, **k: None
    jax.debug = debug
    tree_util = types.ModuleType("jax.tree_util")
    tree_util.tree_map = lambda f, t, *rest: _tree_map(f, t) if not rest else _unsupported("tree_map with rest")()
    jax.tree_util = tree_util
    jax.tree_map = tree_util.tree_map
    jax.Array = jnp.JArr
    jax.devices = lambda *a: ["cpu-shim"]
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    return jax

"""ORACLE SUPPORT (test infrastructure) -- the part of ``jax.numpy`` the reference's hot path uses, on NumPy.

JAX semantics kept: 32-bit default dtypes (x64 disabled: float64 -> float32, int64 -> int32, uint64 -> uint32 on
every result), immutable-style ``arr.at[idx].set(v)``, weak Python scalars (NumPy >= 2 has the same rule),
reductions of float32 over several trailing axes in row-major sequential order (see the package docstring)."""
from __future__ import annotations

import numpy as np

_DOWN = {np.dtype(np.float64): np.float32, np.dtype(np.int64): np.int32, np.dtype(np.uint64): np.uint32}


def _down(x):
    if isinstance(x, np.ndarray):
        t = _DOWN.get(x.dtype)
        if t is not None:
            x = x.astype(t)
        return x.view(JArr)
    if isinstance(x, np.generic):
        t = _DOWN.get(x.dtype)
        return np.asarray(x if t is None else x.astype(t)).view(JArr)
    if isinstance(x, tuple):
        return tuple(_down(v) for v in x)
    if isinstance(x, list):
        return [_down(v) for v in x]
    return x


class _AtIndex:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, value):
        out = np.array(self.arr, copy=True)
        out[self.idx] = value
        return _down(out)

    def add(self, value):
        out = np.array(self.arr, copy=True)
        np.add.at(out, self.idx, value)
        return _down(out)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self.arr, idx)


class JArr(np.ndarray):
    """ndarray with ``.at`` and JAX's 32-bit result types."""

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        ins = tuple(np.asarray(i) if isinstance(i, JArr) else i for i in inputs)
        if out is not None:
            kwargs["out"] = tuple(np.asarray(o) if isinstance(o, JArr) else o for o in out)
        res = getattr(ufunc, method)(*ins, **kwargs)
        return _down(res)

    @property
    def at(self):
        return _At(self)

    def __getitem__(self, idx):
        if isinstance(idx, JArr):
            idx = np.asarray(idx)
        elif isinstance(idx, tuple):
            idx = tuple(np.asarray(i) if isinstance(i, JArr) else i for i in idx)
        return _down(np.asarray(self)[idx])

    def __iter__(self):
        a = np.asarray(self)
        if a.ndim == 0:
            raise TypeError("iteration over a 0-d array")
        return (_down(a[i]) for i in range(a.shape[0]))

    def __bool__(self):
        return bool(np.asarray(self))

    def __int__(self):
        return int(np.asarray(self))

    def __float__(self):
        return float(np.asarray(self))

    def __index__(self):
        return int(np.asarray(self))

    def __hash__(self):
        return id(self)

    def astype(self, dtype, *a, **k):
        return _down(np.asarray(self).astype(dtype, *a, **k))

    def reshape(self, *shape, **k):
        return _down(np.asarray(self).reshape(*shape, **k))

    def sum(self, axis=None, dtype=None, **k):
        return sum(self, axis=axis, dtype=dtype)

    def any(self, axis=None, **k):
        return _down(np.asarray(self).any(axis=axis))

    def all(self, axis=None, **k):
        return _down(np.asarray(self).all(axis=axis))

    def max(self, axis=None, **k):
        return _down(np.asarray(self).max(axis=axis))

    def min(self, axis=None, **k):
        return _down(np.asarray(self).min(axis=axis))

    def block_until_ready(self):
        return self


ndarray = JArr
float32, int32, uint32, uint8, bool_, int8, float16 = np.float32, np.int32, np.uint32, np.uint8, np.bool_, np.int8, np.float16
pi = np.pi
newaxis = None


def _raw(x):
    if isinstance(x, JArr):
        return np.asarray(x)
    if isinstance(x, (list, tuple)):
        return type(x)(_raw(v) for v in x)
    return x


def array(x, dtype=None, **k):
    return _down(np.array(_raw(x), dtype=dtype))


def asarray(x, dtype=None, **k):
    return _down(np.asarray(_raw(x), dtype=dtype))


def sum(x, axis=None, dtype=None, **k):  # noqa: A001 (mirrors jnp.sum)
    a = np.asarray(x)
    if a.dtype == np.bool_:
        a = a.astype(np.int32)
    if a.dtype.kind != "f":
        return _down(a.sum(axis=axis, dtype=dtype))
    # float: sequential row-major accumulation from +0 over the reduced axes (kept last), in the array's precision
    if axis is None:
        axes = tuple(range(a.ndim))
    else:
        axes = tuple(sorted(ax % a.ndim for ax in (axis if isinstance(axis, (tuple, list)) else (axis,))))
    keep = [i for i in range(a.ndim) if i not in axes]
    t = np.transpose(a, keep + list(axes)).reshape(tuple(a.shape[i] for i in keep) + (-1,))
    acc = np.zeros(t.shape[:-1], dtype=a.dtype)
    for j in range(t.shape[-1]):
        acc = (acc + t[..., j]).astype(a.dtype)
    return _down(acc)


def _wrap(fn):
    def f(*args, **kwargs):
        return _down(fn(*[_raw(a) for a in args], **{k: _raw(v) for k, v in kwargs.items()}))
    f.__name__ = fn.__name__
    return f


where = _wrap(np.where)
clip = _wrap(np.clip)
minimum = _wrap(np.minimum)
maximum = _wrap(np.maximum)
exp = _wrap(np.exp)
sin = _wrap(np.sin)
cos = _wrap(np.cos)
tanh = _wrap(np.tanh)
round = _wrap(np.round)  # noqa: A001  half-to-even, like jnp.round
any = _wrap(np.any)  # noqa: A001
all = _wrap(np.all)  # noqa: A001
max = _wrap(np.max)  # noqa: A001
min = _wrap(np.min)  # noqa: A001
argmax = _wrap(np.argmax)
stack = _wrap(np.stack)
concatenate = _wrap(np.concatenate)
zeros_like = _wrap(np.zeros_like)
ones_like = _wrap(np.ones_like)
invert = _wrap(np.invert)
take = _wrap(np.take)
dot = _wrap(np.dot)
linspace = _wrap(np.linspace)
expand_dims = _wrap(np.expand_dims)
squeeze = _wrap(np.squeeze)
logical_and = _wrap(np.logical_and)
logical_or = _wrap(np.logical_or)
logical_not = _wrap(np.logical_not)
abs = _wrap(np.abs)  # noqa: A001
floor = _wrap(np.floor)
mod = _wrap(np.mod)
append = _wrap(np.append)


def modf(x):
    f, i = np.modf(_raw(x))
    return _down(f), _down(i)


def meshgrid(*xs, indexing="xy"):
    return [_down(m) for m in np.meshgrid(*[_raw(x) for x in xs], indexing=indexing)]


def zeros(shape, dtype=None):
    return _down(np.zeros(shape, dtype=np.float32 if dtype is None else dtype))


def ones(shape, dtype=None):
    return _down(np.ones(shape, dtype=np.float32 if dtype is None else dtype))


def full(shape, fill_value, dtype=None):
    return _down(np.full(shape, _raw(fill_value), dtype=dtype))


def arange(*a, dtype=None):
    return _down(np.arange(*a, dtype=dtype))


def pad(x, pad_width, mode="constant", **k):
    k = {kk: _raw(v) for kk, v in k.items()}
    return _down(np.pad(_raw(x), pad_width, mode=mode, **k))

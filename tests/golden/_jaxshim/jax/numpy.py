"""jax.numpy -> numpy (float64 by default, like jax_enable_x64)."""
from numpy import *  # noqa: F401,F403
import numpy as _np
from numpy import linalg, number, integer, issubdtype  # noqa: F401

ndarray = _np.ndarray


class ShimArray(_np.ndarray):
    """ndarray with JAX's functional-update accessor: `x.at[idx].set(v)` / `.add(v)` return a NEW array
    (tatva/lifter uses them, lifter/base.py:216, constraints.py:216-220, :314-318).  Numerics are NumPy's."""

    @property
    def at(self):
        return _At(self)


class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self._arr, idx)


class _AtIndex:
    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def set(self, value):
        out = _np.array(self._arr, copy=True).view(ShimArray)
        out[self._idx] = value
        return out

    def add(self, value):
        out = _np.array(self._arr, copy=True).view(ShimArray)
        _np.add.at(out, self._idx, value)  # duplicates accumulate, like JAX's scatter-add
        return out


def _wrap(a):
    return a.view(ShimArray) if isinstance(a, _np.ndarray) and a.ndim > 0 else a


def array(obj, dtype=None, **kw):
    return _wrap(_np.array(obj, dtype=dtype))


def asarray(obj, dtype=None, **kw):
    return _wrap(_np.asarray(obj, dtype=dtype))


def zeros(shape, dtype=None, **kw):
    return _wrap(_np.zeros(shape, dtype=dtype if dtype is not None else _np.float64))


def zeros_like(a, dtype=None, **kw):
    return _wrap(_np.zeros_like(a, dtype=dtype))


def repeat(a, repeats, axis=None, total_repeat_length=None):
    return _np.repeat(a, repeats, axis=axis)

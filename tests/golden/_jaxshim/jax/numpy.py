"""jax.numpy -> numpy (float64 by default, like jax_enable_x64)."""
from numpy import *  # noqa: F401,F403
import numpy as _np
from numpy import linalg, number, integer, issubdtype  # noqa: F401

ndarray = _np.ndarray


def array(obj, dtype=None, **kw):
    return _np.array(obj, dtype=dtype)


def asarray(obj, dtype=None, **kw):
    return _np.asarray(obj, dtype=dtype)


def repeat(a, repeats, axis=None, total_repeat_length=None):
    return _np.repeat(a, repeats, axis=axis)

"""NumPy-backed stand-in for the handful of `jax` names tatva's hot path imports.

TEST INFRASTRUCTURE ONLY.  JAX is not installable in this image (no network), so the
reference at /root/reference cannot be imported as is.  This shim lets the *unmodified*
reference modules (tatva.element, tatva.mesh, tatva.operator, tatva.sparse._extraction,
tatva.sparse._coloring, tatva.compound, tatva.mpi) execute eagerly on NumPy so that
tests/golden/make_golden.py can record their outputs as fixtures.  Nothing in the
product package imports this.  Semantics reproduced: vmap == stack of per-item calls,
lax.map == the same, jit == identity.  No autodiff (derivative goldens use the
complex-step method on the reference's own energy instead); `jacrev` is the complex-step Jacobian with respect
to the first argument, which is all Operator.interpolate asks of it (operator.py:419-421).
"""
import numpy as _np

from . import numpy  # noqa: F401
from . import lax, tree_util, errors, typing, core  # noqa: F401

Array = _np.ndarray


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def jit(fn=None, **kwargs):
    if fn is None:
        return lambda f: f
    return fn


def _tree_map(f, tree):
    if isinstance(tree, (tuple, list)):
        return type(tree)(_tree_map(f, t) for t in tree)
    return f(tree)


def _tree_stack(items):
    first = items[0]
    if isinstance(first, (tuple, list)):
        return type(first)(_tree_stack([it[i] for it in items]) for i in range(len(first)))
    return _np.stack([_np.asarray(it) for it in items])


def vmap(fn, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.asarray(a).shape[ax]
                break
        outs = []
        for i in range(n):
            call = [a if ax is None else _np.take(_np.asarray(a), i, axis=ax) for a, ax in zip(args, axes)]
            outs.append(fn(*call))
        return _tree_stack(outs)

    return mapped


def jacrev(fn, argnums=0):
    """Complex-step Jacobian wrt argument `argnums` (exact to rounding for the polynomial shape functions)."""
    if argnums != 0:
        raise NotImplementedError("shim jacrev: argnums=0 only")

    def jac(x, *rest):
        x = _np.asarray(x, dtype=_np.float64)
        h = 1e-30
        cols = []
        for j in range(x.size):
            xp = x.astype(_np.complex128).ravel()
            xp[j] += 1j * h
            cols.append(_tree_map(lambda o: _np.imag(_np.asarray(o)) / h, fn(xp.reshape(x.shape), *rest)))
        return _tree_map_stack_last(cols, x.shape)

    return jac


def _tree_map_stack_last(cols, xshape):
    first = cols[0]
    if isinstance(first, (tuple, list)):
        return type(first)(_tree_map_stack_last([c[i] for c in cols], xshape) for i in range(len(first)))
    st = _np.stack([_np.asarray(c) for c in cols], axis=-1)
    return st.reshape(st.shape[:-1] + tuple(xshape))


def jvp(fn, primals, tangents):
    """Complex-step directional derivative (exact to rounding for real-analytic `fn`): what sparse.jacfwd asks of
    jax.jvp (sparse/base.py:264)."""
    (x,), (v,) = primals, tangents
    h = 1e-30
    out = fn(_np.asarray(x, dtype=_np.float64).astype(_np.complex128) + 1j * h * _np.asarray(v, dtype=_np.float64))
    return _np.real(out), _np.imag(out) / h

"""Helpers of tatva/utils.py on torch tensors.

`virtual_work_to_residual` (utils.py:39-115) turns a virtual-work functional that is linear in its first argument (the
test function) into the residual d(fn)/d(test); the reference takes `jax.jacrev`, here it is one reverse pass of
`torch.autograd` (the adjoint kernels of `Operator.grad / eval / integrate` when `fn` is written with them).
`make_project_function` (utils.py:118-183) wraps `Operator.project`; `create_g2l` is re-exported from the lifter.
"""
from __future__ import annotations

from functools import wraps
from typing import Callable

import torch

from .lifter import create_g2l  # noqa: F401  (tatva/utils.py:265-280)


def virtual_work_to_residual(fn: Callable | None = None, /, *, test_arr=None, test_shape=None, test_size=None, jit: bool = False, device=None) -> Callable:
    """residual(*args) = d fn(test, *args) / d test at `test_arr` (zeros of `test_shape` / `test_size` otherwise).
    Usable directly or as a decorator, like the reference; `jit` is accepted and ignored (no tracing compiler).
    The result keeps its autograd graph, so it can be differentiated again (tangent by `sparse.jacfwd`)."""
    if test_arr is not None:
        base = torch.as_tensor(test_arr, dtype=torch.float64, device=device)
    elif test_shape is not None:
        base = torch.zeros(tuple(test_shape), dtype=torch.float64, device=device)
    elif test_size is not None:
        base = torch.zeros(int(test_size), dtype=torch.float64, device=device)
    else:
        raise ValueError("One of 'test_arr', 'test_shape', or 'test_size' must be provided.")
    if fn is None:
        return lambda f: virtual_work_to_residual(f, test_arr=base, jit=jit, device=device)

    @wraps(fn)
    def wrapper(*args, **kwargs):
        dev = device
        if dev is None:
            dev = next((a.device for a in args if isinstance(a, torch.Tensor)), base.device)
        test = base.to(dev).detach().clone().requires_grad_(True)
        with torch.enable_grad():
            work = fn(test, *args, **kwargs)
            if work.ndim != 0:
                raise ValueError("the virtual work must be a scalar")
            needs_graph = any(isinstance(a, torch.Tensor) and a.requires_grad for a in args)
            (res,) = torch.autograd.grad(work, test, create_graph=needs_graph, allow_unused=True)
        return torch.zeros_like(test) if res is None else res

    return wrapper


def make_project_function(nnodes: int, colored_matrix=None, elements=None, lifter=None) -> Callable:
    """Factory of tatva/utils.py:118-183: returns project(op, field) -> nodal values.  The mass system is solved
    matrix-free on the device (`Operator.project`), so `elements` is only checked for presence like the reference does."""
    if colored_matrix is None and elements is None:
        raise ValueError("Must provide elements if colored_matrix is not provided")
    if colored_matrix is not None and lifter is not None and colored_matrix.shape[0] != lifter.size_reduced:
        raise ValueError(f"Colored matrix size does not match lifter reduced size. Expected {lifter.size_reduced}, got {colored_matrix.shape[0]}")

    def _project(op, field):
        if op.n_nodes != nnodes:
            raise ValueError(f"operator has {op.n_nodes} nodes, the projection was built for {nnodes}")
        return op.project(field, colored_matrix, lifter)

    return _project


__all__ = ["virtual_work_to_residual", "make_project_function", "create_g2l"]

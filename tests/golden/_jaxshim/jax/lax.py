import numpy as _np


def map(f, xs, batch_size=None):  # noqa: A001
    from . import _tree_stack

    if isinstance(xs, (tuple, list)):
        n = len(xs[0])
        items = [type(xs)(x[i] for x in xs) for i in range(n)]
    else:
        n = len(xs)
        items = [xs[i] for i in range(n)]
    return _tree_stack([f(it) for it in items])


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)

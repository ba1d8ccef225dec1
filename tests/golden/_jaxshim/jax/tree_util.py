def register_dataclass(cls=None, **kw):
    if cls is None:
        return lambda c: c
    return cls


def register_pytree_node_class(cls):
    return cls

"""Placeholder classes so `tatva.sparse.tracer` (out of scope, never called) imports."""


class Jaxpr:
    pass


class JaxprEqn:
    pass


class Literal:
    pass


class Var:
    pass


class Primitive:
    def __init__(self, *a, **k):
        pass

class Tracer:
    pass

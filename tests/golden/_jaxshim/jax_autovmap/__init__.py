"""Stand-in for jax_autovmap.autovmap: broadcast a per-point function over leading axes."""
import numpy as np


def autovmap(*dargs, **ranks):
    def deco(fn):
        import inspect

        names = list(inspect.signature(fn).parameters)

        def wrapped(*args, **kwargs):
            bound = dict(zip(names, args))
            bound.update(kwargs)
            lead = ()
            for k, v in bound.items():
                r = ranks.get(k)
                if r is None:
                    continue
                shp = np.shape(v)
                l = shp[: len(shp) - r]
                if len(l) > len(lead):
                    lead = l
            if not lead:
                return fn(**bound)
            out = np.empty(lead, dtype=object)
            for idx in np.ndindex(*lead):
                call = {}
                for k, v in bound.items():
                    r = ranks.get(k)
                    if r is None:
                        call[k] = v
                        continue
                    shp = np.shape(v)
                    l = shp[: len(shp) - r]
                    call[k] = np.asarray(v)[idx[len(lead) - len(l):]] if l else v
                out[idx] = fn(**call)
            flat = [np.asarray(o) for o in out.ravel()]
            return np.stack(flat).reshape(lead + flat[0].shape)

        return wrapped

    if dargs and callable(dargs[0]):
        return deco(dargs[0])
    return deco

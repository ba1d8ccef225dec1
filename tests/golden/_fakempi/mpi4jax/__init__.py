"""Importable placeholder for mpi4jax (tatva/mpi.py imports it at module level; the fixtures only build the static
layouts and routing tables, which never call it)."""


def sendrecv(*a, **k):
    raise NotImplementedError("mpi4jax.sendrecv is not available in the golden-fixture stand-in")


def allreduce(*a, **k):
    raise NotImplementedError("mpi4jax.allreduce is not available in the golden-fixture stand-in")

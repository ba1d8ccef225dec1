"""Stand-in for the two mpi4jax calls tatva/mpi.py makes inside its (here un-jitted) exchange functions, on the
thread-based mpi4py stand-in.  TEST INFRASTRUCTURE ONLY."""
import numpy as np


def sendrecv(sendbuf, recvbuf, source, dest, comm=None, **_):
    out = np.empty_like(np.asarray(recvbuf))
    comm.Sendrecv(sendbuf=np.ascontiguousarray(sendbuf), dest=int(dest), recvbuf=out, source=int(source))
    return out


def allreduce(x, op=None, comm=None, **_):
    out = np.empty_like(np.asarray(x))
    comm.Allreduce(np.ascontiguousarray(x), out, op=op)
    return out

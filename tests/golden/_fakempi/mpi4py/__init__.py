"""Thread-based stand-in for the part of mpi4py that tatva/mpi.py uses, so that the UNMODIFIED reference plans can be
built for several ranks inside one process (tests/golden/make_golden.py).  TEST INFRASTRUCTURE ONLY.

Every rank is a thread holding a `Comm`; collectives rendezvous on a shared barrier, point-to-point `Sendrecv` goes
through per-(source, dest) mailboxes.  Only blocking semantics and the calls listed below exist: rank / Get_rank /
Get_size / allreduce / allgather / exscan / Allreduce / Alltoall / Sendrecv / Allgatherv / Barrier."""
from __future__ import annotations

import queue
import threading

import numpy as np


class _Op:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn


class _World:
    def __init__(self, size):
        self.size = size
        self.barrier = threading.Barrier(size)
        self.slots = [None] * size
        self.mail = {(s, d): queue.Queue() for s in range(size) for d in range(size)}


class Comm:
    def __init__(self, world: _World, rank: int):
        self._w, self.rank, self.size = world, rank, world.size

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def Barrier(self):
        self._w.barrier.wait()

    def allgather(self, obj):
        w = self._w
        w.slots[self.rank] = obj
        w.barrier.wait()
        out = list(w.slots)
        w.barrier.wait()  # nobody overwrites a slot before everybody has read
        return out

    def allreduce(self, value, op=None):
        vals = self.allgather(value)
        return (op or MPI.SUM).fn(vals)

    def exscan(self, value, op=None):
        """Exclusive prefix reduction: rank 0 gets None (as mpi4py does), rank r the reduction over ranks < r."""
        vals = self.allgather(value)
        return None if self.rank == 0 else (op or MPI.SUM).fn(vals[: self.rank])

    def Allreduce(self, sendbuf, recvbuf, op=None):
        vals = self.allgather(np.array(sendbuf, copy=True))
        recvbuf[...] = (op or MPI.SUM).fn(vals)

    def Alltoall(self, sendbuf, recvbuf):
        rows = self.allgather(np.array(sendbuf, copy=True))
        recvbuf[...] = np.array([rows[src][self.rank] for src in range(self.size)], dtype=recvbuf.dtype)

    def Sendrecv(self, sendbuf, dest, recvbuf, source, **_):
        self._w.mail[(self.rank, dest)].put(np.array(sendbuf, copy=True))
        got = self._w.mail[(source, self.rank)].get(timeout=60)
        recvbuf[...] = got

    def Allgatherv(self, sendbuf, recv):
        recvbuf, counts = recv[0], recv[1]
        parts = self.allgather(np.array(sendbuf, copy=True))
        recvbuf[...] = np.concatenate(parts)


class _MPI:
    SUM = _Op("SUM", lambda vals: sum(vals[1:], vals[0]) if not isinstance(vals[0], np.ndarray) else np.sum(vals, axis=0))
    MAX = _Op("MAX", lambda vals: max(vals) if not isinstance(vals[0], np.ndarray) else np.max(vals, axis=0))
    Comm = Comm
    COMM_WORLD = None


MPI = _MPI()


def run_ranks(size, fn):
    """Run fn(comm) on `size` threads; returns the per-rank results (re-raises the first exception)."""
    world = _World(size)
    results, errors = [None] * size, []

    def target(r):
        try:
            results[r] = fn(Comm(world, r))
        except BaseException as exc:  # noqa: BLE001
            errors.append(exc)
            world.barrier.abort()

    threads = [threading.Thread(target=target, args=(r,)) for r in range(size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results

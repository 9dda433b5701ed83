"""In-process multi-rank `pace.util.Comm` with real data exchange (one Python thread per rank).

Test infrastructure for running the reference as the oracle.
"""
import copy
import functools
import queue
import threading

import numpy as np

from . import shim  # noqa: F401  (must precede pace imports)
from pace.util.comm import Comm, Request  # noqa: E402


class _Req(Request):
    def __init__(self, fn):
        self._fn = fn

    def wait(self):
        return self._fn()


class World:
    def __init__(self, n):
        self.n = n
        self.q = {}
        self.lock = threading.Lock()
        self.barrier = threading.Barrier(n)
        self.slots = [None] * n
        self.subs = {}

    def chan(self, key):
        with self.lock:
            return self.q.setdefault(key, queue.Queue())


class ThreadComm(Comm):
    def __init__(self, world, rank):
        self.w = world
        self.rank = rank

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.w.n

    def barrier(self):
        self.w.barrier.wait()

    Barrier = barrier

    def _exchange(self, obj):
        self.w.slots[self.rank] = copy.deepcopy(obj)
        self.w.barrier.wait()
        out = list(self.w.slots)
        self.w.barrier.wait()
        return out

    def bcast(self, value, root=0):
        return self._exchange(value)[root]

    def allgather(self, sendobj):
        return self._exchange(sendobj)

    def allreduce(self, sendobj, op=None):
        vals = self._exchange(sendobj)
        if op is None:
            return sum(vals)
        return functools.reduce(op, vals)

    def Send(self, sendbuf, dest, tag=0, **kw):
        self.w.chan((self.rank, dest, tag)).put(np.array(sendbuf, copy=True))

    def Isend(self, sendbuf, dest, tag=0, **kw):
        self.Send(sendbuf, dest, tag)
        return _Req(lambda: None)

    def Recv(self, recvbuf, source, tag=0, **kw):
        recvbuf[...] = self.w.chan((source, self.rank, tag)).get(timeout=600).reshape(recvbuf.shape)

    def Irecv(self, recvbuf, source, tag=0, **kw):
        return _Req(lambda: self.Recv(recvbuf, source, tag))

    def sendrecv(self, sendbuf, dest, **kw):
        raise NotImplementedError

    def Scatter(self, *a, **k):
        raise NotImplementedError

    def Gather(self, *a, **k):
        raise NotImplementedError

    def Split(self, color, key):
        info = self._exchange((color, key, self.rank))
        members = sorted([(k, r) for c, k, r in info if c == color])
        ranks = [r for _, r in members]
        with self.w.lock:
            sub = self.w.subs.setdefault(tuple(ranks), World(len(ranks)))
        return ThreadComm(sub, ranks.index(self.rank))

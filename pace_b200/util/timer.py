"""Timer / NullTimer (util/pace/util/_timing.py:6-100): the `timer.clock(name)` context used by step_dynamics."""
import contextlib
import time

import torch


class Timer:
    def __init__(self, sync_cuda: bool = True):
        self._times = {}
        self._hits = {}
        self._sync = sync_cuda

    @contextlib.contextmanager
    def clock(self, name: str):
        if self._sync and torch.cuda.is_available():
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_push(name)
        t0 = time.perf_counter()
        try:
            yield
        finally:
            if self._sync and torch.cuda.is_available():
                torch.cuda.synchronize()
                torch.cuda.nvtx.range_pop()
            self._times[name] = self._times.get(name, 0.0) + time.perf_counter() - t0
            self._hits[name] = self._hits.get(name, 0) + 1

    @property
    def times(self):
        return dict(self._times)

    @property
    def hits(self):
        return dict(self._hits)

    def reset(self):
        self._times.clear()
        self._hits.clear()


class NullTimer(Timer):
    @contextlib.contextmanager
    def clock(self, name: str):
        yield

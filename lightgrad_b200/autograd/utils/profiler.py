"""Per-op wall-clock accounting (API of lightgrad/autograd/utils/profiler.py:5-40).

``Profiler`` is a context manager that accumulates time and call counts per
operator name, separately for forward and backward; ``Tracker`` brackets one
operator call and only the outermost tracker reports (profiler.py:30-40).

CUDA launches are asynchronous, so a tracker that closes while a ``Profiler``
is active first drains the device through ``Profiler.device_sync`` (installed
by the cuda backend) -- otherwise the time of a kernel would be billed to
whichever later op happens to block.
"""
from time import perf_counter
from collections import defaultdict


class Profiler(object):
    _active_profilers = []
    # hook set by a device backend: callable that blocks until queued work is done
    device_sync = None

    def __init__(self):
        self._fwd_t, self._fwd_n = defaultdict(float), defaultdict(int)
        self._bwd_t, self._bwd_n = defaultdict(float), defaultdict(int)

    def update(self, name, time_delta, backward=False):
        if backward:
            self._bwd_t[name] += time_delta
            self._bwd_n[name] += 1
        else:
            self._fwd_t[name] += time_delta
            self._fwd_n[name] += 1

    def __enter__(self, *args):
        Profiler._active_profilers.append(self)
        return self

    def __exit__(self, *args):
        Profiler._active_profilers.remove(self)

    def table(self):
        """{name: (fwd_seconds, fwd_calls, bwd_seconds, bwd_calls)}"""
        names = set(self._fwd_t) | set(self._bwd_t)
        return {n: (self._fwd_t[n], self._fwd_n[n], self._bwd_t[n], self._bwd_n[n]) for n in names}

    def print(self, topn=-1):
        rows = sorted(self.table().items(), key=lambda kv: -kv[1][0])
        rows = rows[:topn] if topn > 0 else rows
        print(" Function       |   forward      \t|   backward   \n" + "-" * 70)
        for n, (ft, fc, bt, bc) in rows:
            print(" %-15s| %8.4fs (%i)\t| %8.4fs (%i) " % (n, ft, fc, bt, bc))
        print("\n")


class Tracker(object):
    _nesting = 0

    def __init__(self, name, backward=False):
        self.name, self.backward = name, backward
        self.outermost = False

    def __enter__(self, *args):
        self.outermost = (Tracker._nesting == 0)
        Tracker._nesting += 1
        if self.outermost and Profiler.device_sync is not None and Profiler._active_profilers:
            Profiler.device_sync()
        self.t0 = perf_counter()

    def __exit__(self, *args):
        Tracker._nesting = max(0, Tracker._nesting - 1)
        if self.outermost and Profiler._active_profilers:
            if Profiler.device_sync is not None:
                Profiler.device_sync()
            dt = perf_counter() - self.t0
            for p in Profiler._active_profilers:
                p.update(self.name, dt, self.backward)

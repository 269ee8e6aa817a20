"""Wall-clock context manager (the examples wrap ``log_Z`` and MLP fits in it;
/root/reference/boltzmann_machines/utils/stopwatch.py:5-64)."""
import time


class Stopwatch(object):
    def __init__(self, verbose=False):
        self.verbose = verbose
        self._t0 = None
        self._total = 0.

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()
        if self.verbose:
            print('Elapsed time: {0:.3f} sec'.format(self._total))

    def start(self):
        self._t0 = time.perf_counter()
        return self

    def stop(self):
        if self._t0 is not None:
            self._total += time.perf_counter() - self._t0
            self._t0 = None
        return self

    def elapsed(self):
        running = 0. if self._t0 is None else time.perf_counter() - self._t0
        return self._total + running

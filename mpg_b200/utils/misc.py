"""TimerStat with the reference's semantics (utils/misc.py:39-90): sliding window of 10, `.mean`."""
import time

import numpy as np


class TimerStat:
    def __init__(self, window_size=10):
        self._window_size = window_size
        self._samples = []
        self._start_time = None
        self.count = 0

    def __enter__(self):
        assert self._start_time is None, 'concurrent updates not supported'
        self._start_time = time.time()

    def __exit__(self, type, value, tb):
        self.push(time.time() - self._start_time)
        self._start_time = None

    def push(self, dt):
        self._samples.append(dt)
        if len(self._samples) > self._window_size:
            self._samples.pop(0)
        self.count += 1

    @property
    def mean(self):
        return float(np.mean(self._samples)) if self._samples else 0.0

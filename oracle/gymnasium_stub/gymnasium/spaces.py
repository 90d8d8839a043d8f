"""Stub of gymnasium.spaces: only the attributes TradingEnv sets and callers read."""
import numpy as np


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high = low, high
        self.shape = tuple(shape) if shape is not None else ()
        self.dtype = dtype


class MultiDiscrete:
    def __init__(self, nvec, dtype=None, seed=None, start=None):
        import numpy as np
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape, self.dtype = self.nvec.shape, np.dtype(np.int64)

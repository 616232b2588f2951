"""Stub spaces: just the attributes the reference env reads (shape/dtype/low/high/n/contains)."""
import numpy as np


class Space:
    shape = None
    dtype = None


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        self.shape = tuple(int(s) for s in shape)
        self.low = np.broadcast_to(np.asarray(low), self.shape).astype(self.dtype)
        self.high = np.broadcast_to(np.asarray(high), self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        if not np.can_cast(x.dtype, self.dtype):
            return False
        if x.shape != self.shape:
            return False
        return bool(np.all(x >= self.low) and np.all(x <= self.high))


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def contains(self, x):
        return 0 <= int(x) < self.n


class MultiBinary(Space):
    def __init__(self, n):
        self.n = n
        self.shape = (int(n),)
        self.dtype = np.dtype(np.int8)


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape = self.nvec.shape
        self.dtype = np.dtype(np.int64)

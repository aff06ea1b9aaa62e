"""Minimal stand-ins for the two `gym.spaces` classes the hot path consumes (model.py:172-186): only `n`, `shape`,
`dtype`, `low`, `high` are read.  Real gym / gymnasium spaces are accepted everywhere by duck typing."""
import numpy as np


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def __repr__(self):
        return "Discrete(%d)" % self.n


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low)
        self.low = np.broadcast_to(np.asarray(low, dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype), shape).copy()
        self.shape = tuple(shape)
        self.dtype = dtype

    def __repr__(self):
        return "Box(%s, %s)" % (self.shape, self.dtype)


def is_discrete(space):
    return hasattr(space, "n") and not hasattr(space, "low")


def is_box(space):
    return hasattr(space, "low") and hasattr(space, "high")

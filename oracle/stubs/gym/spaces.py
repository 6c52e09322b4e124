import numpy as np


class Box(object):
    """gym 0.10.5 semantics: Box(low, high, shape=None, dtype=float32); scalar bounds are
    broadcast to `shape` and cast to dtype (float32 by default)."""

    def __init__(self, low=None, high=None, shape=None, dtype=None):
        if dtype is None:
            dtype = np.float32
        if shape is None:
            low, high = np.asarray(low), np.asarray(high)
            assert low.shape == high.shape
            shape = low.shape
        else:
            assert np.isscalar(low) and np.isscalar(high)
            low = low + np.zeros(shape)
            high = high + np.zeros(shape)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = low.astype(dtype)
        self.high = high.astype(dtype)

    def sample(self):
        return np.random.uniform(low=self.low, high=self.high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        return x.shape == self.shape and (x >= self.low).all() and (x <= self.high).all()

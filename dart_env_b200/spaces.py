"""Minimal Box space with the semantics of the reference's gym/spaces/box.py (the only space
the Dart envs use: dart_env.py:86,97-100).  gym itself is not a dependency."""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            low, high = np.asarray(low), np.asarray(high)
            assert low.shape == high.shape
            self.shape = low.shape
            self.low, self.high = low.astype(self.dtype), high.astype(self.dtype)
        else:
            assert np.isscalar(low) and np.isscalar(high)
            self.shape = tuple(shape)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)
        self.bounded_below = -np.inf < self.low
        self.bounded_above = np.inf > self.high
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        """Same case analysis as gym/spaces/box.py:67-110 (uniform / exponential / normal)."""
        high = self.high if self.dtype.kind == "f" else self.high.astype("int64") + 1
        sample = np.empty(self.shape)
        unbounded = ~self.bounded_below & ~self.bounded_above
        upp = ~self.bounded_below & self.bounded_above
        low = self.bounded_below & ~self.bounded_above
        bounded = self.bounded_below & self.bounded_above
        sample[unbounded] = self.np_random.normal(size=unbounded[unbounded].shape)
        sample[low] = self.np_random.exponential(size=low[low].shape) + self.low[low]
        sample[upp] = -self.np_random.exponential(size=upp[upp].shape) + self.high[upp]
        sample[bounded] = self.np_random.uniform(low=self.low[bounded], high=high[bounded], size=bounded[bounded].shape)
        return sample.astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)

    def __contains__(self, x):
        return self.contains(x)

    def __repr__(self):
        return "Box" + str(self.shape)

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and np.allclose(self.low, other.low) \
            and np.allclose(self.high, other.high)


def batch_space(space: Box, n: int) -> Box:
    """gym/vector/utils/spaces.py:9-43 for Box."""
    return Box(np.tile(space.low, (n,) + (1,) * len(space.shape)), np.tile(space.high, (n,) + (1,) * len(space.shape)),
               dtype=space.dtype)

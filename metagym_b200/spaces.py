"""gym-independent stand-ins for the two gym.spaces types the reference envs expose (gym is not a dependency).

Mirrors what the reference constructs at metagym/quadrotor/env.py:73-94 and metagym/metamaze/envs/maze_env.py:35-39,
165-169: `.sample() .low .high .shape .n .dtype`, with an optional leading batch axis for sample().
"""
import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = dtype


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        if shape is None:
            shape = self.low.shape
        self.low = np.broadcast_to(self.low, shape).copy()
        self.high = np.broadcast_to(self.high, shape).copy()
        Space.__init__(self, shape, np.dtype(dtype))
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)

    def sample(self, n=None):
        shape = self.shape if n is None else (n,) + self.shape
        return self._rng.uniform(self.low, self.high, size=shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape[-len(self.shape):] == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        Space.__init__(self, (), np.dtype(np.int64))
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)

    def sample(self, n=None):
        if n is None:
            return int(self._rng.randint(self.n))
        return self._rng.randint(self.n, size=n).astype(np.int64)

    def contains(self, x):
        return 0 <= int(x) < self.n

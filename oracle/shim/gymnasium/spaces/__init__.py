"""Discrete.seed/sample(mask)/contains as documented by gymnasium [UPSTREAM-UNVERIFIED]:
seed(s) -> Generator(PCG64(SeedSequence(s))); sample(mask) -> start + rng.choice(where(mask == 1))."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = shape
        self.dtype = dtype
        self._np_random = None
        if seed is not None:
            self.seed(seed)

    @property
    def np_random(self):
        if self._np_random is None:
            self.seed(None)
        return self._np_random

    def seed(self, seed=None):
        ss = np.random.SeedSequence(seed)
        self._np_random = np.random.Generator(np.random.PCG64(ss))
        return seed


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n = int(n)
        self.start = int(start)
        super().__init__((), np.int64, seed)

    def contains(self, x):
        try:
            xi = int(x)
        except (TypeError, ValueError):
            return False
        return xi == x and self.start <= xi < self.start + self.n

    def sample(self, mask=None):
        if mask is not None:
            valid = mask == 1
            if np.any(valid):
                return self.start + self.np_random.choice(np.where(valid)[0])
            return self.start
        return self.start + int(self.np_random.integers(self.n))

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None, start=None):
        self.nvec = np.asarray(nvec)
        super().__init__(self.nvec.shape, dtype, seed)

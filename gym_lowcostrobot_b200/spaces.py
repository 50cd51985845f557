"""Minimal stand-ins for ``gymnasium.spaces.Box`` / ``Dict`` (gymnasium is not a dependency).

Only what the reference envs declare (reach_cube_env.py:95-115): bounds, shape, dtype,
``sample`` and ``contains``.  If gymnasium is importable ``to_gymnasium()`` converts.
"""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def to_gymnasium(self):
        from gymnasium import spaces

        return spaces.Box(self.low, self.high, shape=self.shape, dtype=self.dtype.type)

    def __repr__(self):
        return f"Box({self.low.flat[0]}, {self.high.flat[0]}, {self.shape}, {self.dtype})"


class Dict(dict):
    @property
    def spaces(self):
        """the sub-spaces by key, like ``gymnasium.spaces.Dict.spaces``"""
        return self

    def sample(self):
        return {k: v.sample() for k, v in self.items()}

    def contains(self, x):
        return set(x) == set(self) and all(self[k].contains(x[k]) for k in self)

    def to_gymnasium(self):
        from gymnasium import spaces

        return spaces.Dict({k: v.to_gymnasium() for k, v in self.items()})


def batch_space(space, n):
    if isinstance(space, Dict):
        return Dict({k: batch_space(v, n) for k, v in space.items()})
    b = Box(0, 0, (n,) + space.shape, space.dtype)
    b.low[:] = space.low
    b.high[:] = space.high
    return b

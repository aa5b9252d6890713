"""Light stand-ins for the gym-0.20 spaces the reference env declares (merging_env.py:75-78,101-102).

gym is not a dependency; if it is importable, `as_gym()` converts to the real classes.
"""
from __future__ import annotations

import numpy as np

H, W = 1000, 300   # merging_env.py:23


class Discrete:
    def __init__(self, n: int, seed=None):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)
        self._rng = np.random.default_rng(seed)

    def sample(self) -> int:
        return int(self._rng.integers(self.n))

    def contains(self, x) -> bool:
        try:
            return 0 <= int(x) < self.n
        except (TypeError, ValueError):
            return False

    __contains__ = contains

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def __repr__(self):
        return f"Discrete({self.n})"

    def __eq__(self, o):
        return isinstance(o, Discrete) and o.n == self.n

    def as_gym(self):
        from gym import spaces  # noqa: optional
        return spaces.Discrete(self.n)


class Box:
    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=np.float64)
        self.high = np.asarray(high, dtype=np.float64)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x, dtype=np.float64)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    __contains__ = contains

    def __repr__(self):
        return f"Box({self.shape}, {self.dtype})"

    def as_gym(self):
        from gym import spaces  # noqa: optional
        return spaces.Box(low=self.low, high=self.high, dtype=self.dtype)


class MultiDiscrete:
    """Batched action space of a vector env (gym 0.20 `VectorEnv.action_space`)."""

    def __init__(self, nvec, seed=None):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape = self.nvec.shape
        self.dtype = np.dtype(np.int64)
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return self._rng.integers(0, self.nvec)

    def __repr__(self):
        return f"MultiDiscrete(shape={self.shape}, n={int(self.nvec.flat[0]) if self.nvec.size else 0})"


def merge_observation_space() -> Box:
    """merging_env.py:76-78 (declared float16; values are never clipped to it by the reference)."""
    return Box(low=[-H, -W, -100, 0, 0, -H, -W, -100, 0, 0],
               high=[H, W, 100, H, 100, H, W, 100, H, 100], dtype=np.float16)


def merge_action_space() -> Discrete:
    """merging_env.py:101-102."""
    return Discrete(5)

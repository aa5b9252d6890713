"""Minimal stand-in for gym 0.20 (test infrastructure; see ../README.md)."""
from . import error, spaces, utils          # noqa: F401
from .envs.registration import make, register  # noqa: F401


class Env:
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None

    @property
    def unwrapped(self):
        return self

    def seed(self, seed=None):
        return [seed]

    def close(self):
        pass

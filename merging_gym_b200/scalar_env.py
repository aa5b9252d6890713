"""`MergeEnv` — the reference's scalar env interface on top of a 1-env `MergeVecEnv`.

Lets the reference's per-step scripts (scripts/main.py:190-218, hdqn.py:277-327,
ranbowdqn.py:659-671, human_player.py:93-181) run unchanged: Python lists / floats in and out,
`step(action1, action2=None)`, sticky `done`, `.winner in (None, 1, 2)`, `.r1_accumulate`.
Every call synchronises with the GPU — this wrapper is for compatibility and parity tests, the
vector env is the product.
"""
from __future__ import annotations

import torch

from . import _native as nat
from .spaces import merge_action_space, merge_observation_space
from .vec_env import MergeVecEnv

ENV_ID = "merging_env-v0"          # merging_gym/__init__.py:3-6


class MergeEnv:
    metadata = {}

    def __init__(self, device="cuda", rewards=None):
        self._vec = MergeVecEnv(1, mode="pvp", device=device, auto_reset=False, rewards=rewards,
                                episode_info=False, track_stats=False)
        self.observation_space = merge_observation_space()      # merging_env.py:76-78
        self.action_space = merge_action_space()                # merging_env.py:101-102
        self.action_dict = {0: 0, 1: 10, 2: 20, 3: 30, 4: 40}   # merging_env.py:101
        self.action1 = 1
        self.action2 = 1

    @property
    def unwrapped(self):
        return self

    def show_reward(self):
        return self._vec.show_reward()

    def reset(self):
        return [float(v) for v in self._vec.reset()[0].tolist()]

    def observe(self):
        return [float(v) for v in self._vec.reset(mask=torch.zeros(1, dtype=torch.uint8))[0].tolist()]

    def step(self, action1, action2=None):
        # action_dict[action] raises KeyError for anything outside 0..4 (merging_env.py:147,152)
        self.action_dict[action1]
        if action2 is not None:
            self.action_dict[action2]
        self.action1, self.action2 = action1, action2
        a1 = torch.tensor([int(action1)], dtype=torch.uint8)
        a2 = None if action2 is None else torch.tensor([int(action2)], dtype=torch.uint8)
        obs, rew, done, info = self._vec.step(a1, a2)
        flags = int(info["flags"][0])
        return ([float(v) for v in obs[0].tolist()], [float(v) for v in rew[0].tolist()],
                bool(flags & nat.INFO_DONE), {"collision": bool(flags & nat.INFO_COLLISION)})

    @property
    def done(self):
        return bool(self._vec.done[0])

    @property
    def winner(self):
        w = int(self._vec.winner[0])
        return None if w == 0 else w

    @property
    def r1_accumulate(self):
        return float(self._vec.ret1[0])

    @property
    def r2_accumulate(self):
        return float(self._vec.ret2[0])

    @property
    def time_stamp(self):
        return float(self._vec.steps[0]) * self._vec.constants.dT

    @property
    def state1(self):
        return {"pos": float(self._vec.pos1[0]), "vel": float(self._vec.vel1[0])}

    @property
    def state2(self):
        return {"pos": float(self._vec.pos2[0]), "vel": float(self._vec.vel2[0])}

    def seed(self, seed=None):
        return [seed]

    def render(self, *a, **k):
        raise NotImplementedError("the pygame UI (merging_env.py:241-399) is out of scope")

    def close(self):
        self._vec.close()


def make(env_id: str = ENV_ID, **kwargs):
    """`gym.make("merging_env-v0")` (merging_gym/__init__.py:3-6, scripts/main.py:20)."""
    if env_id != ENV_ID:
        raise KeyError(f"unknown env id {env_id!r}; only {ENV_ID!r} is provided")
    return MergeEnv(**kwargs)


def register_gym(env_id: str = ENV_ID) -> str:
    """Register this env under the reference's id in gym's registry, exactly as
    merging_gym/__init__.py:3-6 does for the Python env, so that the reference scripts'
    `gym.make("merging_env-v0")` (scripts/main.py:20, hdqn.py:26, human_player.py:27,
    ranbowdqn.py:628) returns the GPU-backed env.  Needs gym (or gymnasium) to be importable; it is
    not a dependency of this package."""
    try:
        from gym.envs.registration import register
    except ImportError:
        try:
            from gymnasium.envs.registration import register
        except ImportError as e:
            raise ImportError("register_gym() needs gym or gymnasium") from e
    register(id=env_id, entry_point="merging_gym_b200.scalar_env:MergeEnv")
    return env_id


def make_vec(num_envs: int, **kwargs) -> MergeVecEnv:
    return MergeVecEnv(num_envs, **kwargs)

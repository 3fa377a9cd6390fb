"""TEST INFRASTRUCTURE: the slice of gym 0.13 the reference's env files and MyoSuite 1.2.3 use."""
import importlib

import numpy as np

from . import spaces, utils          # noqa: F401


class Env:
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def seed(self, seed=None):
        return self.env.seed(seed)


from . import wrappers          # noqa: E402,F401  (needs Wrapper)
from .envs import registration  # noqa: E402,F401
from .envs.registration import make, register  # noqa: E402,F401

import importlib


class EnvSpec:
    def __init__(self, id, entry_point=None, max_episode_steps=None, kwargs=None, **_):
        self.id, self.entry_point, self.max_episode_steps = id, entry_point, max_episode_steps
        self._kwargs = dict(kwargs or {})

    def make(self, **kwargs):
        kw = dict(self._kwargs)
        kw.update(kwargs)
        if callable(self.entry_point):
            cls = self.entry_point
        else:
            mod, name = self.entry_point.split(":")
            cls = getattr(importlib.import_module(mod), name)
        env = cls(**kw)
        env.spec = self
        if self.max_episode_steps is not None:
            from gym.wrappers import TimeLimit

            env = TimeLimit(env, max_episode_steps=self.max_episode_steps)
        return env


class _Registry:
    def __init__(self):
        self.env_specs = {}


registry = _Registry()


def register(id, **kwargs):
    registry.env_specs[id] = EnvSpec(id, **kwargs)


def make(id, **kwargs):
    if id not in registry.env_specs:
        raise KeyError("No registered env with id: {}".format(id))
    return registry.env_specs[id].make(**kwargs)

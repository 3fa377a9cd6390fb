"""The winning run's 32-step Baoding curriculum as data: the env kwargs of every training
step (``trained_models/curriculum_steps_complete_baoding_winner/NN_*/config.json`` of the reference, packaged by
tests/golden/make_curriculum.py into assets/curriculum/baoding_winner.json) and the env each step trains on. The reference
runs one ``main.py`` per step by hand, loading the previous step's model; ``iterate`` gives the same sequence as
(step name, env) pairs for a loop such as

    agent = None
    for name, env in curriculum.iterate(num_envs=32768):
        vn = DeviceVecNormalize(env) if agent is None else DeviceVecNormalize.load(prev_env_pkl, env)
        agent = RecurrentPPO("MlpLstmPolicy", vn, ...) if agent is None else RecurrentPPO.load(prev_zip, env=vn)
        agent.learn(total_timesteps=...)
"""
from __future__ import annotations

import json
from typing import Any, Dict, Iterator, List, Tuple

from .assets import asset_path
from .envs import FACTORY_NAMES, EnvironmentFactory, make_task_cfg
from .sim import Model


def load(name: str = "baoding_winner") -> List[Dict[str, Any]]:
    with open(asset_path(f"curriculum/{name}.json")) as f:
        return json.load(f)["steps"]


def task_cfg(step: Dict[str, Any], model: Model = None):
    """The C-ABI task configuration of one curriculum step (what ``EnvironmentFactory.create(env_name, **config)`` builds)."""
    from .envs import REGISTRY

    env_id = FACTORY_NAMES[step["env_name"]]
    model = model or Model(asset_path(REGISTRY[env_id]["model"]))
    return make_task_cfg(model, env_id, **step["config"])


def iterate(num_envs: int, device="cuda:0", seed: int = 0, name: str = "baoding_winner") -> Iterator[Tuple[str, Any]]:
    for k, step in enumerate(load(name)):
        yield step["step"], EnvironmentFactory.create(step["env_name"], num_envs=num_envs, device=device, seed=seed + k, **step["config"])

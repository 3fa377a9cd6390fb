"""B200-native batched myo-simulator: the data-parallel hot path of amathislab/myochallenge
(batched world stepping + recurrent-policy rollout) behind the reference's gym/SB3-style API."""
from . import _capi
from ._capi import MyoError, TaskCfg
from .sim import BatchSim, Model

__all__ = ["BatchSim", "Model", "MyoError", "TaskCfg", "_capi"]

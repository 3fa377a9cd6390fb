"""MyoSuite 1.2.3 envs/myo/base_v0.py ``BaseV0``, restated from memory: 'act' appended to the obs keys of muscle models, tip /
target site ids, and the explicit muscle action remap ctrl = 1 / (1 + exp(-5 (a - 0.5))) in ``step``."""
import numpy as np

from myosuite.envs import env_base

DYN_MUSCLE = 3


class BaseV0(env_base.MujocoEnv):
    def __init__(self, model_path, obsd_model_path=None, seed=None):
        super().__init__(model_path=model_path, obsd_model_path=obsd_model_path, seed=seed)

    def _setup(self, obs_keys, weighted_reward_keys, sites=None, frame_skip=10, muscle_condition="", **kwargs):
        if self.sim.model.na > 0 and "act" not in obs_keys:
            obs_keys = obs_keys.copy()
            obs_keys.append("act")
        self.tip_sids, self.target_sids = [], []
        if sites:
            for site in sites:
                self.tip_sids.append(self.sim.model.site_name2id(site))
                self.target_sids.append(self.sim.model.site_name2id(site + "_target"))
        self.muscle_condition = muscle_condition
        super()._setup(obs_keys=obs_keys, weighted_reward_keys=weighted_reward_keys, frame_skip=frame_skip, **kwargs)

    def step(self, a):
        muscle_a = np.asarray(a, dtype=np.float64).copy()
        if self.sim.model.na and self.normalize_act:
            ind = np.asarray(self.sim.model.actuator_dyntype) == DYN_MUSCLE
            muscle_a[ind] = 1.0 / (1.0 + np.exp(-5.0 * (muscle_a[ind] - 0.5)))
            is_normalized = False
        else:
            is_normalized = self.normalize_act
        self.last_ctrl = self.robot.step(ctrl_desired=muscle_a, ctrl_normalized=is_normalized, step_duration=self.dt)
        return self.forward()

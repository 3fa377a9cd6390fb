"""MyoSuite 1.2.3 envs/myo/myochallenge/reorient_v0.py ``ReorientEnvV0``, restated from memory (SURVEY.md row a11''):
obs = hand_qpos[23] | hand_qvel*dt[23] | obj_pos | goal_pos | pos_err | obj_rot (euler) | goal_rot | rot_err | act."""
import collections

import gym
import numpy as np

from myosuite.envs.myo.base_v0 import BaseV0
from myosuite.utils.quat_math import mat2euler


class ReorientEnvV0(BaseV0):
    DEFAULT_OBS_KEYS = ["hand_qpos", "hand_qvel", "obj_pos", "goal_pos", "pos_err", "obj_rot", "goal_rot", "rot_err"]
    DEFAULT_RWD_KEYS_AND_WEIGHTS = {"pos_dist": 100.0, "rot_dist": 1.0}

    def __init__(self, model_path, obsd_model_path=None, seed=None, **kwargs):
        gym.utils.EzPickle.__init__(self, model_path, obsd_model_path, seed, **kwargs)
        super().__init__(model_path=model_path, obsd_model_path=obsd_model_path, seed=seed)
        self._setup(**kwargs)

    def get_obs_dict(self, sim):
        obs_dict = {}
        obs_dict["t"] = np.array([sim.data.time])
        obs_dict["hand_qpos"] = sim.data.qpos[:-7].copy()
        obs_dict["hand_qvel"] = sim.data.qvel[:-6].copy() * self.dt
        obs_dict["obj_pos"] = sim.data.site_xpos[self.object_sid].copy()
        obs_dict["goal_pos"] = sim.data.site_xpos[self.goal_sid].copy()
        obs_dict["pos_err"] = obs_dict["goal_pos"] - obs_dict["obj_pos"] - self.goal_obj_offset
        obs_dict["obj_rot"] = mat2euler(np.reshape(sim.data.site_xmat[self.object_sid], (3, 3)))
        obs_dict["goal_rot"] = mat2euler(np.reshape(sim.data.site_xmat[self.goal_sid], (3, 3)))
        obs_dict["rot_err"] = obs_dict["goal_rot"] - obs_dict["obj_rot"]
        if sim.model.na > 0:
            obs_dict["act"] = sim.data.act[:].copy()
        return obs_dict

    def get_reward_dict(self, obs_dict):      # the reference's CustomReorientEnv overrides this
        pos_dist = np.abs(np.linalg.norm(self.obs_dict["pos_err"], axis=-1))
        rot_dist = np.abs(np.linalg.norm(self.obs_dict["rot_err"], axis=-1))
        act_mag = np.linalg.norm(self.obs_dict["act"], axis=-1) / self.sim.model.na if self.sim.model.na != 0 else 0
        drop = pos_dist > self.drop_th
        rwd_dict = collections.OrderedDict((
            ("pos_dist", -1.0 * pos_dist),
            ("rot_dist", -1.0 * rot_dist),
            ("act_reg", -1.0 * act_mag),
            ("sparse", -rot_dist - 10.0 * pos_dist),
            ("solved", (pos_dist < self.pos_th) and (rot_dist < self.rot_th) and (not drop)),
            ("done", drop),
        ))
        rwd_dict["dense"] = np.sum([wt * rwd_dict[key] for key, wt in self.rwd_keys_wt.items()], axis=0)
        return rwd_dict

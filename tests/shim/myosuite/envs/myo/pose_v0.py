"""MyoSuite 1.2.3 envs/myo/pose_v0.py ``PoseEnvV0``, restated from memory (SURVEY.md rows a11', a12'')."""
import collections

import gym
import numpy as np

from myosuite.envs.myo.base_v0 import BaseV0


class PoseEnvV0(BaseV0):
    DEFAULT_OBS_KEYS = ["qpos", "qvel", "pose_err"]
    DEFAULT_RWD_KEYS_AND_WEIGHTS = {"pose": 1.0, "bonus": 4.0, "act_reg": 1.0, "penalty": 50}

    def __init__(self, model_path, obsd_model_path=None, seed=None, **kwargs):
        gym.utils.EzPickle.__init__(self, model_path, obsd_model_path, seed, **kwargs)
        super().__init__(model_path=model_path, obsd_model_path=obsd_model_path, seed=seed)
        self._setup(**kwargs)

    def get_obs_dict(self, sim):
        obs_dict = {}
        obs_dict["t"] = np.array([sim.data.time])
        obs_dict["qpos"] = sim.data.qpos[:].copy()
        obs_dict["qvel"] = sim.data.qvel[:].copy() * self.dt
        if sim.model.na > 0:
            obs_dict["act"] = sim.data.act[:].copy()
        obs_dict["pose_err"] = self.target_jnt_value - obs_dict["qpos"]
        return obs_dict

    def get_reward_dict(self, obs_dict):
        pose_dist = np.linalg.norm(obs_dict["pose_err"], axis=-1)
        act_mag = np.linalg.norm(self.obs_dict["act"], axis=-1)
        if self.sim.model.na != 0:
            act_mag = act_mag / self.sim.model.na
        far_th = 4 * np.pi / 2
        rwd_dict = collections.OrderedDict((
            ("pose", -1.0 * pose_dist),
            ("bonus", 1.0 * (pose_dist < self.pose_thd) + 1.0 * (pose_dist < 1.5 * self.pose_thd)),
            ("penalty", -1.0 * (pose_dist > far_th)),
            ("act_reg", -1.0 * act_mag),
            ("sparse", -1.0 * pose_dist),
            ("solved", pose_dist < self.pose_thd),
            ("done", pose_dist > far_th),
        ))
        rwd_dict["dense"] = np.sum([wt * rwd_dict[key] for key, wt in self.rwd_keys_wt.items()], axis=0)
        return rwd_dict

    def get_target_pose(self):
        if self.target_type == "fixed":
            return self.target_jnt_value
        elif self.target_type == "generate":
            return self.np_random.uniform(high=self.target_jnt_range[:, 0], low=self.target_jnt_range[:, 1])
        raise TypeError("Unknown Target type: {}".format(self.target_type))

    def update_target(self, restore_sim=False):
        if restore_sim:
            qpos = self.sim.data.qpos[:].copy()
            qvel = self.sim.data.qvel[:].copy()
        self.target_jnt_value = self.get_target_pose()
        self.sim.data.qpos[:] = self.target_jnt_value.copy()
        self.sim.forward()
        for isite in range(len(self.tip_sids)):
            self.sim.model.site_pos[self.target_sids[isite]] = self.sim.data.site_xpos[self.tip_sids[isite]].copy()
        if restore_sim:
            self.sim.data.qpos[:] = qpos[:]
            self.sim.data.qvel[:] = qvel[:]
        self.sim.forward()

    def reset(self):
        raise NotImplementedError("shim: the reference's CustomPoseEnv overrides reset")

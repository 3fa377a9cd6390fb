"""MyoSuite 1.2.3 envs/myo/myochallenge/baoding_v1.py, restated from memory."""
import collections
import enum

import gym
import numpy as np

from myosuite.envs.myo.base_v0 import BaseV0


class Task(enum.Enum):
    HOLD = 0
    BAODING_CW = 1
    BAODING_CCW = 2


WHICH_TASK = Task.BAODING_CCW


class BaodingEnvV1(BaseV0):
    DEFAULT_OBS_KEYS = ["hand_pos", "object1_pos", "object1_velp", "object2_pos", "object2_velp", "target1_pos", "target2_pos",
                        "target1_err", "target2_err"]
    DEFAULT_RWD_KEYS_AND_WEIGHTS = {"pos_dist_1": 5.0, "pos_dist_2": 5.0}

    def __init__(self, model_path, obsd_model_path=None, seed=None, **kwargs):
        gym.utils.EzPickle.__init__(self, model_path, obsd_model_path, seed, **kwargs)
        super().__init__(model_path=model_path, obsd_model_path=obsd_model_path, seed=seed)
        self._setup(**kwargs)

    def step(self, a):
        if self.which_task in [Task.BAODING_CW, Task.BAODING_CCW]:
            desired_angle_wrt_palm = self.goal[self.counter].copy()
            desired_angle_wrt_palm[0] = desired_angle_wrt_palm[0] + self.ball_1_starting_angle
            desired_angle_wrt_palm[1] = desired_angle_wrt_palm[1] + self.ball_2_starting_angle
            desired_positions_wrt_palm = [0, 0, 0, 0]
            desired_positions_wrt_palm[0] = self.x_radius * np.cos(desired_angle_wrt_palm[0]) + self.center_pos[0]
            desired_positions_wrt_palm[1] = self.y_radius * np.sin(desired_angle_wrt_palm[0]) + self.center_pos[1]
            desired_positions_wrt_palm[2] = self.x_radius * np.cos(desired_angle_wrt_palm[1]) + self.center_pos[0]
            desired_positions_wrt_palm[3] = self.y_radius * np.sin(desired_angle_wrt_palm[1]) + self.center_pos[1]
            for sim in [self.sim, self.sim_obsd]:
                sim.model.site_pos[self.target1_sid, 0] = desired_positions_wrt_palm[0]
                sim.model.site_pos[self.target1_sid, 1] = desired_positions_wrt_palm[1]
                sim.model.site_pos[self.target2_sid, 0] = desired_positions_wrt_palm[2]
                sim.model.site_pos[self.target2_sid, 1] = desired_positions_wrt_palm[3]
        self.counter += 1
        return super().step(a)

    def get_obs_dict(self, sim):
        obs_dict = {}
        obs_dict["t"] = np.array([sim.data.time])
        obs_dict["hand_pos"] = sim.data.qpos[:-14].copy()
        obs_dict["object1_pos"] = sim.data.site_xpos[self.object1_sid].copy()
        obs_dict["object2_pos"] = sim.data.site_xpos[self.object2_sid].copy()
        obs_dict["object1_velp"] = sim.data.qvel[-12:-9].copy() * self.dt
        obs_dict["object2_velp"] = sim.data.qvel[-6:-3].copy() * self.dt
        obs_dict["target1_pos"] = sim.data.site_xpos[self.target1_sid].copy()
        obs_dict["target2_pos"] = sim.data.site_xpos[self.target2_sid].copy()
        obs_dict["target1_err"] = obs_dict["target1_pos"] - obs_dict["object1_pos"]
        obs_dict["target2_err"] = obs_dict["target2_pos"] - obs_dict["object2_pos"]
        if sim.model.na > 0:
            obs_dict["act"] = sim.data.act[:].copy()
        return obs_dict

    def create_goal_trajectory(self, time_step=0.1, time_period=6):
        len_of_goals = 1000
        sign = 0
        if self.which_task == Task.BAODING_CW:
            sign = -1
        if self.which_task == Task.BAODING_CCW:
            sign = 1
        goal_traj = []
        t = 0
        while t < len_of_goals:
            angle_before_shift = sign * 2 * np.pi * (t * time_step / time_period)
            goal_traj.append(np.array([angle_before_shift, angle_before_shift]))
            t += 1
        return np.array(goal_traj)

    def get_reward_dict(self, obs_dict):      # the reference's subclasses override this
        target1_dist = np.linalg.norm(obs_dict["target1_err"], axis=-1)
        target2_dist = np.linalg.norm(obs_dict["target2_err"], axis=-1)
        target_dist = target1_dist + target2_dist
        act_mag = np.linalg.norm(self.obs_dict["act"], axis=-1) / self.sim.model.na if self.sim.model.na != 0 else 0
        object1_pos = obs_dict["object1_pos"][:, :, 2] if obs_dict["object1_pos"].ndim == 3 else obs_dict["object1_pos"][2]
        object2_pos = obs_dict["object2_pos"][:, :, 2] if obs_dict["object2_pos"].ndim == 3 else obs_dict["object2_pos"][2]
        is_fall = np.logical_or(object1_pos < self.drop_th, object2_pos < self.drop_th)
        rwd_dict = collections.OrderedDict((
            ("pos_dist_1", -1.0 * target1_dist),
            ("pos_dist_2", -1.0 * target2_dist),
            ("act_reg", -1.0 * act_mag),
            ("sparse", -target_dist),
            ("solved", (target1_dist < self.proximity_th) * (target2_dist < self.proximity_th) * (~is_fall)),
            ("done", is_fall),
        ))
        rwd_dict["dense"] = np.sum([wt * rwd_dict[key] for key, wt in self.rwd_keys_wt.items()], axis=0)
        return rwd_dict

    def reset(self, reset_pose=None, reset_vel=None, reset_goal=None, time_period=None):
        raise NotImplementedError("shim: the reference's subclasses override reset completely")

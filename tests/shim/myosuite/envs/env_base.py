"""MyoSuite 1.2.3 envs/env_base.py ``MujocoEnv``, restated from memory over ``OracleSim``."""
import gym
import numpy as np
from gym.utils import seeding

from myosuite.robot.robot import Robot
from myosuite.utils.obj_vec_dict import ObsVecDict

from .sim import MjSimState, OracleSim

JNT_SLIDE, JNT_HINGE, TRN_JOINT = 2, 3, 0


class MujocoEnv(gym.Env, gym.utils.EzPickle, ObsVecDict):
    def __init__(self, model_path, obsd_model_path=None, seed=None):
        self.seed(seed)
        self.sim = OracleSim(model_path)
        self.sim_obsd = OracleSim(obsd_model_path) if obsd_model_path else self.sim
        self.sim.forward()
        self.sim_obsd.forward()
        ObsVecDict.__init__(self)

    def _setup(self, obs_keys, weighted_reward_keys, reward_mode="dense", frame_skip=1, normalize_act=True, obs_range=(-10, 10),
               seed=None, rwd_viz=False, device_id=0, **kwargs):
        self.mujoco_render_frames = False
        self.rwd_viz = rwd_viz
        self.robot = Robot(mj_sim=self.sim, random_generator=self.np_random, **kwargs)
        self.frame_skip = frame_skip
        self.normalize_act = normalize_act
        m = self.sim.model
        act_low = -np.ones(m.nu) if self.normalize_act else m.actuator_ctrlrange[:, 0].copy()
        act_high = np.ones(m.nu) if self.normalize_act else m.actuator_ctrlrange[:, 1].copy()
        self.action_space = gym.spaces.Box(act_low, act_high, dtype=np.float32)
        self.init_qvel = self.sim.data.qvel.ravel().copy()
        self.init_qpos = self.sim.data.qpos.ravel().copy()
        if self.normalize_act:      # linear, joint-actuated joints start at the middle of their range (none in the muscle models)
            actuated = np.asarray(m.actuator_trnid)[np.asarray(m.actuator_trntype) == TRN_JOINT, 0]
            linear = np.where(np.logical_or(np.asarray(m.jnt_type) == JNT_SLIDE, np.asarray(m.jnt_type) == JNT_HINGE))[0]
            ids = np.intersect1d(actuated, linear).astype(int)
            if ids.size:
                self.init_qpos[np.asarray(m.jnt_qposadr)[ids]] = np.mean(np.asarray(m.jnt_range)[ids], axis=1)
        self.rwd_dict = {}
        self.rwd_mode = reward_mode
        self.rwd_keys_wt = weighted_reward_keys
        self.obs_dict = {}
        self.obs_keys = obs_keys
        observation, _reward, done, _info = self.step(np.zeros(m.nu))
        assert not done, "Check initialization. Simulation starts in a done state."
        self.obs_dim = observation.size
        self.observation_space = gym.spaces.Box(obs_range[0] * np.ones(self.obs_dim), obs_range[1] * np.ones(self.obs_dim), dtype=np.float32)

    def step(self, a):
        a = np.clip(a, self.action_space.low, self.action_space.high)
        self.last_ctrl = self.robot.step(ctrl_desired=a, ctrl_normalized=self.normalize_act, step_duration=self.dt)
        return self.forward()

    def forward(self):
        obs = self.get_obs()
        self.expand_dims(self.obs_dict)
        self.rwd_dict = self.get_reward_dict(self.obs_dict)
        self.squeeze_dims(self.rwd_dict)
        self.squeeze_dims(self.obs_dict)
        env_info = self.get_env_infos()
        return obs, env_info["rwd_" + self.rwd_mode], bool(env_info["done"]), env_info

    def get_obs(self):
        sen = self.robot.get_sensors()
        self.robot.sensor2sim(sen, self.sim_obsd)
        self.obs_dict = self.get_obs_dict(self.sim_obsd)
        _t, obs = self.obsdict2obsvec(self.obs_dict, self.obs_keys)
        return obs

    def get_env_infos(self):
        return {
            "time": self.obs_dict["t"][()],
            "rwd_dense": self.rwd_dict["dense"][()],
            "rwd_sparse": self.rwd_dict["sparse"][()],
            "solved": self.rwd_dict["solved"][()],
            "done": self.rwd_dict["done"][()],
            "obs_dict": self.obs_dict,
            "rwd_dict": self.rwd_dict,
        }

    def seed(self, seed=None):
        self.input_seed = seed
        self.np_random, seed = seeding.np_random(seed)
        return [seed]

    def reset(self, reset_qpos=None, reset_qvel=None):
        qpos = self.init_qpos.copy() if reset_qpos is None else reset_qpos
        qvel = self.init_qvel.copy() if reset_qvel is None else reset_qvel
        self.robot.reset(qpos, qvel)
        return self.get_obs()

    @property
    def dt(self):
        return self.sim.model.opt.timestep * self.frame_skip

    def set_state(self, qpos, qvel, act=None):
        """mujoco_py semantics: a new MjSimState with the OLD time and activations, then mj_forward."""
        assert qpos.shape == (self.sim.model.nq,) and qvel.shape == (self.sim.model.nv,)
        old = self.sim.get_state()
        self.sim.set_state(MjSimState(old.time, qpos, qvel, old.act if act is None else act, old.udd_state))
        self.sim.forward()

    def close(self):
        pass

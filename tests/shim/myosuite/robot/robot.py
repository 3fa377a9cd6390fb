"""MyoSuite 1.2.3 robot/robot.py, simulation-only subset: ctrl clipping to actuator_ctrlrange, n = int(dt / timestep)
substeps with constant ctrl, reset = mj_resetData + qpos / qvel + mj_forward, sensor2sim = state copy + mj_forward."""
import numpy as np


class Robot:
    def __init__(self, mj_sim, random_generator=None, **kwargs):
        self.sim = mj_sim
        self.np_random = random_generator

    def normalize_actions(self, ctrl):      # not used by muscle envs (BaseV0.step remaps explicitly)
        rng = self.sim.model.actuator_ctrlrange
        return 0.5 * (rng[:, 0] + rng[:, 1]) + ctrl * 0.5 * (rng[:, 1] - rng[:, 0])

    def step(self, ctrl_desired, step_duration, ctrl_normalized=True, realTimeSim=False, render_cbk=None):
        ctrl = np.asarray(ctrl_desired, dtype=np.float64).copy()
        if ctrl_normalized:
            ctrl = self.normalize_actions(ctrl)
        rng = self.sim.model.actuator_ctrlrange
        ctrl_feasible = np.clip(ctrl, rng[:, 0], rng[:, 1])
        self.sim.data.ctrl[:] = ctrl_feasible
        n_frames = int(step_duration / self.sim.model.opt.timestep)
        self.sim.advance(substeps=n_frames, render=False)
        return ctrl_feasible

    def reset(self, reset_pos, reset_vel, blocking=True):
        self.sim.reset()
        self.sim.data.qpos[:] = reset_pos
        self.sim.data.qvel[:] = reset_vel
        self.sim.forward()

    def get_sensors(self):
        d = self.sim.data
        return dict(time=d.time, qpos=d.qpos.copy(), qvel=d.qvel.copy(), act=d.act.copy())

    def sensor2sim(self, sen, sim):
        sim.data.time = sen["time"]
        sim.data.qpos[:] = sen["qpos"]
        sim.data.qvel[:] = sen["qvel"]
        if sim.model.na > 0:
            sim.data.act[:] = sen["act"]
        sim.forward()

    def sync_sims(self, source_sim, destination_sim):
        if source_sim is destination_sim:
            return
        destination_sim.data.time = source_sim.data.time
        destination_sim.data.qpos[:] = source_sim.data.qpos
        destination_sim.data.qvel[:] = source_sim.data.qvel
        destination_sim.data.act[:] = source_sim.data.act
        destination_sim.forward()

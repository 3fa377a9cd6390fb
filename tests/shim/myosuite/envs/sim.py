"""TEST INFRASTRUCTURE: a mujoco_py-``MjSim``-shaped object over the fp64 oracle (one world). ``sim.model.*`` arrays are live,
writable views of the oracle's model (the reference edits body_mass / geom_size / geom_friction / site_pos / body_pos in
place); arrays the oracle has no use for (rgba, groups) are plain numpy copies of the MJB's."""
import collections

import numpy as np

from oracle import mjb, oracle

MjSimState = collections.namedtuple("MjSimState", "time qpos qvel act udd_state")


class _Opt:
    def __init__(self, om):
        self._om = om

    @property
    def timestep(self):
        return self._om.timestep


class SimModel:
    def __init__(self, src: mjb.MjbModel, om: oracle.OracleModel):
        object.__setattr__(self, "_src", src)
        object.__setattr__(self, "_om", om)
        object.__setattr__(self, "_extra", {})
        object.__setattr__(self, "opt", _Opt(om))

    def __getattr__(self, name):
        om, src, extra = self._om, self._src, self._extra
        if name in ("nq", "nv", "nu", "na", "nbody", "njnt", "ngeom", "nsite", "ntendon"):
            return src.sizes[name]
        if name in extra:
            return extra[name]
        try:
            return getattr(om, name)
        except AttributeError:
            pass
        if name in src.arrays:
            extra[name] = np.array(src.arrays[name]).copy()
            return extra[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        getattr(self, name)[...] = value

    def _id(self, group, name):
        i = self._src.name2id(group, name)
        if i is None or i < 0:
            raise ValueError("No %s named %r" % (group, name))
        return i

    def site_name2id(self, n): return self._id("site", n)
    def geom_name2id(self, n): return self._id("geom", n)
    def body_name2id(self, n): return self._id("body", n)
    def joint_name2id(self, n): return self._id("jnt", n)
    def actuator_name2id(self, n): return self._id("actuator", n)


class SimData:
    _FIELDS = ("qpos", "qvel", "act", "ctrl", "qacc", "qacc_warmstart", "site_xpos", "site_xmat", "xpos", "xquat", "xmat", "geom_xpos", "geom_xmat",
               "ten_length", "actuator_force", "qfrc_actuator")

    def __init__(self, od: oracle.OracleData):
        object.__setattr__(self, "_od", od)

    def __getattr__(self, name):
        return getattr(self._od, name)

    def __setattr__(self, name, value):
        if name == "time":
            self._od.time = float(value)
        else:
            getattr(self._od, name)[...] = value

    @property
    def time(self):
        return self._od.time


class OracleSim:
    def __init__(self, model_path):
        from shim import resolve_model_path

        self.model_path = resolve_model_path(model_path)
        self._src = mjb.load(self.model_path)
        self._om = oracle.OracleModel(self._src)
        self._od = oracle.OracleData(self._om)
        self.model = SimModel(self._src, self._om)
        self.data = SimData(self._od)

    def forward(self):
        self._od.forward()

    def step(self):
        self._od.step(1)

    def advance(self, substeps=1, render=False):
        self._od.step(int(substeps))

    def reset(self):
        self._od.reset()          # mj_resetData: qpos0, zeros, time 0, warm start cleared

    def get_state(self):
        d = self.data
        return MjSimState(d.time, d.qpos.copy(), d.qvel.copy(), d.act.copy(), {})

    def set_state(self, state=None, qpos=None, qvel=None, act=None, time=None):
        if state is not None:
            time, qpos, qvel, act = state.time, state.qpos, state.qvel, state.act
        if time is not None:
            self.data.time = time
        if qpos is not None:
            self.data.qpos[:] = qpos
        if qvel is not None:
            self.data.qvel[:] = qvel
        if act is not None and self.model.na:
            self.data.act[:] = act

    def render(self, *a, **k):
        raise NotImplementedError("shim: no rendering")

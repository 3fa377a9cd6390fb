"""ORACLE / TEST INFRASTRUCTURE -- ctypes front end of ``oracle/myo_oracle.c``.

``OracleModel`` mirrors the ``sim.model`` arrays the reference touches
(/root/reference/src/envs/baoding.py:560-604 writes body_mass / geom_friction / geom_size in
place) and ``OracleData`` mirrors ``sim.data``; every field is a live numpy view of the C arrays.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import mjb

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libmyo_oracle.so")
    src = os.path.join(_HERE, "myo_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libmyo_oracle.so"])
    return so


def build_fast() -> str:
    """bench.py's CPU baseline legs: the same source, -O3 -march=native, built ON the machine that runs it (into the system's
    temporary directory, never shipped). Falls back to the strict build when no compiler is available."""
    import tempfile

    so = os.path.join(tempfile.gettempdir(), "libmyo_oracle_fast_%d.so" % os.getuid())
    src = os.path.join(_HERE, "myo_oracle.c")
    try:
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-std=gnu99", "-Wno-unused-function", "-shared", "-o", so, src, "-lm"])
        return so
    except Exception:
        return build()


def lib(path: str | None = None):
    """``path``: load that build of the oracle instead of the strict one (first call only; bench.py's CPU legs)."""
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(path or build())
        vp, ip, cp, dp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)
        L.o_model_new.restype = vp; L.o_model_new.argtypes = [ip]
        L.o_model_free.argtypes = [vp]
        L.o_model_set_excludes.argtypes = [vp, ctypes.c_int, ip]
        L.o_model_set_opt.argtypes = [vp, ctypes.c_double, dp, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_double]
        L.o_model_field.restype = vp; L.o_model_field.argtypes = [vp, cp, ip, ip]
        L.o_data_new.restype = vp; L.o_data_new.argtypes = [vp]
        L.o_data_free.argtypes = [vp]
        L.o_data_field.restype = vp; L.o_data_field.argtypes = [vp, vp, cp, ip, ip]
        L.o_data_time.restype = ctypes.c_double; L.o_data_time.argtypes = [vp]
        L.o_data_set_time.argtypes = [vp, ctypes.c_double]
        for f in ("o_data_ncon", "o_data_nefc", "o_data_solver_iter", "o_data_unsupported", "o_data_overflow"):
            getattr(L, f).restype = ctypes.c_int; getattr(L, f).argtypes = [vp]
        for f in ("o_kinematics", "o_com_pos", "o_tendon", "o_transmission", "o_crb", "o_factor_m", "o_collision",
                  "o_make_constraint", "o_fwd_position", "o_fwd_velocity", "o_fwd_actuation", "o_fwd_acceleration",
                  "o_fwd_constraint", "o_forward", "o_euler", "o_step", "o_reset", "o_set_const"):
            getattr(L, f).argtypes = [vp, vp]; getattr(L, f).restype = None
        L.o_step_n.argtypes = [vp, vp, ctypes.c_int]
        L.o_batch_step.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp]
        L.o_batch_env_step.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp, dp, ip, dp, ctypes.c_double,
                                       ctypes.c_double, dp, dp, ip]
        L.o_set_solver_tol.argtypes = [ctypes.c_double]
        L.o_muscle_gain.restype = ctypes.c_double; L.o_muscle_gain.argtypes = [ctypes.c_double, ctypes.c_double, dp, ctypes.c_double, dp]
        L.o_muscle_bias.restype = ctypes.c_double; L.o_muscle_bias.argtypes = [ctypes.c_double, dp, ctypes.c_double, dp]
        L.o_muscle_dynamics.restype = ctypes.c_double; L.o_muscle_dynamics.argtypes = [ctypes.c_double, ctypes.c_double, dp]
        _LIB = L
    return _LIB


def _view(ptr, count, is_int):
    if count <= 0:
        return np.zeros(0, dtype=np.int32 if is_int else np.float64)
    ct = ctypes.c_int if is_int else ctypes.c_double
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(count,))


class _Fields:
    _cols = {}

    def _get(self, name):
        raise NotImplementedError

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        v = self._get(name)
        if v is None:
            raise AttributeError(name)
        return v


_MODEL_COLS = dict(body_pos=3, body_quat=4, body_ipos=3, body_iquat=4, body_inertia=3, body_invweight0=2,
                   jnt_solref=2, jnt_solimp=5, jnt_pos=3, jnt_axis=3, jnt_range=2, geom_solref=2, geom_solimp=5,
                   geom_size=3, geom_pos=3, geom_quat=4, geom_friction=3, site_pos=3, site_quat=4,
                   tendon_solref_lim=2, tendon_solimp_lim=5, tendon_range=2, actuator_trnid=2, actuator_dynprm=10,
                   actuator_gainprm=10, actuator_biasprm=10, actuator_ctrlrange=2, actuator_forcerange=2,
                   actuator_gear=6, actuator_lengthrange=2)
_DATA_COLS = dict(xpos=3, xquat=4, xmat=9, xipos=3, ximat=9, xanchor=3, xaxis=3, geom_xpos=3, geom_xmat=9,
                  site_xpos=3, site_xmat=9, subtree_com=3, cdof=6, cinert=10, crb=10, cvel=6, cdof_dot=6,
                  contact_pos=3, contact_frame=9, contact_friction=5, contact_solref=2, contact_solimp=5, efc_KBIP=4)


class OracleModel(_Fields):
    """Oracle-side model. ``src`` is an ``mjb.MjbModel`` (kept for names / sizes)."""

    def __init__(self, src: mjb.MjbModel, njmax: int | None = None, nconmax: int | None = None):
        L = lib()
        self._src = src
        s = src.sizes
        self._njmax = int(njmax if njmax is not None else min(max(s["njmax"], 1), 512))
        self._nconmax = int(nconmax if nconmax is not None else min(max(s["nconmax"], 1), 128))
        sz = (ctypes.c_int * 13)(s["nq"], s["nv"], s["nu"], s["na"], s["nbody"], s["njnt"], s["ngeom"], s["nsite"],
                                 s["ntendon"], s["nwrap"], s["nM"], self._njmax, self._nconmax)
        self._p = L.o_model_new(sz)
        self._cache = {}
        o = src.opt
        g = (ctypes.c_double * 3)(o["gravity0"], o["gravity1"], o["gravity2"])
        L.o_model_set_opt(self._p, o["timestep"], g, o["impratio"], o["cone"], o["disableflags"], src.stat["meaninertia"])
        if s.get("nexclude", 0) > 0:
            sig = np.ascontiguousarray(np.asarray(src.arrays["exclude_signature"]).reshape(-1), dtype=np.int32)
            L.o_model_set_excludes(self._p, int(sig.size), sig.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        for name, arr in src.arrays.items():
            v = self._get(name)
            if v is not None:
                v.reshape(-1)[:] = np.asarray(arr).reshape(-1)
        for k in ("nq", "nv", "nu", "na", "nbody", "njnt", "ngeom", "nsite", "ntendon", "nwrap", "nM"):
            object.__setattr__(self, k, s[k])
        self.timestep = o["timestep"]

    def _get(self, name):
        c = self.__dict__.get("_cache")
        if c is None:
            return None
        if name in c:
            return c[name]
        n, isint = ctypes.c_int(), ctypes.c_int()
        p = lib().o_model_field(self._p, name.encode(), ctypes.byref(n), ctypes.byref(isint))
        if not p:
            return None
        v = _view(p, n.value, isint.value)
        if name in _MODEL_COLS:
            v = v.reshape(-1, _MODEL_COLS[name])
        c[name] = v
        return v

    def name2id(self, group, name):
        return self._src.name2id(group, name)

    def __del__(self):
        try:
            lib().o_model_free(self._p)
        except Exception:
            pass


class OracleData(_Fields):
    def __init__(self, model: OracleModel):
        self._m = model
        self._p = lib().o_data_new(model._p)
        self._cache = {}
        lib().o_reset(model._p, self._p)

    def _get(self, name):
        c = self.__dict__.get("_cache")
        if c is None:
            return None
        if name in c:
            return c[name]
        n, isint = ctypes.c_int(), ctypes.c_int()
        p = lib().o_data_field(self._m._p, self._p, name.encode(), ctypes.byref(n), ctypes.byref(isint))
        if not p:
            return None
        v = _view(p, n.value, isint.value)
        if name in _DATA_COLS:
            v = v.reshape(-1, _DATA_COLS[name])
        elif name in ("ten_J", "actuator_moment", "efc_J", "Mdense", "Lchol"):
            v = v.reshape(-1, self._m.nv) if self._m.nv else v
        c[name] = v
        return v

    @property
    def time(self):
        return lib().o_data_time(self._p)

    @time.setter
    def time(self, t):
        lib().o_data_set_time(self._p, float(t))

    ncon = property(lambda s: lib().o_data_ncon(s._p))
    nefc = property(lambda s: lib().o_data_nefc(s._p))
    solver_iter = property(lambda s: lib().o_data_solver_iter(s._p))
    unsupported = property(lambda s: lib().o_data_unsupported(s._p))
    overflow = property(lambda s: lib().o_data_overflow(s._p))

    def call(self, fn):
        getattr(lib(), fn)(self._m._p, self._p)

    def forward(self):
        self.call("o_forward")

    def step(self, n=1):
        lib().o_step_n(self._m._p, self._p, int(n))

    def reset(self):
        self.call("o_reset")

    def __del__(self):
        try:
            lib().o_data_free(self._p)
        except Exception:
            pass


def load(path) -> tuple[OracleModel, OracleData]:
    m = OracleModel(mjb.load(path))
    return m, OracleData(m)

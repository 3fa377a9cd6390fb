"""Host-side handles over the C ABI: ``Model`` (parsed MJB, ``sim.model``-like arrays) and
``BatchSim`` (n independent worlds on one GPU, stepped by the fused CUDA world kernel).

``BatchSim`` is the batched stand-in for the per-process ``MjSim`` that MyoSuite's ``BaseV0`` builds
from the ``model_path`` kwarg (/root/reference/src/envs/__init__.py:17,29,44,62) and that
/root/reference/src/envs/baoding.py:183,206,560-608 drives through ``step`` / ``set_state`` /
in-place ``sim.model`` writes.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi
from ._capi import TaskCfg, check

_NP_DTYPES = {0: np.float64, 1: np.int32, 2: np.uint8, 3: np.float32, 4: np.uint8}


class Model:
    """Parsed MuJoCo 2.1.0 ``.mjb`` model. Arrays are writable numpy views (``sim.model.*`` semantics):
    edits made before a ``BatchSim`` is created are baked into that batch."""

    def __init__(self, path: str, lib=None):
        self._L = lib if lib is not None else _capi.lib()
        h = C.c_void_p()
        check(self._L, self._L.myo_model_load_mjb(str(path).encode(), C.byref(h)))
        self._h = h
        self.path = str(path)

    def size(self, name: str) -> int:
        v = C.c_int()
        check(self._L, self._L.myo_model_size(self._h, name.encode(), C.byref(v)))
        return v.value

    def opt(self, name: str) -> float:
        v = C.c_double()
        check(self._L, self._L.myo_model_opt(self._h, name.encode(), C.byref(v)))
        return v.value

    def array(self, name: str) -> np.ndarray:
        p, r, c, dt = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        check(self._L, self._L.myo_model_array(self._h, name.encode(), C.byref(p), C.byref(r), C.byref(c), C.byref(dt)))
        n = r.value * c.value
        dtype = np.dtype(_NP_DTYPES[dt.value])
        if n == 0:
            return np.zeros((r.value, c.value) if c.value != 1 else (0,), dtype=dtype)
        buf = (C.c_uint8 * (n * dtype.itemsize)).from_address(p.value)
        a = np.frombuffer(buf, dtype=dtype, count=n)
        return a.reshape(r.value, c.value) if c.value != 1 else a

    def name2id(self, group: str, name: str) -> int:
        i = self._L.myo_model_name2id(self._h, group.encode(), name.encode())
        if i < 0:
            raise KeyError(f"no {group} named {name!r}")
        return i

    def id2name(self, group: str, i: int) -> str:
        return self._L.myo_model_id2name(self._h, group.encode(), int(i)).decode()

    def check(self) -> str:
        """'' when the kernels can run this model, else one line per unsupported feature (``myo_model_check``)."""
        buf = C.create_string_buffer(8192)
        self._L.myo_model_check(self._h, buf, len(buf))
        return buf.value.decode()

    def default_task_cfg(self, kind: int) -> TaskCfg:
        cfg = TaskCfg()
        check(self._L, self._L.myo_task_cfg_default(self._h, int(kind), C.byref(cfg)))
        return cfg

    def __getattr__(self, k):
        if k.startswith("n") and not k.startswith("_"):
            try:
                return self.size(k)
            except _capi.MyoError:
                pass
        raise AttributeError(k)

    def __del__(self):
        try:
            self._L.myo_model_free(self._h)
        except Exception:
            pass


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchSim:
    """``n_worlds`` independent worlds of one model on one device."""

    def __init__(self, model: Model, n_worlds: int, cfg: TaskCfg, device="cuda:0", seed: int = 0):
        self._L = model._L
        self.model = model
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type == "cuda":
            if not torch.cuda.is_available():
                raise _capi.MyoError("BatchSim needs a CUDA device: there is no CPU path")
            torch.cuda.set_device(self.device)
            dev_index = self.device.index or 0
        else:
            dev_index = 0   # only reachable with the test-only host emulation library
            if self._L is _capi._LIB:
                raise _capi.MyoError("the product library steps worlds on CUDA devices only")
        h = C.c_void_p()
        check(self._L, self._L.myo_batch_create(model._h, int(n_worlds), dev_index, C.byref(cfg), C.c_uint64(seed), C.byref(h)))
        self._h = h
        d = [C.c_int() for _ in range(7)]
        check(self._L, self._L.myo_batch_dims(h, *[C.byref(x) for x in d]))
        self.n, self.nq, self.nv, self.na, self.nu, self.nobs, self.nparam = [x.value for x in d]
        f32 = dict(dtype=torch.float32, device=self.device)
        u8 = dict(dtype=torch.uint8, device=self.device)
        self.obs = torch.zeros(self.n, self.nobs, **f32)
        self.reward = torch.zeros(self.n, **f32)
        self.done = torch.zeros(self.n, **u8)
        self.truncated = torch.zeros(self.n, **u8)
        self.terminal_obs = torch.zeros(self.n, self.nobs, **f32)
        self.info = torch.zeros(self.n, _capi.MYO_INFO_TERMS, **f32)

    # -- plumbing ---------------------------------------------------------------------------
    def _stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return None

    def _f32(self, x, shape):
        t = torch.as_tensor(x, dtype=torch.float32, device=self.device).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def launch_info(self):
        v = [C.c_int() for _ in range(4)]
        check(self._L, self._L.myo_batch_launch_info(self._h, *[C.byref(x) for x in v]))
        return dict(lanes_per_world=v[0].value, worlds_per_cta=v[1].value, smem_bytes=v[2].value, regs_per_thread=v[3].value)

    @property
    def launch_count(self) -> int:
        return int(self._L.myo_batch_launch_count(self._h))

    # -- env-level API ------------------------------------------------------------------------
    def reset(self, mask=None) -> torch.Tensor:
        m = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        check(self._L, self._L.myo_batch_reset(self._h, _ptr(m), _ptr(self.obs), self._stream()))
        return self.obs

    def step(self, actions):
        a = self._f32(actions, (self.n, self.nu))
        check(self._L, self._L.myo_batch_step(self._h, _ptr(a), _ptr(self.obs), _ptr(self.reward), _ptr(self.done),
                                              _ptr(self.truncated), _ptr(self.terminal_obs), _ptr(self.info), self._stream()))
        return self.obs, self.reward, self.done, self.truncated

    def get_obs(self) -> torch.Tensor:
        check(self._L, self._L.myo_batch_get_obs(self._h, _ptr(self.obs), self._stream()))
        return self.obs

    # -- sim-level API (MjSim.step / forward / set_state) ---------------------------------------
    def mj_step(self, ctrl=None, nsub: int = 1):
        c = None if ctrl is None else self._f32(ctrl, (self.n, self.nu))
        check(self._L, self._L.myo_batch_mj_step(self._h, _ptr(c), int(nsub), self._stream()))

    def forward(self, ctrl=None):
        c = None if ctrl is None else self._f32(ctrl, (self.n, self.nu))
        check(self._L, self._L.myo_batch_forward(self._h, _ptr(c), self._stream()))

    def get_state(self):
        f32 = dict(dtype=torch.float32, device=self.device)
        qpos, qvel = torch.empty(self.n, self.nq, **f32), torch.empty(self.n, self.nv, **f32)
        act, time = torch.empty(self.n, max(self.na, 1), **f32), torch.empty(self.n, **f32)
        check(self._L, self._L.myo_batch_get_state(self._h, _ptr(qpos), _ptr(qvel), _ptr(act) if self.na else None, _ptr(time), self._stream()))
        return qpos, qvel, act[:, : self.na], time

    def set_state(self, qpos=None, qvel=None, act=None, time=None):
        q = None if qpos is None else self._f32(qpos, (self.n, self.nq))
        v = None if qvel is None else self._f32(qvel, (self.n, self.nv))
        a = None if act is None or not self.na else self._f32(act, (self.n, self.na))
        t = None if time is None else self._f32(time, (self.n,))
        check(self._L, self._L.myo_batch_set_state(self._h, _ptr(q), _ptr(v), _ptr(a), _ptr(t), self._stream()))

    def set_param(self, kind: int, obj_id: int, values):
        ncomp = _capi.PARAM_NCOMP[int(kind)]
        v = self._f32(values, (self.n, ncomp))
        check(self._L, self._L.myo_batch_set_param(self._h, int(kind), int(obj_id), _ptr(v), self._stream()))

    def get_param(self, kind: int, obj_id: int) -> torch.Tensor:
        ncomp = _capi.PARAM_NCOMP[int(kind)]
        v = torch.empty(self.n, ncomp, dtype=torch.float32, device=self.device)
        check(self._L, self._L.myo_batch_get_param(self._h, int(kind), int(obj_id), _ptr(v), self._stream()))
        return v

    def get_task_state(self):
        """Per-world task state (see ``myo_batch_get_task_state``): (ti int32[n, 4] = elapsed, episode, task, flags;
        tf float32[n, 8] = angle1, angle2, x_radius, y_radius, period, pos_dist, rot_dist, -; pose target float32[n, nq])."""
        ti = torch.empty(self.n, _capi.TASK_STATE_I, dtype=torch.int32, device=self.device)
        tf = torch.empty(self.n, _capi.TASK_STATE_F, dtype=torch.float32, device=self.device)
        pt = torch.empty(self.n, self.nq, dtype=torch.float32, device=self.device)
        check(self._L, self._L.myo_batch_get_task_state(self._h, _ptr(ti), _ptr(tf), _ptr(pt), self._stream()))
        return ti, tf, pt

    def set_task_state(self, ti=None, tf=None, pose_target=None):
        i = None if ti is None else torch.as_tensor(ti, device=self.device).to(torch.int32).contiguous()
        f = None if tf is None else self._f32(tf, (self.n, _capi.TASK_STATE_F))
        p = None if pose_target is None else self._f32(pose_target, (self.n, self.nq))
        if i is not None and tuple(i.shape) != (self.n, _capi.TASK_STATE_I):
            raise ValueError("ti must be [n, %d]" % _capi.TASK_STATE_I)
        check(self._L, self._L.myo_batch_set_task_state(self._h, _ptr(i), _ptr(f), _ptr(p), self._stream()))

    def stage(self, name: str) -> torch.Tensor:
        """Per-component parity hook: a stage result of the last ``forward`` / ``mj_step`` substep."""
        sid = _capi.STAGES[name]
        w = C.c_int()
        check(self._L, self._L.myo_batch_stage_dump(self._h, sid, None, C.byref(w), self._stream()))
        dt = torch.int32 if name in _capi.INT_STAGES else torch.float32
        out = torch.zeros(self.n, w.value, dtype=dt, device=self.device)
        check(self._L, self._L.myo_batch_stage_dump(self._h, sid, _ptr(out), C.byref(w), self._stream()))
        return out

    def status(self) -> int:
        v = C.c_int()
        check(self._L, self._L.myo_batch_status(self._h, C.byref(v), self._stream()))
        return v.value

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._L.myo_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

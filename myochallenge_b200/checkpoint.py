"""Checkpoint / normaliser interchange with the reference's artifacts.

The reference saves and loads two kinds of files:

* SB3 zips written by ``RecurrentPPO.save`` / read by ``RecurrentPPO.load`` (``trained_models/**/*.zip``,
  /root/reference/src/main_eval.py:60-75, /root/reference/src/train/trainer.py:49-64 through ``custom_objects``):
  ``data`` (JSON; non-JSON values as base64 cloudpickle under ``":serialized:"``), ``policy.pth`` (torch state dict
  with the keys ``RecurrentPolicy.state_dict_shapes`` lists), ``policy.optimizer.pth``, ``pytorch_variables.pth``,
  ``_stable_baselines3_version``, ``system_info.txt``;
* ``VecNormalize`` pickles written by ``VecNormalize.save`` (``env.pkl``, ``training_env.pkl``,
  ``rl_model_vecnormalize_*.pkl``; /root/reference/src/main_eval.py:65-67, /root/reference/src/main_baoding.py:75).

Neither SB3 nor gym is installed next to this package, so reading maps their classes onto attribute bags and
writing emits pickles that refer to the SB3 / gym classes BY NAME (a pickle stores ``module`` + ``qualname`` for a
class, never its code), built from stand-in classes registered under those module names for the duration of the
dump. What could be verified here: every file this module writes is read back by this module, and every shipped
artifact of the reference is read (tests/test_checkpoint_host.py). What could not: loading the written files with
SB3 itself.
"""
from __future__ import annotations

import base64
import contextlib
import io
import json
import pickle
import sys
import types
import zipfile
from collections import OrderedDict
from typing import Any, Dict, Optional

import numpy as np
import torch

SB3_VERSION = "1.6.2"       # /root/reference/requirements.txt:135
# Explicit (module, name) allowlist for unpickling foreign files: array / scalar reconstruction and plain containers
# only. Everything else - including builtins such as eval / exec / getattr / __import__ - resolves to an inert _Bag.
_SAFE_GLOBALS = {
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy", "float32"), ("numpy", "float64"), ("numpy", "int64"), ("numpy", "int32"),
    ("numpy", "bool_"), ("numpy", "uint8"),
    ("collections", "OrderedDict"), ("collections", "deque"), ("copyreg", "_reconstructor"), ("_codecs", "encode"),
    ("builtins", "object"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "set"), ("builtins", "frozenset"),
    ("builtins", "int"), ("builtins", "float"), ("builtins", "bool"), ("builtins", "str"), ("builtins", "bytes"), ("builtins", "bytearray"),
    ("builtins", "complex"), ("builtins", "slice"), ("builtins", "range"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_parameter"), ("torch", "FloatStorage"), ("torch", "DoubleStorage"),
    ("torch", "LongStorage"), ("torch", "device"), ("torch", "Size"),
}


class _Bag:
    """Stand-in for an SB3 / gym object: keeps constructor arguments and state as attributes."""

    def __init__(self, *a, **k):
        self._args, self._kwargs = a, k

    def __setstate__(self, s):
        if isinstance(s, dict):
            self.__dict__.update(s)
        else:
            self._state = s


class RNNStates(tuple):
    """sb3_contrib.common.recurrent.type_aliases.RNNStates: (pi, vf), each an (h, c) pair of [n_layers, n_envs, H]."""

    def __new__(cls, pi, vf):
        return tuple.__new__(cls, (pi, vf))

    pi = property(lambda self: self[0])
    vf = property(lambda self: self[1])


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _SAFE_GLOBALS:
            return super().find_class(module, name)
        if name == "RNNStates":
            return RNNStates
        if (module, name) == ("torch.storage", "_load_from_bytes"):      # tensors inside cloudpickled attributes
            return lambda b: torch.load(io.BytesIO(b), map_location="cpu", weights_only=True)
        return type(name, (_Bag,), {"__module__": module})


def _loads(b: bytes):
    return _Unpickler(io.BytesIO(b)).load()


def _decode_data(raw: Dict[str, Any]) -> Dict[str, Any]:
    out = {}
    for k, v in raw.items():
        if isinstance(v, dict) and ":serialized:" in v:
            try:
                out[k] = _loads(base64.b64decode(v[":serialized:"]))
            except Exception:       # functions / closures (lr schedules) carry code objects: keep the JSON-side summary
                out[k] = {kk: vv for kk, vv in v.items() if kk != ":serialized:"}
        else:
            out[k] = v
    return out


def sb3_save_path(path: str) -> str:
    """SB3's ``open_path`` in write mode: '.zip' is appended only when the path has no suffix at all (the reference's
    trainer saves 'final_model.pkl', which stays 'final_model.pkl')."""
    import os

    return path + ".zip" if os.path.splitext(path)[1] == "" else path


def sb3_load_path(path: str) -> str:
    """SB3's ``open_path`` in read mode: the path as given if it exists, else with '.zip' appended."""
    import os

    if os.path.exists(path) or os.path.splitext(path)[1] != "":
        return path
    return path + ".zip"


def load_sb3_zip(path: str) -> Dict[str, Any]:
    """-> dict(state_dict, data, optimizer (or None), version). ``data`` holds the constructor / training attributes
    (``policy_kwargs``, ``n_steps``, ``batch_size``, ``gamma``, ..., ``_last_obs``, ``_last_lstm_states``)."""
    with zipfile.ZipFile(path) as z:
        names = set(z.namelist())
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
        raw = json.loads(z.read("data")) if "data" in names else {}
        opt = torch.load(io.BytesIO(z.read("policy.optimizer.pth")), map_location="cpu", weights_only=True) if "policy.optimizer.pth" in names else None
        ver = z.read("_stable_baselines3_version").decode().strip() if "_stable_baselines3_version" in names else None
    data = _decode_data(raw)
    pk = data.get("policy_kwargs")
    if not isinstance(pk, dict):        # stored as {":type:": "<class 'dict'>", ":serialized:": ..., **summary}
        pk = {k: v for k, v in raw.get("policy_kwargs", {}).items() if not k.startswith(":")}
        data["policy_kwargs"] = pk
    return dict(state_dict=OrderedDict((k, v.float()) for k, v in sd.items()), data=data, optimizer=opt, version=ver)


def architecture_of(state_dict) -> Dict[str, Any]:
    """(obs_dim, act_dim, lstm_hidden, pi widths, vf widths) read off the tensor shapes of an SB3 MlpLstmPolicy."""
    H = state_dict["lstm_actor.weight_hh_l0"].shape[1]
    O = state_dict["lstm_actor.weight_ih_l0"].shape[1]
    A = state_dict["action_net.weight"].shape[0]
    widths = {}
    for net in ("policy_net", "value_net"):
        w, l = [], 0
        while f"mlp_extractor.{net}.{2 * l}.weight" in state_dict:
            w.append(state_dict[f"mlp_extractor.{net}.{2 * l}.weight"].shape[0]); l += 1
        widths[net] = tuple(w)
    d_pi = widths["policy_net"][-1] if widths["policy_net"] else H
    ls = tuple(state_dict["log_std"].shape)
    if ls not in ((A,), (d_pi, A)):
        raise ValueError("log_std has shape %s: expected (%d,) or, with use_sde=True, (%d, %d)" % (ls, A, d_pi, A))
    if "lstm_critic.weight_hh_l0" not in state_dict:
        raise ValueError("the checkpoint has no separate critic LSTM (enable_critic_lstm=False is not built)")
    return dict(obs_dim=O, act_dim=A, lstm_hidden=H, pi=widths["policy_net"], vf=widths["value_net"], use_sde=len(ls) == 2)


def sb3_parameter_order(names):
    """The order ``policy.named_parameters()`` yields an MlpLstmPolicy's tensors in (= key order of ``policy.pth`` and
    parameter indices of ``policy.optimizer.pth``): direct parameters first (``log_std``), then the sub-modules in
    registration order - mlp_extractor, action_net, value_net (ActorCriticPolicy._build), lstm_actor, lstm_critic
    (RecurrentActorCriticPolicy.__init__) - as the shipped phase-1 checkpoint shows."""
    rank = {"log_std": 0, "mlp_extractor.policy_net": 1, "mlp_extractor.value_net": 2, "action_net": 3, "value_net": 4, "lstm_actor": 5, "lstm_critic": 6}

    def key(item):
        i, k = item
        for prefix, r in rank.items():
            if k == prefix or k.startswith(prefix + "."):
                return (r, i)
        return (99, i)

    return [k for _, k in sorted(enumerate(names), key=key)]


@contextlib.contextmanager
def _named_classes(spec):
    """Register stand-in classes under foreign module names so pickle refers to them by those names."""
    created, saved = {}, {}
    try:
        for module, name in spec:
            parts = module.split(".")
            for i in range(1, len(parts) + 1):
                mn = ".".join(parts[:i])
                if mn not in sys.modules:
                    sys.modules[mn] = types.ModuleType(mn); saved[mn] = None
            cls = type(name, (), {"__module__": module, "__qualname__": name})
            setattr(sys.modules[module], name, cls)
            created[(module, name)] = cls
        yield created
    finally:
        for mn in saved:
            sys.modules.pop(mn, None)


def _by_reference(module: str, name: str) -> bytes:
    """Pickle of a class/function *reference* (protocol 2 GLOBAL opcode), as cloudpickle emits for importable objects."""
    return b"\x80\x02c" + module.encode() + b"\n" + name.encode() + b"\n."


def _ser(b: bytes, **summary) -> Dict[str, Any]:
    d = {":serialized:": base64.b64encode(b).decode()}
    d.update(summary)
    return d


def save_sb3_zip(path: str, state_dict, data: Optional[Dict[str, Any]] = None, optimizer: Optional[Dict[str, Any]] = None) -> None:
    """Write an SB3-layout zip. ``data``: plain (JSON-able) training attributes; ``policy_class`` and ``policy_kwargs``
    are added in SB3's serialised form (class by reference)."""
    d = dict(data or {})
    arch = architecture_of(state_dict)
    pk = dict(d.pop("policy_kwargs", {}) or {})
    pk.setdefault("lstm_hidden_size", arch["lstm_hidden"])
    pk.setdefault("net_arch", [dict(pi=list(arch["pi"]), vf=list(arch["vf"]))])
    pk.setdefault("enable_critic_lstm", True)
    pk_plain = {k: v for k, v in pk.items() if isinstance(v, (int, float, bool, str, list, dict, type(None)))}
    out = {"policy_class": _ser(_by_reference("sb3_contrib.common.recurrent.policies", "RecurrentActorCriticPolicy"),
                                **{":type:": "<class 'abc.ABCMeta'>", "__module__": "sb3_contrib.common.recurrent.policies"}),
           "policy_kwargs": _ser(pickle.dumps(pk_plain, protocol=2), **{":type:": "<class 'dict'>"}, **{k: v for k, v in pk_plain.items()})}
    for k, v in d.items():
        if isinstance(v, np.ndarray):
            out[k] = _ser(pickle.dumps(v, protocol=2), **{":type:": "<class 'numpy.ndarray'>"})
        else:
            json.dumps(v)        # raises for anything that is not JSON-able: pass arrays or plain values
            out[k] = v
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("data", json.dumps(out, indent=4))
        b = io.BytesIO()
        torch.save(OrderedDict((k, torch.as_tensor(state_dict[k]).detach().cpu().float().clone()) for k in sb3_parameter_order(list(state_dict))), b)
        z.writestr("policy.pth", b.getvalue())
        if optimizer is not None:
            b = io.BytesIO(); torch.save(optimizer, b); z.writestr("policy.optimizer.pth", b.getvalue())
        b = io.BytesIO(); torch.save({}, b); z.writestr("pytorch_variables.pth", b.getvalue())
        z.writestr("_stable_baselines3_version", SB3_VERSION)
        z.writestr("system_info.txt", "written by myochallenge_b200.checkpoint (SB3 zip layout)\n")


# ---- VecNormalize pickles --------------------------------------------------------------------------------------------------
def load_vecnormalize(path: str) -> Dict[str, Any]:
    """-> dict(obs_mean, obs_var, obs_count, ret_mean, ret_var, ret_count, clip_obs, clip_reward, gamma, epsilon,
    training, norm_obs, norm_reward, obs_shape)."""
    with open(path, "rb") as f:
        v = _Unpickler(f).load()
    d = v.__dict__
    o, r = d["obs_rms"].__dict__, d["ret_rms"].__dict__
    return dict(obs_mean=np.asarray(o["mean"], np.float64), obs_var=np.asarray(o["var"], np.float64), obs_count=float(o["count"]),
                ret_mean=float(np.asarray(r["mean"])), ret_var=float(np.asarray(r["var"])), ret_count=float(r["count"]),
                clip_obs=float(d["clip_obs"]), clip_reward=float(d["clip_reward"]), gamma=float(d["gamma"]), epsilon=float(d["epsilon"]),
                training=bool(d.get("training", True)), norm_obs=bool(d.get("norm_obs", True)), norm_reward=bool(d.get("norm_reward", True)),
                obs_shape=tuple(np.asarray(o["mean"]).shape))


def save_vecnormalize(path: str, stats: Dict[str, Any], num_envs: int, act_dim: int) -> None:
    """Write a pickle with the attribute set of SB3 1.6.2's ``VecNormalize.__getstate__`` (no ``venv``, ``class_attributes``
    or ``returns``), referring to the SB3 / gym classes by name."""
    spec = [("stable_baselines3.common.vec_env.vec_normalize", "VecNormalize"), ("stable_baselines3.common.running_mean_std", "RunningMeanStd"),
            ("gym.spaces.box", "Box")]
    with _named_classes(spec) as cls:
        VN, RMS, BoxT = cls[spec[0]], cls[spec[1]], cls[spec[2]]

        def rms(mean, var, count):
            o = RMS.__new__(RMS); o.mean, o.var, o.count = mean, var, count
            return o

        def box(lo, hi, n):
            b = BoxT.__new__(BoxT)
            b.dtype = np.dtype(np.float32); b._shape = (n,); b.low = np.full(n, lo, np.float32); b.high = np.full(n, hi, np.float32)
            b.bounded_below = np.full(n, np.isfinite(lo)); b.bounded_above = np.full(n, np.isfinite(hi)); b._np_random = None
            return b

        n_obs = int(np.asarray(stats["obs_mean"]).size)
        v = VN.__new__(VN)
        v.obs_rms = rms(np.asarray(stats["obs_mean"], np.float64), np.asarray(stats["obs_var"], np.float64), float(stats["obs_count"]))
        v.ret_rms = rms(np.float64(stats["ret_mean"]), np.float64(stats["ret_var"]), float(stats["ret_count"]))
        v.clip_obs, v.clip_reward = float(stats.get("clip_obs", 10.0)), float(stats.get("clip_reward", 10.0))
        v.gamma, v.epsilon = float(stats.get("gamma", 0.99)), float(stats.get("epsilon", 1e-8))
        v.training, v.norm_obs, v.norm_reward = bool(stats.get("training", True)), bool(stats.get("norm_obs", True)), bool(stats.get("norm_reward", True))
        v.norm_obs_keys = None
        v.num_envs = int(num_envs)
        v.observation_space, v.action_space = box(-v.clip_obs, v.clip_obs, n_obs), box(-1.0, 1.0, act_dim)
        v.old_obs, v.old_reward = np.zeros((0,)), np.zeros((0,))
        with open(path, "wb") as f:
            pickle.dump(v, f, protocol=4)

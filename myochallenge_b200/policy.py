"""Host-side handle of the recurrent policy forward (``myo_policy_*`` in the C ABI).

Mirrors the pieces of sb3-contrib's ``RecurrentActorCriticPolicy`` the reference's rollout touches
(constructed by /root/reference/src/train/trainer.py:49-64 as ``RecurrentPPO("MlpLstmPolicy", ...)``
with ``policy_kwargs`` such as
/root/reference/trained_models/curriculum_steps_complete_baoding_winner/01_rsi_static/main.py:194-199):
state-dict keys, ``forward(obs, lstm_states, episode_starts, deterministic)`` returning
``(actions, values, log_prob, lstm_states)``, ``predict_values`` and ``get_distribution``-free sampling.
PyTorch supplies device memory and streams only; the arithmetic is the tcgen05 kernel in
``csrc/myo_policy.cu``.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _capi
from ._capi import PolicyCfg, check


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class RecurrentPolicy:
    """``MlpLstmPolicy`` forward on one GPU: separate actor / critic LSTMs (``enable_critic_lstm=True``,
    ``shared_lstm=False``, one layer), ReLU ``mlp_extractor`` with ``net_arch=[dict(pi=[...], vf=[...])]``,
    ``action_net`` + state-independent ``log_std`` (DiagGaussian) and ``value_net``."""

    def __init__(self, obs_dim: int, act_dim: int, lstm_hidden: int = 256, pi: Sequence[int] = (256, 256),
                 vf: Sequence[int] = (256, 256), max_batch: int = 32768, device="cuda:0", lib=None, use_sde: bool = False,
                 precision: str = "bf16"):
        self._L = lib if lib is not None else _capi.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _capi.MyoError("RecurrentPolicy needs a CUDA device: there is no CPU path")
        torch.cuda.set_device(self.device)
        self.obs_dim, self.act_dim, self.lstm_hidden = int(obs_dim), int(act_dim), int(lstm_hidden)
        self.pi, self.vf = tuple(int(x) for x in pi), tuple(int(x) for x in vf)
        cfg = PolicyCfg()
        cfg.obs_dim, cfg.act_dim, cfg.lstm_hidden = self.obs_dim, self.act_dim, self.lstm_hidden
        cfg.n_pi_layers, cfg.n_vf_layers = len(self.pi), len(self.vf)
        for i, wdt in enumerate(self.pi):
            cfg.pi_layers[i] = wdt
        for i, wdt in enumerate(self.vf):
            cfg.vf_layers[i] = wdt
        self.use_sde = bool(use_sde)
        cfg.use_sde = int(self.use_sde)
        h = C.c_void_p()
        check(self._L, self._L.myo_policy_create(C.byref(cfg), int(max_batch), self.device.index or 0, C.byref(h)))
        self._h = h
        self.max_batch = int(max_batch)
        self.set_precision(precision)
        self._state: Dict[str, torch.Tensor] = {}
        # generalised state-dependent exploration (use_sde=True): the kernel produces the mean and latent_pi; the noise
        # latent_pi . E[w] with one exploration matrix per world and the log-prob come from myo_sde_sample
        self.latent_dim = self.pi[-1] if self.pi else self.lstm_hidden
        self._latent = self._noise_mat = self._std2 = None
        self._noise_epoch, self._sde_seed = 0, 0x5DE
        if self.use_sde:
            if self.act_dim > 64:
                raise ValueError("use_sde supports act_dim <= 64")
            if self.pi:
                self._latent = torch.zeros(self.max_batch, self.latent_dim, dtype=torch.float32, device=self.device)
                check(self._L, self._L.myo_policy_set_latent_out(self._h, _ptr(self._latent)))

    # -- parameters ---------------------------------------------------------------------------
    def state_dict_shapes(self) -> Dict[str, Tuple[int, ...]]:
        H, O, A = self.lstm_hidden, self.obs_dim, self.act_dim
        shp = {"log_std": (self.latent_dim, A) if self.use_sde else (A,)}
        for net in ("lstm_actor", "lstm_critic"):
            shp[f"{net}.weight_ih_l0"] = (4 * H, O)
            shp[f"{net}.weight_hh_l0"] = (4 * H, H)
            shp[f"{net}.bias_ih_l0"] = (4 * H,)
            shp[f"{net}.bias_hh_l0"] = (4 * H,)
        for name, widths in (("policy_net", self.pi), ("value_net", self.vf)):
            d = H
            for l, wdt in enumerate(widths):
                shp[f"mlp_extractor.{name}.{2 * l}.weight"] = (wdt, d)
                shp[f"mlp_extractor.{name}.{2 * l}.bias"] = (wdt,)
                d = wdt
        shp["action_net.weight"] = (A, self.pi[-1] if self.pi else H)
        shp["action_net.bias"] = (A,)
        shp["value_net.weight"] = (1, self.vf[-1] if self.vf else H)
        shp["value_net.bias"] = (1,)
        return shp

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """``sd``: tensors under the SB3 state-dict keys (``policy.pth`` of a RecurrentPPO zip)."""
        for k, shape in self.state_dict_shapes().items():
            if k not in sd:
                raise KeyError(f"state dict lacks {k}")
            t = torch.as_tensor(sd[k]).to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{k}: expected shape {shape}, got {tuple(t.shape)}")
            self._state[k] = t
            if k == "log_std" and self.use_sde:      # the kernel's own Gaussian head is unused: zero log_std, zero noise -> mean
                t = torch.zeros(self.act_dim, dtype=torch.float32, device=self.device)
                self._noise_mat = None                # exploration matrices follow log_std: resample before the next forward
            check(self._L, self._L.myo_policy_set_weight(self._h, k.encode(), _ptr(t), t.numel(), self._stream()))

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self._state)

    def init_random(self, seed: int = 0, log_std_init: float = -2.0) -> Dict[str, torch.Tensor]:
        """Random initialisation the way torch / SB3 do it with ``ortho_init=False``: LSTM and Linear
        parameters ~ U(-1/sqrt(fan), 1/sqrt(fan)), ``log_std`` = ``log_std_init``."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        sd = {}
        H = self.lstm_hidden
        for k, shape in self.state_dict_shapes().items():
            if k == "log_std":
                sd[k] = torch.full(shape, float(log_std_init))
                continue
            if k.startswith("lstm_"):
                bound = 1.0 / math.sqrt(H)
            else:
                fan_in = self.state_dict_shapes()[k.rsplit(".", 1)[0] + ".weight"][1]
                bound = 1.0 / math.sqrt(fan_in)
            sd[k] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        self.load_state_dict(sd)
        return sd

    def set_obs_norm(self, mean: Optional[torch.Tensor], var: Optional[torch.Tensor], epsilon: float = 1e-8, clip_obs: float = 10.0):
        """Fuse ``VecNormalize.normalize_obs`` into the kernel's input load (``None`` switches it off)."""
        if mean is None:
            check(self._L, self._L.myo_policy_set_obs_norm(self._h, None, None, C.c_float(epsilon), C.c_float(clip_obs), self._stream()))
            return
        m = torch.as_tensor(mean).to(device=self.device, dtype=torch.float32).contiguous()
        v = torch.as_tensor(var).to(device=self.device, dtype=torch.float32).contiguous()
        check(self._L, self._L.myo_policy_set_obs_norm(self._h, _ptr(m), _ptr(v), C.c_float(epsilon), C.c_float(clip_obs), self._stream()))
        if m is not mean or v is not var:
            torch.cuda.current_stream(self.device).synchronize()   # m, v are temporaries: keep them alive until the copy ran

    def seed(self, seed: int) -> None:
        self._sde_seed = int(seed) ^ 0x5DE5DE
        if not self.use_sde:
            check(self._L, self._L.myo_policy_seed(self._h, C.c_uint64(seed)))

    def reset_noise(self, n_envs: Optional[int] = None) -> None:
        """``policy.reset_noise(n_envs)``: draw one exploration matrix per world from the current ``log_std`` (SB3 calls it at the
        start of every rollout and every ``sde_sample_freq`` steps)."""
        if not self.use_sde:
            return
        n = self.max_batch if n_envs is None else int(n_envs)
        L, A = self.latent_dim, self.act_dim
        if self._noise_mat is None or self._noise_mat.shape[0] < n:
            self._noise_mat = torch.empty(n, L, A, dtype=torch.bfloat16, device=self.device)
            self._std2 = torch.empty(L, A, dtype=torch.float32, device=self.device)
        self._noise_epoch += 1
        check(self._L, self._L.myo_sde_reset_noise(_ptr(self._noise_mat), _ptr(self._std2), _ptr(self._state["log_std"]), n, L, A,
                                                   C.c_uint64(self._sde_seed), C.c_uint32(self._noise_epoch), self._stream()))

    # -- forward ------------------------------------------------------------------------------
    def set_precision(self, precision: str):
        """"bf16": bf16 operands with fp32 accumulation on the tensor cores (the rollout path). "fp32": plain fp32 arithmetic,
        the precision sb3-contrib's torch policy runs at - for evaluating trained reference checkpoints (src/main_eval.py)."""
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        check(self._L, self._L.myo_policy_set_precision(self._h, 1 if precision == "fp32" else 0))
        self.precision = precision

    def initial_state(self, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(h, c), each ``[2, n, H]`` (0 = actor LSTM, 1 = critic LSTM), zero-initialised."""
        z = torch.zeros(2, n, self.lstm_hidden, dtype=torch.float32, device=self.device)
        return z, z.clone()

    def forward(self, obs: torch.Tensor, lstm_states: Tuple[torch.Tensor, torch.Tensor], episode_starts: torch.Tensor,
                deterministic: bool = False, noise: Optional[torch.Tensor] = None, out=None):
        """One rollout step for ``n`` worlds. ``lstm_states`` are updated in place. Returns
        ``(actions[n, act_dim] (unclipped), values[n], log_prob[n], lstm_states)``."""
        n = obs.shape[0]
        h, c = lstm_states
        if obs.dtype != torch.float32 or not obs.is_contiguous() or obs.device != self.device:
            obs = obs.to(device=self.device, dtype=torch.float32).contiguous()
        es = episode_starts
        if es.dtype != torch.uint8 or es.device != self.device or not es.is_contiguous():
            es = (es != 0).to(device=self.device, dtype=torch.uint8).contiguous()
        if out is None:
            actions = torch.empty(n, self.act_dim, dtype=torch.float32, device=self.device)
            values = torch.empty(n, dtype=torch.float32, device=self.device)
            logp = torch.empty(n, dtype=torch.float32, device=self.device)
        else:
            actions, values, logp = out
        if self.use_sde:
            if noise is not None:
                raise ValueError("use_sde: the noise is state dependent (latent_pi . exploration matrix); pass deterministic=True or nothing")
            nz = None
            deterministic_kernel = True
        else:
            deterministic_kernel = deterministic
        if deterministic:
            nz = None
        elif noise is not None:
            nz = noise.to(device=self.device, dtype=torch.float32).contiguous()
        else:
            nz = None   # in-kernel sampling when seed() was called, else the mean
        if deterministic_kernel and noise is None:
            # the kernel samples only when a seed is set; make "deterministic" win over a seeded policy
            zero = getattr(self, "_zero_noise", None)
            if zero is None or zero.shape[0] < n:
                zero = torch.zeros(max(n, 1), self.act_dim, dtype=torch.float32, device=self.device)
                self._zero_noise = zero
            nz = zero[:n]
        check(self._L, self._L.myo_policy_forward(self._h, int(n), _ptr(obs), _ptr(h), _ptr(c), _ptr(es), _ptr(nz), _ptr(actions),
                                                  _ptr(values), _ptr(logp), self._stream()))
        if self.use_sde and not deterministic:
            if self._noise_mat is None or self._noise_mat.shape[0] < n:
                self.reset_noise(max(n, self.max_batch if n <= self.max_batch else n))
            lat = self._latent if self._latent is not None else h[0]
            check(self._L, self._L.myo_sde_sample(_ptr(lat), int(lat.shape[1]), _ptr(self._noise_mat), _ptr(self._std2), _ptr(actions), _ptr(logp), int(n),
                                                  self.latent_dim, self.act_dim, self._stream()))
        return actions, values, logp, (h, c)

    def predict_values(self, obs: torch.Tensor, lstm_states: Tuple[torch.Tensor, torch.Tensor], episode_starts: Optional[torch.Tensor] = None):
        """``RecurrentActorCriticPolicy.predict_values``: critic value of ``obs`` from the given LSTM states, which are
        left untouched (the forward runs on copies)."""
        h, c = lstm_states
        h, c = h.clone().contiguous(), c.clone().contiguous()
        n = obs.shape[0]
        es = torch.zeros(n, dtype=torch.uint8, device=self.device) if episode_starts is None else episode_starts
        _, values, _, _ = self.forward(obs, (h, c), es, deterministic=True)
        return values

    @property
    def launch_count(self) -> int:
        return int(self._L.myo_policy_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._L.myo_policy_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def torch_reference_forward(sd: Dict[str, torch.Tensor], obs, h, c, episode_starts, noise=None, pi=(), vf=()):
    """fp32 PyTorch restatement of ``RecurrentActorCriticPolicy.forward`` for one step (floating-point
    reference of the tcgen05 kernel; used by tests and the smoke check). All tensors on one device."""
    outs = []
    hs, cs = [], []
    keep = (1.0 - episode_starts.float()).unsqueeze(1)
    for net, name, widths in ((0, "lstm_actor", pi), (1, "lstm_critic", vf)):
        h0, c0 = h[net] * keep, c[net] * keep
        gates = obs @ sd[f"{name}.weight_ih_l0"].T + sd[f"{name}.bias_ih_l0"] + h0 @ sd[f"{name}.weight_hh_l0"].T + sd[f"{name}.bias_hh_l0"]
        i, f, g, o = gates.chunk(4, dim=1)
        c1 = torch.sigmoid(f) * c0 + torch.sigmoid(i) * torch.tanh(g)
        h1 = torch.sigmoid(o) * torch.tanh(c1)
        hs.append(h1); cs.append(c1)
        x = h1
        mlp = "policy_net" if net == 0 else "value_net"
        for l in range(len(widths)):
            x = torch.relu(x @ sd[f"mlp_extractor.{mlp}.{2 * l}.weight"].T + sd[f"mlp_extractor.{mlp}.{2 * l}.bias"])
        outs.append(x)
    mean = outs[0] @ sd["action_net.weight"].T + sd["action_net.bias"]
    values = (outs[1] @ sd["value_net.weight"].T + sd["value_net.bias"]).squeeze(1)
    log_std = sd["log_std"]
    z = torch.zeros_like(mean) if noise is None else noise
    actions = mean + torch.exp(log_std) * z
    logp = (-0.5 * z * z - log_std - 0.5 * math.log(2 * math.pi)).sum(1)
    return actions, values, logp, torch.stack(hs), torch.stack(cs)

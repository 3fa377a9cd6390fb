"""Policy update of the PPO loop on the device (``myo_ppo_*`` in the C ABI) and the
``RecurrentPPO`` front end the reference's trainer drives.

Mirrors what /root/reference/src/train/trainer.py:49-71 builds and calls - ``RecurrentPPO("MlpLstmPolicy", env,
n_steps=..., batch_size=..., n_epochs=..., learning_rate=..., clip_range=..., ent_coef=..., vf_coef=...,
max_grad_norm=..., policy_kwargs=...)`` then ``agent.learn(total_timesteps=...)`` - with sb3-contrib's argument names
and defaults (``RecurrentPPO.__init__``: lr 3e-4, n_steps 128, batch_size 128, n_epochs 10, gamma 0.99, gae_lambda
0.95, clip_range 0.2, ent_coef 0, vf_coef 0.5, max_grad_norm 0.5; Adam eps 1e-5 from ActorCriticPolicy).

Differences by design, documented in DESIGN.md: a minibatch is ``batch_size // n_steps`` whole world sequences (chosen
by a random permutation of the worlds each epoch) instead of ``batch_size`` consecutive samples of the flattened
buffer cut at a random offset; the forward / backward GEMMs run on bf16 operands with fp32 accumulation, the same
rounding the rollout kernel applies, unless ``precision="fp32"``.

Across GPUs every rank updates from its own worlds' minibatch and the flat gradient bucket is summed with ONE
all-reduce per optimiser step (NCCL over NVLink); the gradient norm for clipping is taken after the reduction, on
identical data on every rank, so it needs no second collective.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _capi
from ._capi import PolicyCfg, PPOHyper, check
from .policy import RecurrentPolicy
from .rollout import RecurrentRolloutBuffer, collect_rollouts

STAT_NAMES = ("policy_loss", "value_loss", "entropy_loss", "approx_kl", "clip_fraction", "loss", "adv_mean", "adv_std")


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def allreduce_flat(grad: torch.Tensor) -> float:
    """Sum the flat gradient bucket over the ranks in place (one collective per optimiser step); returns the scale that
    turns the sum into the mean (1 / world_size), which the Adam kernel applies together with the clip factor."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()
    return 1.0


class PPOUpdate:
    """Flat-parameter PPO optimiser for a ``RecurrentPolicy`` architecture: loss + gradient of a minibatch of whole
    sequences, gradient all-reduce, clip_grad_norm_ and Adam."""

    def __init__(self, policy: RecurrentPolicy, n_steps: int, batch_worlds: int, precision: str = "bf16", learning_rate: float = 3e-4,
                 clip_range: float = 0.2, clip_range_vf: Optional[float] = None, ent_coef: float = 0.0, vf_coef: float = 0.5,
                 max_grad_norm: float = 0.5, normalize_advantage: bool = True, betas=(0.9, 0.999), adam_eps: float = 1e-5):
        self._L = _capi.lib()
        self.device = policy.device
        self.policy = policy
        self.n_steps, self.batch_worlds = int(n_steps), int(batch_worlds)
        cfg = PolicyCfg()
        cfg.obs_dim, cfg.act_dim, cfg.lstm_hidden = policy.obs_dim, policy.act_dim, policy.lstm_hidden
        cfg.n_pi_layers, cfg.n_vf_layers = len(policy.pi), len(policy.vf)
        for i, w in enumerate(policy.pi):
            cfg.pi_layers[i] = w
        for i, w in enumerate(policy.vf):
            cfg.vf_layers[i] = w
        cfg.use_sde = int(getattr(policy, "use_sde", False))
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        h = C.c_void_p()
        check(self._L, self._L.myo_ppo_create(C.byref(cfg), self.n_steps, self.batch_worlds, 1 if precision == "bf16" else 0,
                                              self.device.index or 0, C.byref(h)))
        self._h = h
        self.n_params = int(self._L.myo_ppo_param_count(h))
        f = dict(dtype=torch.float32, device=self.device)
        self.params = torch.zeros(self.n_params, **f)
        self._bucket = torch.zeros(self.n_params + 1, **f)       # flat gradient + one slot for approx_kl (reduced with it)
        self.grad = self._bucket[: self.n_params]
        self.exp_avg = torch.zeros(self.n_params, **f)
        self.exp_avg_sq = torch.zeros(self.n_params, **f)
        self.stats = torch.zeros(len(STAT_NAMES), **f)
        self.grad_norm = torch.zeros(1, **f)
        self.learning_rate, self.betas, self.adam_eps, self.max_grad_norm = learning_rate, betas, adam_eps, max_grad_norm
        self.hyper = PPOHyper(clip_range, -1.0 if clip_range_vf is None else clip_range_vf, ent_coef, vf_coef, int(normalize_advantage))
        self.step_count = 0
        self._views: Dict[str, torch.Tensor] = {}
        off, num = C.c_int64(), C.c_int64()
        for k, shape in policy.state_dict_shapes().items():
            check(self._L, self._L.myo_ppo_param_offset(h, k.encode(), C.byref(off), C.byref(num)))
            self._views[k] = self.params[off.value: off.value + num.value].view(shape)
            self._views["grad/" + k] = self.grad[off.value: off.value + num.value].view(shape)
        if policy.state_dict():
            self.load_state_dict(policy.state_dict())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- parameters ---------------------------------------------------------------------------------------------------
    def load_state_dict(self, sd):
        for k in self.policy.state_dict_shapes():
            self._views[k].copy_(torch.as_tensor(sd[k]).to(device=self.device, dtype=torch.float32))

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Views into the flat parameter vector under the SB3 state-dict keys."""
        return {k: self._views[k] for k in self.policy.state_dict_shapes()}

    def grad_dict(self) -> Dict[str, torch.Tensor]:
        return {k: self._views["grad/" + k] for k in self.policy.state_dict_shapes()}

    def push_to_policy(self):
        """Hand the updated parameters to the rollout kernel (re-packs its bf16 weight images)."""
        self.policy.load_state_dict({k: v.clone() for k, v in self.state_dict().items()})

    # -- one optimiser step -----------------------------------------------------------------------------------------------
    def minibatch_grad(self, buf: RecurrentRolloutBuffer, world_idx: torch.Tensor, n_steps: Optional[int] = None):
        """Loss statistics + flat gradient of the sequences of ``world_idx`` (int32 device tensor)."""
        idx = world_idx.to(device=self.device, dtype=torch.int32).contiguous()
        T = buf.n_steps if n_steps is None else n_steps
        check(self._L, self._L.myo_ppo_minibatch_grad(
            self._h, _p(self.params), int(T), int(buf.n_envs), _p(idx), int(idx.numel()), _p(buf.observations), _p(buf.actions),
            _p(buf.episode_starts), _p(buf.values), _p(buf.log_probs), _p(buf.advantages), _p(buf.returns), _p(buf.h0), _p(buf.c0),
            C.byref(self.hyper), _p(self.grad), _p(self.stats), self._stream()))
        return self.stats

    def all_reduce_grad(self, with_kl: bool = False) -> float:
        """One collective per optimiser step. ``with_kl``: approx_kl of this rank's minibatch rides in the last slot of
        the bucket, so every rank sees the same (mean) value and takes the same early-stop branch."""
        if with_kl:
            self._bucket[self.n_params] = self.stats[3]
            return allreduce_flat(self._bucket)
        return allreduce_flat(self.grad)

    def adam_step(self, grad_scale: float = 1.0, learning_rate: Optional[float] = None):
        self.step_count += 1
        lr = self.learning_rate if learning_rate is None else learning_rate
        check(self._L, self._L.myo_ppo_adam_step(self._h, _p(self.params), _p(self.grad), _p(self.exp_avg), _p(self.exp_avg_sq), self.step_count,
                                                 lr, self.betas[0], self.betas[1], self.adam_eps,
                                                 -1.0 if self.max_grad_norm is None else self.max_grad_norm, grad_scale, _p(self.grad_norm),
                                                 self._stream()))

    # -- RecurrentPPO.train ----------------------------------------------------------------------------------------------------
    def train(self, buf: RecurrentRolloutBuffer, n_epochs: int = 10, generator: Optional[torch.Generator] = None, target_kl: Optional[float] = None,
              learning_rate: Optional[float] = None) -> Dict[str, float]:
        """``n_epochs`` passes over the rollout in minibatches of ``batch_worlds`` world sequences; returns the mean of the
        per-minibatch statistics (SB3's ``train/*`` log entries)."""
        n = buf.n_envs
        acc = torch.zeros_like(self.stats)
        count = 0
        stop = False
        for _ in range(n_epochs):
            perm = torch.randperm(n, generator=generator, device="cpu").to(self.device, dtype=torch.int32)
            for s in range(0, n, self.batch_worlds):
                self.minibatch_grad(buf, perm[s: s + self.batch_worlds])
                scale = self.all_reduce_grad(with_kl=target_kl is not None)
                # SB3 stops before the optimiser step of the offending minibatch; the test is on the rank-mean approx_kl
                # (reduced with the gradient) so no rank leaves the loop while the others wait in the collective
                if target_kl is not None and float(self._bucket[self.n_params]) * scale > 1.5 * target_kl:
                    stop = True
                    break
                self.adam_step(scale, learning_rate)
                acc += self.stats
                count += 1
            if stop:
                break
        self.push_to_policy()
        mean = (acc / max(count, 1)).tolist()
        out = {"train/" + k: v for k, v in zip(STAT_NAMES, mean)}
        out["train/n_updates"] = count
        out["train/grad_norm"] = float(self.grad_norm)
        return out

    @property
    def launch_count(self) -> int:
        return int(self._L.myo_ppo_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._L.myo_ppo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RecurrentPPO:
    """``sb3_contrib.RecurrentPPO`` as /root/reference/src/train/trainer.py:49-71 uses it, over the device path:
    ``env`` is a ``rollout.DeviceVecNormalize`` (or a bare ``MyoVecEnv``); ``learn`` alternates ``collect_rollouts`` and
    ``train`` without leaving HBM."""

    def __init__(self, policy="MlpLstmPolicy", env=None, learning_rate=3e-4, n_steps=128, batch_size=128, n_epochs=10, gamma=0.99,
                 gae_lambda=0.95, clip_range=0.2, clip_range_vf=None, normalize_advantage=True, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5,
                 target_kl=None, policy_kwargs=None, seed=0, precision="bf16", device=None, use_sde=False, sde_sample_freq=-1, **_ignored):
        if policy != "MlpLstmPolicy":
            raise ValueError("only MlpLstmPolicy is built (the policy every reference run uses)")
        if env is None:
            raise ValueError("env is required")
        self.env = env
        venv = getattr(env, "venv", env)
        sim = venv.sim
        self.device = sim.device if device is None else torch.device(device)
        kw = dict(policy_kwargs or {})
        arch = kw.get("net_arch", [dict(pi=[64, 64], vf=[64, 64])])
        arch = arch[0] if isinstance(arch, (list, tuple)) and arch and isinstance(arch[0], dict) else dict(pi=list(arch), vf=list(arch))
        self.policy = RecurrentPolicy(sim.nobs, sim.nu, kw.get("lstm_hidden_size", 256), tuple(arch.get("pi", ())), tuple(arch.get("vf", ())),
                                      max_batch=venv.num_envs, device=self.device, use_sde=use_sde)
        self.use_sde, self.sde_sample_freq = bool(use_sde), int(sde_sample_freq)
        self.policy.init_random(seed, log_std_init=kw.get("log_std_init", 0.0))
        self.policy.seed(seed + 1 + 7919 * self._rank())       # same weights on every rank, different action noise
        if callable(clip_range):           # SB3 accepts schedules of the remaining progress; the kernels take the value at 1.0
            clip_range = float(clip_range(1.0))
        if callable(clip_range_vf):
            clip_range_vf = float(clip_range_vf(1.0))
        self.n_steps, self.n_epochs, self.gamma, self.gae_lambda, self.target_kl = n_steps, n_epochs, gamma, gae_lambda, target_kl
        self.n_envs = venv.num_envs
        self.batch_worlds = max(1, min(self.n_envs, batch_size // n_steps))
        self.update = PPOUpdate(self.policy, n_steps, self.batch_worlds, precision, learning_rate if not callable(learning_rate) else learning_rate(1.0),
                                clip_range, clip_range_vf, ent_coef, vf_coef, max_grad_norm, normalize_advantage)
        self._lr_schedule = learning_rate if callable(learning_rate) else None
        self.buffer = RecurrentRolloutBuffer(n_steps, self.n_envs, sim.nobs, sim.nu, self.policy.lstm_hidden, self.device, gamma, gae_lambda)
        if hasattr(env, "attach"):
            env.attach(self.policy)
        self._gen = torch.Generator(device="cpu").manual_seed(seed)
        self.num_timesteps = 0
        self._state = None
        self.logs = []

    # -- SB3 zip interchange (RecurrentPPO.save / RecurrentPPO.load; /root/reference/src/main_eval.py:60-75) -----------------------
    def get_parameters(self):
        return {"policy": {k: v.detach().cpu().clone() for k, v in self.update.state_dict().items()}}

    def set_parameters(self, state_dict):
        sd = state_dict.get("policy", state_dict)
        self.update.load_state_dict(sd)
        self.update.push_to_policy()

    def save(self, path: str):
        from . import checkpoint

        u = self.update
        data = dict(n_envs=self.n_envs, num_timesteps=self.num_timesteps, n_steps=self.n_steps, batch_size=self.batch_worlds * self.n_steps,
                    n_epochs=self.n_epochs, gamma=self.gamma, gae_lambda=self.gae_lambda, ent_coef=u.hyper.ent_coef, vf_coef=u.hyper.vf_coef,
                    max_grad_norm=u.max_grad_norm, learning_rate=u.learning_rate, clip_range=u.hyper.clip_range,
                    clip_range_vf=None if u.hyper.clip_range_vf <= 0 else u.hyper.clip_range_vf, normalize_advantage=bool(u.hyper.normalize_advantage),
                    target_kl=self.target_kl, use_sde=self.use_sde, sde_sample_freq=self.sde_sample_freq, _n_updates=u.step_count,
                    policy_kwargs=dict(lstm_hidden_size=self.policy.lstm_hidden, net_arch=[dict(pi=list(self.policy.pi), vf=list(self.policy.vf))],
                                       enable_critic_lstm=True, ortho_init=False))
        # torch.optim.Adam state over the parameters in state-dict order, as policy.optimizer.pth holds it
        names = checkpoint.sb3_parameter_order(list(self.policy.state_dict_shapes()))
        views = lambda flat: [flat[o: o + n].view(shp).detach().cpu().clone() for (o, n, shp) in
                              ((u._views[k].storage_offset(), u._views[k].numel(), u._views[k].shape) for k in names)]
        m, v = views(u.exp_avg), views(u.exp_avg_sq)
        opt = {"state": {i: {"step": torch.tensor(float(u.step_count)), "exp_avg": m[i], "exp_avg_sq": v[i]} for i in range(len(names))} if u.step_count else {},
               "param_groups": [{"lr": u.learning_rate, "betas": tuple(u.betas), "eps": u.adam_eps, "weight_decay": 0, "amsgrad": False,
                                 "params": list(range(len(names)))}]}
        checkpoint.save_sb3_zip(checkpoint.sb3_save_path(path), self.get_parameters()["policy"], data, opt)

    @classmethod
    def load(cls, path: str, env=None, custom_objects=None, **kwargs):
        """Rebuild an agent from an SB3 zip (the reference's ``RecurrentPPO.load(path, env=..., custom_objects=...)``):
        architecture from the tensors, hyper-parameters from ``data`` (overridden by ``custom_objects`` / kwargs), Adam
        moments from ``policy.optimizer.pth`` when present."""
        from . import checkpoint

        ck = checkpoint.load_sb3_zip(checkpoint.sb3_load_path(path))
        d = dict(ck["data"])
        d.update(custom_objects or {})
        d.update(kwargs)
        arch = checkpoint.architecture_of(ck["state_dict"])

        def num(key, default):
            v = d.get(key, default)
            if callable(v):
                return v
            return default if isinstance(v, dict) or v is None else v       # serialised schedules come back as dict summaries

        agent = cls("MlpLstmPolicy", env, learning_rate=num("learning_rate", 3e-4), n_steps=int(num("n_steps", 128)),
                    batch_size=int(num("batch_size", 128)), n_epochs=int(num("n_epochs", 10)), gamma=float(num("gamma", 0.99)),
                    gae_lambda=float(num("gae_lambda", 0.95)), clip_range=num("clip_range", 0.2), clip_range_vf=d.get("clip_range_vf") if not isinstance(d.get("clip_range_vf"), dict) else None,
                    normalize_advantage=bool(num("normalize_advantage", True)), ent_coef=float(num("ent_coef", 0.0)), vf_coef=float(num("vf_coef", 0.5)),
                    max_grad_norm=float(num("max_grad_norm", 0.5)), target_kl=d.get("target_kl") if not isinstance(d.get("target_kl"), dict) else None,
                    policy_kwargs=dict(lstm_hidden_size=arch["lstm_hidden"], net_arch=[dict(pi=list(arch["pi"]), vf=list(arch["vf"]))]),
                    seed=int(d["seed"]) if isinstance(d.get("seed"), int) else 0, precision=d.get("precision", "bf16"),
                    use_sde=arch["use_sde"], sde_sample_freq=int(num("sde_sample_freq", -1)))
        agent.set_parameters(ck["state_dict"])
        agent.num_timesteps = int(d.get("num_timesteps", 0)) if not isinstance(d.get("num_timesteps"), dict) else 0
        opt = ck.get("optimizer")
        if opt and opt.get("state"):
            u = agent.update
            names = list(agent.policy.state_dict_shapes())
            # SB3 registers parameters in module order, which is the state-dict order of policy.pth
            order = list(ck["state_dict"].keys())
            for i, k in enumerate(order):
                st = opt["state"].get(i)
                if st is None or k not in names:
                    continue
                o, n = u._views[k].storage_offset(), u._views[k].numel()
                u.exp_avg[o: o + n].copy_(st["exp_avg"].reshape(-1)); u.exp_avg_sq[o: o + n].copy_(st["exp_avg_sq"].reshape(-1))
                u.step_count = int(st["step"])
        return agent

    @staticmethod
    def _rank() -> int:
        import torch.distributed as dist

        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0

    @staticmethod
    def _world_size() -> int:
        import torch.distributed as dist

        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def predict(self, observation, state=None, episode_start=None, deterministic: bool = False):
        """``model.predict(obs, state=lstm_states, episode_start=..., deterministic=...)`` -> (actions clipped to the action space,
        lstm_states), on device tensors of all worlds (observations already normalised unless the env's moments are attached)."""
        obs = torch.as_tensor(observation, dtype=torch.float32, device=self.device)
        n = obs.shape[0]
        if state is None:
            state = self.policy.initial_state(n)
        if episode_start is None:
            episode_start = torch.zeros(n, dtype=torch.uint8, device=self.device)
        es = torch.as_tensor(episode_start).to(self.device)
        actions, _, _, state = self.policy.forward(obs, state, es, deterministic=deterministic)
        return actions.clamp(-1.0, 1.0), state

    def evaluate(self, n_episodes: int = 2000, deterministic: bool = True):
        """The reference's evaluation loop (src/main_eval.py) over this agent's env; see ``evaluate.evaluate_policy``."""
        from .evaluate import evaluate_policy
        from .rollout import DeviceVecNormalize

        norm = self.env if isinstance(self.env, DeviceVecNormalize) else None
        out = evaluate_policy(self.policy, getattr(self.env, "venv", self.env), n_episodes, deterministic, norm)
        self._state = None          # the env was reset by the evaluation: the next learn() starts from a fresh reset
        return out

    def learn(self, total_timesteps: int, callback=None, reset_num_timesteps: bool = True, **_ignored):
        """``agent.learn(total_timesteps, callback=[...], reset_num_timesteps=True)`` (/root/reference/src/train/trainer.py:67-71).
        ``callback``: None, a plain ``callback(agent, log) -> bool``, a ``callbacks.BaseCallback`` or a list of them."""
        from .callbacks import BaseCallback, CallbackList

        cb = None
        if isinstance(callback, (list, tuple)):
            cb = CallbackList(list(callback))
        elif isinstance(callback, BaseCallback):
            cb = callback
        if cb is not None:
            cb.init_callback(self)
            cb.on_training_start()
        if reset_num_timesteps:
            self.num_timesteps = 0
        norm = self.env if hasattr(self.env, "obs_rms") and hasattr(self.env.obs_rms, "sync") else None

        def fresh_reset():
            # reset_device() folds this rank's reset observations into obs_rms: merge them across ranks at once, so the
            # state every rank enters the next interval with (the next ``base``) is the same everywhere
            base = norm.obs_rms.state.clone() if norm is not None else None
            self._obs = self.env.reset_device()
            if norm is not None:
                norm.obs_rms.sync(base)
                norm._push_obs_norm()
            self._starts = torch.ones(self.n_envs, dtype=torch.uint8, device=self.device)
            self._state = self.policy.initial_state(self.n_envs)

        if self._state is None:
            fresh_reset()
        start = self.num_timesteps
        while self.num_timesteps - start < total_timesteps:
            base = (norm.obs_rms.state.clone(), norm.ret_rms.state.clone()) if norm is not None else None
            self._obs, self._starts = collect_rollouts(self.env, self.policy, self.buffer, self._state, self._obs, self._starts,
                                                               sde_sample_freq=self.sde_sample_freq)
            if norm is not None:          # ranks saw different worlds: merge what each added to the running moments (no-op on one GPU)
                norm.obs_rms.sync(base[0]); norm.ret_rms.sync(base[1])
                norm._push_obs_norm()
            self.num_timesteps += self.n_steps * self.n_envs * self._world_size()
            lr = None
            if self._lr_schedule is not None:
                lr = self._lr_schedule(max(0.0, 1.0 - (self.num_timesteps - start) / float(total_timesteps)))
            log = self.update.train(self.buffer, self.n_epochs, self._gen, self.target_kl, lr)
            log["time/total_timesteps"] = self.num_timesteps
            log["rollout/reward_mean"] = float(self.buffer.rewards.mean())
            self.logs.append(log)
            if cb is not None:
                if not cb.on_rollout_end(self.n_steps, log):
                    break
            elif callback is not None and callback(self, log) is False:
                break
            if self._state is None:        # a callback evaluated on the training env and reset it
                fresh_reset()
        if cb is not None:
            cb.on_training_end()
        return self

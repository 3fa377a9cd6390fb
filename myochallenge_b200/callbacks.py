"""The SB3 callbacks the reference's training scripts pass to ``agent.learn`` (/root/reference/src/main_baoding.py:84-104:
``EvalCallback`` and ``CheckpointCallback``; ``BaseCallback`` is what /root/reference/src/metrics/custom_callbacks.py subclasses),
restated for ``ppo.RecurrentPPO``. SB3 calls ``_on_step`` after every ``VecEnv.step``; the device loop hands control back once per
rollout, so ``n_calls`` advances by ``n_steps`` at a time and frequency triggers fire when a multiple of the frequency was crossed
during the rollout (at most once per rollout).
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional

import numpy as np


class BaseCallback:
    def __init__(self, verbose: int = 0):
        self.verbose = verbose
        self.model = None
        self.n_calls = 0
        self.num_timesteps = 0
        self.locals: Dict[str, Any] = {}
        self.parent: Optional["BaseCallback"] = None

    def init_callback(self, model) -> None:
        self.model = model
        self._init_callback()

    def _init_callback(self) -> None:
        pass

    def on_training_start(self, locals_=None, globals_=None) -> None:
        self._on_training_start()

    def _on_training_start(self) -> None:
        pass

    def on_rollout_end(self, n_steps: int, log: Dict[str, Any]) -> bool:
        """Called by ``RecurrentPPO.learn`` after a rollout + update; returns False to stop training."""
        prev = self.n_calls
        self.n_calls += int(n_steps)
        self.num_timesteps = self.model.num_timesteps
        self.locals = {"log": log, "prev_n_calls": prev}
        return bool(self._on_step())

    def _crossed(self, freq: int) -> bool:
        prev = self.locals.get("prev_n_calls", 0)
        return freq > 0 and (self.n_calls // freq) > (prev // freq)

    def _on_step(self) -> bool:
        return True

    def on_training_end(self) -> None:
        self._on_training_end()

    def _on_training_end(self) -> None:
        pass


class CallbackList(BaseCallback):
    def __init__(self, callbacks: List[BaseCallback]):
        super().__init__()
        self.callbacks = list(callbacks)

    def init_callback(self, model) -> None:
        self.model = model
        for cb in self.callbacks:
            cb.init_callback(model)

    def on_training_start(self, locals_=None, globals_=None) -> None:
        for cb in self.callbacks:
            cb.on_training_start(locals_, globals_)

    def on_rollout_end(self, n_steps, log) -> bool:
        ok = True
        for cb in self.callbacks:
            ok = cb.on_rollout_end(n_steps, log) and ok
        return ok

    def on_training_end(self) -> None:
        for cb in self.callbacks:
            cb.on_training_end()


class CheckpointCallback(BaseCallback):
    """``CheckpointCallback(save_freq, save_path, name_prefix="rl_model", save_vecnormalize=False)``: writes
    ``{prefix}_{num_timesteps}_steps.zip`` (SB3 zip layout) and, if asked, ``{prefix}_vecnormalize_{num_timesteps}_steps.pkl``."""

    def __init__(self, save_freq: int, save_path: str, name_prefix: str = "rl_model", save_replay_buffer: bool = False,
                 save_vecnormalize=False, verbose: int = 0):
        super().__init__(verbose)
        self.save_freq, self.save_path, self.name_prefix = int(save_freq), save_path, name_prefix
        self.save_vecnormalize = bool(save_vecnormalize) and str(save_vecnormalize) != "False"

    def _init_callback(self) -> None:
        os.makedirs(self.save_path, exist_ok=True)

    def _on_step(self) -> bool:
        if self._crossed(self.save_freq):
            path = os.path.join(self.save_path, f"{self.name_prefix}_{self.num_timesteps}_steps.zip")
            self.model.save(path)
            if self.save_vecnormalize and hasattr(self.model.env, "save"):
                self.model.env.save(os.path.join(self.save_path, f"{self.name_prefix}_vecnormalize_{self.num_timesteps}_steps.pkl"))
            if self.verbose:
                print(f"Saving model checkpoint to {path}")
        return True


class EvalCallback(BaseCallback):
    """``EvalCallback(eval_env, n_eval_episodes, eval_freq, best_model_save_path, log_path, deterministic, callback_on_new_best)``:
    every ``eval_freq`` calls, plays ``n_eval_episodes`` per the batched evaluation loop, appends to ``evaluations.npz`` (keys
    ``timesteps``, ``results``, ``ep_lengths`` as SB3 writes them, here with the mean per evaluation) and keeps the best model."""

    def __init__(self, eval_env, callback_on_new_best: Optional[BaseCallback] = None, n_eval_episodes: int = 5, eval_freq: int = 10000,
                 log_path: Optional[str] = None, best_model_save_path: Optional[str] = None, deterministic: bool = True, render: bool = False,
                 verbose: int = 1, warn: bool = True, callback_after_eval: Optional[BaseCallback] = None):
        super().__init__(verbose)
        self.eval_env, self.n_eval_episodes, self.eval_freq, self.deterministic = eval_env, int(n_eval_episodes), int(eval_freq), deterministic
        self.log_path, self.best_model_save_path = log_path, best_model_save_path
        self.callback_on_new_best, self.callback_after_eval = callback_on_new_best, callback_after_eval
        self.best_mean_reward, self.last_mean_reward = -np.inf, -np.inf
        self.evaluations_timesteps, self.evaluations_results, self.evaluations_length = [], [], []

    def _init_callback(self) -> None:
        for d in (self.log_path, self.best_model_save_path):
            if d:
                os.makedirs(d, exist_ok=True)
        for cb in (self.callback_on_new_best, self.callback_after_eval):
            if cb is not None:
                cb.init_callback(self.model); cb.parent = self

    def _on_step(self) -> bool:
        if not self._crossed(self.eval_freq):
            return True
        from .evaluate import evaluate_policy
        from .rollout import DeviceVecNormalize

        env = self.eval_env
        norm = env if isinstance(env, DeviceVecNormalize) else (self.model.env if isinstance(self.model.env, DeviceVecNormalize) else None)
        if isinstance(env, DeviceVecNormalize) and isinstance(self.model.env, DeviceVecNormalize):      # sync_envs_normalization
            env.obs_rms.state.copy_(self.model.env.obs_rms.state); env.ret_rms.state.copy_(self.model.env.ret_rms.state)
            env.obs_rms.load(env.obs_rms.mean.cpu().numpy(), env.obs_rms.var.cpu().numpy(), float(env.obs_rms.count))
        out = evaluate_policy(self.model.policy, getattr(env, "venv", env), self.n_eval_episodes, self.deterministic, norm)
        if norm is not None and norm is not self.model.env and isinstance(self.model.env, DeviceVecNormalize):
            self.model.env._push_obs_norm()              # the policy's fused normalisation goes back to the training env's moments
        self.last_mean_reward = out["mean_reward"]
        self.evaluations_timesteps.append(self.num_timesteps)
        self.evaluations_results.append(out["mean_reward"]); self.evaluations_length.append(out["mean_length"])
        if self.log_path:
            np.savez(os.path.join(self.log_path, "evaluations"), timesteps=self.evaluations_timesteps,
                     results=np.asarray(self.evaluations_results)[:, None], ep_lengths=np.asarray(self.evaluations_length)[:, None])
        if self.verbose:
            print(f"Eval num_timesteps={self.num_timesteps}, episode_reward={out['mean_reward']:.2f} +/- {out['reward_sem']:.2f}, "
                  f"episode length {out['mean_length']:.2f}, score {out['score']:.4f}")
        self.model.logs[-1].update({"eval/mean_reward": out["mean_reward"], "eval/mean_ep_length": out["mean_length"], "eval/score": out["score"]})
        ok = True
        if out["mean_reward"] > self.best_mean_reward:
            self.best_mean_reward = out["mean_reward"]
            if self.best_model_save_path:
                self.model.save(os.path.join(self.best_model_save_path, "best_model.zip"))
            if self.callback_on_new_best is not None:
                ok = self.callback_on_new_best.on_rollout_end(0, self.locals.get("log", {}))
        if self.callback_after_eval is not None:
            ok = self.callback_after_eval.on_rollout_end(0, self.locals.get("log", {})) and ok
        return ok


class EvaluateLSTM(BaseCallback):
    """/root/reference/src/metrics/custom_callbacks.py:7-48: every ``eval_freq`` timesteps play ``num_episodes`` deterministic
    episodes on ``eval_env`` with the TRAINING model, observations normalised by the TRAINING env's moments, and record the mean
    cumulative reward under ``name``. Batched: the episodes are spread over the worlds of ``eval_env`` (a ``MyoVecEnv``)."""

    def __init__(self, eval_freq: int, eval_env, name: str, num_episodes: int = 20, verbose: int = 0):
        super().__init__(verbose)
        self.eval_freq, self.eval_env, self.name, self.num_episodes = int(eval_freq), eval_env, name, int(num_episodes)

    def _on_step(self) -> bool:
        if not self._crossed(self.eval_freq):
            return True
        from .evaluate import evaluate_policy
        from .rollout import DeviceVecNormalize

        norm = self.model.env if isinstance(self.model.env, DeviceVecNormalize) else None
        out = evaluate_policy(self.model.policy, getattr(self.eval_env, "venv", self.eval_env), self.num_episodes, True, norm)
        if norm is not None:
            norm._push_obs_norm()
        self.model.logs[-1][self.name] = out["mean_reward"]
        if self.verbose:
            print(f"{self.name}: {out['mean_reward']:.3f} over {out['episodes']} episodes")
        return True


class EnvDumpCallback(BaseCallback):
    """/root/reference/src/metrics/custom_callbacks.py:51-61: ``training_env.save(save_path/training_env.pkl)`` when triggered
    (the reference hangs it on ``EvalCallback(callback_on_new_best=...)``, /root/reference/src/main_baoding.py:84-95)."""

    def __init__(self, save_path: str, verbose: int = 0):
        super().__init__(verbose)
        self.save_path = save_path

    def _on_step(self) -> bool:
        env_path = os.path.join(self.save_path, "training_env.pkl")
        if self.verbose > 0:
            print("Saving the training environment to path ", env_path)
        os.makedirs(self.save_path, exist_ok=True)
        self.model.env.save(env_path)
        return True


class TensorboardCallback(BaseCallback):
    """/root/reference/src/metrics/custom_callbacks.py:64-81: the mean over a rollout of ``info[key]`` for every key in
    ``info_keywords``, recorded as ``rollout/<key>``. The device rollout keeps the per-step reward terms in its buffer
    (``RecurrentRolloutBuffer.infos`` [T, n, MYO_INFO_TERMS]); the means are taken there, no per-step host round trip. The scalars
    land in ``model.logs[-1]`` and, when ``log_dir`` is given, in ``<log_dir>/progress.csv`` (one row per rollout; there is no
    TensorBoard writer in this image, the CSV has the columns SB3's csv logger writes)."""

    def __init__(self, info_keywords, verbose: int = 0, log_dir: Optional[str] = None):
        super().__init__(verbose)
        self.info_keywords = tuple(info_keywords)
        self.rollout_info: Dict[str, float] = {}
        self.log_dir = log_dir
        self._columns = None

    def _on_step(self) -> bool:
        env = getattr(self.model.env, "venv", self.model.env)
        keys = list(env.info_keys)
        means = self.model.buffer.info_means()           # [MYO_INFO_TERMS] on the host
        self.rollout_info = {}
        for key in self.info_keywords:
            if key not in keys:
                raise KeyError(f"info keyword {key!r} is not one of {keys}")
            self.rollout_info[key] = float(means[keys.index(key)])
            self.model.logs[-1]["rollout/" + key] = self.rollout_info[key]
        if self.log_dir:
            os.makedirs(self.log_dir, exist_ok=True)
            row = {k: v for k, v in self.model.logs[-1].items() if isinstance(v, (int, float))}
            path = os.path.join(self.log_dir, "progress.csv")
            with open(path, "a") as fh:
                if self._columns is None:           # columns are fixed by the first row; later rows fill them by name
                    self._columns = list(row.keys())
                    fh.write(",".join(self._columns) + "\n")
                fh.write(",".join(repr(row[k]) if k in row else "" for k in self._columns) + "\n")
        return True

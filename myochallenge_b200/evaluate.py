"""Batched evaluation of a recurrent policy (the loop of /root/reference/src/main_eval.py:79-120 - deterministic
``model.predict`` on ``VecNormalize.normalize_obs(obs)``, one episode after the other, mean length and mean cumulative reward
with their standard errors - run over all worlds of a ``MyoVecEnv`` at once, on the device).

Every world plays whole episodes back to back; the first ``n_episodes`` episodes to FINISH would over-represent short episodes,
so the harness fixes the quota per world up front (``ceil(n_episodes / num_envs)`` episodes each) and counts exactly those.
Besides the reference's two numbers it reports what its callbacks / logs track: fraction of steps ``solved`` (the challenge
score), mean muscle effort (``act_mag``, the ``act_reg`` term) and the drop rate.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import _capi


def evaluate_policy(policy, env, n_episodes: int = 2000, deterministic: bool = True, obs_norm=None, max_steps: Optional[int] = None) -> Dict[str, float]:
    """``policy``: ``RecurrentPolicy``; ``env``: ``MyoVecEnv``; ``obs_norm``: ``DeviceVecNormalize`` (its moments are fused into
    the policy's input load and frozen: ``training`` is not touched, nothing is updated) or None for raw observations."""
    n = env.num_envs
    dev = env.device
    quota = max(1, math.ceil(n_episodes / n))
    if obs_norm is not None and obs_norm.norm_obs:
        policy.set_obs_norm(obs_norm.obs_rms.mean_f, obs_norm.obs_rms.var_f, obs_norm.epsilon, obs_norm.clip_obs)
    obs = env.reset_device()
    h, c = policy.initial_state(n)
    starts = torch.ones(n, dtype=torch.uint8, device=dev)
    ep_ret = torch.zeros(n, dtype=torch.float64, device=dev)
    ep_len = torch.zeros(n, dtype=torch.int64, device=dev)
    played = torch.zeros(n, dtype=torch.int64, device=dev)
    rets, lens = [], []
    solved = torch.zeros((), dtype=torch.float64, device=dev)
    effort = torch.zeros((), dtype=torch.float64, device=dev)
    drops = torch.zeros((), dtype=torch.float64, device=dev)
    steps = torch.zeros((), dtype=torch.float64, device=dev)
    baoding = env.cfg.kind == _capi.TASK_BAODING
    horizon = int(env.cfg.max_episode_steps) if env.cfg.max_episode_steps > 0 else 1000
    limit = max_steps if max_steps is not None else quota * horizon + 1
    for _ in range(limit):
        actions, _, _, _ = policy.forward(obs, (h, c), starts, deterministic=deterministic)
        obs, rew, done, trunc = env.step_device(actions.clamp(-1.0, 1.0))
        live = played < quota                                  # worlds still inside their quota
        ep_ret += torch.where(live, rew.double(), torch.zeros_like(ep_ret))
        ep_len += live.long()
        info = env.sim.info                                    # reward terms of this step [n, 8]
        lf = live.double()
        steps += lf.sum()
        solved += (info[:, 5].double() * lf).sum()             # "solved" term (both tasks keep it in slot 5)
        effort += (-(info[:, 2 if baoding else 3]).double() * lf).sum()      # act_reg = -act_mag
        fin = done.bool() & live
        if baoding:
            drops += (fin & ~trunc.bool()).double().sum()
        if bool(fin.any()):
            rets.append(ep_ret[fin].clone()); lens.append(ep_len[fin].clone())
            ep_ret.masked_fill_(fin, 0.0); ep_len.masked_fill_(fin, 0)
            played += fin.long()
        starts = done
        if bool((played >= quota).all()):
            break
    r = torch.cat(rets) if rets else torch.zeros(0, dtype=torch.float64, device=dev)
    l = torch.cat(lens).double() if lens else torch.zeros(0, dtype=torch.float64, device=dev)
    k = max(int(r.numel()), 1)
    sem = lambda x: float(x.std(unbiased=False) / math.sqrt(k)) if x.numel() > 1 else 0.0      # np.std / sqrt(n), as the reference prints
    return {"episodes": int(r.numel()), "mean_reward": float(r.mean()) if r.numel() else float("nan"), "reward_sem": sem(r),
            "mean_length": float(l.mean()) if l.numel() else float("nan"), "length_sem": sem(l),
            "score": float(solved / steps.clamp(min=1)), "effort": float(effort / steps.clamp(min=1)),
            "drop_rate": float(drops / k) if baoding else 0.0}

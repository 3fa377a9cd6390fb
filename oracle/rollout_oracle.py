"""TEST INFRASTRUCTURE - numpy restatement of the rollout-side arithmetic of the reference's training loop.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module; the product
(``myochallenge_b200/``) never does.

The arithmetic lives in third-party code that is NOT in /root/reference (pinned there by requirements.txt:126,135:
``sb3-contrib==1.6.2``, ``stable-baselines3==1.6.2``) and is reached from /root/reference/src/main_baoding.py:75
(``VecNormalize(envs)``) and /root/reference/src/train/trainer.py:67-71 (``RecurrentPPO.learn``).  Neither package is
installed or installable here, so these functions restate the published algorithms:

* ``RunningMeanStd``        stable_baselines3/common/running_mean_std.py (update, update_from_moments)
* ``vecnormalize_step``     stable_baselines3/common/vec_env/vec_normalize.py (step_wait, normalize_obs / _reward)
* ``gae``                   stable_baselines3/common/buffers.py RolloutBuffer.compute_returns_and_advantage, which
                            sb3_contrib's RecurrentRolloutBuffer inherits

PARITY UNPINNED for the update rules themselves (no SB3 binary or golden trajectory in the container).  Pinned: the
constants (clip_obs = clip_reward = 10, gamma = 0.99, epsilon = 1e-8) and the stored moments, by the reference's own
VecNormalize pickles (tests/golden/vecnormalize_baoding_step32.npz), and the closed-form properties the tests check
(moments of a concatenation, GAE against its defining double sum).
"""
import numpy as np


class RunningMeanStd:
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = float(epsilon)

    def update(self, arr):
        arr = np.asarray(arr, np.float64)
        self.update_from_moments(arr.mean(axis=0), arr.var(axis=0), arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot
        m2 = self.var * self.count + batch_var * batch_count + np.square(delta) * self.count * batch_count / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


def vecnormalize_step(ret_rms, returns, rewards, dones, gamma=0.99, epsilon=1e-8, clip_reward=10.0, training=True, norm_reward=True):
    """Reward path of VecNormalize.step_wait; mutates ret_rms / returns, returns the normalised rewards."""
    if training:
        returns[:] = returns * gamma + rewards
        ret_rms.update(returns)
    out = np.clip(rewards / np.sqrt(ret_rms.var + epsilon), -clip_reward, clip_reward) if norm_reward else rewards
    returns[np.asarray(dones, bool)] = 0
    return out.astype(np.float32)


def normalize_obs(obs_rms, obs, epsilon=1e-8, clip_obs=10.0):
    return np.clip((obs - obs_rms.mean) / np.sqrt(obs_rms.var + epsilon), -clip_obs, clip_obs).astype(np.float32)


def gae(rewards, values, episode_starts, last_values, dones, gamma=0.99, gae_lambda=0.95):
    """rewards / values / episode_starts: [T, n]; returns (advantages, returns), fp32 as SB3's buffers are."""
    rewards, values = np.asarray(rewards, np.float32), np.asarray(values, np.float32)
    T = rewards.shape[0]
    adv = np.zeros_like(rewards)
    last = np.zeros(rewards.shape[1], np.float32)
    for step in reversed(range(T)):
        if step == T - 1:
            nnt = 1.0 - np.asarray(dones, np.float32)
            nv = np.asarray(last_values, np.float32)
        else:
            nnt = 1.0 - np.asarray(episode_starts[step + 1], np.float32)
            nv = values[step + 1]
        delta = rewards[step] + gamma * nv * nnt - values[step]
        last = delta + gamma * gae_lambda * nnt * last
        adv[step] = last
    return adv, adv + values


def gae_by_definition(rewards, values, episode_starts, last_values, dones, gamma, lam):
    """O(T^2) double sum A_t = sum_k (gamma lam)^k delta_{t+k} cut at episode boundaries, fp64: the known answer."""
    r, v = np.asarray(rewards, np.float64), np.asarray(values, np.float64)
    T, n = r.shape
    vnext = np.concatenate([v[1:], np.asarray(last_values, np.float64)[None]], 0)
    nnt = 1.0 - np.concatenate([np.asarray(episode_starts, np.float64)[1:], np.asarray(dones, np.float64)[None]], 0)
    delta = r + gamma * vnext * nnt - v
    adv = np.zeros((T, n))
    for t in range(T):
        w = np.ones(n)
        for k in range(t, T):
            adv[t] += w * delta[k]
            w = w * gamma * lam * nnt[k]
    return adv

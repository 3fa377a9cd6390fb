"""Ensemble / mixture-of-ensembles evaluation and the reference's custom callbacks on the GPU
(/root/reference/src/eval_mixture_of_ensembles.py:125-345, /root/reference/src/metrics/custom_callbacks.py:7-81)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from myochallenge_b200 import ensemble
from myochallenge_b200.envs import make_vec_env
from myochallenge_b200.policy import RecurrentPolicy

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _policy(seed, n):
    p = RecurrentPolicy(86, 39, lstm_hidden=64, pi=(64,), vf=(64,), max_batch=n, device=DEV)
    p.init_random(seed=seed, log_std_init=-2.0)
    return p


def test_ensemble_action_is_the_mean_of_its_members(product_lib):
    n = 96
    rng = np.random.default_rng(0)
    norms = [dict(obs_mean=rng.normal(0, 0.1, 86), obs_var=rng.uniform(0.5, 2.0, 86), epsilon=1e-8, clip_obs=10.0) for _ in range(3)]
    ens = ensemble.Ensemble([_policy(s, n) for s in range(3)], norms)
    solo = [_policy(s, n) for s in range(3)]
    for p, nm in zip(solo, norms):
        p.set_obs_norm(nm["obs_mean"], nm["obs_var"], 1e-8, 10.0)
    states = [p.initial_state(n) for p in solo]
    starts = torch.ones(n, dtype=torch.uint8, device=DEV)
    for t in range(4):
        obs = torch.from_numpy(rng.normal(0, 1, (n, 86)).astype(np.float32)).to(DEV)
        got = ens.act(obs, starts)
        want = sum(p.forward(obs, st, starts, deterministic=True)[0] for p, st in zip(solo, states)) / 3.0
        assert torch.allclose(got, want, atol=1e-6)
        starts = torch.from_numpy((rng.random(n) < 0.2).astype(np.uint8)).to(DEV)


def test_mixture_evaluation_on_baoding(product_lib):
    """``eval_perf`` with the reference's OWN classifier weights (golden fixture) on random-init ensembles: quota of episodes
    played, the classifier consulted exactly once per episode that reaches step 13, hold worlds handed to the hold ensemble."""
    g = np.load(os.path.join(GOLDEN, "task_classifier.npz"))
    clf = ensemble.TaskClassifier({k[2:]: g[k] for k in g.files if k.startswith("w:")}, g["scaler_mean"], g["scaler_scale"], device=DEV)
    assert float((clf.logits(torch.from_numpy(g["x"]).to(DEV)).cpu() - torch.from_numpy(g["logits"])).abs().max()) < 1e-3
    n = 64
    env = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, device=DEV, seed=3, clip_actions=True, max_episode_steps=20)
    base = ensemble.Ensemble([_policy(s, n) for s in (0, 1)], [None, None])
    hold = ensemble.Ensemble([_policy(s, n) for s in (2, 3)], [None, None])
    mix = ensemble.MixtureOfEnsembles(base, hold, clf)
    out = ensemble.evaluate_mixture(mix, env, n_episodes=2 * n)
    assert out["episodes"] == 2 * n and 1 <= out["mean_length"] <= 20 and np.isfinite(out["mean_reward"]) and out["mean_effort"] > 0
    full = out["classified_episodes"]
    assert 0 < full <= 2 * n and 0.0 <= out["classifier_inaccuracy"] <= 1.0
    # a plain ensemble goes through the same loop
    out2 = ensemble.evaluate_mixture(base, env, n_episodes=n)
    assert out2["episodes"] == n and np.isnan(out2["classifier_inaccuracy"])


def test_reference_custom_callbacks(product_lib, tmp_path):
    """``TensorboardCallback(info_keywords)``, ``EvaluateLSTM`` and ``EnvDumpCallback`` as the reference's training scripts build
    them (/root/reference/src/main_baoding.py:84-125)."""
    from myochallenge_b200.callbacks import EnvDumpCallback, EvalCallback, EvaluateLSTM, TensorboardCallback
    from myochallenge_b200.ppo import RecurrentPPO
    from myochallenge_b200.rollout import DeviceVecNormalize

    n, T = 64, 8
    env = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, device=DEV, seed=2, clip_actions=True, max_episode_steps=6)
    vn = DeviceVecNormalize(env, gamma=0.99)
    eval_env = make_vec_env("CustomMyoChallengeBaodingP2-v1", 32, device=DEV, seed=9, clip_actions=True, max_episode_steps=6)
    agent = RecurrentPPO("MlpLstmPolicy", vn, n_steps=T, batch_size=T * 32, n_epochs=1, learning_rate=1e-4,
                         policy_kwargs=dict(lstm_hidden_size=64, net_arch=[dict(pi=[32], vf=[32])], log_std_init=-2.0), seed=1)
    tb = TensorboardCallback(info_keywords=("pos_dist_1", "pos_dist_2", "act_reg", "alive", "solved"), log_dir=str(tmp_path))
    cbs = [tb, EvaluateLSTM(eval_freq=16, eval_env=eval_env, name="eval/lstm", num_episodes=32),
           EvalCallback(DeviceVecNormalize(eval_env, gamma=0.99, training=False, norm_reward=False), callback_on_new_best=EnvDumpCallback(str(tmp_path)),
                        n_eval_episodes=32, eval_freq=16, verbose=0)]
    agent.learn(total_timesteps=4 * n * T, callback=cbs)
    for log in agent.logs:
        assert -1.0 <= log["rollout/pos_dist_1"] <= 0.0 and 0.0 <= log["rollout/alive"] <= 1.0
        assert log["rollout/act_reg"] <= 0.0
    assert "eval/lstm" in agent.logs[1] and "eval/lstm" not in agent.logs[0]
    assert os.path.exists(os.path.join(tmp_path, "training_env.pkl"))
    rows = open(os.path.join(tmp_path, "progress.csv")).read().strip().splitlines()
    assert len(rows) == 5 and "rollout/solved" in rows[0]
    # the info means are means of the env's own reward terms: compare one rollout by hand
    from myochallenge_b200.rollout import collect_rollouts
    o = vn.reset_device().clone()
    st = torch.ones(n, dtype=torch.uint8, device=DEV)
    state = agent.policy.initial_state(n)
    collect_rollouts(vn, agent.policy, agent.buffer, state, o, st)
    m = agent.buffer.info_means()
    assert agent.buffer.info_count == n * T and m.shape[0] == env.sim.info.shape[1] and np.isfinite(m).all()


def test_ensemble_loads_sb3_zips_and_vecnormalize_pickles(product_lib, tmp_path):
    """``Ensemble.load(PATH_TO_*_NET, PATH_TO_NORMALIZED_*_ENV)``: the files the reference lists are SB3 zips and VecNormalize pickles.
    Two are written here from the reference's phase-1 weights (golden fixture) with different normalisers; the loaded ensemble must
    act like the two policies evaluated by hand at fp32 with their own moments."""
    from myochallenge_b200 import checkpoint

    g = np.load(os.path.join(GOLDEN, "policy_phase1.npz"))
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}
    rng = np.random.default_rng(5)
    zips, pkls, norms = [], [], []
    for i in range(2):
        sdi = {k: (v + 0.01 * i * torch.randn(v.shape, generator=torch.Generator().manual_seed(i))) for k, v in sd.items()}
        zp, pp = str(tmp_path / f"rl_model_{i}.zip"), str(tmp_path / f"rl_model_vecnormalize_{i}.pkl")
        checkpoint.save_sb3_zip(zp, sdi)
        st = dict(obs_mean=rng.normal(0, 0.2, 86), obs_var=rng.uniform(0.5, 1.5, 86), obs_count=1e6, ret_mean=0.0, ret_var=1.0, ret_count=1e4,
                  clip_obs=10.0, clip_reward=10.0, gamma=0.99, epsilon=1e-8, training=False, norm_obs=True, norm_reward=False)
        checkpoint.save_vecnormalize(pp, st, num_envs=1, act_dim=39)
        zips.append(zp); pkls.append(pp); norms.append((sdi, st))
    n = 64
    ens = ensemble.Ensemble.load(zips, pkls, device=DEV, max_batch=n, precision="fp32")
    obs = torch.from_numpy(g["obs"][:n]).to(DEV)
    starts = torch.ones(n, dtype=torch.uint8, device=DEV)
    got = ens.act(obs, starts)
    want = 0
    for sdi, st in norms:
        p = RecurrentPolicy(86, 39, lstm_hidden=128, pi=(), vf=(), max_batch=n, device=DEV, precision="fp32")
        p.load_state_dict(sdi)
        p.set_obs_norm(st["obs_mean"], st["obs_var"], 1e-8, 10.0)
        want = want + p.forward(obs, p.initial_state(n), starts, deterministic=True)[0]
    assert torch.allclose(got, want / 2, atol=1e-5)

"""GPU parity of the PPO update (``myo_ppo_*`` through the C ABI) against the torch restatement
of sb3-contrib's RecurrentPPO.train in oracle/ppo_oracle.py (fp64 autograd, split-and-pad sequences).

Tolerances (floating point; the loss is a mean over T x B samples, gradients are sums of ~T x B products):
  precision="fp32" (fp32 cuBLAS GEMMs, fp32 elementwise): loss terms |err| <= 2e-5 + 1e-4 |ref|; each gradient tensor
      max |err| <= 2e-4 * max |ref| + 1e-7 (fp32 accumulation order vs fp64 autograd);
  precision="bf16" (bf16 operands, fp32 accumulation - the production mode, same rounding as the rollout kernel):
      loss terms |err| <= 2e-2 + 5e-2 |ref|; each gradient tensor: cosine similarity with the oracle >= 0.98 and
      norm within 10 % (operand rounding 2^-8 per product, no systematic bias) on the random-init winning architecture
      and on the critic of the trained checkpoint. The ACTOR of the trained checkpoint has sigma ~ e^-2..e^-3, so the
      rounding of the mean moves log_prob by z dmu / sigma ~ 0.1-0.5 against old log-probs the oracle computed WITHOUT
      that rounding: ratios cross the clip boundary and the two gradients are gradients of different clip patterns
      (bar there: cosine >= 0.5). The relevant bf16 bar is the end-to-end test at the bottom: rollout kernel and update
      forward round alike, so approx_kl ~ 0 and clip_fraction ~ 0 before the first optimiser step.
Adam + clip_grad_norm_: max |err| <= 1e-6 + 1e-5 |ref| against torch.optim.Adam / torch.nn.utils.clip_grad_norm_.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from myochallenge_b200.policy import RecurrentPolicy
from myochallenge_b200.ppo import PPOUpdate, STAT_NAMES
from myochallenge_b200.rollout import RecurrentRolloutBuffer
from oracle import ppo_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(O, A, H, pi, vf, T, B, n_envs, seed, precision, sd=None, start_prob=0.15, use_sde=False, **hyper):
    pol = RecurrentPolicy(O, A, lstm_hidden=H, pi=pi, vf=vf, max_batch=64, device=DEV, use_sde=use_sde)
    if sd is None:
        sd = pol.init_random(seed=seed, log_std_init=-0.7)
    else:
        pol.load_state_dict(sd)
    sd64 = {k: np.asarray(v.cpu() if torch.is_tensor(v) else v, np.float64) for k, v in sd.items()}
    batch = ppo_oracle.synthetic_batch(sd64, T, B, seed=seed + 1, start_prob=start_prob, dtype=np.float32)
    upd = PPOUpdate(pol, T, B, precision=precision, **hyper)
    buf = RecurrentRolloutBuffer(T, n_envs, O, A, H, DEV)
    rng = np.random.default_rng(seed + 2)
    idx = rng.permutation(n_envs)[:B].astype(np.int32)            # the minibatch worlds sit scattered in a wider buffer

    def scatter(dst, src):                                        # src [T][B][..] -> dst [T][n_envs][..]
        dst.normal_() if dst.dtype.is_floating_point else dst.zero_()
        dst[:, torch.from_numpy(idx).long().to(DEV)] = torch.from_numpy(np.ascontiguousarray(src)).to(DEV).to(dst.dtype)

    scatter(buf.observations, batch["obs"]); scatter(buf.actions, batch["actions"]); scatter(buf.episode_starts, batch["episode_starts"])
    scatter(buf.values, batch["old_values"]); scatter(buf.log_probs, batch["old_log_prob"]); scatter(buf.advantages, batch["advantages"])
    scatter(buf.returns, batch["returns"])
    buf.h0.normal_(); buf.c0.normal_()
    buf.h0[:, torch.from_numpy(idx).long().to(DEV)] = torch.from_numpy(batch["h0"]).to(DEV)
    buf.c0[:, torch.from_numpy(idx).long().to(DEV)] = torch.from_numpy(batch["c0"]).to(DEV)
    return pol, upd, buf, torch.from_numpy(idx).to(DEV), sd64, batch


CASES = [
    # O, A, H, pi, vf, T, B, n_envs
    (17, 5, 64, (32,), (48, 16), 12, 7, 19),
    (86, 39, 128, (), (), 8, 16, 16),
    (86, 39, 256, (256, 256), (256, 256), 6, 33, 40),
]


@pytest.mark.parametrize("O,A,H,pi,vf,T,B,n_envs", CASES)
def test_gradient_matches_oracle_fp32(product_lib, O, A, H, pi, vf, T, B, n_envs):
    hyper = dict(clip_range=0.2, ent_coef=0.01, vf_coef=0.7, normalize_advantage=True)
    pol, upd, buf, idx, sd64, batch = _setup(O, A, H, pi, vf, T, B, n_envs, 3, "fp32", **hyper)
    stats = upd.minibatch_grad(buf, idx).cpu().numpy()
    ref_grads, ref_stats = ppo_oracle.gradients(sd64, batch, **hyper)
    for i, k in enumerate(STAT_NAMES[:6]):
        assert abs(stats[i] - ref_stats[k]) <= 2e-5 + 1e-4 * abs(ref_stats[k]), (k, stats[i], ref_stats[k])
    assert 0.0 < ref_stats["clip_fraction"] < 1.0
    got = upd.grad_dict()
    for k, r in ref_grads.items():
        g = got[k].cpu().double()
        tol = 2e-4 * float(r.abs().max()) + 1e-7
        assert float((g - r).abs().max()) <= tol, (k, float((g - r).abs().max()), tol)


def test_gradient_matches_oracle_bf16_winning_architecture(product_lib):
    O, A, H, pi, vf, T, B, n_envs = CASES[2]
    hyper = dict(clip_range=0.2, ent_coef=0.01, vf_coef=0.7, normalize_advantage=True)
    pol, upd, buf, idx, sd64, batch = _setup(O, A, H, pi, vf, T, B, n_envs, 3, "bf16", **hyper)
    stats = upd.minibatch_grad(buf, idx).cpu().numpy()
    ref_grads, ref_stats = ppo_oracle.gradients(sd64, batch, **hyper)
    for i, k in enumerate(STAT_NAMES[:6]):
        assert abs(stats[i] - ref_stats[k]) <= 2e-2 + 5e-2 * abs(ref_stats[k]), (k, stats[i], ref_stats[k])
    got = upd.grad_dict()
    for k, r in ref_grads.items():
        gk = got[k].cpu().double().flatten(); rk = r.flatten()
        cos = float(torch.dot(gk, rk) / (gk.norm() * rk.norm() + 1e-30))
        assert cos >= 0.98 and 0.9 <= float(gk.norm() / rk.norm()) <= 1.1, (k, cos, float(gk.norm() / rk.norm()))


@pytest.mark.parametrize("H,B,T", [(64, 7, 12), (128, 150, 9), (256, 33, 6), (256, 260, 5)])
def test_persistent_recurrent_kernels_match_the_launch_chain(product_lib, monkeypatch, H, B, T):
    """bf16 mode runs the LSTM recurrence in the persistent tcgen05 cluster kernels (csrc/myo_lstm_seq.cu); MYO_PPO_SEQ=0 keeps round
    1's chain of cuBLAS GEMM + cell launches. Same operands (bf16-rounded W_hh and h, fp32 accumulation), so both must give the
    same loss and gradients up to accumulation order, the fast exp of the fused cell and the bf16 roundings those flip: loss terms
    to 2e-3 relative, every gradient tensor to 1 % in the L2 norm and 3 % of its max entrywise (the oracle bars - cosine >= 0.98 -
    hold for both and are far looser)."""
    O, A, pi, vf = 30, 6, (64,), (64,)
    hyper = dict(clip_range=0.2, ent_coef=0.01, vf_coef=0.7, normalize_advantage=True)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("MYO_PPO_SEQ", mode)
        pol, upd, buf, idx, sd64, batch = _setup(O, A, H, pi, vf, T, B, B + 5, 11, "bf16", **hyper)
        stats = upd.minibatch_grad(buf, idx).cpu().numpy().copy()
        grads = {k: v.cpu().double().clone() for k, v in upd.grad_dict().items()}
        out[mode] = (stats, grads, upd.launch_count)
    (s0, g0, l0), (s1, g1, l1) = out["0"], out["1"]
    assert l1 < l0 - T, (l0, l1)                       # the recurrent chain no longer launches per step
    for i in range(6):
        assert abs(s0[i] - s1[i]) <= 2e-3 * abs(s0[i]) + 2e-4, (STAT_NAMES[i], s0[i], s1[i])
    for k in g0:
        err = float((g0[k] - g1[k]).abs().max()); ref = float(g0[k].abs().max())
        l2 = float((g0[k] - g1[k]).norm() / (g0[k].norm() + 1e-30))
        assert err <= 3e-2 * ref + 1e-7 and l2 <= 1e-2, (k, err, ref, l2)


def test_value_clipping_and_raw_advantages_fp32(product_lib):
    hyper = dict(clip_range=0.1, clip_range_vf=0.05, ent_coef=0.0, vf_coef=1.0, normalize_advantage=False)
    pol, upd, buf, idx, sd64, batch = _setup(17, 5, 64, (), (32,), 10, 9, 12, 5, "fp32", **hyper)
    stats = upd.minibatch_grad(buf, idx).cpu().numpy()
    ref_grads, ref_stats = ppo_oracle.gradients(sd64, batch, **hyper)
    for i, k in enumerate(STAT_NAMES[:6]):
        assert abs(stats[i] - ref_stats[k]) <= 2e-5 + 1e-4 * abs(ref_stats[k]), (k, stats[i], ref_stats[k])
    got = upd.grad_dict()
    for k, r in ref_grads.items():
        g = got[k].cpu().double()
        assert float((g - r).abs().max()) <= 2e-4 * float(r.abs().max()) + 1e-7, k


def test_phase1_checkpoint_gradient_fp32_and_bf16(product_lib):
    """Real weights: the reference's shipped phase-1 policy (LSTM-128, no MLP layers)."""
    g = np.load(os.path.join(GOLDEN, "policy_phase1.npz"))
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}
    hyper = dict(clip_range=0.2, ent_coef=0.001, vf_coef=0.5)
    ref = None
    for precision in ("fp32", "bf16"):
        pol, upd, buf, idx, sd64, batch = _setup(86, 39, 128, (), (), 16, 24, 32, 7, precision, sd=sd, **hyper)
        stats = upd.minibatch_grad(buf, idx).cpu().numpy()
        if ref is None:
            ref = ppo_oracle.gradients(sd64, batch, **hyper)
        ref_grads, ref_stats = ref
        got = upd.grad_dict()
        for k, r in ref_grads.items():
            gk = got[k].cpu().double().flatten(); rk = r.flatten()
            if precision == "fp32":
                assert float((gk - rk).abs().max()) <= 2e-4 * float(rk.abs().max()) + 1e-7, k
            else:
                cos = float(torch.dot(gk, rk) / (gk.norm() * rk.norm() + 1e-30))
                if "critic" in k or "value_net" in k:
                    assert cos >= 0.98 and 0.9 <= float(gk.norm() / rk.norm()) <= 1.1, (k, cos, float(gk.norm() / rk.norm()))
                else:                   # actor side on the trained checkpoint: see the module docstring
                    assert cos >= 0.5, (k, cos)
        if precision == "bf16":     # critic-side and parameter-only terms (the policy terms depend on the clip pattern, see above)
            for i, k in ((1, "value_loss"), (2, "entropy_loss")):
                assert abs(stats[i] - ref_stats[k]) <= 2e-2 + 5e-2 * abs(ref_stats[k]), (k, stats[i], ref_stats[k])


def test_adam_and_clip_match_torch(product_lib):
    pol, upd, buf, idx, sd64, batch = _setup(17, 5, 64, (32,), (), 6, 5, 5, 11, "fp32", learning_rate=3e-3, max_grad_norm=0.5)
    p = upd.params.cpu().clone(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step in (1, 2, 3):
        upd.minibatch_grad(buf, idx)
        g = upd.grad.cpu().clone()
        upd.adam_step()
        p, m, v, norm = ppo_oracle.adam_step(p, g, m, v, step, 3e-3, (0.9, 0.999), 1e-5, 0.5)
        assert abs(float(upd.grad_norm) - norm) <= 1e-5 * norm
        assert float((upd.params.cpu() - p).abs().max()) <= 1e-6 + 1e-5 * float(p.abs().max())
        assert torch.allclose(upd.exp_avg.cpu(), m, rtol=1e-4, atol=1e-9) and torch.allclose(upd.exp_avg_sq.cpu(), v, rtol=1e-4, atol=1e-12)
    # the scale argument averages a summed bucket: a doubled gradient with scale 1/2 takes the same step
    upd2 = PPOUpdate(pol, 6, 5, precision="fp32", learning_rate=3e-3, max_grad_norm=0.5)
    upd2.load_state_dict(upd.state_dict()); upd2.exp_avg.copy_(upd.exp_avg); upd2.exp_avg_sq.copy_(upd.exp_avg_sq); upd2.step_count = upd.step_count
    upd.minibatch_grad(buf, idx); upd2.grad.copy_(upd.grad * 2)
    upd.adam_step(); upd2.adam_step(grad_scale=0.5)
    assert torch.allclose(upd.params, upd2.params, rtol=1e-6, atol=1e-8)


def test_train_epochs_lower_the_loss_and_push_weights(product_lib):
    """A few epochs of the whole update loop on one synthetic rollout: the surrogate + value loss goes down, the policy
    handle receives the new weights (its forward changes), gradient is deterministic run to run."""
    hyper = dict(learning_rate=1e-3, clip_range=0.2, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5)
    pol, upd, buf, idx, sd64, batch = _setup(17, 5, 64, (32,), (32,), 10, 8, 32, 13, "bf16", **hyper)
    all_idx = torch.arange(32, dtype=torch.int32, device=DEV)
    s0 = upd.minibatch_grad(buf, all_idx[:8]).clone(); g0 = upd.grad.clone()
    s1 = upd.minibatch_grad(buf, all_idx[:8]).clone()
    assert torch.equal(g0, upd.grad) and torch.equal(s0, s1)           # bitwise reproducible (no atomics)
    before = {k: v.clone() for k, v in upd.state_dict().items()}
    first = float(upd.minibatch_grad(buf, idx)[1])
    log = upd.train(buf, n_epochs=4, generator=torch.Generator().manual_seed(0))
    assert log["train/n_updates"] == 4 * 4
    last = float(upd.minibatch_grad(buf, idx)[1])
    assert last < first, (first, last)                                   # value loss on the same minibatch went down
    assert any(not torch.equal(before[k], v) for k, v in upd.state_dict().items())
    assert torch.equal(pol.state_dict()["action_net.weight"], upd.state_dict()["action_net.weight"])


def test_recurrent_ppo_learn_on_baoding_is_on_policy_at_the_first_minibatch(product_lib):
    """The whole loop (reference: RecurrentPPO(...).learn, /root/reference/src/train/trainer.py:49-71) on Baoding worlds.
    Before any optimiser step, re-evaluating the rollout's own actions must reproduce the rollout's log-probs and values:
    the update's forward (cuBLAS, bf16 operands, sequences from h0 with in-sequence resets) and the rollout kernel
    (tcgen05, one step at a time, VecNormalize fused) are two implementations of one function. Bars: approx_kl <= 2e-3,
    clip_fraction <= 0.02, value loss vs stored returns consistent with GAE (returns - values = advantages)."""
    from myochallenge_b200.envs import make_vec_env
    from myochallenge_b200.ppo import RecurrentPPO
    from myochallenge_b200.rollout import DeviceVecNormalize, collect_rollouts

    n, T = 256, 16
    env = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, device=DEV, seed=5, clip_actions=True, max_episode_steps=7)
    vn = DeviceVecNormalize(env, gamma=0.99)
    agent = RecurrentPPO("MlpLstmPolicy", vn, n_steps=T, batch_size=T * 64, n_epochs=2, learning_rate=1e-4, ent_coef=0.001,
                         policy_kwargs=dict(lstm_hidden_size=64, net_arch=[dict(pi=[64], vf=[64])], log_std_init=-2.0, enable_critic_lstm=True),
                         seed=1)
    assert agent.batch_worlds == 64
    # one rollout, then the loss at the unchanged parameters
    agent._obs = vn.reset_device().clone()
    agent._starts = torch.ones(n, dtype=torch.uint8, device=DEV)
    agent._state = agent.policy.initial_state(n)
    agent._obs, agent._starts = collect_rollouts(vn, agent.policy, agent.buffer, agent._state, agent._obs, agent._starts)
    assert int(agent.buffer.episode_starts[1:].sum()) > n                  # resets inside the sequences (horizon 7)
    stats = agent.update.minibatch_grad(agent.buffer, torch.arange(64, dtype=torch.int32, device=DEV)).cpu().numpy()
    assert stats[3] <= 2e-3 and stats[4] <= 0.02, stats
    adv = agent.buffer.advantages[:, :64]
    assert abs(stats[1] - float((adv * adv).mean())) <= 5e-2 * float((adv * adv).mean()) + 1e-3, stats
    # and the loop itself
    agent.learn(total_timesteps=2 * n * T)
    assert len(agent.logs) == 2 and agent.num_timesteps == 2 * n * T
    for log in agent.logs:
        assert log["train/n_updates"] == 2 * 4 and np.isfinite(log["train/loss"]) and log["train/approx_kl"] < 0.05


def test_recurrent_ppo_save_load_round_trip(product_lib, tmp_path):
    """RecurrentPPO.save -> RecurrentPPO.load (SB3 zip layout) and DeviceVecNormalize.save -> load: parameters, Adam moments, step
    count, hyper-parameters and running moments survive; the reloaded agent takes the same next optimiser step."""
    from myochallenge_b200.envs import make_vec_env
    from myochallenge_b200.ppo import RecurrentPPO
    from myochallenge_b200.rollout import DeviceVecNormalize

    n, T = 128, 8
    env = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, device=DEV, seed=2, clip_actions=True, max_episode_steps=6)
    vn = DeviceVecNormalize(env, gamma=0.99)
    kw = dict(n_steps=T, batch_size=T * 32, n_epochs=1, learning_rate=3e-4, clip_range=0.25, ent_coef=0.002, vf_coef=0.8, max_grad_norm=0.7, gae_lambda=0.9,
              policy_kwargs=dict(lstm_hidden_size=64, net_arch=[dict(pi=[32], vf=[48, 16])], log_std_init=-1.0))
    agent = RecurrentPPO("MlpLstmPolicy", vn, seed=4, **kw)
    agent.learn(total_timesteps=n * T)
    zp, ep = str(tmp_path / "agent.zip"), str(tmp_path / "env.pkl")
    agent.save(zp); vn.save(ep)
    vn2 = DeviceVecNormalize.load(ep, env)
    assert torch.equal(vn2.obs_rms.state, vn.obs_rms.state) and torch.equal(vn2.ret_rms.state, vn.ret_rms.state)
    agent2 = RecurrentPPO.load(zp, env=vn2)
    assert agent2.n_steps == T and agent2.batch_worlds == 32 and agent2.gae_lambda == 0.9 and agent2.num_timesteps == n * T
    assert agent2.policy.pi == (32,) and agent2.policy.vf == (48, 16) and agent2.policy.lstm_hidden == 64
    h1, h2 = agent.update.hyper, agent2.update.hyper
    assert (h1.clip_range, h1.ent_coef, h1.vf_coef) == (h2.clip_range, h2.ent_coef, h2.vf_coef) and agent2.update.max_grad_norm == pytest.approx(0.7)
    assert agent2.update.step_count == agent.update.step_count == 4
    for a, b in ((agent.update.params, agent2.update.params), (agent.update.exp_avg, agent2.update.exp_avg), (agent.update.exp_avg_sq, agent2.update.exp_avg_sq)):
        assert torch.equal(a, b)
    # same next step from the same rollout data
    idx = torch.arange(32, dtype=torch.int32, device=DEV)
    agent.update.minibatch_grad(agent.buffer, idx); agent.update.adam_step()
    agent2.update.minibatch_grad(agent.buffer, idx); agent2.update.adam_step()
    assert torch.equal(agent.update.params, agent2.update.params)


def test_evaluate_policy_counts_fixed_quota_per_world(product_lib):
    """The reference's evaluation loop (src/main_eval.py:79-120) batched: every world contributes exactly its quota of episodes;
    lengths respect the horizon; a horizon-limited, never-dropping setting gives length == horizon and drop_rate 0."""
    from myochallenge_b200.envs import make_vec_env
    from myochallenge_b200.evaluate import evaluate_policy

    n = 96
    env = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, device=DEV, seed=8, clip_actions=True, max_episode_steps=9, drop_th=-10.0)
    pol = RecurrentPolicy(env.sim.nobs, env.sim.nu, lstm_hidden=64, pi=(32,), vf=(32,), max_batch=n, device=DEV)
    pol.init_random(seed=0, log_std_init=-2.0)
    out = evaluate_policy(pol, env, n_episodes=250, deterministic=True)
    assert out["episodes"] == 3 * n                       # ceil(250 / 96) = 3 per world
    assert out["mean_length"] == 9.0 and out["length_sem"] == 0.0 and out["drop_rate"] == 0.0
    assert np.isfinite(out["mean_reward"]) and 0.0 <= out["score"] <= 1.0 and out["effort"] > 0
    env2 = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, device=DEV, seed=8, clip_actions=True, max_episode_steps=200, drop_th=1.40)
    out2 = evaluate_policy(pol, env2, n_episodes=n, deterministic=True)      # balls start at z ~ 1.44: a high drop threshold ends episodes early
    assert out2["episodes"] == n and out2["mean_length"] < 200 and out2["drop_rate"] > 0.05


# ---- generalised state-dependent exploration (use_sde=True, the winning runs' setting: /root/reference/docs/summary.md:86-117) ----
def test_sde_sampling_matches_the_distribution(product_lib):
    """myo_sde_reset_noise / myo_sde_sample against StateDependentNoiseDistribution: action = mean + latent . E[w] with the stored
    exploration matrices, log_prob = log N(action; mean, sqrt(latent^2 . exp(log_std)^2 + 1e-6)); E ~ exp(log_std) * N(0, 1)."""
    n, H, A = 512, 64, 39
    pol = RecurrentPolicy(86, A, lstm_hidden=H, pi=(32,), vf=(), max_batch=n, device=DEV, use_sde=True)
    sd = pol.init_random(seed=2, log_std_init=-1.5)
    assert tuple(sd["log_std"].shape) == (32, A)
    g = torch.Generator().manual_seed(0)
    obs = torch.randn(n, 86, generator=g).to(DEV)
    starts = torch.ones(n, dtype=torch.uint8, device=DEV)
    h, c = pol.initial_state(n)
    mean, v0, _, _ = pol.forward(obs, (h, c), starts, deterministic=True)
    mean = mean.clone()
    pol.seed(11); pol.reset_noise(n)
    h, c = pol.initial_state(n)
    act, v1, logp, _ = pol.forward(obs, (h, c), starts)
    lat = pol._latent[:n].double(); E = pol._noise_mat[:n].double(); ls = pol.state_dict()["log_std"].double()
    noise = torch.bmm(lat.unsqueeze(1), E).squeeze(1)
    assert torch.allclose((act - mean).double(), noise, rtol=1e-4, atol=1e-5) and torch.equal(v0, v1)
    var = (lat ** 2) @ (ls.exp() ** 2) + 1e-6
    ref_lp = torch.distributions.Normal(mean.double(), var.sqrt()).log_prob(act.double()).sum(1)
    assert torch.allclose(logp.double(), ref_lp, rtol=1e-4, atol=1e-3)
    z = (E / ls.exp()).flatten()                           # the draws behind the matrices: standard normal (bf16 storage: 2^-9 relative)
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3
    pol.reset_noise(n)
    assert not torch.equal(E, pol._noise_mat[:n].double())   # a new epoch draws new matrices


@pytest.mark.parametrize("pi,vf,precision", [((32,), (48, 16), "fp32"), ((), (), "fp32"), ((256, 256), (256, 256), "bf16")])
def test_sde_gradient_matches_oracle(product_lib, pi, vf, precision):
    O, A, H, T, B, n_envs = (86, 39, 256, 6, 33, 40) if precision == "bf16" else (17, 5, 64, 10, 9, 12)
    hyper = dict(clip_range=0.2, ent_coef=0.01, vf_coef=0.7, normalize_advantage=True)
    pol = RecurrentPolicy(O, A, lstm_hidden=H, pi=pi, vf=vf, max_batch=64, device=DEV, use_sde=True)
    sd = pol.init_random(seed=5, log_std_init=-1.0)
    sd["log_std"] = sd["log_std"] + 0.3 * torch.randn(sd["log_std"].shape, generator=torch.Generator().manual_seed(1))
    pol, upd, buf, idx, sd64, batch = _setup(O, A, H, pi, vf, T, B, n_envs, 9, precision, sd=sd, use_sde=True, **hyper)
    stats = upd.minibatch_grad(buf, idx).cpu().numpy()
    ref_grads, ref_stats = ppo_oracle.gradients(sd64, batch, **hyper)
    got = upd.grad_dict()
    assert tuple(got["log_std"].shape) == (pi[-1] if pi else H, A)
    for i, k in enumerate(STAT_NAMES[:6]):
        tol = (2e-5 + 1e-4 * abs(ref_stats[k])) if precision == "fp32" else (2e-2 + 5e-2 * abs(ref_stats[k]))
        assert abs(stats[i] - ref_stats[k]) <= tol, (k, stats[i], ref_stats[k])
    for k, r in ref_grads.items():
        gk = got[k].cpu().double().flatten(); rk = r.flatten()
        if precision == "fp32":
            assert float((gk - rk).abs().max()) <= 2e-4 * float(rk.abs().max()) + 1e-7, (k, float((gk - rk).abs().max()), float(rk.abs().max()))
        else:
            cos = float(torch.dot(gk, rk) / (gk.norm() * rk.norm() + 1e-30))
            assert cos >= 0.98 and 0.9 <= float(gk.norm() / rk.norm()) <= 1.1, (k, cos, float(gk.norm() / rk.norm()))


def test_sde_recurrent_ppo_is_on_policy_and_learns(product_lib):
    """RecurrentPPO(use_sde=True) end to end on Baoding: before the first optimiser step the update reproduces the rollout's
    log-probs (approx_kl <= 2e-3, clip_fraction <= 0.02), then two iterations run."""
    from myochallenge_b200.envs import make_vec_env
    from myochallenge_b200.ppo import RecurrentPPO
    from myochallenge_b200.rollout import DeviceVecNormalize, collect_rollouts

    n, T = 256, 16
    env = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, device=DEV, seed=5, clip_actions=True, max_episode_steps=7)
    vn = DeviceVecNormalize(env, gamma=0.99)
    agent = RecurrentPPO("MlpLstmPolicy", vn, n_steps=T, batch_size=T * 64, n_epochs=2, learning_rate=1e-4, ent_coef=0.001, use_sde=True,
                         policy_kwargs=dict(lstm_hidden_size=64, net_arch=[dict(pi=[64], vf=[64])], log_std_init=-2.0), seed=1)
    agent._obs = vn.reset_device().clone()
    agent._starts = torch.ones(n, dtype=torch.uint8, device=DEV)
    agent._state = agent.policy.initial_state(n)
    agent._obs, agent._starts = collect_rollouts(vn, agent.policy, agent.buffer, agent._state, agent._obs, agent._starts)
    acts = agent.buffer.actions
    assert float(acts.std()) > 1e-3 and torch.isfinite(agent.buffer.log_probs).all()
    stats = agent.update.minibatch_grad(agent.buffer, torch.arange(64, dtype=torch.int32, device=DEV)).cpu().numpy()
    assert stats[3] <= 2e-3 and stats[4] <= 0.02, stats
    before = agent.update.state_dict()["log_std"].clone()
    agent.learn(total_timesteps=2 * n * T)
    assert len(agent.logs) == 2 and all(np.isfinite(l["train/loss"]) for l in agent.logs)
    assert not torch.equal(before, agent.update.state_dict()["log_std"]) and tuple(before.shape) == (64, 39)


@pytest.mark.parametrize("env_id,n,H", [("CustomMyoElbowPoseRandom-v0", 64, 64), ("CustomMyoFingerPoseRandom-v0", 512, 64), ("CustomMyoHandPoseRandom-v0", 256, 128)])
def test_recurrent_ppo_on_the_pose_configs(product_lib, env_id, n, H):
    """BASELINE configs[0..2] (elbow, finger, hand pose) through the same loop: rollout + update run, the first minibatch is on
    policy, episodes end by the horizon (100)."""
    from myochallenge_b200.envs import make_vec_env
    from myochallenge_b200.ppo import RecurrentPPO
    from myochallenge_b200.rollout import DeviceVecNormalize, collect_rollouts

    T = 12
    env = make_vec_env(env_id, n, device=DEV, seed=1, clip_actions=True, max_episode_steps=10)
    vn = DeviceVecNormalize(env, gamma=0.99)
    agent = RecurrentPPO("MlpLstmPolicy", vn, n_steps=T, batch_size=T * (n // 2), n_epochs=1, learning_rate=1e-4,
                         policy_kwargs=dict(lstm_hidden_size=H, net_arch=[dict(pi=[64], vf=[64])], log_std_init=-1.0), seed=3)
    agent._obs = vn.reset_device().clone()
    agent._starts = torch.ones(n, dtype=torch.uint8, device=DEV)
    agent._state = agent.policy.initial_state(n)
    agent._obs, agent._starts = collect_rollouts(vn, agent.policy, agent.buffer, agent._state, agent._obs, agent._starts)
    assert float(agent.buffer.episode_starts[10].float().mean()) > 0.9      # horizon 10: (nearly) every world restarts at step 10 (pose envs may end earlier: far_th)
    stats = agent.update.minibatch_grad(agent.buffer, torch.arange(n // 2, dtype=torch.int32, device=DEV)).cpu().numpy()
    assert stats[3] <= 2e-3 and stats[4] <= 0.02, stats
    agent.learn(total_timesteps=n * T)
    assert np.isfinite(agent.logs[-1]["train/loss"]) and agent.logs[-1]["train/n_updates"] == 2


def test_sb3_style_callbacks(product_lib, tmp_path):
    """``agent.learn(total_timesteps, callback=[EvalCallback(...), CheckpointCallback(...)], reset_num_timesteps=True)`` as
    /root/reference/src/main_baoding.py:84-125 calls it: checkpoints (zip + VecNormalize pickle) appear when a multiple of save_freq is
    crossed, evaluations.npz grows every eval_freq calls, the best model is kept, and training continues from a fresh reset afterwards."""
    from myochallenge_b200.callbacks import BaseCallback, CheckpointCallback, EvalCallback
    from myochallenge_b200.envs import make_vec_env
    from myochallenge_b200.ppo import RecurrentPPO
    from myochallenge_b200.rollout import DeviceVecNormalize

    n, T = 64, 8
    env = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, device=DEV, seed=2, clip_actions=True, max_episode_steps=6)
    vn = DeviceVecNormalize(env, gamma=0.99)
    eval_env = DeviceVecNormalize(make_vec_env("CustomMyoChallengeBaodingP2-v1", 32, device=DEV, seed=9, clip_actions=True, max_episode_steps=6), gamma=0.99,
                                  training=False, norm_reward=False)
    agent = RecurrentPPO("MlpLstmPolicy", vn, n_steps=T, batch_size=T * 32, n_epochs=1, learning_rate=1e-4,
                         policy_kwargs=dict(lstm_hidden_size=64, net_arch=[dict(pi=[32], vf=[32])], log_std_init=-2.0), seed=1)

    class Count(BaseCallback):
        def _on_step(self):
            self.seen = getattr(self, "seen", 0) + 1
            return True

    new_best = Count()
    cbs = [EvalCallback(eval_env, callback_on_new_best=new_best, n_eval_episodes=40, eval_freq=16, log_path=str(tmp_path), best_model_save_path=str(tmp_path),
                        deterministic=True, verbose=0),
           CheckpointCallback(save_freq=24, save_path=str(tmp_path), save_vecnormalize="True")]
    agent.learn(total_timesteps=6 * n * T, callback=cbs, reset_num_timesteps=True)          # 6 rollouts = 48 calls
    assert len(agent.logs) == 6 and agent.num_timesteps == 6 * n * T
    zips = sorted(f for f in os.listdir(tmp_path) if f.startswith("rl_model_") and f.endswith(".zip"))
    pkls = sorted(f for f in os.listdir(tmp_path) if f.startswith("rl_model_vecnormalize_"))
    assert len(zips) == 2 and len(pkls) == 2 and zips[0] == f"rl_model_{3 * n * T}_steps.zip"      # calls 24 and 48
    ev = np.load(os.path.join(tmp_path, "evaluations.npz"))
    assert list(ev["timesteps"]) == [2 * n * T, 4 * n * T, 6 * n * T] and ev["results"].shape == (3, 1) and ev["ep_lengths"].max() <= 6
    assert os.path.exists(os.path.join(tmp_path, "best_model.zip")) and new_best.seen >= 1
    assert "eval/mean_reward" in agent.logs[1] and "eval/mean_reward" not in agent.logs[0]
    agent2 = RecurrentPPO.load(os.path.join(tmp_path, zips[-1]), env=vn)
    assert torch.equal(agent2.update.params, agent.update.params)

"""GPU: the rollout-side kernels (csrc/myo_rollout.cu) through the C ABI against the numpy oracle, and the device
rollout loop (policy forward -> env step -> buffer -> GAE) end to end on Baoding worlds."""
import numpy as np
import pytest
import torch

from conftest import HAND_BAODING
from myochallenge_b200 import _capi
from oracle import rollout_oracle as ro

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_gae_kernel_matches_oracle(product_lib):
    from myochallenge_b200.rollout import RecurrentRolloutBuffer

    rng = np.random.default_rng(0)
    for T, n in ((1, 5), (17, 1000), (128, 4096)):
        buf = RecurrentRolloutBuffer(T, n, 3, 2, 4, DEV, gamma=0.99, gae_lambda=0.9)
        r, v = rng.normal(0, 1, (T, n)).astype(np.float32), rng.normal(0, 3, (T, n)).astype(np.float32)
        starts = (rng.random((T, n)) < 0.05).astype(np.uint8)
        lv, dn = rng.normal(0, 3, n).astype(np.float32), (rng.random(n) < 0.2).astype(np.uint8)
        buf.rewards.copy_(torch.from_numpy(r)); buf.values.copy_(torch.from_numpy(v)); buf.episode_starts.copy_(torch.from_numpy(starts))
        buf.full = True
        buf.compute_returns_and_advantage(torch.from_numpy(lv).to(DEV), torch.from_numpy(dn).to(DEV))
        adv, ret = ro.gae(r, v, starts, lv, dn, 0.99, 0.9)
        # fp32 both sides; the kernel contracts a * b + c into FMAs, numpy does not: 1e-6 of the value scale
        np.testing.assert_allclose(buf.advantages.cpu().numpy(), adv, rtol=1e-5, atol=2e-5)
        np.testing.assert_allclose(buf.returns.cpu().numpy(), ret, rtol=1e-5, atol=2e-5)


def test_running_moments_kernel_matches_oracle(product_lib):
    from myochallenge_b200.rollout import DeviceRunningMeanStd

    rng = np.random.default_rng(1)
    for d, sizes in ((86, (32768, 100, 1, 4097)), (1, (32768, 7)), (300, (513, 64))):
        dev, ref = DeviceRunningMeanStd(d, DEV), ro.RunningMeanStd(shape=(d,))
        for n in sizes:
            x = (rng.normal(0.3, 2.0, (n, d)) * rng.uniform(0.01, 10, d) + 1.4).astype(np.float32)
            dev.update(torch.from_numpy(x).to(DEV))
            ref.update(x)
            np.testing.assert_allclose(dev.mean.cpu().numpy(), ref.mean, rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(dev.var.cpu().numpy(), ref.var, rtol=1e-9)
            assert float(dev.count) == pytest.approx(ref.count)
            np.testing.assert_allclose(dev.mean_f.cpu().numpy(), ref.mean.astype(np.float32), rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(dev.var_f.cpu().numpy(), ref.var.astype(np.float32), rtol=1e-6)
    # bit-identical when repeated (fixed merge order, no atomics)
    a, b = DeviceRunningMeanStd(86, DEV), DeviceRunningMeanStd(86, DEV)
    x = torch.randn(32768, 86, device=DEV)
    a.update(x); b.update(x)
    assert torch.equal(a.state, b.state)


def test_vecnorm_reward_kernel_matches_oracle(product_lib):
    from myochallenge_b200.rollout import DeviceRunningMeanStd, _p, _stream_ptr

    L = product_lib
    n = 5000
    rng = np.random.default_rng(2)
    rms_d, rms = DeviceRunningMeanStd(1, DEV), ro.RunningMeanStd(shape=())
    ret_d, ret = torch.zeros(n, dtype=torch.float64, device=DEV), np.zeros(n)
    out = torch.empty(n, device=DEV)
    for t in range(12):
        rew = rng.normal(2, 30, n).astype(np.float32)
        done = (rng.random(n) < 0.1).astype(np.uint8)
        s = rms_d._scratch_for(n)
        rew_d, done_d = torch.from_numpy(rew).to(DEV), torch.from_numpy(done).to(DEV)      # keep alive across the async launches
        _capi.check(L, L.myo_vecnorm_reward(_p(rms_d.state), _p(ret_d), _p(rew_d), _p(done_d), _p(out),
                                            n, 0.99, 1e-8, 10.0, 1, 1, _p(s), _stream_ptr(torch.device(DEV))))
        want = ro.vecnormalize_step(rms, ret, rew.astype(np.float64), done, 0.99, 1e-8, 10.0)
        np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(ret_d.cpu().numpy(), ret, rtol=1e-10)     # the kernel fuses ret * gamma + r into one FMA
        np.testing.assert_allclose(rms_d.var.cpu().numpy(), [rms.var], rtol=1e-10)


def test_collect_rollouts_on_baoding(product_lib):
    """RecurrentPPO.collect_rollouts on device: buffer contents are consistent with what the env and the policy
    returned, truncated worlds were bootstrapped, GAE ran."""
    from myochallenge_b200.envs import make_vec_env
    from myochallenge_b200.policy import RecurrentPolicy
    from myochallenge_b200.rollout import DeviceVecNormalize, RecurrentRolloutBuffer, collect_rollouts

    n, T = 256, 12
    env = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, device=DEV, seed=3, clip_actions=True, max_episode_steps=5)
    pol = RecurrentPolicy(env.sim.nobs, env.sim.nu, lstm_hidden=64, pi=(64,), vf=(64,), max_batch=n, device=DEV)
    pol.init_random(seed=0, log_std_init=-2.0)
    pol.seed(9)
    vn = DeviceVecNormalize(env, pol, gamma=0.99)
    buf = RecurrentRolloutBuffer(T, n, env.sim.nobs, env.sim.nu, 64, DEV, gamma=0.99, gae_lambda=0.95)
    h, c = pol.initial_state(n)
    obs = vn.reset_device().clone()
    starts = torch.ones(n, dtype=torch.uint8, device=DEV)
    obs, starts = collect_rollouts(vn, pol, buf, (h, c), obs, starts)
    torch.cuda.synchronize()
    assert buf.full and torch.isfinite(buf.advantages).all() and torch.isfinite(buf.returns).all()
    assert torch.allclose(buf.returns, buf.advantages + buf.values, atol=1e-5)
    es = buf.episode_starts.cpu().numpy()
    assert es[0].all()                                     # first step after reset
    assert es[5].mean() > 0.9 and es[10].mean() > 0.9      # horizon 5: TimeLimit ends (nearly) every world every 5 steps
    assert es[1:5].mean() < 0.2
    # the observation moments saw reset + T steps of n worlds
    assert float(vn.obs_rms.count) == pytest.approx((T + 1) * n + 1e-4)
    assert float(vn.ret_rms.count) == pytest.approx(T * n + 1e-4)
    # GAE of the stored rollout reproduces on the host
    lv = pol.predict_values(obs, (h, c), starts).cpu().numpy()
    adv, _ = ro.gae(buf.rewards.cpu().numpy(), buf.values.cpu().numpy(), es, lv, starts.cpu().numpy(), 0.99, 0.95)
    np.testing.assert_allclose(buf.advantages.cpu().numpy(), adv, rtol=1e-4, atol=1e-4)


def test_pipelined_stepper_equals_plain_stepping(product_lib):
    """``rollout.PipelinedStepper`` (sub-batches on their own streams, one shared policy handle) changes the schedule, not the
    results: every sub-batch ends bit-identical to the same env stepped alone on the default stream."""
    from myochallenge_b200.envs import make_vec_env
    from myochallenge_b200.policy import RecurrentPolicy
    from myochallenge_b200.rollout import PipelinedStepper

    dev, n, T = "cuda:0", 96, 12
    mk = lambda seed: make_vec_env("CustomMyoChallengeBaodingP2-v1", n, device=dev, seed=seed, clip_actions=True, max_episode_steps=7)
    pol = RecurrentPolicy(86, 39, lstm_hidden=64, pi=(64,), vf=(64,), max_batch=n, device=dev)
    pol.init_random(seed=3, log_std_init=-1.0)
    st = PipelinedStepper([mk(5), mk(6)], pol, deterministic=True)
    st.reset()
    for _ in range(T):
        st.step()
    st.join()
    torch.cuda.synchronize()
    for i, seed in enumerate((5, 6)):
        env = mk(seed)
        obs = env.reset_device()
        state = pol.initial_state(n)
        starts = torch.ones(n, dtype=torch.uint8, device=dev)
        for _ in range(T):
            a, _, _, _ = pol.forward(obs, state, starts, deterministic=True)
            obs, rew, done, _ = env.step_device(a.clamp(-1.0, 1.0))
            starts = done
        torch.cuda.synchronize()
        assert torch.equal(obs, st.obs[i]) and torch.equal(rew, st.rewards[i]) and torch.equal(done, st.dones[i])
        assert torch.equal(state[0], st.states[i][0]) and torch.equal(state[1], st.states[i][1])

"""CPU (host emulation of the kernel sources): the curriculum knobs of CustomBaodingP2Env.reset
(/root/reference/src/envs/baoding.py:494-647) - reference-state initialisation (RSI) and the beta-distributed
start angle / ball mass / ball size - as the device reset implements them."""
import numpy as np
import pytest
import torch

from conftest import HAND_BAODING
from myochallenge_b200 import _capi
from myochallenge_b200.envs import make_task_cfg
from myochallenge_b200.sim import BatchSim, Model

ENV = "CustomMyoChallengeBaodingP2-v1"


def _sim(emul_lib, n, seed=0, **kw):
    m = Model(HAND_BAODING, lib=emul_lib)
    cfg = make_task_cfg(m, ENV, **kw)
    return m, cfg, BatchSim(m, n, cfg, device="cpu", seed=seed)


def test_rsi_puts_the_balls_on_their_targets(emul_lib):
    n = 12
    m, cfg, sim = _sim(emul_lib, n, seed=3, enable_rsi=True, rsi_probability=1, balls_overlap=True)
    obs = sim.reset().numpy()
    # obs layout: object1_pos 23:26, object2_pos 29:32, target1_pos 35:38, target2_pos 38:41, errors 41:47
    # the balls sit on the targets' world xy of the in-reset step's observation; the targets ride on the palm, which the
    # restored hand pose moves back by a fraction of a millimetre (tests/env_fixture_checks.py compares these offsets with
    # the reference's own)
    np.testing.assert_allclose(obs[:, 23:25], obs[:, 35:37], atol=1e-3)
    np.testing.assert_allclose(obs[:, 29:31], obs[:, 38:40], atol=1e-3)
    assert np.abs(obs[:, [41, 42, 44, 45]]).max() > 1e-5
    # rotation-task worlds had their targets moved onto the ellipse at the RSI angles: the two targets sit opposite
    q, v, a, _ = [t.numpy() for t in sim.get_state()]
    assert np.all(v == 0)
    # muscle activations after the in-reset zero-action env step: 10 Euler steps of the first-order dynamics at ctrl = sigmoid(-2.5)
    dyn = m.array("actuator_dynprm")[:, :2]
    u = 1.0 / (1.0 + np.exp(2.5))
    ref = np.zeros(sim.na)
    for _ in range(10):
        tau = dyn[:, 0] * (0.5 + 1.5 * ref)
        ref = np.clip(ref + 0.002 * (u - ref) / tau, 0, 1)
    np.testing.assert_allclose(a, np.tile(ref, (n, 1)), rtol=1e-5)
    assert np.all(obs[:, 47:86] > 0)
    # a few worlds are rotation tasks (task ~ choice of 3): their targets differ between worlds, hold worlds keep the model's sites
    assert len(np.unique(np.round(obs[:, 35], 5))) > 2


def _target_angle(sim, xr=0.025, yr=0.028, center=(-0.0125, -0.07)):
    """Angle of target 1 on the goal ellipse, from the target site's position in its body frame (the per-world site_pos the
    env step writes: x = xr cos(a) + cx, y = yr sin(a) + cy, /root/reference/src/envs/baoding.py:357 + BaodingEnvV1.step)."""
    p = sim.get_param(_capi.PARAM_SITE_POS, sim.cfg.target_site[0]).numpy()
    return np.arctan2((p[:, 1] - center[1]) / yr, (p[:, 0] - center[0]) / xr)


def test_rsi_counter_runs_one_step_ahead(emul_lib):
    """self.counter is 1 after an RSI reset (the in-reset env.step advanced it and placed the targets at goal[0]), so the
    first real step uses goal[1]: the targets advance by one goal step from the reset observation. Without RSI the first
    step uses goal[0], i.e. the start angle itself."""
    n = 6
    kw = dict(task_choice="fixed", goal_time_period=(5, 5), goal_xrange=(0.025, 0.025), goal_yrange=(0.028, 0.028))
    _, _, plain = _sim(emul_lib, n, seed=1, **kw)
    _, _, rsi = _sim(emul_lib, n, seed=1, enable_rsi=True, rsi_probability=1, balls_overlap=True, **kw)
    plain.reset()
    rsi.reset()
    a0 = _target_angle(rsi)
    zero = torch.zeros(n, plain.nu)
    plain.step(zero); a_plain = _target_angle(plain)
    rsi.step(zero); a1 = _target_angle(rsi)
    rsi.step(zero); a2 = _target_angle(rsi)
    step = 2 * np.pi * 0.02 / 5.0                       # fixed task = CCW: goal[t] = +2 pi t dt / period
    wrap = lambda x: np.mod(x + np.pi, 2 * np.pi) - np.pi
    np.testing.assert_allclose(wrap(a1 - a0), step, atol=3e-4)
    np.testing.assert_allclose(wrap(a2 - a1), step, atol=3e-4)
    np.testing.assert_allclose(a_plain, 0.25 * np.pi, atol=3e-4)       # fixed task start angle pi / 4, goal[0] = 0
    assert len(np.unique(np.round(a0, 3))) == n                        # RSI start angles are random per world


def test_rsi_probability_zero_and_redraw_of_angles(emul_lib):
    n = 8
    _, _, off = _sim(emul_lib, n, seed=5, enable_rsi=True, rsi_probability=0)
    _, _, ref = _sim(emul_lib, n, seed=5)
    np.testing.assert_array_equal(off.reset().numpy()[:, :35], ref.reset().numpy()[:, :35])    # never taken: plain reset (one extra RNG draw only)


@pytest.mark.parametrize("knob,a,b", [("beta_ball_mass", 2.0, 5.0), ("beta_ball_size", 0.5, 0.5), ("beta_ball_mass", 5.0, 1.0)])
def test_beta_knobs_follow_the_beta_distribution(emul_lib, knob, a, b):
    n = 1500
    m, cfg, sim = _sim(emul_lib, n, seed=7, **{knob: (a, b)})
    sim.reset()
    if knob == "beta_ball_mass":
        lo, hi = 0.03, 0.3
        x = sim.get_param(_capi.PARAM_BODY_MASS, m.name2id("body", "ball1")).numpy().reshape(-1)
    else:
        lo, hi = 0.018, 0.024
        x = sim.get_param(_capi.PARAM_GEOM_SIZE, cfg.ball_geom[1]).numpy()[:, 0]
    z = (x - lo) / (hi - lo)
    assert z.min() >= 0 and z.max() <= 1
    mean, var = a / (a + b), a * b / ((a + b) ** 2 * (a + b + 1))
    assert abs(z.mean() - mean) < 4 * np.sqrt(var / n)
    assert abs(z.var() - var) < 0.15 * var + 1e-3


def test_beta_init_angle(emul_lib):
    n = 1500
    _, _, sim = _sim(emul_lib, n, seed=9, limit_init_angle=np.pi, beta_init_angle=(2.0, 2.0), task_choice="random",
                     goal_xrange=(0.025, 0.025), goal_yrange=(0.028, 0.028))
    sim.reset()
    sim.step(torch.zeros(n, sim.nu))
    # the start angle is sampled only with task_choice "random" (:497-537). After the first step the targets of the rotation worlds
    # sit at the start angle (goal[0] = 0); hold worlds keep the model's sites (angle 3 pi / 4, phase 0) and are left out
    ang = _target_angle(sim)
    phase = np.mod(ang - 0.75 * np.pi + np.pi, 2 * np.pi) - np.pi          # random_phase in [-pi, pi)
    phase = phase[np.abs(phase) > 1e-4]
    n = len(phase)
    assert 800 < n < 1200                                # two of three tasks rotate
    z = (phase + np.pi) / (2 * np.pi)
    assert abs(z.mean() - 0.5) < 4 * np.sqrt(0.05 / n) and abs(z.var() - 0.05) < 0.01       # beta(2, 2): mean 1/2, var 1/20


def test_phase1_reset(emul_lib):
    """CustomBaodingEnv.reset (phase 1, /root/reference/src/envs/baoding.py:146-215): start angles 3 pi / 4 and - pi / 4 (+ the RSI phase),
    ball placement gated by rsi_probability, then noise on the balls / palm / fingers."""
    n = 400
    m = Model(HAND_BAODING, lib=emul_lib)
    P1 = "CustomMyoChallengeBaodingP1-v1"
    # plain P1 (task None = CCW): after the first step the targets sit at 3 pi / 4 (goal[0] = 0)
    sim = BatchSim(m, 8, make_task_cfg(m, P1), device="cpu", seed=2)
    sim.reset()
    sim.step(torch.zeros(8, sim.nu))
    np.testing.assert_allclose(_target_angle(sim), 0.75 * np.pi, atol=3e-4)
    # RSI with probability 0.5 + noises
    cfg = make_task_cfg(m, P1, task="random", enable_rsi=True, rsi_probability=0.5, noise_palm=1.0, noise_fingers=0.5, noise_balls=0.002, drop_th=1.3)
    sim = BatchSim(m, n, cfg, device="cpu", seed=3)
    obs = sim.reset().numpy()
    q = sim.get_state()[0].numpy()
    ti, _, _ = sim.get_task_state()
    took_rsi = (ti.numpy()[:, 3] & 1) == 1
    assert (q[:, 0] >= -np.pi / 2 - 1e-6).all() and (q[:, 0] <= -np.pi / 2 + np.pi / 18 + 1e-6).all() and q[:, 0].std() > 0.02     # palm noise
    assert (np.abs(q[:, 1:3]) <= np.pi / 18 + 1e-6).all()
    assert np.allclose(q[:, 3], q[:, 6]) and (np.abs(q[:, 3]) <= np.pi / 36 + 1e-6).all()                       # thumb: one draw, noise 0.5
    assert np.allclose(q[:, 7], q[:, 22]) and (q[:, 7] >= 0).all() and (q[:, 7] <= np.pi / 12 + 1e-6).all()     # flexions
    assert np.allclose(q[:, 8], q[:, 20]) and (np.abs(q[:, 8]) <= np.pi / 72 + 1e-6).all()                      # abductions
    on_target = took_rsi
    assert 0.4 < on_target.mean() < 0.6
    assert (obs[on_target][:, 47:] > 0).all() and (obs[~on_target][:, 47:] == 0).all()      # activations only after the in-reset step
    # without palm noise the RSI worlds' balls are within the ball noise (+ the palm's sag during the in-reset step) of the targets
    cfg = make_task_cfg(m, P1, task="random", enable_rsi=True, rsi_probability=0.5, noise_balls=0.002, drop_th=1.3)
    sim = BatchSim(m, n, cfg, device="cpu", seed=4)
    obs = sim.reset().numpy()
    took = (sim.get_task_state()[0].numpy()[:, 3] & 1) == 1
    err = np.abs(obs[:, [41, 42, 44, 45]]).max(1)
    assert (err[took] <= 0.002 + 1e-3).all() and (err[~took] > 0.004).mean() > 0.9

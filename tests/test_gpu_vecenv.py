"""GPU: the SB3-shaped front end (MyoVecEnv) over the world kernel: shapes, auto-reset infos as the SubprocVecEnv
worker + Monitor produce them, host-array API == device API, seeds, and the reset distribution of the P2 env
(/root/reference/src/envs/baoding.py:494-604) checked through its observable effects."""
import numpy as np
import pytest
import torch

from myochallenge_b200 import _capi
from myochallenge_b200.envs import EnvironmentFactory, make_vec_env

pytestmark = pytest.mark.gpu


def test_vecenv_contract_and_autoreset(product_lib):
    n = 256
    env = EnvironmentFactory.create("CustomMyoBaodingBallsP2", num_envs=n, seed=3)
    assert env.num_envs == n and env.observation_space.shape == (86,) and env.action_space.shape == (39,)
    obs = env.reset()
    assert obs.shape == (n, 86) and obs.dtype == np.float32 and np.isfinite(obs).all()
    # just-reset Baoding observation: hand pose = init_qpos (palm up), act = 0   [src/envs/baoding.py:400-401]
    assert np.allclose(obs[:, 0], -1.57) and np.allclose(obs[:, 1:23], 0) and np.allclose(obs[:, 47:], 0)
    rng = np.random.default_rng(0)
    lengths, seen_done, seen_trunc = np.zeros(n, int), 0, 0
    for t in range(210):
        obs, rew, done, infos = env.step(rng.uniform(-1, 1, (n, 39)).astype(np.float32))
        lengths += 1
        assert obs.shape == (n, 86) and rew.shape == (n,) and done.dtype == bool and len(infos) == n
        for i in np.nonzero(done)[0][:4]:
            d = infos[int(i)]
            assert d["terminal_observation"].shape == (86,) and d["episode"]["l"] == lengths[i] and "TimeLimit.truncated" in d
            assert lengths[i] <= 200 and (d["TimeLimit.truncated"] == (lengths[i] == 200 and not d["done"]))
            seen_trunc += int(d["TimeLimit.truncated"])
            # after auto-reset the returned observation is the new episode's first one
            assert np.allclose(obs[i, 47:], 0) and obs[i, 0] == pytest.approx(-1.57)
        i = int(np.nonzero(~done)[0][0])
        assert "terminal_observation" not in infos[i] and set(infos[i]["rwd_dict"]) >= {"pos_dist_1", "alive", "solved", "done"}
        seen_done += int(done.sum())
        lengths[done] = 0
    assert seen_done >= n       # every world finished at least once within 210 steps (horizon 200)
    env.close()


def test_host_api_equals_device_api_and_seeds(product_lib):
    n = 64
    a = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, seed=11)
    b = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, seed=11)
    c = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, seed=12)
    oa, ob, oc = a.reset(), b.reset_device().cpu().numpy(), c.reset()
    assert np.array_equal(oa, ob)                                        # same seed -> same worlds
    act = np.random.default_rng(1).uniform(-1, 1, (n, 39)).astype(np.float32)
    oc1, _, _, _ = c.step(act)
    for _ in range(5):
        oa, ra, da, _ = a.step(act)
        ob, rb, db, _ = [t.cpu().numpy() for t in b.step_device(torch.from_numpy(act).cuda())]
        assert np.array_equal(oa, ob) and np.array_equal(ra, rb) and np.array_equal(da, db.astype(bool))
    a2 = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, seed=11)
    a2.reset()
    oa1, _, _, _ = a2.step(act)
    assert not np.array_equal(oa1, oc1)                                  # another seed -> other tasks / radii / ball physics


def test_p2_reset_distribution(product_lib):
    """Physics randomisation ranges of the P2 registration show up in the per-world parameters after reset, and the
    three tasks (hold / cw / ccw) are drawn uniformly."""
    n = 4096
    env = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, seed=5)
    env.reset()
    sim, cfg = env.sim, env.cfg
    for k in range(2):
        mass = sim.get_param(_capi.PARAM_BODY_MASS, cfg.ball_body[k]).cpu().numpy()[:, 0]
        size = sim.get_param(_capi.PARAM_GEOM_SIZE, cfg.ball_geom[k]).cpu().numpy()[:, 0]
        fri = sim.get_param(_capi.PARAM_GEOM_FRICTION, cfg.ball_geom[k]).cpu().numpy()
        assert 0.03 <= mass.min() and mass.max() <= 0.3 and abs(mass.mean() - 0.165) < 0.01
        assert 0.018 <= size.min() and size.max() <= 0.024 and abs(size.mean() - 0.021) < 3e-4
        assert 0.8 <= fri[:, 0].min() and fri[:, 0].max() <= 1.2 and fri[:, 0].std() > 0.1
    # the target sites (positions in their body frame) move only for cw / ccw: after one step a third of the worlds keep theirs (hold)
    s0 = sim.get_param(_capi.PARAM_SITE_POS, cfg.target_site[0]).cpu().numpy().copy()
    env.step(np.zeros((n, 39), np.float32))
    s1 = sim.get_param(_capi.PARAM_SITE_POS, cfg.target_site[0]).cpu().numpy()
    moved = np.abs(s1[:, :2] - s0[:, :2]).max(1) > 1e-7
    assert 0.6 < moved.mean() < 0.73
    env.close()


def test_rsi_and_beta_reset_matches_host_emulation(product_lib, emul_lib):
    """The curriculum knobs of the reset (RSI, beta-distributed angle / mass / size; /root/reference/src/envs/baoding.py:504-520,
    563-638) on the GPU against the same kernel sources run on the host (tests/test_reset_knobs_emul.py checks those against
    the reference's semantics): same counter-based draws, so observations and parameters agree to float rounding."""
    from conftest import HAND_BAODING
    from myochallenge_b200.envs import make_task_cfg
    from myochallenge_b200.sim import BatchSim, Model

    n = 64
    kw = dict(enable_rsi=True, rsi_probability=0.7, limit_init_angle=np.pi, beta_init_angle=(2.0, 3.0), beta_ball_mass=(2.0, 5.0),
              beta_ball_size=(0.5, 0.5), noise_fingers=0.5)
    outs = []
    for lib, dev in ((product_lib, "cuda:0"), (emul_lib, "cpu")):
        m = Model(HAND_BAODING, lib=lib)
        cfg = make_task_cfg(m, "CustomMyoChallengeBaodingP2-v1", **kw)
        sim = BatchSim(m, n, cfg, device=dev, seed=21)
        obs = sim.reset().cpu().numpy().copy()
        mass = sim.get_param(_capi.PARAM_BODY_MASS, m.name2id("body", "ball2")).cpu().numpy().copy()
        size = sim.get_param(_capi.PARAM_GEOM_SIZE, cfg.ball_geom[0]).cpu().numpy().copy()
        o1 = sim.step(torch.zeros(n, sim.nu, device=dev))[0].cpu().numpy().copy()
        outs.append((obs, mass, size, o1))
    for a, b in zip(*outs):      # (the RSI reset runs an env step: 32-lane and single-lane summation orders differ by a few fp32 ulps of the forces)
        np.testing.assert_allclose(a, b, rtol=5e-4, atol=5e-4)
    obs = outs[0][0]
    on_target = np.abs(obs[:, 41:43]).max(1) < 1e-3       # the in-reset step leaves the palm-mounted targets a fraction of a millimetre off
    assert 0.5 < on_target.mean() < 0.9            # rsi_probability 0.7


def test_phase1_reset_and_curriculum_steps_on_gpu(product_lib, emul_lib):
    """The phase-1 env's reset (RSI gate, ball / palm / finger noise, task selector) on the GPU against the host build of the same
    sources, and a few steps of the packaged 32-step curriculum created through the factory and stepped."""
    from conftest import HAND_BAODING
    from myochallenge_b200 import curriculum
    from myochallenge_b200.envs import make_task_cfg
    from myochallenge_b200.sim import BatchSim, Model

    n = 64
    kw = dict(task="random", enable_rsi=True, rsi_probability=0.6, noise_palm=0.5, noise_fingers=0.3, noise_balls=0.001, drop_th=1.3)
    outs = []
    for lib, dev in ((product_lib, "cuda:0"), (emul_lib, "cpu")):
        m = Model(HAND_BAODING, lib=lib)
        sim = BatchSim(m, n, make_task_cfg(m, "CustomMyoChallengeBaodingP1-v1", **kw), device=dev, seed=4)
        obs = sim.reset().cpu().numpy().copy()
        o1 = sim.step(torch.zeros(n, sim.nu, device=dev))[0].cpu().numpy().copy()
        outs.append((obs, o1))
    for a, b in zip(*outs):
        np.testing.assert_allclose(a, b, rtol=5e-4, atol=5e-4)
    steps = curriculum.load()
    rng = np.random.default_rng(0)
    for k in range(len(steps)):                       # all 32 steps of the winning curriculum
        env = EnvironmentFactory.create(steps[k]["env_name"], num_envs=32, seed=k, **steps[k]["config"])
        obs = env.reset()
        assert np.isfinite(obs).all(), steps[k]["step"]
        for _ in range(12):
            obs, rew, done, infos = env.step(rng.uniform(-1, 1, (32, 39)).astype(np.float32))
        assert obs.shape == (32, 86) and np.isfinite(obs).all() and np.isfinite(rew).all(), steps[k]["step"]
        assert env.sim.status() & 8 == 0, steps[k]["step"]          # no non-finite state
        env.close()

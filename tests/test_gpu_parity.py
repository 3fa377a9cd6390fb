"""Parity tests proper: the CUDA world kernel, called through the C ABI, against the oracle on the same
seeded inputs (one step, per-stage, env step), plus size-independent properties at the BASELINE sizes."""
import numpy as np
import pytest
import torch

from conftest import ELBOW, FINGER, HAND_BAODING, HAND_DIE, HAND_POSE
from myochallenge_b200 import _capi
import parity_common as pc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = [("elbow", ELBOW, _capi.TASK_POSE, 16), ("finger", FINGER, _capi.TASK_POSE, 64),
         ("hand_pose", HAND_POSE, _capi.TASK_POSE, 16), ("baoding", HAND_BAODING, _capi.TASK_BAODING, 32),
         ("die", HAND_DIE, _capi.TASK_REORIENT, 32)]


@pytest.mark.parametrize("name,path,kind,n", CASES, ids=[c[0] for c in CASES])
def test_one_step_state_and_contact_parity(product_lib, name, path, kind, n):
    pc.check_one_step(product_lib, DEV, path, kind, n, seed=21)


@pytest.mark.parametrize("name,path,kind,n", CASES, ids=[c[0] for c in CASES])
def test_stage_parity(product_lib, name, path, kind, n):
    pc.check_stages(product_lib, DEV, path, kind, n, seed=22)


@pytest.mark.parametrize("name,path,kind,n", [CASES[1], CASES[3]], ids=["finger", "baoding"])
def test_env_step(product_lib, name, path, kind, n):
    pc.check_env_step_matches_mj_steps(product_lib, DEV, path, kind, 8)


def test_episode_return_distribution_matches_oracle(product_lib):
    """north_star's multi-step bar: 128 fully randomised Baoding worlds, 40 env steps (400 mj_steps) under seeded random actions,
    against the fp64 oracle replaying every world: episode lengths, returns, and a KS test on the two return samples."""
    out = pc.check_episode_returns(product_lib, DEV, 128, 40)
    assert out["dropped"] > 0.05, out            # the sample contains early terminations (drops), not just full-length episodes
    print(out)


def test_lane_width_invariance(product_lib, monkeypatch):
    """The tile width only changes which lane does the work and the order of the tile-wide reductions:
    8-, 16- and 32-lane runs of the same worlds agree to fp32 rounding (1e-5 relative after 5 substeps)."""
    from conftest import random_states
    qpos, qvel, act, ctrl = random_states(HAND_BAODING, 16, 5)
    outs = []
    for lanes in ("8", "16", "32"):
        monkeypatch.setenv("MYO_LANES", lanes)
        _, _, B = pc.make_batch(product_lib, HAND_BAODING, _capi.TASK_BAODING, 16, DEV)
        B.set_state(qpos, qvel, act)
        B.mj_step(ctrl, 5)
        outs.append(torch.cat([t.reshape(16, -1) for t in B.get_state()[:3]], 1).cpu())
    for o in outs[1:]:
        assert torch.allclose(outs[0], o, rtol=1e-5, atol=1e-5), float((outs[0] - o).abs().max())


def test_full_size_properties(product_lib):
    """BASELINE config 5 size (32k worlds): determinism across runs, independence of worlds from the batch they
    sit in, TimeLimit / auto-reset bookkeeping and finite outputs."""
    n = 32768
    def run(nw, steps):
        _, cfg, B = pc.make_batch(product_lib, HAND_BAODING, _capi.TASK_BAODING, nw, DEV, task_choice_random=1)
        B.reset()
        g = torch.Generator(device="cpu").manual_seed(0)
        ep_len = torch.zeros(nw, dtype=torch.int32, device=DEV)
        ndone = 0
        for s in range(steps):
            a = (torch.rand(n, B.nu, generator=g) * 2 - 1)[:nw].to(DEV)
            obs, rew, done, trunc = B.step(a)
            assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
            ep_len += 1
            assert (ep_len[done.bool()] <= cfg.max_episode_steps).all()
            assert (trunc.bool() <= done.bool()).all()                 # truncated implies done
            ndone += int(done.sum())
            ep_len[done.bool()] = 0
        return obs.clone(), rew.clone(), ndone, B.status()
    o1, r1, nd1, st1 = run(n, 12)
    o2, r2, nd2, st2 = run(n, 12)
    assert torch.equal(o1, o2) and torch.equal(r1, r2) and nd1 == nd2          # deterministic
    o3, r3, _, _ = run(1000, 12)
    assert torch.equal(o1[:1000], o3) and torch.equal(r1[:1000], r3)            # worlds are independent
    assert st1 & 8 == 0                                                          # no non-finite state


def test_more_contacts_than_the_fast_layout_holds(product_lib):
    """Worlds with 17+ contacts: pairs bit-exact through the full-capacity parity hooks, and the env step (fast kernel + redo
    pass) reproduces the oracle's step with no overflow reported."""
    pc.check_many_contacts(product_lib, "cuda:0", n=32)

"""CPU: the rollout-side oracle (oracle/rollout_oracle.py: SB3 RunningMeanStd / VecNormalize reward path / GAE,
restated) against closed-form known answers, the reference's stored VecNormalize constants, and the host logic of
the cross-rank moment merge (rollout.merge_moment_states; gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN
from oracle import rollout_oracle as ro


def test_gae_matches_its_definition():
    rng = np.random.default_rng(0)
    T, n = 24, 9
    r, v = rng.normal(0, 1, (T, n)).astype(np.float32), rng.normal(0, 2, (T, n)).astype(np.float32)
    starts = rng.random((T, n)) < 0.15
    last_v, dones = rng.normal(0, 2, n).astype(np.float32), rng.random(n) < 0.3
    adv, ret = ro.gae(r, v, starts, last_v, dones, 0.99, 0.95)
    ref = ro.gae_by_definition(r, v, starts, last_v, dones, 0.99, 0.95)
    np.testing.assert_allclose(adv, ref, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(ret, adv + v, rtol=0, atol=0)
    # an episode boundary cuts the sum: the advantage before a start does not see rewards after it
    r2 = r.copy(); r2[10:] += 100.0
    starts2 = starts.copy(); starts2[10] = True
    a1, _ = ro.gae(r, v, starts2, last_v, dones, 0.99, 0.95)
    a2, _ = ro.gae(r2, v, starts2, last_v, dones, 0.99, 0.95)
    np.testing.assert_array_equal(a1[:9], a2[:9])


def test_running_mean_std_equals_moments_of_the_concatenation():
    rng = np.random.default_rng(1)
    batches = [rng.normal(3, 2, (k, 5)) for k in (7, 64, 33, 1)]
    rms = ro.RunningMeanStd(epsilon=1e-4, shape=(5,))
    for b in batches:
        rms.update(b)
    x = np.concatenate(batches)
    # the prior (mean 0, var 1, count 1e-4) is one more tiny batch
    n, eps = len(x), 1e-4
    mean = x.sum(0) / (n + eps)
    var = ((x - mean) ** 2).sum(0) / (n + eps) + eps * (1 + mean ** 2) / (n + eps)
    np.testing.assert_allclose(rms.mean, mean, rtol=1e-12)
    np.testing.assert_allclose(rms.var, var, rtol=1e-10)
    assert rms.count == pytest.approx(n + eps)


def test_vecnormalize_reward_path_and_reference_constants():
    g = np.load(os.path.join(GOLDEN, "vecnormalize_baoding_step32.npz"))
    # constants stored in the reference's own VecNormalize pickle
    assert float(g["clip_obs"]) == 10.0 and float(g["epsilon"]) == 1e-8
    rng = np.random.default_rng(2)
    n = 50
    rms, returns = ro.RunningMeanStd(shape=()), np.zeros(n)
    tot = np.zeros(n)
    for t in range(20):
        rew = rng.normal(1, 5, n)
        dones = rng.random(n) < 0.1
        tot = tot * 0.99 + rew
        out = ro.vecnormalize_step(rms, returns, rew, dones, gamma=0.99, epsilon=1e-8, clip_reward=10.0)
        assert np.abs(out).max() <= 10.0
        np.testing.assert_allclose(out, np.clip(rew / np.sqrt(rms.var + 1e-8), -10, 10).astype(np.float32))
        assert (returns[dones] == 0).all()
        np.testing.assert_allclose(returns[~dones], tot[~dones])
        tot[dones] = 0
    # obs normalisation with the reference's stored moments maps its stored mean to 0
    rmo = ro.RunningMeanStd(shape=g["obs_mean"].shape)
    rmo.mean, rmo.var = g["obs_mean"].astype(np.float64), g["obs_var"].astype(np.float64)
    np.testing.assert_allclose(ro.normalize_obs(rmo, g["obs_mean"][None]), 0, atol=1e-6)


def test_merge_moment_states_matches_single_process():
    from myochallenge_b200.rollout import merge_moment_states

    rng = np.random.default_rng(3)
    d = 4
    base = ro.RunningMeanStd(shape=(d,))
    base.update(rng.normal(0, 1, (10, d)))
    shards = [rng.normal(2, 3, (16, d)), rng.normal(-1, 0.5, (24, d))]

    def state(r):
        return np.concatenate([r.mean, r.var, [r.count]])

    per_rank = []
    for sh in shards:
        r = ro.RunningMeanStd(shape=(d,))
        r.mean, r.var, r.count = base.mean.copy(), base.var.copy(), base.count
        r.update(sh)
        per_rank.append(state(r))
    merged = merge_moment_states(per_rank, d, base=state(base))
    ref = ro.RunningMeanStd(shape=(d,))
    ref.mean, ref.var, ref.count = base.mean.copy(), base.var.copy(), base.count
    ref.update(np.concatenate(shards))
    np.testing.assert_allclose(merged[:d], ref.mean, rtol=1e-11)
    np.testing.assert_allclose(merged[d:2 * d], ref.var, rtol=1e-9)
    assert merged[2 * d] == pytest.approx(ref.count)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _sync_worker(rank, world, port, q):
    from myochallenge_b200.rollout import _all_gather, merge_moment_states

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    x = rng.normal(1, 2, (40, 3))
    r = ro.RunningMeanStd(shape=(3,))
    base = np.concatenate([r.mean, r.var, [r.count]])
    r.update(x[rank * 20:(rank + 1) * 20])                 # this rank's worlds
    mine = torch.from_numpy(np.concatenate([r.mean, r.var, [r.count]]))
    merged = merge_moment_states([t.numpy() for t in _all_gather(mine)], 3, base=base)
    q.put((rank, merged))
    dist.destroy_process_group()


def test_two_rank_rollout_moment_sync_gloo():
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    ps = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
    rng = np.random.default_rng(11)
    ref = ro.RunningMeanStd(shape=(3,))
    ref.update(rng.normal(1, 2, (40, 3)))
    for _, merged in res:
        np.testing.assert_allclose(merged[:3], ref.mean, rtol=1e-11)
        np.testing.assert_allclose(merged[3:6], ref.var, rtol=1e-9)
        assert merged[6] == pytest.approx(ref.count)


def test_merge_with_per_rank_bases_is_rank_consistent_and_exact():
    """ranks whose bases differ (each folded its own reset observations in before the base was
    taken) must still end with one common state, and it must not drift when the merge is repeated rollout after rollout."""
    from myochallenge_b200.rollout import merge_moment_states

    rng = np.random.default_rng(5)
    d, K = 3, 4

    def state(r):
        return np.concatenate([r.mean, r.var, [r.count]])

    ranks = []
    for k in range(K):
        r = ro.RunningMeanStd(shape=(d,))
        r.update(rng.normal(k, 1 + k, (8, d)))           # rank-specific reset batch: bases differ
        ranks.append(r)
    spread = []
    for it in range(8):
        bases = [state(r) for r in ranks]
        batches = [rng.normal(0.5 * k, 2.0, (32, d)) for k in range(K)]
        for r, b in zip(ranks, batches):
            r.update(b)
        states = [state(r) for r in ranks]
        merged = [merge_moment_states(states, d, bases=bases) for _ in range(K)]      # what each rank computes
        for m in merged[1:]:
            np.testing.assert_array_equal(m, merged[0])
        # exact: rank 0's base + every rank's batch of this interval
        ref = ro.RunningMeanStd(shape=(d,))
        ref.mean, ref.var, ref.count = bases[0][:d].copy(), bases[0][d:2 * d].copy(), bases[0][2 * d]
        ref.update(np.concatenate(batches))
        np.testing.assert_allclose(merged[0][:d], ref.mean, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(merged[0][d:2 * d], ref.var, rtol=1e-8)
        assert merged[0][2 * d] == pytest.approx(ref.count)
        for r in ranks:
            r.mean, r.var, r.count = merged[0][:d].copy(), merged[0][d:2 * d].copy(), float(merged[0][2 * d])
        spread.append(float(np.abs(merged[0][:d]).max()))
    assert max(spread) < 5.0 and np.isfinite(spread).all()          # the r1 formula reached 2e3 here


def _sync3_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from myochallenge_b200.rollout import _all_gather, merge_moment_states

    rng = np.random.default_rng(100 + rank)
    r = ro.RunningMeanStd(shape=(2,))
    r.update(rng.normal(rank, 1, (5, 2)))                 # per-rank reset observations: a different base on every rank
    out = []
    for it in range(3):
        base = torch.from_numpy(np.concatenate([r.mean, r.var, [r.count]]))
        r.update(rng.normal(0, 1, (16, 2)))
        mine = torch.cat([torch.from_numpy(np.concatenate([r.mean, r.var, [r.count]])), base])
        got = torch.stack(_all_gather(mine))
        m = merge_moment_states(got[:, :5], 2, bases=got[:, 5:]).numpy()
        r.mean, r.var, r.count = m[:2].copy(), m[2:4].copy(), float(m[4])
        out.append(m)
    q.put((rank, np.stack(out)))
    dist.destroy_process_group()


def test_three_rank_sync_with_different_bases_gloo():
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    ps = [ctx.Process(target=_sync3_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
    for _, m in res[1:]:
        np.testing.assert_array_equal(m, res[0][1])          # every rank holds the same moments after every sync
    counts = res[0][1][:, 4]
    np.testing.assert_allclose(np.diff(counts), 48.0)          # 3 ranks x 16 samples per interval, nothing double counted

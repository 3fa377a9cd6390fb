"""CPU checks around the PPO update: the oracle's two formulations agree (sb3-contrib's
split-and-pad sequences vs the product's whole sequences with the state masked at episode starts - the equivalence the
CUDA design rests on), the Adam / clip restatement is torch's own, and the flat-bucket gradient all-reduce used across
ranks (gloo, world_size 2)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT
from oracle import ppo_oracle


def _random_sd(O, A, H, pi, vf, seed):
    g = torch.Generator().manual_seed(seed)
    u = lambda *s, b=1.0: ((torch.rand(*s, generator=g) * 2 - 1) * b).numpy().astype(np.float64)
    sd = {"log_std": np.full(A, -0.7)}
    for name in ("lstm_actor", "lstm_critic"):
        b = H ** -0.5
        sd.update({f"{name}.weight_ih_l0": u(4 * H, O, b=b), f"{name}.weight_hh_l0": u(4 * H, H, b=b), f"{name}.bias_ih_l0": u(4 * H, b=b),
                   f"{name}.bias_hh_l0": u(4 * H, b=b)})
    for net, widths in (("policy_net", pi), ("value_net", vf)):
        d = H
        for l, w in enumerate(widths):
            sd[f"mlp_extractor.{net}.{2 * l}.weight"] = u(w, d, b=d ** -0.5); sd[f"mlp_extractor.{net}.{2 * l}.bias"] = u(w, b=d ** -0.5)
            d = w
    sd["action_net.weight"] = u(A, pi[-1] if pi else H, b=0.1); sd["action_net.bias"] = u(A, b=0.1)
    sd["value_net.weight"] = u(1, vf[-1] if vf else H, b=0.1); sd["value_net.bias"] = u(1, b=0.1)
    return sd


@pytest.mark.parametrize("T,B,start_prob", [(9, 4, 0.25), (6, 3, 0.0), (7, 5, 0.6)])
def test_masked_recurrence_equals_split_and_pad(T, B, start_prob):
    sd = _random_sd(11, 4, 16, (24,), (8, 16), seed=T)
    batch = ppo_oracle.synthetic_batch(sd, T, B, seed=B, start_prob=start_prob, dtype=np.float64)
    hyper = dict(clip_range=0.2, ent_coef=0.01, vf_coef=0.7, normalize_advantage=True)
    grads, stats = ppo_oracle.gradients(sd, batch, **hyper)
    loss2, grads2 = ppo_oracle.masked_recurrence_loss(sd, batch, **hyper)
    assert abs(stats["loss"] - loss2) < 1e-10
    assert 0.0 < stats["clip_fraction"] < 1.0          # both branches of the clipped surrogate are exercised
    for k, g in grads.items():
        assert torch.allclose(g, grads2[k], rtol=1e-9, atol=1e-12), k


def test_value_clipping_and_no_adv_normalisation_agree_too():
    sd = _random_sd(7, 3, 8, (), (), seed=3)
    batch = ppo_oracle.synthetic_batch(sd, 5, 6, seed=1, start_prob=0.3, dtype=np.float64)
    hyper = dict(clip_range=0.1, clip_range_vf=0.05, ent_coef=0.0, vf_coef=1.0, normalize_advantage=False)
    grads, stats = ppo_oracle.gradients(sd, batch, **hyper)
    loss2, grads2 = ppo_oracle.masked_recurrence_loss(sd, batch, **hyper)
    assert abs(stats["loss"] - loss2) < 1e-10
    for k, g in grads.items():
        assert torch.allclose(g, grads2[k], rtol=1e-9, atol=1e-12), k


def test_adam_restatement_follows_the_textbook_update():
    rng = np.random.default_rng(0)
    p = rng.normal(size=50); g = rng.normal(size=50) * 3; m = rng.normal(size=50) * 0.1; v = rng.random(50) * 0.01
    lr, b1, b2, eps, step, mx = 1e-3, 0.9, 0.999, 1e-5, 7, 0.5
    p1, m1, v1, norm = ppo_oracle.adam_step(p, g, m, v, step, lr, (b1, b2), eps, mx)
    gc = g * min(1.0, mx / (np.linalg.norm(g) + 1e-6))
    m_ref = b1 * m + (1 - b1) * gc; v_ref = b2 * v + (1 - b2) * gc * gc
    p_ref = p - lr / (1 - b1 ** step) * m_ref / (np.sqrt(v_ref) / np.sqrt(1 - b2 ** step) + eps)
    assert abs(norm - np.linalg.norm(g)) < 1e-9
    assert np.allclose(p1.numpy(), p_ref, rtol=1e-10) and np.allclose(m1.numpy(), m_ref) and np.allclose(v1.numpy(), v_ref)


def test_golden_checkpoint_gradients_are_finite_and_nonzero():
    g = np.load(os.path.join(GOLDEN, "policy_phase1.npz"))
    sd = {k[2:]: g[k].astype(np.float64) for k in g.files if k.startswith("w:")}
    batch = ppo_oracle.synthetic_batch(sd, 4, 3, seed=0, dtype=np.float64)
    grads, stats = ppo_oracle.gradients(sd, batch, ent_coef=0.001)
    assert np.isfinite(stats["loss"])
    for k, v in grads.items():
        assert torch.isfinite(v).all() and float(v.abs().max()) > 0, k


def _rank_main(rank, world, port, out):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from myochallenge_b200.ppo import allreduce_flat

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)
    scale = allreduce_flat(g)
    out[rank] = (g * scale).tolist()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    port = 29500 + os.getpid() % 500
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_rank_main, args=(2, port, out), nprocs=2, join=True)
        want = (torch.arange(10, dtype=torch.float32) * 1.5).tolist()       # mean of 1x and 2x
        assert out[0] == want and out[1] == want

"""CPU: the fp32 restatement of the policy forward (myochallenge_b200.policy.torch_reference_forward, the
floating-point reference the GPU tests compare the tcgen05 kernel against) reproduces the outputs that stock
torch.nn.LSTM / Linear give on the reference's shipped checkpoint (tests/golden/policy_phase1.npz, made by
tests/golden/make_policy_golden.py from /root/reference/trained_models/phase_1/phase1_final.zip)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from myochallenge_b200.policy import torch_reference_forward


def load_golden():
    g = np.load(os.path.join(GOLDEN, "policy_phase1.npz"))
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}
    return g, sd


def test_fixture_shapes():
    g, sd = load_golden()
    assert g["obs"].shape == (16, 86) and g["h"].shape == (2, 16, 128) and g["mean"].shape == (16, 39)
    assert sd["lstm_actor.weight_ih_l0"].shape == (512, 86) and sd["action_net.weight"].shape == (39, 128)
    assert sum(v.numel() for v in sd.values()) == 226383     # parameter count of the shipped checkpoint


def test_restatement_matches_torch_lstm():
    g, sd = load_golden()
    a, v, lp, h1, c1 = torch_reference_forward(sd, torch.from_numpy(g["obs"]), torch.from_numpy(g["h"]), torch.from_numpy(g["c"]),
                                               torch.from_numpy(g["starts"]))
    np.testing.assert_allclose(a.numpy(), g["mean"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(v.numpy(), g["value"], rtol=1e-5, atol=1e-5)
    for net in range(2):
        np.testing.assert_allclose(h1[net].numpy(), g[f"h_out_{net}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(c1[net].numpy(), g[f"c_out_{net}"], rtol=1e-5, atol=1e-6)
    # deterministic log-prob of the mean action: -sum(log_std) - 39/2 log(2 pi)
    ref = -(sd["log_std"].sum() + 0.5 * 39 * np.log(2 * np.pi))
    assert abs(float(lp[0]) - float(ref)) < 1e-4

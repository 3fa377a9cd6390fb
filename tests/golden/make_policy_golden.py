"""Generates tests/golden/policy_phase1.npz from the reference's shipped checkpoint
(/root/reference/trained_models/phase_1/phase1_final.zip: SB3 RecurrentPPO zip, LSTM 86->128 x2,
no MLP layers). Run in the build container, where /root/reference exists:

    python tests/golden/make_policy_golden.py

The expected outputs are computed with stock ``torch.nn.LSTM`` / ``torch.nn.Linear`` modules on CPU in
fp32 from the checkpoint's own tensors: weights, the stored rollout batch ``_last_obs[16,86]``
(VecNormalize-normalised), ``_last_episode_starts`` and ``_last_lstm_states`` (pi / vf h, c [1,16,128]).
"""
import base64
import io
import json
import os
import pickle
import sys
import zipfile

import numpy as np
import torch

REF = "/root/reference/trained_models/phase_1/phase1_final.zip"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "policy_phase1.npz")


class _Stub:
    def __init__(self, *a, **k):
        self.args, self.kw = a, k

    def __setstate__(self, s):
        self.state = s


class _RNNStates(tuple):
    def __new__(cls, pi, vf):
        return tuple.__new__(cls, (pi, vf))


class _Unpickler(pickle.Unpickler):
    """sb3_contrib / gym are not installed: their classes unpickle to tuples / stubs."""

    def find_class(self, module, name):
        if module.startswith("numpy") or module in ("builtins", "collections", "copyreg", "_codecs"):
            return super().find_class(module, name)
        if module.startswith("torch"):
            return super().find_class(module, name)
        if name == "RNNStates":
            return _RNNStates
        return _Stub


def _load(blob):
    return _Unpickler(io.BytesIO(base64.b64decode(blob[":serialized:"]))).load()


def main():
    z = zipfile.ZipFile(REF)
    sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu")
    data = json.loads(z.read("data"))
    obs = np.asarray(_load(data["_last_obs"]), np.float32)
    starts = np.asarray(_load(data["_last_episode_starts"]), np.float32)
    pi_state, vf_state = _load(data["_last_lstm_states"])
    h = np.stack([np.asarray(pi_state[0], np.float32)[0], np.asarray(vf_state[0], np.float32)[0]])   # [2,16,128]
    c = np.stack([np.asarray(pi_state[1], np.float32)[0], np.asarray(vf_state[1], np.float32)[0]])
    H = sd["lstm_actor.weight_hh_l0"].shape[1]
    outs = {}
    with torch.no_grad():
        feats = []
        for net, name in enumerate(("lstm_actor", "lstm_critic")):
            lstm = torch.nn.LSTM(obs.shape[1], H, num_layers=1)
            lstm.load_state_dict({k.split(".", 1)[1]: v for k, v in sd.items() if k.startswith(name)})
            keep = torch.from_numpy(1.0 - starts).view(1, -1, 1)
            y, (h1, c1) = lstm(torch.from_numpy(obs).unsqueeze(0), (torch.from_numpy(h[net:net + 1]) * keep, torch.from_numpy(c[net:net + 1]) * keep))
            feats.append(y[0])
            outs[f"h_out_{net}"] = h1[0].numpy()
            outs[f"c_out_{net}"] = c1[0].numpy()
        mean = torch.nn.functional.linear(feats[0], sd["action_net.weight"], sd["action_net.bias"])
        value = torch.nn.functional.linear(feats[1], sd["value_net.weight"], sd["value_net.bias"]).squeeze(1)
    np.savez_compressed(OUT, obs=obs, starts=starts, h=h, c=c, mean=mean.numpy(), value=value.numpy(),
                        **outs, **{"w:" + k: v.numpy() for k, v in sd.items()})
    print("wrote", OUT, os.path.getsize(OUT), "bytes; obs", obs.shape, "starts", starts, "H", H)
    print("mean[0,:5]", mean[0, :5].numpy(), "value[:4]", value[:4].numpy())


if __name__ == "__main__":
    sys.exit(main())

"""Generates tests/golden/vecnormalize_baoding_step32.npz from the reference's shipped VecNormalize pickle
(/root/reference/trained_models/curriculum_steps_complete_baoding_winner/32_phase_2_smaller_rate_resume/env.pkl):
obs_rms / ret_rms running moments and the wrapper constants (clip_obs, clip_reward, gamma, epsilon).
Run in the build container, where /root/reference exists:   python tests/golden/make_vecnormalize_golden.py
SB3 / gym are not installed here: their classes unpickle to attribute bags."""
import os
import pickle

import numpy as np

REF = "/root/reference/trained_models/curriculum_steps_complete_baoding_winner/32_phase_2_smaller_rate_resume/env.pkl"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vecnormalize_baoding_step32.npz")


class _Bag:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, s):
        self.__dict__.update(s if isinstance(s, dict) else {"state": s})


class _U(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("numpy") or module in ("builtins", "collections", "copyreg", "_codecs"):
            return super().find_class(module, name)
        return type(name, (_Bag,), {})


def main():
    with open(REF, "rb") as f:
        v = _U(f).load()
    d = v.__dict__
    o, r = d["obs_rms"].__dict__, d["ret_rms"].__dict__
    np.savez(OUT, obs_mean=np.asarray(o["mean"], np.float64), obs_var=np.asarray(o["var"], np.float64), obs_count=float(o["count"]),
             ret_mean=float(r["mean"]), ret_var=float(r["var"]), ret_count=float(r["count"]), clip_obs=float(d["clip_obs"]),
             clip_reward=float(d["clip_reward"]), gamma=float(d["gamma"]), epsilon=float(d["epsilon"]))
    print("wrote", OUT, "obs_mean[23:26]", np.asarray(o["mean"])[23:26], "count", o["count"], "clip", d["clip_obs"], d["clip_reward"], d["gamma"], d["epsilon"])


if __name__ == "__main__":
    main()

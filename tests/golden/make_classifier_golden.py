"""TEST INFRASTRUCTURE - writes tests/golden/task_classifier.npz from the reference's own classifier.

The reference's ``models/classifier.py::TaskClassifier`` (imported unmodified through tests/shim) is loaded with the shipped
``trained_models/winning_ensemble/classifier/classifier.pt``; the shipped ``scaler.pkl`` (sklearn ``StandardScaler``) is
unpickled with sklearn itself. Inputs are drawn around the scaler's own statistics; the recorded logits / decisions are what
``myochallenge_b200.ensemble.TaskClassifier`` has to reproduce. Only runs where /root/reference exists.

    python tests/golden/make_classifier_golden.py
"""
import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import shim  # noqa: E402

shim.install()
from models.classifier import TaskClassifier  # noqa: E402

REF = "/root/reference/trained_models/winning_ensemble/classifier"
clf = TaskClassifier(13)
sd = torch.load(os.path.join(REF, "classifier.pt"), map_location="cpu")
clf.load_state_dict(sd)
with open(os.path.join(REF, "scaler.pkl"), "rb") as f:
    scaler = pickle.load(f)
rng = np.random.default_rng(0)
x = scaler.mean_ + scaler.scale_ * rng.normal(0, 1.5, (256, 13 * 18))
with torch.no_grad():
    logits = clf(torch.FloatTensor(scaler.transform(x))).squeeze(-1).numpy()
    task = torch.round(torch.sigmoid(torch.from_numpy(logits))).numpy().astype(np.int64)
out = {"w:" + k: v.numpy() for k, v in sd.items()}
out.update(scaler_mean=scaler.mean_, scaler_scale=scaler.scale_, x=x.astype(np.float32), logits=logits, task=task)
np.savez_compressed(os.path.join(HERE, "task_classifier.npz"), **out)
print("wrote task_classifier.npz:", {k: v.shape for k, v in out.items()}, "hold fraction", float((task == 0).mean()))

"""Test configuration. ``-m "not gpu"`` runs on a CPU-only box: oracle vs golden vectors, host logic,
C-ABI symbol checks and the single-lane host emulation of the kernel sources (tests/emul).
``-m gpu`` runs the parity tests proper through the C ABI on a B200."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ASSETS = os.path.join(ROOT, "myochallenge_b200", "assets")
FINGER = os.path.join(ASSETS, "finger", "myo_finger_v0.mjb")
MOTOR_FINGER = os.path.join(ASSETS, "finger", "motor_finger_v0.mjb")
MYO_LOAD = os.path.join(ASSETS, "basic", "myo_load.mjb")
HAND_BAODING = os.path.join(ASSETS, "hand", "myo_hand_baoding.mjb")
HAND_POSE = os.path.join(ASSETS, "hand", "myo_hand_pose.mjb")
HAND_DIE = os.path.join(ASSETS, "hand", "myo_hand_die.mjb")
ELBOW = os.path.join(ASSETS, "arm", "myo_elbow_1dof6muscles.mjb")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def product_lib():
    """The product library. Built by __graft_entry__.build(); model functions work without a GPU."""
    from myochallenge_b200 import _capi

    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return _capi.lib()


@pytest.fixture(scope="session")
def emul_lib():
    """TEST INFRASTRUCTURE: the kernel sources compiled for the host, one lane per world (tests/emul)."""
    from myochallenge_b200 import _capi

    d = os.path.join(ROOT, "tests", "emul")
    so = os.path.join(d, "libmyo_emul.so")
    srcs = [os.path.join(d, "emul.cpp")] + [os.path.join(ROOT, "myochallenge_b200", "csrc", f)
                                            for f in os.listdir(os.path.join(ROOT, "myochallenge_b200", "csrc")) if "." in f]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.isfile(s)):
        subprocess.check_call([os.path.join(d, "build.sh")])
    return _capi.bind(so, only=("myo_model", "myo_batch", "myo_task", "myo_last", "myo_version"))


def random_states(model_path, n, seed, steps=(5, 120), ctrl_hi=0.6):
    """Physically plausible states: roll the oracle forward from the model's initial pose under random
    constant controls. Returns float32 arrays (qpos, qvel, act, ctrl)."""
    from oracle import oracle

    om, od = oracle.load(model_path)
    rng = np.random.default_rng(seed)
    out = []
    init = np.array(om.qpos0).copy()
    if om.nq == 37:      # baoding: reference init pose (/root/reference/src/envs/baoding.py:400-401)
        init[:23] = 0
        init[0] = -1.57
    if om.nq == 30:      # die: reference init pose (/root/reference/src/envs/reorient.py:123-124)
        init[:23] = 0
        init[0] = -1.5
    for _ in range(n):
        od.reset()
        od.qpos[:] = init
        od.ctrl[:] = rng.uniform(0, ctrl_hi, om.nu)
        od.step(int(rng.integers(*steps)))
        out.append([np.array(od.qpos), np.array(od.qvel), np.array(od.act), np.array(od.ctrl)])
    return [np.array([o[k] for o in out], np.float32) for k in range(4)]

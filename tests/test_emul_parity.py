"""CPU-side check of the kernel *logic*: the product's kernel sources compiled for the host, one lane per
world (tests/emul), against the oracle. The GPU run of the same checks is tests/test_gpu_parity.py."""
import pytest

from conftest import ELBOW, FINGER, HAND_BAODING, HAND_DIE, HAND_POSE
from myochallenge_b200 import _capi
import parity_common as pc

CASES = [("elbow", ELBOW, _capi.TASK_POSE, 8), ("finger", FINGER, _capi.TASK_POSE, 24),
         ("hand_pose", HAND_POSE, _capi.TASK_POSE, 6), ("baoding", HAND_BAODING, _capi.TASK_BAODING, 10),
         ("die", HAND_DIE, _capi.TASK_REORIENT, 10)]


@pytest.mark.parametrize("name,path,kind,n", CASES, ids=[c[0] for c in CASES])
def test_one_step_state_and_contact_parity(emul_lib, name, path, kind, n):
    pc.check_one_step(emul_lib, "cpu", path, kind, n, seed=11)


@pytest.mark.parametrize("name,path,kind,n", CASES, ids=[c[0] for c in CASES])
def test_stage_parity(emul_lib, name, path, kind, n):
    pc.check_stages(emul_lib, "cpu", path, kind, n, seed=12)


@pytest.mark.parametrize("name,path,kind,n", [CASES[1], CASES[3]], ids=["finger", "baoding"])
def test_env_step(emul_lib, name, path, kind, n):
    pc.check_env_step_matches_mj_steps(emul_lib, "cpu", path, kind, min(n, 6))


def test_episode_returns_small(emul_lib):
    """Logic check of the multi-step return comparison on the host build (6 worlds, 10 env steps); the statistical version runs on the GPU."""
    out = pc.check_episode_returns(emul_lib, "cpu", 6, 10, min_same_length=0.8)
    assert out["same_length"] >= 0.8


def test_more_contacts_than_the_fast_layout_holds(emul_lib):
    pc.check_many_contacts(emul_lib, "cpu", n=6)

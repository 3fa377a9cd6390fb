"""CPU: the host build of the kernel sources (tests/emul) against the fixtures recorded from the reference's own env code
(see tests/env_fixture_checks.py). The same checks run on the GPU in tests/test_gpu_env_fixtures.py."""
import pytest

import env_fixture_checks as fx

STEP_TAGS = ["p2_default", "p2_cur32", "p2_rsi", "p1_cur02", "p2_drop", "finger_random", "elbow_random", "hand_random", "die_p2", "die_drop"]
RESET_TAGS_CPU = ["cur02", "cur17", "cur26", "cur32", "p2_default", "p1_default", "p2_knobs", "p2_fixed_task", "p1_noise", "finger_random",
                  "finger_fixed", "elbow_random", "hand_random", "hand_fixed", "finger_distance", "die_p1", "die_p2", "die_axes", "elbow_sds", "finger_sds0", "elbow_weight"]


@pytest.mark.parametrize("tag", STEP_TAGS)
def test_env_step_reproduces_the_reference_run(emul_lib, tag):
    fx.check_step_cases(emul_lib, "cpu", tag)


@pytest.mark.parametrize("tag", RESET_TAGS_CPU)
def test_reset_reproduces_the_reference_distribution(emul_lib, tag):
    fx.check_reset_distribution(emul_lib, "cpu", tag, n=768)

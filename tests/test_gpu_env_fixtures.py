"""GPU: the CUDA kernels through the C ABI against the fixtures recorded from the reference's own env code
(tests/golden/make_env_fixtures.py; checks in tests/env_fixture_checks.py): every curriculum step's reset distribution, and
the deterministic env-step cases."""
import pytest

import env_fixture_checks as fx
from test_env_fixtures_emul import STEP_TAGS

pytestmark = pytest.mark.gpu

RESET_TAGS = ["cur%02d" % i for i in range(1, 33)] + ["p2_default", "p1_default", "p2_knobs", "p2_fixed_task", "p1_noise", "finger_random",
                                                       "finger_fixed", "elbow_random", "hand_random", "hand_fixed", "finger_distance", "die_p1", "die_p2", "die_axes", "elbow_sds", "finger_sds0", "elbow_weight"]


@pytest.mark.parametrize("tag", STEP_TAGS)
def test_env_step_reproduces_the_reference_run(product_lib, tag):
    fx.check_step_cases(product_lib, "cuda:0", tag)


@pytest.mark.parametrize("tag", RESET_TAGS)
def test_reset_reproduces_the_reference_distribution(product_lib, tag):
    fx.check_reset_distribution(product_lib, "cuda:0", tag, n=4096)

"""Pins the oracle (oracle/myo_oracle.c) against the known answers MuJoCo 2.1.0 itself stored in the
reference's shipped model binaries: dof_M0, dof_invweight0, body_invweight0,
body_subtreemass, tendon_length0, tendon_invweight0, actuator_length0, actuator_acc0 -- i.e.
kinematics, tendon wrapping (sphere / cylinder / pulley), moment arms, CRB mass matrix and M^-1."""
import numpy as np
import pytest

from conftest import FINGER, MOTOR_FINGER, MYO_LOAD
from oracle import mjb, oracle

DERIVED = ["dof_M0", "dof_invweight0", "body_invweight0", "body_subtreemass", "tendon_length0", "tendon_invweight0",
           "actuator_length0", "actuator_acc0"]


@pytest.mark.parametrize("path", [FINGER, MOTOR_FINGER, MYO_LOAD])
def test_set_const_reproduces_mujoco_constants(path):
    m, d = oracle.load(path)
    ref = {k: np.array(getattr(m, k)).copy() for k in DERIVED}
    for k in DERIVED:               # wipe, so a pass cannot come from the stored values
        getattr(m, k)[...] = 0
    d.call("o_set_const")
    for k in DERIVED:
        got = np.array(getattr(m, k))
        np.testing.assert_allclose(got, ref[k], rtol=1e-9, atol=1e-12, err_msg=k)


def test_finger_printed_digits():
    """The values quoted in SURVEY.md 8c (read from the MJB with an independent parser)."""
    m, d = oracle.load(FINGER)
    for k in DERIVED:
        getattr(m, k)[...] = 0
    d.call("o_set_const")
    np.testing.assert_allclose(m.dof_M0, [0.01507419, 0.01507618, 0.01113436, 0.0100749], atol=5e-9)
    np.testing.assert_allclose(m.tendon_length0, [0.192135, 0.181620, 0.262120, 0.040868, 0.040868], atol=5e-7)
    np.testing.assert_allclose(m.actuator_acc0, [2.2572, 1.5921, 1.5921, 2.2246, 2.7664], atol=5e-5)
    np.testing.assert_allclose(m.dof_invweight0, [66.3385, 68.2750, 92.4084, 99.3609], atol=5e-5)


@pytest.mark.parametrize("path", [FINGER, MOTOR_FINGER, MYO_LOAD])
def test_mjb_roundtrip_is_byte_exact(path):
    raw = open(path, "rb").read()
    assert mjb.dump(mjb.load(raw)) == raw


def test_moment_arm_is_length_gradient():
    """ten_J = d ten_length / d qpos (finite differences) -- a size-independent property of mj_tendon."""
    m, d = oracle.load(FINGER)
    rng = np.random.default_rng(0)
    for _ in range(5):
        q = rng.uniform([-0.3, -0.3, 0.1, 0.1], [0.3, 0.9, 0.9, 0.9])
        d.qpos[:] = q
        d.call("o_fwd_position")
        J = np.array(d.ten_J).copy()
        for k in range(4):
            e = np.zeros(4); e[k] = 1e-6
            d.qpos[:] = q + e; d.call("o_fwd_position"); lp = np.array(d.ten_length).copy()
            d.qpos[:] = q - e; d.call("o_fwd_position"); lm = np.array(d.ten_length).copy()
            np.testing.assert_allclose((lp - lm) / 2e-6, J[:, k], atol=2e-6)


def test_muscle_curves_known_points():
    """MuJoCo 2.1.0 mju_muscleGain / Bias / Dynamics at points whose values follow from the published
    piecewise definitions."""
    L = oracle.lib()
    import ctypes
    prm = (ctypes.c_double * 10)(0.75, 1.05, 100.0, 200.0, 0.5, 1.6, 1.5, 1.3, 1.2, 0)
    lr = (ctypes.c_double * 2)(0.0, 0.3)      # L0 = 1, L = 0.75 + len
    # optimal length, zero velocity: FL = 1, FV = 1 -> gain = -F0
    assert L.o_muscle_gain(0.25, 0.0, lr, 1.0, prm) == pytest.approx(-100.0)
    # V <= -1 -> FV = 0
    assert L.o_muscle_gain(0.25, -1.5 * 1.0, lr, 1.0, prm) == 0.0
    # lengthening plateau: FV = fvmax
    assert L.o_muscle_gain(0.25, 10.0, lr, 1.0, prm) == pytest.approx(-120.0)
    # passive force: zero up to L = 1, fpmax*0.5 at L = b = 1.3
    assert L.o_muscle_bias(0.20, lr, 1.0, prm) == 0.0
    assert L.o_muscle_bias(0.55, lr, 1.0, prm) == pytest.approx(-100.0 * 1.3 * 0.5)
    # force < 0 -> scale / acc0
    prm2 = (ctypes.c_double * 10)(0.75, 1.05, -1.0, 200.0, 0.5, 1.6, 1.5, 1.3, 1.2, 0)
    assert L.o_muscle_gain(0.25, 0.0, lr, 4.0, prm2) == pytest.approx(-50.0)
    # activation dynamics: tau_act*(0.5+1.5 act) when ctrl > act
    dyn = (ctypes.c_double * 10)(0.01, 0.04, 0, 0, 0, 0, 0, 0, 0, 0)
    assert L.o_muscle_dynamics(1.0, 0.0, dyn) == pytest.approx(1.0 / (0.01 * 0.5))
    assert L.o_muscle_dynamics(0.0, 1.0, dyn) == pytest.approx(-1.0 / (0.04 / 2.0))


def test_energy_is_conserved_without_dissipation():
    """Property test of the dynamics stages the MJB constants cannot pin: with damping, actuation, limits and
    contacts off, kinetic + potential energy drifts only at the integrator's O(h) rate."""
    m, d = oracle.load(MOTOR_FINGER)
    m.dof_damping[:] = 0
    m.jnt_limited[:] = 0
    m.tendon_limited[:] = 0
    m.geom_contype[:] = 0
    m.geom_conaffinity[:] = 0
    d.qpos[:] = [0.1, 0.6, 0.5, 0.3]

    def energy():
        d.call("o_fwd_position")
        M = np.array(d.Mdense).reshape(4, 4)
        ke = 0.5 * d.qvel @ M @ d.qvel
        pe = sum(m.body_mass[b] * 9.81 * d.xipos[b, 2] for b in range(1, m.nbody))
        return ke + pe

    e0 = energy()
    ke_max = 0.0
    for _ in range(200):
        d.step(1)
        M = np.array(d.Mdense).reshape(4, 4)
        ke_max = max(ke_max, 0.5 * d.qvel @ M @ d.qvel)
    assert ke_max > 1e-4                                # something moved
    assert abs(energy() - e0) < 0.05 * ke_max           # semi-implicit Euler drift stays a few % of the exchange


def test_lengthrange_brackets_the_tendon_lengths_over_the_joint_box():
    """SURVEY.md 8c (weaker fixture): MuJoCo stored ``actuator_lengthrange`` = the shortest / longest muscle length it reached by
    simulation (pulling each muscle until equilibrium, so slightly past the soft joint limits and not necessarily at the global
    extremum over the other joints). The oracle's tendon path over the whole joint box (corners + 3000 random poses, i.e. with
    the sphere / cylinder wraps and pulleys engaged far from qpos0) must reproduce those extremes: within 1.5 % of each bound."""
    import itertools

    m, d = oracle.load(FINGER)
    lr = np.array(m.actuator_lengthrange).reshape(-1, 2)
    jr = np.array(m.jnt_range).reshape(-1, 2)
    rng = np.random.default_rng(0)
    pts = [np.array(c) for c in itertools.product(*[(a, b) for a, b in jr])] + [rng.uniform(jr[:, 0], jr[:, 1]) for _ in range(3000)]
    mn, mx = np.full(lr.shape[0], np.inf), np.full(lr.shape[0], -np.inf)
    for q in pts:
        d.qpos[:] = q
        d.call("o_fwd_position")
        L = np.array(d.actuator_length)
        mn, mx = np.minimum(mn, L), np.maximum(mx, L)
    np.testing.assert_allclose(mn, lr[:, 0], rtol=1.5e-2)
    np.testing.assert_allclose(mx, lr[:, 1], rtol=1.5e-2)
    assert np.abs(mn[1:] / lr[1:, 0] - 1).max() < 5e-4        # four of the five lower bounds are reproduced to 0.05 %

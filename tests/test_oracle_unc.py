"""Pins the oracle's sensor / actuator arithmetic to the reference's OWN simple_uncertainty_models.hpp, compiled unmodified
into oracle/_ref/libref_unc.so against a stand-in for the one un-vendored name it needs (arc_helpers::TruncatedNormalDistribution,
oracle/ref_standins): given the same standard draw z (in standard deviations, |z| <= 2), the reference's
TruncatedNormalUncertainSensor::GetSensorValue (:38-45) and TruncatedNormalUncertainVelocityActuator::GetControlValue (:68-88)
must return, bit for bit, what the oracle returns for the tape draw the specification derives from z:
sensor draw = z * (max_sensor_noise / 2), actuator draw = z / 2.  On the GPU box oracle/_ref travels prebuilt."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref(oracle):
    lib = oracle.load_ref_unc()
    if lib is None:
        pytest.skip("oracle/_ref/libref_unc.so not built (reference tree absent)")
    return lib


def _ref_sensor(ref, lower, upper, process, z):
    out = np.zeros_like(process)
    ref.ref_sensor_values(lower, upper, process.size, process.ctypes.data, z.ctypes.data, out.ctypes.data)
    return out


def _ref_actuator(ref, noise_bound, limit, inputs, z):
    out, noiseless = np.zeros_like(inputs), np.zeros_like(inputs)
    ref.ref_actuator_values(noise_bound, limit, inputs.size, inputs.ctypes.data, z.ctypes.data, out.ctypes.data, noiseless.ctypes.data)
    return out, noiseless


def test_sensor_matches_reference_header(oracle, ref):
    rng = np.random.default_rng(11)
    n = 20000
    process = np.ascontiguousarray(rng.normal(scale=3.0, size=n))
    z = np.ascontiguousarray(np.clip(rng.normal(size=n), -2.0, 2.0))
    z[:4] = [-2.0, 2.0, 0.0, -0.0]
    for bound in (0.01, 0.002, 0.37, 1.0, 0.0):
        # the robots construct their sensors as (-max_sensor_noise, +max_sensor_noise): simple_robot_models.hpp:139-141, 566-571, 1277
        want = _ref_sensor(ref, -bound, bound, process, z)
        got = oracle.sensor_values(process, z * (bound * 0.5))
        assert np.array_equal(got, want), bound
        assert np.max(np.abs(want - process)) <= bound + 1e-14


def test_actuator_matches_reference_header(oracle, ref):
    rng = np.random.default_rng(12)
    n = 20000
    inputs = np.ascontiguousarray(rng.normal(scale=1.5, size=n))
    z = np.ascontiguousarray(np.clip(rng.normal(size=n), -2.0, 2.0))
    z[:4] = [-2.0, 2.0, 0.0, -0.0]
    for noise_bound, limit in ((0.1, 1.0), (0.3, 0.9), (-0.25, -2.0), (0.0, 1.0), (0.15, 0.0), (1.0, 1e-3)):
        inputs[4:10] = [limit, -limit, abs(limit), -abs(limit), np.nextafter(abs(limit), 10.0), -np.nextafter(abs(limit), 10.0)]   # clamp edges
        want, want_noiseless = _ref_actuator(ref, noise_bound, limit, inputs, z)
        got, got_noiseless = oracle.actuator_values(noise_bound, limit, inputs, z * 0.5)
        assert np.array_equal(got_noiseless, want_noiseless), (noise_bound, limit)
        assert np.array_equal(got, want), (noise_bound, limit)
        assert np.max(np.abs(want_noiseless)) <= abs(limit)
        # noise is proportional to the commanded value (simple_uncertainty_models.hpp:84): never more than |bound * command|
        assert np.all(np.abs(want - want_noiseless) <= abs(noise_bound) * np.abs(want_noiseless) * (1 + 1e-15))
    assert ref.ref_actuator_max_velocity_noise(0.1, 2.0) == 0.1 * 2.0


def test_tape_draws_are_the_reference_distributions_values(oracle, ref):
    """the counter-based tape holds exactly the values the reference's distributions assign to the standard draws:
    sensor N(0, bound/2) truncated to +-bound, actuator N(0, 1/2) truncated to +-1"""
    import cases
    sc = cases.se2_pid_sensor_noise()
    tape = oracle.fill_noise_tape_counter(9, sc["robot"], sc["params"], 16)
    d = sc["robot"].desc
    for axis in range(3):
        z = np.ascontiguousarray(np.stack([oracle.standard_draws(9, 0, 16, s, 0, axis) for s in range(sc["params"].num_steps)]).reshape(-1))
        bound = d.axis[axis].max_sensor_noise
        assert np.array_equal(_ref_sensor(ref, -bound, bound, np.zeros_like(z), z), tape[:, 0, axis, :].reshape(-1))
        z1 = np.ascontiguousarray(np.stack([oracle.standard_draws(9, 0, 16, s, 1, axis) for s in range(sc["params"].num_steps)]).reshape(-1))
        lim = d.axis[axis].velocity_limit
        want, _ = _ref_actuator(ref, 1.0, lim, np.full_like(z1, lim), z1)      # noise_bound 1, command = limit: value = lim + draw * lim
        assert np.allclose((want - lim) / lim, tape[:, 1, axis, :].reshape(-1), rtol=0, atol=1e-15)

"""Pins the oracle's PID against the reference's own simple_pid_controller.hpp (the one reference file that
compiles standalone): committed golden vectors produced from it, plus a live comparison when oracle/_ref exists."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _cases():
    with open(os.path.join(HERE, "golden", "pid_golden.json")) as f:
        return json.load(f)["cases"]


def test_known_answer_from_survey(oracle):
    # SimplePIDController(1.0, 0.1, 0.01, 5.0): ComputeFeedbackTerm(0.5, 0.01) then (0.4, 0.01)
    terms = oracle.pid_sequence(1.0, 0.1, 0.01, 5.0, [0.5, 0.4], [0.01, 0.01])
    assert terms[0] == 1.0002499999999999
    assert terms[1] == 0.30070000000000002


def test_against_reference_golden_vectors(oracle):
    for c in _cases():
        errors = np.array([float.fromhex(v) for v in c["errors"]])
        dts = np.array([float.fromhex(v) for v in c["dts"]])
        want = np.array([float.fromhex(v) for v in c["terms"]])
        got = oracle.pid_sequence(c["kp"], c["ki"], c["kd"], c["clamp"], errors, dts)
        assert np.array_equal(got, want), c


def test_against_live_reference_build(oracle):
    if oracle.load_ref_pid() is None:
        pytest.skip("oracle/_ref/libref_pid.so not built (reference tree absent)")
    rng = np.random.default_rng(7)
    for _ in range(50):
        kp, ki, kd, clamp = rng.normal(size=4) * np.array([3.0, 1.0, 0.05, 0.5])
        errors = rng.normal(size=64)
        dts = rng.uniform(0.001, 0.05, size=64)
        assert np.array_equal(oracle.pid_sequence(kp, ki, kd, clamp, errors, dts),
                              oracle.ref_pid_sequence(kp, ki, kd, clamp, errors, dts))

"""The shared deterministic elementary functions (include/upc_math.h) against numpy / mpmath."""
import math

import mpmath
import numpy as np


def _ulp_err(got, want):
    want = np.asarray(want, dtype=np.float64)
    return np.max(np.abs(got - want) / np.maximum(np.spacing(np.abs(want)), 5e-324))


def test_sin_cos_accuracy(oracle):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-20, 20, 20000), rng.uniform(-1e-3, 1e-3, 2000), np.linspace(-math.pi, math.pi, 1001)])
    want_s = np.array([float(mpmath.sin(mpmath.mpf(v))) for v in x[:3000]])
    want_c = np.array([float(mpmath.cos(mpmath.mpf(v))) for v in x[:3000]])
    assert np.max(np.abs(oracle.math_fn(0, x[:3000]) - want_s)) < 2.3e-16
    assert np.max(np.abs(oracle.math_fn(1, x[:3000]) - want_c)) < 2.3e-16
    assert np.max(np.abs(oracle.math_fn(0, x) - np.sin(x))) < 4.5e-16
    assert np.max(np.abs(oracle.math_fn(1, x) - np.cos(x))) < 4.5e-16


def test_atan2_acos_accuracy(oracle):
    rng = np.random.default_rng(2)
    y, x = rng.normal(size=20000), rng.normal(size=20000)
    assert _ulp_err(oracle.math_fn(2, y, x), np.arctan2(y, x)) <= 2.0
    u = rng.uniform(-1, 1, 20000)
    assert np.max(np.abs(oracle.math_fn(3, u) - np.arccos(u))) < 1e-15
    # axes and corners
    pts = np.array([[0.0, 1.0], [0.0, -1.0], [1.0, 0.0], [-1.0, 0.0], [0.0, 0.0], [1.0, 1.0], [-1.0, -1.0]])
    assert np.max(np.abs(oracle.math_fn(2, pts[:, 0], pts[:, 1]) - np.arctan2(pts[:, 0], pts[:, 1]))) < 3e-16
    assert oracle.math_fn(3, np.array([1.0]))[0] == 0.0
    assert abs(oracle.math_fn(3, np.array([-1.0]))[0] - math.pi) < 3e-16


def test_wrap_and_signed_distance(oracle):
    rng = np.random.default_rng(3)
    v = rng.uniform(-30, 30, 10000)
    w = oracle.math_fn(4, v)
    assert np.all(w > -math.pi) and np.all(w <= math.pi)        # simple_robot_models.hpp:115 invariant
    assert np.max(np.abs(np.sin(w) - np.sin(v))) < 1e-13
    inside = rng.uniform(-3.1, 3.1, 1000)
    assert np.array_equal(oracle.math_fn(4, inside), inside)     # untouched when already in range
    a, b = rng.uniform(-10, 10, 10000), rng.uniform(-10, 10, 10000)
    d = oracle.math_fn(5, a, b)
    assert np.all(d > -math.pi) and np.all(d <= math.pi)
    assert np.max(np.abs(np.sin(a + d) - np.sin(b))) < 1e-13
    # antisymmetric up to the +-pi edge
    assert np.array_equal(np.abs(oracle.math_fn(5, b, a)), np.abs(d))


def test_se3_log_exp_round_trip(oracle):
    rng = np.random.default_rng(4)
    for scale in (1e-9, 1e-4, 1e-2, 0.5, 2.0, 3.1):
        for _ in range(40):
            w = rng.normal(size=3)
            w = w / np.linalg.norm(w) * scale * rng.uniform(0.5, 1.0)
            tw = np.concatenate([rng.normal(size=3), w])
            T = oracle.se3_exp(tw)
            R = T[:9].reshape(3, 3)
            assert np.max(np.abs(R @ R.T - np.eye(3))) < 1e-14
            back = oracle.se3_log(T)
            assert np.max(np.abs(back - tw)) < 2e-9 * max(1.0, 1.0 / (math.pi - min(scale, 3.14))), (scale, tw, back)


def test_se3_log_near_pi_branch(oracle):
    from uncertainty_planning_core_b200 import scenarios
    for ang in (math.pi - 5e-4, math.pi - 1e-6, math.pi):
        R = scenarios.rotation_from_axis_angle([0.2, -0.7, 0.4], ang)
        tw = oracle.se3_log(np.concatenate([R.reshape(9), [0.1, 0.2, 0.3]]))
        T = oracle.se3_exp(tw)
        assert np.max(np.abs(T[:9].reshape(3, 3) - R)) < 1e-6
        assert np.max(np.abs(T[9:] - np.array([0.1, 0.2, 0.3]))) < 1e-6

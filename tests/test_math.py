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


# ------------------------------------------------------------------ SE(3) closed forms against an independent implementation (mpmath)

def _mp_twist_matrix(tw):
    import mpmath as mp
    v, w = tw[:3], tw[3:]
    return mp.matrix([[0, -w[2], w[1], v[0]], [w[2], 0, -w[0], v[1]], [-w[1], w[0], 0, v[2]], [0, 0, 0, 0]])


def _twists(rng):
    """rotation angles from 1e-12 up to pi - 1e-3, across every branch threshold of DESIGN.md 2.3 (1e-8, 1e-3, 0.05, pi - 1e-3)"""
    out = []
    for theta in (0.0, 1e-12, 0.9e-8, 1.1e-8, 1e-6, 0.9e-3, 1.1e-3, 0.01, 0.049, 0.051, 0.3, 1.0, 2.0, 3.0, math.pi - 0.05, math.pi - 2e-3):
        for _ in range(6):
            axis = rng.normal(size=3)
            axis /= np.linalg.norm(axis)
            out.append(np.concatenate([rng.normal(size=3), axis * theta]))
    return out


def test_se3_exp_matches_mpmath_expm(oracle):
    import mpmath as mp
    mp.mp.dps = 50
    rng = np.random.default_rng(21)
    for tw in _twists(rng):
        want = mp.expm(_mp_twist_matrix([mp.mpf(float(x)) for x in tw]), method="taylor")
        got = oracle.se3_exp(tw)
        R, t = got[:9].reshape(3, 3), got[9:]
        err = max(max(abs(mp.mpf(float(R[i, j])) - want[i, j]) for i in range(3) for j in range(3)),
                  max(abs(mp.mpf(float(t[i])) - want[i, 3]) for i in range(3)))
        assert err < 1e-13 * max(1.0, float(np.linalg.norm(tw[:3]))), (tw, float(err))


def test_se3_log_matches_mpmath(oracle):
    """Log of the exactly exponentiated twist (mpmath, 50 digits, rounded once to double) gives the twist back to 1e-13 (scaled by
    the conditioning of the logarithm, 1 / (pi - theta), near pi)"""
    import mpmath as mp
    mp.mp.dps = 50
    rng = np.random.default_rng(22)
    for tw in _twists(rng):
        T = mp.expm(_mp_twist_matrix([mp.mpf(float(x)) for x in tw]), method="taylor")
        cfg = np.array([float(T[i, j]) for i in range(3) for j in range(3)] + [float(T[i, 3]) for i in range(3)])
        got = oracle.se3_log(cfg)
        theta = float(np.linalg.norm(tw[3:]))
        tol = 1e-13 * max(1.0, float(np.linalg.norm(tw[:3]))) * max(1.0, 1.0 / (math.pi - theta))
        assert np.max(np.abs(got - tw)) < tol, (tw, got)


def test_se3_log_near_pi_matches_mpmath(oracle):
    """theta within 1e-3 of pi (the symmetric-part branch of DESIGN.md 2.3): the returned twist has |omega| <= pi and its exact
    exponential (mpmath, 50 digits) reproduces the input transform - which characterises the principal logarithm"""
    import mpmath as mp
    from uncertainty_planning_core_b200 import scenarios
    mp.mp.dps = 50
    for ang in (math.pi - 9e-4, math.pi - 1e-4, math.pi - 1e-5, math.pi - 1e-7):
        for axis in ([0.2, -0.7, 0.4], [1.0, 0.0, 0.0], [-0.3, 0.1, 0.9], [0.0, 0.0, -1.0]):
            R = scenarios.rotation_from_axis_angle(axis, ang)
            t = np.array([0.1, -0.2, 0.3])
            got = oracle.se3_log(np.concatenate([R.reshape(9), t]))
            theta = float(np.linalg.norm(got[3:]))
            assert abs(theta - ang) < 1e-9 and theta <= math.pi
            T = mp.expm(_mp_twist_matrix([mp.mpf(float(x)) for x in got]), method="taylor")
            err_R = max(abs(T[i, j] - mp.mpf(float(R[i, j]))) for i in range(3) for j in range(3))
            err_t = max(abs(T[i, 3] - mp.mpf(float(t[i]))) for i in range(3))
            # the rotation is reproduced to the accuracy of the input matrix (built in double: ~1e-16 / sin(theta) on the axis)
            assert err_R < 1e-9 and err_t < 1e-9, (ang, axis, float(err_R), float(err_t))


def test_quaternion_and_se3_distance_match_mpmath(oracle):
    import mpmath as mp
    from uncertainty_planning_core_b200 import abi, scenarios
    import cases
    mp.mp.dps = 40
    rng = np.random.default_rng(23)
    robot = cases.cl_se3_blobs()["robot"]
    cfgs = []
    for ang in (0.0, 1e-9, 1e-4, 0.3, 1.5, 2.5, 3.0, 3.14, math.pi - 1e-7):
        for _ in range(4):
            R = scenarios.rotation_from_axis_angle(rng.normal(size=3), ang)
            cfgs.append(scenarios.se3_config(R, rng.normal(size=3)))
    cfgs = np.array(cfgs)
    for c in cfgs:
        q = oracle.quaternion_from_rotation(c)
        x, y, z, w = q
        assert abs(np.linalg.norm(q) - 1.0) < 1e-14
        Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.max(np.abs(Rq - c[:9].reshape(3, 3))) < 1e-14
    D = oracle.pair_distances(robot, cfgs)
    for i in range(0, len(cfgs), 3):
        for j in range(len(cfgs)):
            Ri, Rj = mp.matrix(cfgs[i][:9].reshape(3, 3).tolist()), mp.matrix(cfgs[j][:9].reshape(3, 3).tolist())
            rel = Ri.T * Rj
            c = (rel[0, 0] + rel[1, 1] + rel[2, 2] - 1) / 2
            c = max(mp.mpf(-1), min(mp.mpf(1), c))
            # angle through atan2 of the skew part for accuracy near 0 and pi
            s = mp.sqrt((rel[2, 1] - rel[1, 2]) ** 2 + (rel[0, 2] - rel[2, 0]) ** 2 + (rel[1, 0] - rel[0, 1]) ** 2) / 2
            angle = mp.atan2(s, c)
            dt = mp.sqrt(sum((mp.mpf(float(cfgs[i][9 + k])) - mp.mpf(float(cfgs[j][9 + k]))) ** 2 for k in range(3)))
            want = float(dt + angle)
            # acos(2 dq^2 - 1) loses half the digits near zero angle (at most 2e-8 absolute)
            assert abs(D[i, j] - want) < 3e-8, (i, j, D[i, j], want)
            if float(angle) > 1e-3:
                assert abs(D[i, j] - want) < 1e-12 * max(1.0, 1.0 / max(float(mp.pi - angle), 1e-9) ** 0.5), (i, j, D[i, j], want)

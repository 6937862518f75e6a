"""Pins the oracle's robot models to the reference's OWN simple_robot_models.hpp (SimpleSE2Robot :92-431, SimpleSE3Robot :507-875,
SimpleLinkedRobot :1274-2218), compiled unmodified into oracle/_ref/libref_rob.so against the stand-in Eigen / EigenHelpers headers
of oracle/ref_standins (neither library is vendored by the reference or installed here).  The stand-ins are written from the
libraries' documented semantics with libm - independently of the oracle's closed forms and of include/upc_math.h - so both sides
compute the same quantities through different arithmetic: agreement is to rounding (tolerances below), not bit for bit.

What this pins: the order and use of the draws, which PID / sensor / actuator serves which axis, the clamp-then-noise order, angle
wrapping of SE(2) headings and continuous joints, limits applied to sensed and commanded joint values, right-multiplication by the
exponential of the noisy twist, forward kinematics joint by joint, the point Jacobians (including the prismatic column of
simple_robot_models.hpp:2064), the configuration distances the clustering uses and the per-dimension distances / averages the
per-cluster statistics use.  On the GPU box oracle/_ref travels prebuilt."""
import math

import numpy as np
import pytest

import cases
from uncertainty_planning_core_b200 import abi, scenarios


@pytest.fixture(scope="module")
def ref(oracle):
    lib = oracle.load_ref_rob()
    if lib is None:
        pytest.skip("oracle/_ref/libref_rob.so not built (reference tree absent)")
    return lib


def _draws(rng, n_seq, n_steps, dof):
    z = np.clip(rng.normal(size=(n_seq, n_steps, 2, dof)), -2.0, 2.0)
    z[0, 0] = 2.0
    z[0, 1] = -2.0
    return z


def _se3(rng, spread=1.0, max_angle=2.5):
    axis = rng.normal(size=3)
    return scenarios.se3_config(scenarios.rotation_from_axis_angle(axis, rng.uniform(0.0, max_angle)), rng.uniform(-spread, spread, size=3))


def _robots():
    """(name, robot, random configuration generator)"""
    se2 = scenarios.se2_robot(kp=4.0, velocity_limit=0.8, r_velocity_limit=0.6, max_sensor_noise=0.01, max_actuator_noise=0.2, lattice=3)
    se2_pid = scenarios.se2_robot(kp=2.0, max_sensor_noise=0.02, max_actuator_noise=0.1, lattice=2)
    for a in range(3):      # integral and derivative terms, integral clamp that binds (simple_pid_controller.hpp:122-135)
        se2_pid.desc.axis[a].ki, se2_pid.desc.axis[a].kd, se2_pid.desc.axis[a].integral_clamp = 0.8, 0.01, 0.002
    se3 = scenarios.se3_robot(kp=3.0, velocity_limit=0.7, r_velocity_limit=0.9, max_sensor_noise=0.005, max_actuator_noise=0.15, lattice=(2, 2, 2))
    se3_quiet = scenarios.se3_robot(kp=3.0, max_sensor_noise=0.0, max_actuator_noise=0.25, lattice=(2, 2, 2))
    arm = scenarios.linked_arm(num_joints=7, points_per_link=2, kp=4.0, velocity_limit=0.9, max_sensor_noise=0.01, max_actuator_noise=0.2, limit=1.2)
    J = abi
    types = [J.JOINT_CONTINUOUS, J.JOINT_REVOLUTE, J.JOINT_FIXED, J.JOINT_PRISMATIC, J.JOINT_REVOLUTE, J.JOINT_CONTINUOUS, J.JOINT_REVOLUTE]
    mixed = scenarios.linked_arm(joint_types=types, points_per_link=2, limit=0.25, weights=[1.0, 2.0, 0.5, -1.5, 1.0, 0.25],
                                 max_sensor_noise=0.02, max_actuator_noise=0.3)

    def se2_cfg(rng):
        return np.array([rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-math.pi, math.pi)])

    def arm_cfg(rng):
        return rng.uniform(-1.1, 1.1, size=7)

    def mixed_cfg(rng):
        q = rng.uniform(-0.24, 0.24, size=6)
        q[0], q[4] = rng.uniform(-3.1, 3.1), rng.uniform(-3.1, 3.1)      # the continuous joints (active indices 0 and 4)
        return q

    return [("se2", se2, se2_cfg), ("se2_pid", se2_pid, se2_cfg), ("se3", se3, _se3), ("se3_noise_free_sensors", se3_quiet, _se3),
            ("arm7", arm, arm_cfg), ("arm_mixed_joints", mixed, mixed_cfg)]


ROBOTS = _robots()


@pytest.mark.parametrize("name,robot,gen", ROBOTS, ids=[r[0] for r in ROBOTS])
def test_control_and_apply_sequences_match_reference_header(oracle, ref, name, robot, gen):
    """sense -> PID -> clamp (GenerateControlAction :239-259 / :685-714 / :1823-1851) then noisy apply (ApplyControlInput :282-295 /
    :746-763 / :1882-1911), twelve controller intervals per edge, with and without noise"""
    rng = np.random.default_rng(101)
    n_seq, n_steps, dt = 48, 12, 0.05
    starts = np.stack([gen(rng) for _ in range(n_seq)])
    targets = np.stack([gen(rng) for _ in range(n_seq)])
    if name.startswith("se2"):
        starts[0, 2], targets[0, 2] = 3.1, -3.1                         # the short way round crosses +-pi
        starts[1, 2], targets[1, 2] = -3.13, 3.0
    if robot.desc.kind == abi.ROBOT_SE3:
        # the SE(3) logarithm is ill-conditioned at a half turn; both sides use different closed forms there, so stay below 2.5 rad
        targets = np.stack([_se3(rng, max_angle=0.8) for _ in range(n_seq)])
        starts = np.stack([_se3(rng, max_angle=0.8) for _ in range(n_seq)])
    z = _draws(rng, n_seq, n_steps, robot.desc.dof)
    for draws in (None, z):
        want_u, want_c = oracle.rob_sequence(robot, starts, targets, dt, n_steps, draws, reference=True)
        got_u, got_c = oracle.rob_sequence(robot, starts, targets, dt, n_steps, draws)
        assert np.isfinite(want_u).all() and np.isfinite(want_c).all()
        assert np.abs(want_u).max() > 0.1
        assert np.allclose(got_u, want_u, rtol=0, atol=5e-12), np.abs(got_u - want_u).max()
        assert np.allclose(got_c, want_c, rtol=0, atol=1e-12), np.abs(got_c - want_c).max()
    # noise is really applied: the noisy and the noise-free trajectories differ
    assert np.abs(oracle.rob_sequence(robot, starts, targets, dt, n_steps, z)[1] - oracle.rob_sequence(robot, starts, targets, dt, n_steps)[1]).max() > 1e-4


def test_commands_saturate_and_noise_scales_with_the_command(oracle, ref):
    """far targets: every command sits on its velocity limit after the clamp (simple_uncertainty_models.hpp:68-75), and the applied
    motion is limit * dt * (1 + draw * noise_bound) with the draw in [-1, 1] (:77-88)"""
    robot = ROBOTS[0][1]
    rng = np.random.default_rng(5)
    starts = np.zeros((16, 3))
    targets = np.tile(np.array([50.0, -50.0, 0.0]), (16, 1))
    z = _draws(rng, 16, 4, 3)
    for reference in (True, False):
        u, c = oracle.rob_sequence(robot, starts, targets, 0.05, 4, z, reference=reference)
        assert np.all(u[:, :, 0] == 0.8) and np.all(u[:, :, 1] == -0.8)
        step0 = c[:, 0, 0]
        assert np.allclose(step0, 0.8 * 0.05 * (1.0 + 0.5 * z[:, 0, 1, 0] * 0.2), rtol=0, atol=1e-15)


@pytest.mark.parametrize("name,robot,gen", ROBOTS, ids=[r[0] for r in ROBOTS])
def test_kinematics_and_point_jacobians_match_reference_header(oracle, ref, name, robot, gen):
    """link transforms (ComputePose :363-371, UpdateTransforms :1299-1343) and ComputeLinkPointJacobian of every body point
    (:334-361, :813-832, :2011-2084)"""
    rng = np.random.default_rng(202)
    configs = np.stack([gen(rng) for _ in range(64)])
    want_t, want_j = oracle.rob_kinematics(robot, configs, reference=True)
    got_t, got_j = oracle.rob_kinematics(robot, configs)
    assert np.allclose(got_t, want_t, rtol=0, atol=5e-15), np.abs(got_t - want_t).max()
    assert np.allclose(got_j, want_j, rtol=0, atol=5e-15), np.abs(got_j - want_j).max()
    assert np.abs(want_j).max() > 0.01
    if name == "arm_mixed_joints":
        # the prismatic joint's column applies the WHOLE child transform to the axis, translation included (:2064) - reproduced, not fixed
        d = robot.desc
        prismatic_active = 2
        last_point = d.num_points - 1
        col = want_j[:, last_point, :, prismatic_active]
        assert np.abs(np.linalg.norm(col, axis=1) - 1.0).max() > 0.1


@pytest.mark.parametrize("name,robot,gen", ROBOTS, ids=[r[0] for r in ROBOTS])
def test_configuration_distances_match_reference_header(oracle, ref, name, robot, gen):
    """ComputeConfigurationDistance (:424-430, :871-874, :2214-2217: what both clustering passes and the goal tests use) and
    ComputePerDimensionConfigurationDistance (:419-422, :866-869, :2206-2209: what the per-cluster variances use)"""
    rng = np.random.default_rng(303)
    a = np.stack([gen(rng) for _ in range(400)])
    b = np.stack([gen(rng) for _ in range(400)])
    b[:20] = a[:20]                                                     # identical configurations: distance exactly 0 on both sides
    if robot.desc.kind == abi.ROBOT_SE3:
        # near-identical rotations: arccos(2 dq^2 - 1) loses half the digits near dq = 1 on both sides
        for i in range(20, 60):
            a[i] = _se3(rng)
            b[i] = a[i].copy()
            b[i][9:] += rng.uniform(-0.05, 0.05, size=3)
    want_d, want_dims = oracle.rob_distances(robot, a, b, reference=True)
    got_d, got_dims = oracle.rob_distances(robot, a, b)
    tol = 5e-8 if robot.desc.kind == abi.ROBOT_SE3 else 1e-14
    assert np.allclose(got_d, want_d, rtol=0, atol=tol), np.abs(got_d - want_d).max()
    assert np.allclose(got_dims, want_dims, rtol=0, atol=tol), np.abs(got_dims - want_dims).max()
    if robot.desc.kind == abi.ROBOT_SE3:
        # a rotation matrix that is orthonormal only to rounding gives |q1.q2| a few ulps below 1, and arccos(2 dq^2 - 1) turns that
        # into ~1e-7 - on both sides (EigenHelpers::Distance has the same 1 - epsilon guard)
        assert got_d[:20].max() < 2e-7 and want_d[:20].max() < 2e-7
    else:
        assert np.all(got_d[:20] == 0.0) and np.all(want_d[:20] == 0.0)
    assert want_d[60:].min() > 1e-3
    if robot.desc.kind == abi.ROBOT_SE3:
        # away from the arccosine's flat spot the two agree to rounding
        far = want_dims[:, 3] > 0.05
        assert far.sum() > 100
        assert np.allclose(got_d[far], want_d[far], rtol=0, atol=1e-13)


@pytest.mark.parametrize("name,robot,gen", ROBOTS, ids=[r[0] for r in ROBOTS])
def test_average_configuration_matches_reference_header(oracle, ref, name, robot, gen):
    """AverageConfigurations (:373-399, :834-844, :2093-2148) of a tight cluster: the expectation of a child state"""
    rng = np.random.default_rng(404)
    centre = gen(rng)
    if robot.desc.kind == abi.ROBOT_SE3:
        members = []
        for _ in range(33):
            m = centre.copy()
            R = m[:9].reshape(3, 3) @ scenarios.rotation_from_axis_angle(rng.normal(size=3), rng.uniform(0, 0.05))
            members.append(scenarios.se3_config(R, m[9:] + rng.normal(scale=0.01, size=3)))
        members = np.stack(members)
    else:
        members = centre[None, :] + rng.normal(scale=0.01, size=(33, centre.size))
        if name.startswith("se2"):
            members[:, 2] = 3.14 + rng.normal(scale=0.01, size=33)       # heading cluster straddling +-pi: circular mean
    want = oracle.rob_average(robot, members, reference=True)
    got = oracle.rob_average(robot, members)
    assert np.allclose(got, want, rtol=0, atol=1e-13), np.abs(got - want).max()


@pytest.mark.parametrize("name,robot,gen", ROBOTS, ids=[r[0] for r in ROBOTS])
def test_child_state_statistics_match_reference_planner_state(oracle, ref, name, robot, gen):
    """The per-cluster statistics (next row (f)-1): the reference's OWN uncertainty_planner_state.hpp, compiled unmodified next to its
    robot models - UncertaintyPlannerState(particles)::UpdateStatistics(robot) (:339-351): ComputeExpectation, ComputeVariance,
    ComputeSpaceIndependentVariance and the two directional forms (:591-714) - against the oracle's cluster statistics (which the
    product's upc_cluster_statistics reproduces to 0 ulp on the GPU) for one cluster: many members, two members, a single member."""
    rng = np.random.default_rng(505)
    centre = gen(rng)
    step_size = 0.35
    for count in (41, 2, 1):
        if robot.desc.kind == abi.ROBOT_SE3:
            members = []
            for _ in range(count):
                R = centre[:9].reshape(3, 3) @ scenarios.rotation_from_axis_angle(rng.normal(size=3), rng.uniform(0.01, 0.2))
                members.append(scenarios.se3_config(R, centre[9:] + rng.normal(scale=0.05, size=3)))
            members = np.stack(members)
        else:
            members = centre[None, :] + rng.normal(scale=0.03, size=(count, centre.size))
            if name.startswith("se2"):
                members[:, 2] = -3.12 + rng.normal(scale=0.03, size=count)     # headings straddling +-pi
            if name == "arm_mixed_joints":
                # joint values are configurations only inside their limits (SimpleJointModel clamps on construction, :903-916);
                # active joints 0 and 4 are continuous
                limited = [1, 2, 3, 5]
                members[:, limited] = np.clip(members[:, limited], -0.25, 0.25)
        want = oracle.ref_state_statistics(robot, members, step_size)
        got = oracle.cluster_statistics(robot, members, np.zeros(count, dtype=np.int32), np.zeros(count, dtype=bool), True, step_size, 1)
        tol = 2e-9 if robot.desc.kind == abi.ROBOT_SE3 else 1e-13          # SE(3): arccos near zero angle, two closed forms of the quaternion
        assert int(got["counts"][0]) == count
        assert np.allclose(got["expectations"][0], want["expectation"], rtol=0, atol=1e-13), (count, np.abs(got["expectations"][0] - want["expectation"]).max())
        assert abs(got["variance"][0] - want["variance"]) <= tol, count
        assert abs(got["si_variance"][0] - want["si_variance"]) <= tol * 10, count
        assert np.allclose(got["variances"][0], want["variances"], rtol=0, atol=tol), count
        assert np.allclose(got["si_variances"][0], want["si_variances"], rtol=0, atol=tol * 10), count
        if count > 2:
            assert want["variance"] > 1e-5 and np.all(want["variances"] >= 0.0)

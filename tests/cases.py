"""Small seeded cases shared by the CPU parity tests (oracle vs emulated device program), the GPU parity
tests (oracle vs CUDA through the C ABI) and tests/golden/make_golden.py.  Sized so the oracle needs
about a second per case."""
import math

import numpy as np

from uncertainty_planning_core_b200 import abi, scenarios

SIM_STAT_KEYS = ("particle_steps", "collided_particles", "resolver_iterations", "failed_resolves")
# seed / first particle of the committed seeded-noise fixtures (tests/golden/sim_seeded_*.npz)
SEEDED_GOLDEN_SEED = 0x9E3779B97F4A7C15
SEEDED_GOLDEN_FIRST = 12345


# ------------------------------------------------------------------------------------------- simulation

def se2_contact():
    return scenarios.c1_se2(seed=11, n=96, num_steps=48, shape=(128, 128, 1), resolution=0.08)


def se2_no_contacts_allowed():
    sc = scenarios.c1_se2(seed=12, n=80, num_steps=48, shape=(128, 128, 1), resolution=0.08)
    sc["params"].allow_contacts = 0
    return sc


def se2_microsteps():
    sc = scenarios.c1_se2(seed=13, n=64, num_steps=24, max_microsteps=3, shape=(128, 128, 1), resolution=0.08)
    sc["robot"] = scenarios.se2_robot(kp=20.0, velocity_limit=6.0, r_velocity_limit=3.0)
    return sc


def se2_pid_sensor_noise():
    sc = scenarios.c1_se2(seed=14, n=77, num_steps=40, shape=(128, 128, 1), resolution=0.08)
    lin = abi.make_axis(kp=4.0, ki=0.8, kd=0.02, integral_clamp=0.05, velocity_limit=0.9, max_sensor_noise=0.01,
                        max_actuator_noise=0.2)
    rot = abi.make_axis(kp=3.0, ki=0.5, kd=0.01, integral_clamp=0.1, velocity_limit=1.2, max_sensor_noise=0.02,
                        max_actuator_noise=0.3)
    base = scenarios.se2_robot()
    sc["robot"] = abi.Robot(abi.ROBOT_SE2, [base.points], [lin, lin, rot])
    return sc


def se2_shortcut_per_target():
    """targets.size() == starts.size() (uncertainty_contact_planning.hpp:1826) + early exit"""
    sc = scenarios.c5_batch(seed=15, edges=9, particles=7, num_steps=64)
    sc["params"].simulation_shortcut_distance = 0.08
    sc["seed"] = 15
    return sc


def se3_contact():
    return scenarios.c2_se3(seed=21, n=64, num_steps=48, shape=(64, 64, 64), resolution=0.0625)


def se3_individual_jacobians():
    sc = scenarios.c2_se3(seed=22, n=48, num_steps=40, shape=(64, 64, 64), resolution=0.0625)
    sc["params"].use_individual_jacobians = 1
    return sc


def se3_sensor_noise_spheres():
    sc = scenarios.c2_se3(seed=23, n=40, num_steps=40, shape=(64, 64, 64), resolution=0.0625)
    sc["robot"] = scenarios.se3_robot(kp=4.0, max_sensor_noise=0.01, max_actuator_noise=0.25, radius=0.03)
    sc["params"].simulation_shortcut_distance = 0.05
    return sc


def se3_large_rotation():
    """start orientation close to pi from identity: exercises the near-pi branch of the SE(3) log"""
    sc = scenarios.c2_se3(seed=24, n=24, num_steps=30, shape=(64, 64, 64), resolution=0.0625)
    R0 = scenarios.rotation_from_axis_angle([0.3, -0.5, 0.8], math.pi - 2.0e-4)
    R1 = scenarios.rotation_from_axis_angle([0.3, -0.5, 0.8], math.pi - 0.2)
    sc["starts"][:, :9] = R0.reshape(9)
    sc["targets"][:, :9] = R1.reshape(9)
    return sc


def _arm_environment(resolution=0.0625, shape=(64, 64, 64)):
    boxes = np.array([[2.45, 2.0, 2.1, 0.2, 0.4, 0.25],
                      [1.55, 2.0, 1.4, 0.15, 0.4, 0.2],
                      [2.0, 2.5, 2.6, 0.4, 0.15, 0.2]])
    return scenarios.make_environment(shape, resolution, boxes), boxes


def linked_contact():
    env, boxes = _arm_environment()
    robot = scenarios.linked_arm()
    n = 48
    q0 = np.array([0.05, 0.02, -0.03, 0.04, 0.0, 0.02, 0.0])
    q1 = np.array([0.1, 0.35, 0.0, 0.3, 0.1, 0.2, 0.0])
    params = abi.make_sim_params(48, 0.01, env.resolution, allow_contacts=True)
    return dict(name="linked_contact", env=env, robot=robot, starts=np.tile(q0, (n, 1)), targets=q1[None, :], params=params,
                seed=31, boxes=boxes)


def linked_mixed_joints():
    """continuous, prismatic and fixed joints, joint limits that bind, non-unit distance weights"""
    env, boxes = _arm_environment()
    J = abi
    types = [J.JOINT_CONTINUOUS, J.JOINT_REVOLUTE, J.JOINT_FIXED, J.JOINT_PRISMATIC, J.JOINT_REVOLUTE, J.JOINT_CONTINUOUS, J.JOINT_REVOLUTE]
    robot = scenarios.linked_arm(joint_types=types, limit=0.25, weights=[1.0, 2.0, 0.5, -1.5, 1.0, 0.25],
                                 max_sensor_noise=0.02, max_actuator_noise=0.3)
    n = 40
    q0 = np.array([3.0, 0.1, 0.05, 0.1, -3.05, 0.0])
    q1 = np.array([-3.0, 0.6, 0.2, 0.5, 3.0, 0.3])
    params = abi.make_sim_params(40, 0.01, env.resolution, allow_contacts=True, max_microsteps=2)
    params.simulation_shortcut_distance = 0.02
    return dict(name="linked_mixed", env=env, robot=robot, starts=np.tile(q0, (n, 1)), targets=q1[None, :], params=params,
                seed=32, boxes=boxes)


def linked_ten_joints():
    """a 10-DoF arm: more than 8 degrees of freedom, so the product runs the full-capacity linked instantiation (16 / 17)"""
    env, boxes = _arm_environment()
    robot = scenarios.linked_arm(num_joints=10, link_length=0.2, points_per_link=8)
    n = 36
    q0 = np.array([0.05, 0.02, -0.03, 0.04, 0.0, 0.02, 0.0, 0.01, -0.02, 0.0])
    q1 = np.array([0.1, 0.35, 0.0, 0.3, 0.1, 0.2, 0.0, 0.15, -0.1, 0.05])
    params = abi.make_sim_params(40, 0.01, env.resolution, allow_contacts=True)
    return dict(name="linked_ten_joints", env=env, robot=robot, starts=np.tile(q0, (n, 1)), targets=q1[None, :], params=params,
                seed=37, boxes=boxes)


GOLDEN_SIM_CASES = {
    "se2_contact": se2_contact,
    "se2_no_contacts_allowed": se2_no_contacts_allowed,
    "se2_microsteps": se2_microsteps,
    "se2_pid_sensor_noise": se2_pid_sensor_noise,
    "se2_shortcut_per_target": se2_shortcut_per_target,
    "se3_contact": se3_contact,
    "se3_individual_jacobians": se3_individual_jacobians,
    "se3_sensor_noise_spheres": se3_sensor_noise_spheres,
    "se3_large_rotation": se3_large_rotation,
    "linked_contact": linked_contact,
    "linked_mixed_joints": linked_mixed_joints,
    "linked_ten_joints": linked_ten_joints,
}


# ------------------------------------------------------------------------------------------- clustering

def _cparams(feature=abi.FEATURE_NONE, sig_thr=0.125, dist_thr=0.1):
    p = abi.ClusterParams()
    p.feature_type = feature
    p.signature_matching_threshold = sig_thr
    p.distance_clustering_threshold = dist_thr
    return p


def _se2_env_robot():
    sc = scenarios.c1_se2(seed=11, n=1, shape=(128, 128, 1), resolution=0.08)
    return sc["env"], sc["robot"]


def cl_se2_blobs():
    env, robot = _se2_env_robot()
    cfg, _ = scenarios.c4_blobs(abi.ROBOT_SE2, 300, k=5, seed=41)
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams())


def cl_se2_single_blob():
    env, robot = _se2_env_robot()
    cfg, _ = scenarios.c4_blobs(abi.ROBOT_SE2, 130, k=1, seed=42)
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams())


def cl_se2_chain():
    """points 0.06 apart on a line, threshold 0.1: one connected component that is not a clique"""
    env, robot = _se2_env_robot()
    n = 70
    cfg = np.stack([1.0 + 0.06 * np.arange(n), np.full(n, 2.0), 0.01 * np.sin(np.arange(n))], axis=1)
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams())


def cl_se2_ties():
    """equally spaced points with exactly representable gaps: every nearest pair ties"""
    env, robot = _se2_env_robot()
    n = 33
    cfg = np.stack([1.0 + 0.125 * np.arange(n), np.full(n, 2.0), np.zeros(n)], axis=1)
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams(dist_thr=0.125))


def cl_se2_random():
    env, robot = _se2_env_robot()
    rng = np.random.default_rng(43)
    cfg = np.stack([rng.uniform(1.0, 1.6, 220), rng.uniform(1.0, 1.5, 220), rng.uniform(-0.2, 0.2, 220)], axis=1)
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams())


def cl_se2_identical():
    env, robot = _se2_env_robot()
    cfg = np.tile(np.array([1.5, 2.5, 3.0]), (45, 1))
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams())


def cl_se2_wrap():
    """angles either side of +-pi are close"""
    env, robot = _se2_env_robot()
    rng = np.random.default_rng(44)
    th = np.where(rng.random(90) < 0.5, math.pi - 0.01, -math.pi + 0.01) + rng.uniform(-0.005, 0.005, 90)
    cfg = np.stack([np.full(90, 1.0), np.full(90, 1.0), th], axis=1)
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams())


def cl_se2_crs():
    """particles straddling convex-segment slab borders: the signature pass splits first"""
    env, robot = _se2_env_robot()
    rng = np.random.default_rng(45)
    n = 160
    # x-slab border of the 128-cell grid at cell 32 (2.56 m), overlap 2 cells
    x = np.where(rng.random(n) < 0.5, 2.0, 3.3) + rng.uniform(-0.03, 0.03, n)
    y = np.where(rng.random(n) < 0.3, 2.0, 3.3) + rng.uniform(-0.03, 0.03, n)
    cfg = np.stack([x, y, rng.uniform(-0.02, 0.02, n)], axis=1)
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams(feature=abi.FEATURE_CRS, sig_thr=0.125, dist_thr=0.1))


def cl_se2_crs_messy():
    env, robot = _se2_env_robot()
    rng = np.random.default_rng(46)
    n = 120
    cfg = np.stack([rng.uniform(2.2, 2.9, n), rng.uniform(2.2, 2.9, n), rng.uniform(-0.3, 0.3, n)], axis=1)
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams(feature=abi.FEATURE_CRS, sig_thr=0.05, dist_thr=0.15))


def cl_se2_batch():
    env, robot = _se2_env_robot()
    rng = np.random.default_rng(47)
    sets = []
    for s in range(7):
        if s % 3 == 0:
            c, _ = scenarios.c4_blobs(abi.ROBOT_SE2, 50, k=1, seed=100 + s)
        elif s % 3 == 1:
            c, _ = scenarios.c4_blobs(abi.ROBOT_SE2, 50, k=3, seed=100 + s)
        else:
            c = np.stack([rng.uniform(1.0, 1.4, 50), rng.uniform(1.0, 1.3, 50), rng.uniform(-0.1, 0.1, 50)], axis=1)
        sets.append(c)
    return dict(env=env, robot=robot, configs=np.concatenate(sets, axis=0), cparams=_cparams(), n_sets=7)


def _se3_env_robot():
    sc = scenarios.c2_se3(seed=21, n=1, shape=(64, 64, 64), resolution=0.0625)
    return sc["env"], sc["robot"]


def cl_se3_blobs():
    env, robot = _se3_env_robot()
    cfg, _ = scenarios.c4_blobs(abi.ROBOT_SE3, 200, k=4, seed=51)
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams())


def cl_se3_random():
    env, robot = _se3_env_robot()
    rng = np.random.default_rng(52)
    n = 150
    cfg = np.zeros((n, 12))
    for i in range(n):
        R = scenarios.rotation_from_axis_angle(rng.normal(size=3), rng.uniform(0.0, 0.12))
        if i % 10 == 0:
            R = scenarios.rotation_from_axis_angle(rng.normal(size=3), rng.uniform(2.9, 3.14))  # negative-trace branch
        cfg[i] = scenarios.se3_config(R, [1.0 + rng.uniform(0, 0.25), 1.0 + rng.uniform(0, 0.2), 1.0])
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams(dist_thr=0.12))


def cl_se3_crs():
    env, robot = _se3_env_robot()
    rng = np.random.default_rng(53)
    n = 96
    cfg = np.zeros((n, 12))
    for i in range(n):
        R = scenarios.rotation_from_axis_angle(rng.normal(size=3), rng.uniform(0.0, 0.05))
        x = (1.7 if rng.random() < 0.5 else 2.6) + rng.uniform(-0.02, 0.02)
        cfg[i] = scenarios.se3_config(R, [x, 1.5 + rng.uniform(-0.02, 0.02), 1.5])
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams(feature=abi.FEATURE_CRS))


def _linked_env_robot(mixed=False):
    sc = linked_mixed_joints() if mixed else linked_contact()
    return sc["env"], sc["robot"]


def cl_linked_blobs():
    env, robot = _linked_env_robot()
    cfg, _ = scenarios.c4_blobs(abi.ROBOT_LINKED, 180, k=4, seed=61, dof=7)
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams())


def cl_linked_random_mixed():
    env, robot = _linked_env_robot(mixed=True)
    rng = np.random.default_rng(62)
    n = 140
    cfg = rng.uniform(-0.05, 0.05, size=(n, 6))
    cfg[:, 0] = np.where(rng.random(n) < 0.5, math.pi - 0.01, -math.pi + 0.01)  # continuous joint wraps
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams(dist_thr=0.08))


def cl_linked_crs():
    env, robot = _linked_env_robot()
    rng = np.random.default_rng(63)
    n = 64
    cfg = rng.uniform(-0.02, 0.02, size=(n, 7))
    cfg[:, 1] += np.where(rng.random(n) < 0.5, 0.0, 0.6)
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams(feature=abi.FEATURE_CRS, sig_thr=0.05, dist_thr=0.1))


def _with_ptpm(case, num_steps=30, goal=0.06, seed=71):
    """attach POINT_TO_POINT_MOVEMENT inputs: simulation parameters and the tape of the n*(n-1) pair simulations"""
    import ctypes as C
    n = case["configs"].shape[0]
    sp = abi.make_sim_params(num_steps, 0.01, case["env"].resolution, allow_contacts=True)
    sp.simulation_shortcut_distance = goal * 0.5
    tape = abi.fill_noise_tape(seed, case["robot"], sp, n * n - n)
    cp = case["cparams"]
    cp.feature_type = abi.FEATURE_PTPM
    cp.goal_distance_threshold = goal
    cp.ptpm_sim = C.pointer(sp)
    cp.ptpm_tape = tape.ctypes.data
    case["ptpm"] = (sp, tape, goal)  # keeps the buffers alive
    return case


def cl_se2_ac():
    """two groups of particles on either side of an obstacle: the straight lines between them are blocked"""
    sc = scenarios.c1_se2(seed=11, n=1, shape=(128, 128, 1), resolution=0.08)
    env, robot, boxes = sc["env"], sc["robot"], sc["boxes"]
    b = boxes[0]
    rng = np.random.default_rng(48)
    n = 90
    side = np.where(rng.random(n) < 0.5, -1.0, 1.0)
    x = b[0] + side * (b[3] + 0.35) + rng.uniform(-0.03, 0.03, n)
    y = b[1] + rng.uniform(-0.03, 0.03, n)
    cfg = np.stack([x, y, rng.uniform(-0.1, 0.1, n)], axis=1)
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams(feature=abi.FEATURE_AC, dist_thr=0.2))


def cl_se3_ac_messy():
    sc = scenarios.c2_se3(seed=21, n=1, shape=(64, 64, 64), resolution=0.0625)
    rng = np.random.default_rng(49)
    n = 70
    cfg = np.zeros((n, 12))
    for i in range(n):
        R = scenarios.rotation_from_axis_angle(rng.normal(size=3), rng.uniform(0.0, 0.3))
        cfg[i] = scenarios.se3_config(R, rng.uniform(0.8, 3.2, size=3))
    return dict(env=sc["env"], robot=sc["robot"], configs=cfg, cparams=_cparams(feature=abi.FEATURE_AC, dist_thr=1.5))


def cl_linked_ac():
    env, robot = _linked_env_robot()
    rng = np.random.default_rng(50)
    n = 48
    cfg = rng.uniform(-0.05, 0.05, size=(n, 7))
    cfg[:, 1] += np.where(rng.random(n) < 0.5, -0.5, 0.55)
    return dict(env=env, robot=robot, configs=cfg, cparams=_cparams(feature=abi.FEATURE_AC, dist_thr=0.3))


def cl_se2_ptpm():
    """particles on two sides of a wall corner: moving between the groups fails, moving within succeeds"""
    c = cl_se2_ac()
    c["configs"] = c["configs"][:26]
    c["cparams"] = _cparams(dist_thr=0.2)
    return _with_ptpm(c, num_steps=40, goal=0.05)


def cl_se3_ptpm():
    env, robot = _se3_env_robot()
    rng = np.random.default_rng(54)
    n = 14
    cfg = np.zeros((n, 12))
    for i in range(n):
        R = scenarios.rotation_from_axis_angle(rng.normal(size=3), rng.uniform(0.0, 0.05))
        x = (1.7 if i % 2 == 0 else 2.0) + rng.uniform(-0.02, 0.02)
        cfg[i] = scenarios.se3_config(R, [x, 1.5 + rng.uniform(-0.02, 0.02), 1.5])
    return _with_ptpm(dict(env=env, robot=robot, configs=cfg, cparams=_cparams(dist_thr=0.1)), num_steps=25, goal=0.08)


def cl_linked_ptpm():
    """the 7-DoF arm, two groups of joint configurations on either side of an obstacle in the arm's sweep"""
    env, robot = _linked_env_robot()
    rng = np.random.default_rng(55)
    n = 12
    cfg = rng.uniform(-0.01, 0.01, size=(n, 7))
    cfg[:, 1] += np.where(np.arange(n) % 2 == 0, 0.0, 0.45)
    return _with_ptpm(dict(env=env, robot=robot, configs=cfg, cparams=_cparams(dist_thr=0.1)), num_steps=30, goal=0.05, seed=72)


GOLDEN_CLUSTER_CASES = {
    "linked_ptpm": cl_linked_ptpm,
    "se2_ac": cl_se2_ac,
    "se3_ac_messy": cl_se3_ac_messy,
    "linked_ac": cl_linked_ac,
    "se2_ptpm": cl_se2_ptpm,
    "se3_ptpm": cl_se3_ptpm,
    "se2_blobs": cl_se2_blobs,
    "se2_single_blob": cl_se2_single_blob,
    "se2_chain": cl_se2_chain,
    "se2_ties": cl_se2_ties,
    "se2_random": cl_se2_random,
    "se2_identical": cl_se2_identical,
    "se2_wrap": cl_se2_wrap,
    "se2_crs": cl_se2_crs,
    "se2_crs_messy": cl_se2_crs_messy,
    "se2_batch": cl_se2_batch,
    "se3_blobs": cl_se3_blobs,
    "se3_random": cl_se3_random,
    "se3_crs": cl_se3_crs,
    "linked_blobs": cl_linked_blobs,
    "linked_random_mixed": cl_linked_random_mixed,
    "linked_crs": cl_linked_crs,
}

"""Seeded synthetic workloads restating the five BASELINE.json configurations.

Pure numpy; used by tests/ and bench.py on both the CPU-oracle and the GPU side so that they see
bit-identical inputs.  Nothing here calls the oracle or the CUDA library.

  c1_se2      SE(2) robot, 8x8 body-point lattice, 256x256x1 SDF @ 0.04 m, 12 boxes, N = 1,000
  c2_se3      SE(3) free flyer, 4x4x8 lattice (P = 128), 128^3 SDF @ 0.03125 m, 24 boxes, N = 10,000
  c3_linked   7 revolute joints, 8 links x 16 spheres (P = 128), 256^3 SDF @ 0.015625 m, N = 100,000
  c4_blobs    clustering sweep inputs: K Gaussian blobs of configurations
  c5_batch    E edges x 1,000 particles on the C1 environment, one (start, target) pair per edge
"""
import math

import numpy as np

from . import abi


# ----------------------------------------------------------------------------------------- environments

def box_sdf(shape, resolution, boxes, origin=(0.0, 0.0, 0.0), planar=False):
    """Signed distance (float32) to a union of axis-aligned boxes, sampled at voxel centres.
    boxes: [k, 6] (cx, cy, cz, hx, hy, hz).  planar=True ignores z (2-D worlds)."""
    nx, ny, nz = shape
    xs = origin[0] + (np.arange(nx) + 0.5) * resolution
    ys = origin[1] + (np.arange(ny) + 0.5) * resolution
    zs = origin[2] + (np.arange(nz) + 0.5) * resolution
    sdf = np.full(shape, np.inf, dtype=np.float32)
    for b in np.asarray(boxes, dtype=np.float64):
        qx = (np.abs(xs - b[0]) - b[3]).astype(np.float32)[:, None, None]
        qy = (np.abs(ys - b[1]) - b[4]).astype(np.float32)[None, :, None]
        if planar:
            qz = np.full((1, 1, nz), -np.inf, dtype=np.float32)
        else:
            qz = (np.abs(zs - b[2]) - b[5]).astype(np.float32)[None, None, :]
        outside = np.sqrt(np.maximum(qx, 0) ** 2 + np.maximum(qy, 0) ** 2 + np.maximum(qz, 0) ** 2)
        inside = np.minimum(np.maximum(np.maximum(qx, qy), qz), 0)
        np.minimum(sdf, (outside + inside).astype(np.float32), out=sdf)
    return sdf


def slab_segments(shape, resolution, sdf, slabs=4, overlap_cells=2):
    """Synthetic convex-segment bitmasks: the free space is covered by `slabs` overlapping x-slabs
    (bits 0..) and `slabs` overlapping y-slabs (bits slabs..); obstacle cells carry 0."""
    nx, ny, nz = shape
    seg = np.zeros(shape, dtype=np.uint32)
    ix = np.arange(nx)[:, None, None]
    iy = np.arange(ny)[None, :, None]
    for s in range(slabs):
        lo = s * nx // slabs - overlap_cells
        hi = (s + 1) * nx // slabs + overlap_cells
        seg |= np.where((ix >= lo) & (ix < hi), np.uint32(1 << s), np.uint32(0)).astype(np.uint32)
        lo = s * ny // slabs - overlap_cells
        hi = (s + 1) * ny // slabs + overlap_cells
        seg |= np.where((iy >= lo) & (iy < hi), np.uint32(1 << (slabs + s)), np.uint32(0)).astype(np.uint32)
    seg = np.broadcast_to(seg, shape).copy()
    seg[sdf < 0] = 0
    return seg


def random_boxes(rng, count, extent, half_lo, half_hi, planar=False):
    boxes = np.zeros((count, 6))
    boxes[:, 0:3] = rng.uniform(0.15 * extent, 0.85 * extent, size=(count, 3))
    boxes[:, 3:6] = rng.uniform(half_lo, half_hi, size=(count, 3))
    if planar:
        boxes[:, 2] = 0.0
        boxes[:, 5] = 1.0
    return boxes


def make_environment(shape, resolution, boxes, planar=False, with_collision_map=True):
    origin = (0.0, 0.0, -0.5 * resolution) if planar else (0.0, 0.0, 0.0)
    sdf = box_sdf(shape, resolution, boxes, origin, planar)
    gfw = list(abi.IDENTITY34)
    gfw[3], gfw[7], gfw[11] = -origin[0], -origin[1], -origin[2]
    seg = occ = None
    if with_collision_map:
        seg = slab_segments(shape, resolution, sdf)
        occ = (sdf < 0).astype(np.float32)
    return abi.Environment(sdf, resolution, gfw, seg, occ)


def sample_sdf(env, points):
    """Nearest-voxel SDF value at world points [k, 3] (scenario construction only)."""
    g = np.array(env.desc.grid_from_world[:]).reshape(3, 4)
    p = points @ g[:, :3].T + g[:, 3]
    idx = np.floor(p / env.resolution).astype(int)
    shape = np.array(env.sdf.shape)
    ok = np.all((idx >= 0) & (idx < shape), axis=1)
    idx = np.clip(idx, 0, shape - 1)
    out = env.sdf[idx[:, 0], idx[:, 1], idx[:, 2]].astype(np.float64)
    out[~ok] = 1.0e3
    return out


# ----------------------------------------------------------------------------------------- robots

def se2_robot(kp=5.0, velocity_limit=1.0, r_velocity_limit=1.0, max_sensor_noise=0.0, max_actuator_noise=0.1,
              lattice=8, spacing=0.04, radius=0.0):
    c = (np.arange(lattice) - 0.5 * (lattice - 1)) * spacing
    xx, yy = np.meshgrid(c, c, indexing="ij")
    pts = np.stack([xx.ravel(), yy.ravel(), np.zeros(lattice * lattice), np.full(lattice * lattice, radius)], axis=1)
    lin = abi.make_axis(kp=kp, velocity_limit=velocity_limit, max_sensor_noise=max_sensor_noise,
                        max_actuator_noise=max_actuator_noise)
    rot = abi.make_axis(kp=kp, velocity_limit=r_velocity_limit, max_sensor_noise=max_sensor_noise,
                        max_actuator_noise=max_actuator_noise)
    return abi.Robot(abi.ROBOT_SE2, [pts], [lin, lin, rot])


def se3_robot(kp=5.0, velocity_limit=1.0, r_velocity_limit=1.0, max_sensor_noise=0.0, max_actuator_noise=0.1,
              lattice=(4, 4, 8), spacing=0.04, radius=0.0):
    axes = [(np.arange(k) - 0.5 * (k - 1)) * spacing for k in lattice]
    xx, yy, zz = np.meshgrid(*axes, indexing="ij")
    n = xx.size
    pts = np.stack([xx.ravel(), yy.ravel(), zz.ravel(), np.full(n, radius)], axis=1)
    lin = abi.make_axis(kp=kp, velocity_limit=velocity_limit, max_sensor_noise=max_sensor_noise,
                        max_actuator_noise=max_actuator_noise)
    rot = abi.make_axis(kp=kp, velocity_limit=r_velocity_limit, max_sensor_noise=max_sensor_noise,
                        max_actuator_noise=max_actuator_noise)
    return abi.Robot(abi.ROBOT_SE3, [pts], [lin, lin, lin, rot, rot, rot])


def translation34(x, y, z):
    return [1.0, 0.0, 0.0, x, 0.0, 1.0, 0.0, y, 0.0, 0.0, 1.0, z]


def linked_arm(num_joints=7, link_length=0.3, points_per_link=16, radius=0.02, kp=5.0, velocity_limit=1.0,
               max_sensor_noise=0.0, max_actuator_noise=0.1, limit=2.9, base=(2.0, 2.0, 0.5),
               joint_types=None, weights=None):
    """Serial arm: joint k connects link k to link k+1; axes alternate z / y; every link carries
    `points_per_link` spheres spread along its local z axis (the root link's sit at its origin column)."""
    links = []
    for l in range(num_joints + 1):
        t = (np.arange(points_per_link) + 0.5) / points_per_link * link_length
        pts = np.stack([np.zeros_like(t), np.zeros_like(t), t, np.full_like(t, radius)], axis=1)
        links.append(pts)
    joints = []
    axes = []
    for k in range(num_joints):
        jt = abi.JOINT_REVOLUTE if joint_types is None else joint_types[k]
        joints.append(dict(type=jt, parent=k, child=k + 1, axis=(0.0, 0.0, 1.0) if k % 2 == 0 else (0.0, 1.0, 0.0),
                           lower=-limit, upper=limit, transform=translation34(0.0, 0.0, link_length)))
        if jt != abi.JOINT_FIXED:
            axes.append(abi.make_axis(kp=kp, velocity_limit=velocity_limit, max_sensor_noise=max_sensor_noise,
                                      max_actuator_noise=max_actuator_noise))
    return abi.Robot(abi.ROBOT_LINKED, links, axes, joints=joints, distance_weights=weights,
                     base_transform=translation34(*base))


# ----------------------------------------------------------------------------------------- configurations

def rotation_from_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(angle) * K + (1 - math.cos(angle)) * (K @ K)


def se3_config(R, t):
    return np.concatenate([np.asarray(R, dtype=np.float64).reshape(9), np.asarray(t, dtype=np.float64).reshape(3)])


def _pick_contact_edge(rng, boxes, env, clearance, travel, push, planar):
    """A (start, target) translation pair that slides along a box face while pushing `push` metres into it."""
    shape = np.array(env.sdf.shape) * env.resolution
    for _ in range(1000):
        b = boxes[rng.integers(len(boxes))]
        axis = int(rng.integers(2 if planar else 3))
        sign = 1.0 if rng.random() < 0.5 else -1.0
        tang = [a for a in range(2 if planar else 3) if a != axis]
        start = b[0:3].copy()
        start[axis] = b[axis] + sign * (b[3 + axis] + clearance)
        target = start.copy()
        target[axis] = b[axis] + sign * (b[3 + axis] - push)
        ta = tang[int(rng.integers(len(tang)))]
        off = rng.uniform(-0.5, 0.5) * b[3 + ta]
        start[ta] += off - 0.5 * travel
        target[ta] += off + 0.5 * travel
        if planar:
            start[2] = target[2] = 0.0
        lo, hi = 0.3, shape[: (2 if planar else 3)] - 0.3
        if np.any(start[: len(hi)] < lo) or np.any(start[: len(hi)] > hi):
            continue
        if sample_sdf(env, start[None, :])[0] < clearance * 0.9:
            continue
        return start, target
    raise RuntimeError("could not place a contact edge")


def c1_se2(seed=1, n=1000, num_steps=64, max_microsteps=1, shape=(256, 256, 1), resolution=0.04, num_boxes=12):
    rng = np.random.default_rng(seed)
    extent = shape[0] * resolution
    boxes = random_boxes(rng, num_boxes, extent, 0.3, 0.8, planar=True)
    env = make_environment(shape, resolution, boxes, planar=True)
    robot = se2_robot()
    start, target = _pick_contact_edge(rng, boxes, env, clearance=0.45, travel=0.35, push=0.05, planar=True)
    th0, th1 = rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5)
    starts = np.tile(np.array([start[0], start[1], th0]), (n, 1))
    targets = np.array([[target[0], target[1], th1]])
    params = abi.make_sim_params(num_steps, 0.01, resolution, allow_contacts=True, max_microsteps=max_microsteps)
    return dict(name="c1_se2", env=env, robot=robot, starts=starts, targets=targets, params=params, seed=seed, boxes=boxes)


def c2_se3(seed=2, n=10000, num_steps=64, max_microsteps=1, shape=(128, 128, 128), resolution=0.03125, num_boxes=24, edge=0):
    """`edge` > 0 picks another planner edge (start, target) in the same environment (multi-GPU edge sharding)."""
    rng = np.random.default_rng(seed)
    extent = shape[0] * resolution
    boxes = random_boxes(rng, num_boxes, extent, 0.2, 0.5)
    env = make_environment(shape, resolution, boxes)
    robot = se3_robot()
    if edge:
        rng = np.random.default_rng([seed, 1000 + edge])
    start, target = _pick_contact_edge(rng, boxes, env, clearance=0.35, travel=0.35, push=0.05, planar=False)
    R0 = rotation_from_axis_angle(rng.normal(size=3), rng.uniform(0.2, 1.0))
    R1 = rotation_from_axis_angle(rng.normal(size=3), rng.uniform(0.2, 1.0))
    starts = np.tile(se3_config(R0, start), (n, 1))
    targets = se3_config(R1, target)[None, :]
    params = abi.make_sim_params(num_steps, 0.01, resolution, allow_contacts=True, max_microsteps=max_microsteps)
    return dict(name="c2_se3", env=env, robot=robot, starts=starts, targets=targets, params=params, seed=seed, boxes=boxes)


def c3_linked(seed=3, n=100000, num_steps=64, max_microsteps=1, shape=(256, 256, 256), resolution=0.015625, num_boxes=24):
    rng = np.random.default_rng(seed)
    extent = shape[0] * resolution
    boxes = random_boxes(rng, num_boxes, extent, 0.15, 0.4)
    # keep the base column free of the random clutter, then place three obstacles in the arm's sweep
    keep = np.linalg.norm(boxes[:, 0:2] - np.array([2.0, 2.0]), axis=1) > 0.9
    boxes = np.concatenate([boxes[keep], np.array([[2.45, 2.0, 2.1, 0.2, 0.4, 0.25],
                                                   [1.55, 2.0, 1.4, 0.15, 0.4, 0.2],
                                                   [2.0, 2.5, 2.6, 0.4, 0.15, 0.2]])], axis=0)
    env = make_environment(shape, resolution, boxes)
    robot = linked_arm()
    q0 = np.array([0.05, 0.02, -0.03, 0.04, 0.0, 0.02, 0.0]) + rng.uniform(-0.02, 0.02, size=7)
    q1 = np.array([0.1, 0.2, 0.0, 0.0, 0.1, 0.0, 0.0]) + rng.uniform(-0.005, 0.005, size=7)
    starts = np.tile(q0, (n, 1))
    targets = q1[None, :]
    params = abi.make_sim_params(num_steps, 0.01, resolution, allow_contacts=True, max_microsteps=max_microsteps)
    return dict(name="c3_linked", env=env, robot=robot, starts=starts, targets=targets, params=params, seed=seed, boxes=boxes)


def c4_blobs(kind, n, k=8, threshold=0.1, seed=4, dof=7):
    """Configurations drawn as k blobs whose diameter stays below `threshold` and whose centres are at
    least 4 x threshold apart: complete-link clustering at `threshold` returns exactly the blobs."""
    rng = np.random.default_rng(seed)
    owner = rng.integers(k, size=n)
    owner[:k] = np.arange(k)  # every blob is inhabited
    radius = 0.2 * threshold  # per-component spread keeps blob diameters < threshold
    if kind == abi.ROBOT_SE2:
        centres = np.stack([np.arange(k) * 4.0 * threshold + 1.0, np.full(k, 1.0), np.linspace(-1.0, 1.0, k)], axis=1)
        cfg = centres[owner] + rng.uniform(-radius, radius, size=(n, 3)) * np.array([1.0, 1.0, 0.5])
        return cfg, owner
    if kind == abi.ROBOT_SE3:
        cfg = np.zeros((n, 12))
        axes = rng.normal(size=(k, 3))
        angles = rng.uniform(0.2, 1.2, size=k)
        for i in range(n):
            b = owner[i]
            R = rotation_from_axis_angle(axes[b], angles[b] + rng.uniform(-radius, radius) * 0.5)
            t = np.array([b * 4.0 * threshold + 1.0, 1.0, 1.0]) + rng.uniform(-radius, radius, size=3) * 0.5
            cfg[i] = se3_config(R, t)
        return cfg, owner
    centres = rng.uniform(-1.0, 1.0, size=(k, dof))
    centres[:, 0] = np.arange(k) * 4.0 * threshold - 1.4
    cfg = centres[owner] + rng.uniform(-radius, radius, size=(n, dof)) / math.sqrt(dof)
    return cfg, owner


def c5_batch(seed=5, edges=4096, particles=1000, num_steps=64):
    """E candidate edges x `particles` particles on the C1 environment, as one batched call with
    targets.size() == starts.size() (the form used at uncertainty_contact_planning.hpp:1826)."""
    base = c1_se2(seed=1, n=1, num_steps=num_steps)
    env, robot, boxes = base["env"], base["robot"], base["boxes"]
    rng = np.random.default_rng(seed)
    starts = np.zeros((edges, 3))
    targets = np.zeros((edges, 3))
    for e in range(edges):
        if e % 2 == 0:
            s, t = _pick_contact_edge(rng, boxes, env, clearance=0.45, travel=0.35, push=0.05, planar=True)
        else:
            while True:
                s = np.array([rng.uniform(0.6, 9.6), rng.uniform(0.6, 9.6), 0.0])
                if sample_sdf(env, s[None, :])[0] > 0.45:
                    break
            ang = rng.uniform(-math.pi, math.pi)
            t = s + 0.4 * np.array([math.cos(ang), math.sin(ang), 0.0])
        starts[e] = [s[0], s[1], rng.uniform(-0.5, 0.5)]
        targets[e] = [t[0], t[1], rng.uniform(-0.5, 0.5)]
    return dict(name="c5_batch", env=env, robot=robot, starts=np.repeat(starts, particles, axis=0),
                targets=np.repeat(targets, particles, axis=0), params=base["params"], seed=seed, edges=edges,
                particles=particles, boxes=boxes)

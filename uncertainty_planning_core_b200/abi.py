"""ctypes mirror of include/upc_b200.h and loader for the in-tree CUDA library.

The structures here are shared by the product bindings (libupc_b200.so) and by the test-side oracle
bindings (oracle/binding.py), which take the very same POD descriptors.
"""
import ctypes as C
import os

import numpy as np

UPC_MAX_DOF = 16
UPC_MAX_JOINTS = 16
UPC_MAX_LINKS = 17

UPC_OK = 0
UPC_ERR_INVALID_ARGUMENT = 1
UPC_ERR_CUDA = 2
UPC_ERR_NO_ROBOT = 3
UPC_ERR_UNSUPPORTED = 4

ROBOT_SE2, ROBOT_SE3, ROBOT_LINKED = 0, 1, 2
JOINT_FIXED, JOINT_REVOLUTE, JOINT_CONTINUOUS, JOINT_PRISMATIC = 0, 1, 2, 4
FEATURE_NONE, FEATURE_CRS, FEATURE_AC, FEATURE_PTPM = 0, 1, 2, 3


class EnvDesc(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("reserved0", C.c_int32),
        ("resolution", C.c_double),
        ("grid_from_world", C.c_double * 12),
        ("sdf", C.c_void_p),
        ("convex_segment", C.c_void_p),
        ("occupancy", C.c_void_p),
        ("sdf_oob_value", C.c_float),
        ("occupancy_oob_value", C.c_float),
    ]


class AxisParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ("kp", "ki", "kd", "integral_clamp", "velocity_limit", "max_sensor_noise", "max_actuator_noise")]


class JointDesc(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("parent_link", C.c_int32), ("child_link", C.c_int32), ("reserved0", C.c_int32),
        ("axis", C.c_double * 3),
        ("lower_limit", C.c_double), ("upper_limit", C.c_double),
        ("transform", C.c_double * 12),
    ]


class RobotDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("dof", C.c_int32), ("num_links", C.c_int32), ("num_joints", C.c_int32),
        ("num_points", C.c_int32), ("reserved0", C.c_int32),
        ("link_point_offset", C.c_int32 * (UPC_MAX_LINKS + 1)),
        ("points", C.c_void_p),
        ("axis", AxisParams * UPC_MAX_DOF),
        ("distance_weights", C.c_double * UPC_MAX_DOF),
        ("base_transform", C.c_double * 12),
        ("joints", JointDesc * UPC_MAX_JOINTS),
    ]


class SimParams(C.Structure):
    _fields_ = [
        ("num_steps", C.c_int32), ("max_microsteps", C.c_int32), ("allow_contacts", C.c_int32),
        ("use_individual_jacobians", C.c_int32), ("max_resolver_iterations", C.c_int32),
        ("resolver_decay_iterations", C.c_int32), ("reserved0", C.c_int32), ("reserved1", C.c_int32),
        ("controller_interval", C.c_double), ("simulation_shortcut_distance", C.c_double),
        ("contact_distance", C.c_double), ("resolve_distance", C.c_double), ("microstep_distance", C.c_double),
        ("resolver_initial_scale", C.c_double), ("resolver_decay_rate", C.c_double), ("resolver_damping", C.c_double),
    ]


class ClusterParams(C.Structure):
    _fields_ = [
        ("feature_type", C.c_int32), ("reserved0", C.c_int32),
        ("signature_matching_threshold", C.c_double),
        ("distance_clustering_threshold", C.c_double),
        ("goal_distance_threshold", C.c_double),
        ("ptpm_sim", C.POINTER(SimParams)),
        ("ptpm_tape", C.c_void_p),
        ("ptpm_seed", C.c_uint64),
    ]


class OutcomeLayout(C.Structure):
    _fields_ = [(k, C.c_int64) for k in
                ("particles", "sets", "labels_offset", "n_clusters_offset", "collided_offset", "configs_offset", "block_bytes")]


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in
                ("kernel_launches", "particles_simulated", "particle_steps", "collided_particles",
                 "resolver_iterations", "failed_resolves", "cluster_calls", "cluster_fallback_calls",
                 "h2d_bytes", "d2h_bytes")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def variance_width(kind, dof):
    if kind == ROBOT_SE2:
        return 3
    if kind == ROBOT_SE3:
        return 4
    return dof


def config_width(kind, dof):
    if kind == ROBOT_SE2:
        return 3
    if kind == ROBOT_SE3:
        return 12
    return dof


IDENTITY34 = (1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0)


class Environment:
    """Host-side voxel environment (SDF + collision-map planes) and its upc_env_desc."""

    def __init__(self, sdf, resolution, grid_from_world=IDENTITY34, convex_segment=None, occupancy=None,
                 sdf_oob_value=1.0e3, occupancy_oob_value=0.0):
        sdf = np.ascontiguousarray(sdf, dtype=np.float32)
        assert sdf.ndim == 3
        self.sdf = sdf
        self.convex_segment = None if convex_segment is None else np.ascontiguousarray(convex_segment, dtype=np.uint32)
        self.occupancy = None if occupancy is None else np.ascontiguousarray(occupancy, dtype=np.float32)
        self.resolution = float(resolution)
        d = EnvDesc()
        d.nx, d.ny, d.nz = sdf.shape
        d.resolution = self.resolution
        d.grid_from_world = (C.c_double * 12)(*[float(v) for v in grid_from_world])
        d.sdf = sdf.ctypes.data
        d.convex_segment = None if self.convex_segment is None else self.convex_segment.ctypes.data
        d.occupancy = None if self.occupancy is None else self.occupancy.ctypes.data
        d.sdf_oob_value = sdf_oob_value
        d.occupancy_oob_value = occupancy_oob_value
        self.desc = d


def make_axis(kp=1.0, ki=0.0, kd=0.0, integral_clamp=0.0, velocity_limit=1.0, max_sensor_noise=0.0,
              max_actuator_noise=0.0):
    a = AxisParams()
    a.kp, a.ki, a.kd, a.integral_clamp = kp, ki, kd, integral_clamp
    a.velocity_limit, a.max_sensor_noise, a.max_actuator_noise = velocity_limit, max_sensor_noise, max_actuator_noise
    return a


class Robot:
    """Host-side robot model and its upc_robot_desc.

    points_per_link: list (one entry per link) of arrays [k, 3] or [k, 4] (x, y, z[, radius]).
    axes: list of AxisParams, one per degree of freedom.
    joints (linked only): list of dicts(type, parent, child, axis, lower, upper, transform(12)).
    """

    def __init__(self, kind, points_per_link, axes, joints=None, distance_weights=None, base_transform=IDENTITY34):
        self.kind = kind
        pts = []
        offsets = [0]
        for lp in points_per_link:
            lp = np.asarray(lp, dtype=np.float64)
            if lp.size == 0:
                lp = np.zeros((0, 4))
            elif lp.shape[-1] == 3:
                lp = np.concatenate([lp.reshape(-1, 3), np.zeros((lp.reshape(-1, 3).shape[0], 1))], axis=1)
            else:
                lp = lp.reshape(-1, 4)
            pts.append(lp)
            offsets.append(offsets[-1] + lp.shape[0])
        self.points = np.ascontiguousarray(np.concatenate(pts, axis=0), dtype=np.float64)
        d = RobotDesc()
        d.kind = kind
        d.dof = len(axes)
        d.num_links = len(points_per_link)
        d.num_points = self.points.shape[0]
        for i, o in enumerate(offsets):
            d.link_point_offset[i] = o
        d.points = self.points.ctypes.data
        for i, a in enumerate(axes):
            d.axis[i] = a
        weights = distance_weights if distance_weights is not None else [1.0] * len(axes)
        for i, w in enumerate(weights):
            d.distance_weights[i] = float(w)
        d.base_transform = (C.c_double * 12)(*[float(v) for v in base_transform])
        joints = joints or []
        d.num_joints = len(joints)
        for i, j in enumerate(joints):
            jd = d.joints[i]
            jd.type = j["type"]
            jd.parent_link = j["parent"]
            jd.child_link = j["child"]
            jd.axis = (C.c_double * 3)(*[float(v) for v in j["axis"]])
            jd.lower_limit = float(j.get("lower", 0.0))
            jd.upper_limit = float(j.get("upper", 0.0))
            jd.transform = (C.c_double * 12)(*[float(v) for v in j["transform"]])
        self.desc = d
        self.dof = d.dof
        self.width = config_width(kind, d.dof)
        self.num_points = d.num_points


def make_sim_params(num_steps, controller_interval, resolution, allow_contacts=True, max_microsteps=1,
                    simulation_shortcut_distance=0.0, use_individual_jacobians=False, contact_distance=0.0,
                    resolve_margin_ratio=0.1, microstep_ratio=0.5, max_resolver_iterations=25,
                    resolver_decay_iterations=5, resolver_initial_scale=1.0, resolver_decay_rate=0.5,
                    resolver_damping=1.0e-6):
    p = SimParams()
    p.num_steps = num_steps
    p.max_microsteps = max_microsteps
    p.allow_contacts = 1 if allow_contacts else 0
    p.use_individual_jacobians = 1 if use_individual_jacobians else 0
    p.max_resolver_iterations = max_resolver_iterations
    p.resolver_decay_iterations = resolver_decay_iterations
    p.controller_interval = controller_interval
    p.simulation_shortcut_distance = simulation_shortcut_distance
    p.contact_distance = contact_distance
    p.resolve_distance = contact_distance + resolve_margin_ratio * resolution
    p.microstep_distance = microstep_ratio * resolution
    p.resolver_initial_scale = resolver_initial_scale
    p.resolver_decay_rate = resolver_decay_rate
    p.resolver_damping = resolver_damping
    return p


def tape_shape(params, dof, n):
    return (params.num_steps, 1 + params.max_microsteps, dof, n)


class UpcError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("upc error %d: %s" % (code, message))
        self.code = code


_LIB = None


def library_path():
    """In-tree library; UPC_B200_LIBRARY names an alternative build of the same sources (kernel tuning experiments)."""
    override = os.environ.get("UPC_B200_LIBRARY")
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libupc_b200.so")


def load_library():
    """Load libupc_b200.so.  There is no fallback: a missing library is an error."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C uncertainty_planning_core_b200/csrc`). This package has no CPU fallback." % path)
    lib = C.CDLL(path)
    vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.upc_last_error.restype = C.c_char_p
    lib.upc_version.restype = C.c_char_p
    lib.upc_config_width.argtypes = [i32, i32]
    lib.upc_config_width.restype = i32
    lib.upc_tape_doubles_per_particle.argtypes = [C.POINTER(SimParams), i32]
    lib.upc_tape_doubles_per_particle.restype = i64
    lib.upc_create.argtypes = [C.POINTER(EnvDesc), i32, C.POINTER(vp)]
    lib.upc_destroy.argtypes = [vp]
    lib.upc_set_robot.argtypes = [vp, C.POINTER(RobotDesc)]
    lib.upc_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.upc_reset_stats.argtypes = [vp]
    lib.upc_set_lanes_per_particle.argtypes = [vp, i32]
    lib.upc_host_alloc.argtypes = [C.POINTER(vp), i64]
    lib.upc_host_free.argtypes = [vp]
    lib.upc_synchronize.argtypes = [vp]
    lib.upc_forward_simulate.argtypes = [vp, C.POINTER(SimParams), i64, vp, i64, vp, vp, vp, vp]
    lib.upc_forward_simulate_device.argtypes = [vp, C.POINTER(SimParams), i64, vp, i64, vp, vp, vp, vp, vp]
    lib.upc_check_config_collision.argtypes = [vp, vp, dp, dp, C.POINTER(i32)]
    lib.upc_cluster_particles.argtypes = [vp, C.POINTER(ClusterParams), i64, i64, vp, vp, vp]
    lib.upc_cluster_particles_device.argtypes = [vp, C.POINTER(ClusterParams), i64, i64, vp, vp, vp, vp]
    lib.upc_fill_noise_tape.argtypes = [C.c_uint64, C.POINTER(RobotDesc), C.POINTER(SimParams), i64, i64, i64, vp]
    lib.upc_count_within.argtypes = [vp, i64, vp, i64, vp, dp, C.POINTER(i64), vp]
    lib.upc_nearest_neighbors.argtypes = [vp, i64, vp, vp, vp, vp, i64, vp, dp, vp, vp]
    lib.upc_variance_width.argtypes = [i32, i32]
    lib.upc_variance_width.restype = i32
    lib.upc_cluster_statistics.argtypes = [vp, i64, vp, vp, vp, i32, dp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.upc_measure_l2_gather_gbs.argtypes = [vp, i64, C.POINTER(dp)]
    lib.upc_measure_fp64_tflops.argtypes = [vp, C.POINTER(dp)]
    u64 = C.c_uint64
    lib.upc_forward_simulate_seeded.argtypes = [vp, C.POINTER(SimParams), u64, i64, i64, i64, vp, i64, vp, vp, vp]
    lib.upc_forward_simulate_seeded_device.argtypes = [vp, C.POINTER(SimParams), u64, i64, i64, i64, vp, i64, vp, vp, vp, vp]
    lib.upc_fill_noise_tape_seeded.argtypes = [u64, C.POINTER(RobotDesc), C.POINTER(SimParams), i64, i64, i64, vp]
    lib.upc_forward_simulate_traced.argtypes = [vp, C.POINTER(SimParams), i64, vp, i64, vp, vp, u64, i64, vp, vp, vp, vp, vp]
    lib.upc_simulate_and_cluster_edges.argtypes = [vp, C.POINTER(SimParams), C.POINTER(ClusterParams), u64, i64, i64, i64, i64, vp, vp, vp, vp, vp, vp]
    lib.upc_device_scratch.argtypes = [vp, i32, i64, C.POINTER(vp)]
    lib.upc_memcpy_to_device.argtypes = [vp, vp, vp, i64]
    lib.upc_memcpy_to_host.argtypes = [vp, vp, vp, i64]
    lib.upc_comm_get_unique_id.argtypes = [vp]
    lib.upc_comm_create.argtypes = [vp, i32, i32, vp, C.POINTER(vp)]
    lib.upc_comm_create_all.argtypes = [i32, C.POINTER(vp), C.POINTER(vp)]
    lib.upc_comm_destroy.argtypes = [vp]
    lib.upc_comm_world.argtypes = [vp]
    lib.upc_comm_rank.argtypes = [vp]
    lib.upc_comm_nccl_version.argtypes = [C.POINTER(i32)]
    lib.upc_shard_bounds.argtypes = [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    lib.upc_shard_bounds.restype = None
    lib.upc_outcome_layout_for.argtypes = [i32, i64, i64, C.POINTER(OutcomeLayout)]
    lib.upc_comm_all_gather_outcomes.argtypes = [vp, vp, i64, vp]
    lib.upc_comm_wait.argtypes = [vp, vp]
    lib.upc_comm_synchronize.argtypes = [vp]
    lib.upc_sharded_edge_batch.argtypes = [vp, C.POINTER(SimParams), C.POINTER(ClusterParams), u64, i64, i64, vp, vp, vp, vp]
    lib.upc_edge_batch_submit.argtypes = [vp, C.POINTER(SimParams), u64, i64, i64, i64, i64, vp, vp, i32]
    lib.upc_edge_batch_collect.argtypes = [vp, C.POINTER(ClusterParams), i32, vp, vp, vp, vp]
    lib.upc_sharded_edge_batch_simulate.argtypes = [vp, C.POINTER(SimParams), u64, i64, i64, vp, vp, vp, vp]
    lib.upc_sharded_edge_batch_finish.argtypes = [vp, C.POINTER(ClusterParams), i64, i64, vp, vp]
    lib.upc_sharded_edge_batch_all.argtypes = [i32, vp, C.POINTER(SimParams), C.POINTER(ClusterParams), u64, i64, i64, vp, vp, vp, vp]
    for name in ("upc_forward_simulate_seeded", "upc_forward_simulate_seeded_device", "upc_fill_noise_tape_seeded", "upc_forward_simulate_traced",
                 "upc_simulate_and_cluster_edges", "upc_comm_get_unique_id", "upc_comm_create", "upc_comm_create_all", "upc_comm_destroy",
                 "upc_comm_world", "upc_comm_rank", "upc_comm_nccl_version", "upc_outcome_layout_for", "upc_comm_all_gather_outcomes", "upc_comm_wait",
                 "upc_comm_synchronize", "upc_sharded_edge_batch", "upc_sharded_edge_batch_simulate", "upc_sharded_edge_batch_finish", "upc_sharded_edge_batch_all",
                 "upc_edge_batch_submit", "upc_edge_batch_collect", "upc_device_scratch", "upc_memcpy_to_device", "upc_memcpy_to_host"):
        getattr(lib, name).restype = i32
    for name in ("upc_cluster_statistics", "upc_nearest_neighbors", "upc_count_within", "upc_measure_l2_gather_gbs", "upc_measure_fp64_tflops", "upc_create", "upc_destroy", "upc_set_robot", "upc_get_stats", "upc_reset_stats",
                 "upc_set_lanes_per_particle", "upc_host_alloc", "upc_host_free", "upc_synchronize",
                 "upc_forward_simulate", "upc_forward_simulate_device", "upc_check_config_collision",
                 "upc_cluster_particles", "upc_cluster_particles_device", "upc_fill_noise_tape"):
        getattr(lib, name).restype = i32
    _LIB = lib
    return lib


EXPORTED_SYMBOLS = (
    "upc_config_width", "upc_tape_doubles_per_particle", "upc_last_error", "upc_version", "upc_create", "upc_destroy",
    "upc_set_robot", "upc_get_stats", "upc_reset_stats", "upc_set_lanes_per_particle", "upc_host_alloc", "upc_host_free",
    "upc_forward_simulate", "upc_forward_simulate_device", "upc_check_config_collision", "upc_cluster_particles",
    "upc_cluster_particles_device", "upc_fill_noise_tape", "upc_synchronize", "upc_measure_l2_gather_gbs",
    "upc_measure_fp64_tflops", "upc_count_within", "upc_variance_width", "upc_cluster_statistics", "upc_nearest_neighbors",
    "upc_forward_simulate_seeded", "upc_forward_simulate_seeded_device", "upc_fill_noise_tape_seeded", "upc_forward_simulate_traced",
    "upc_simulate_and_cluster_edges", "upc_comm_get_unique_id", "upc_comm_create", "upc_comm_create_all", "upc_comm_destroy",
    "upc_comm_world", "upc_comm_rank", "upc_comm_nccl_version", "upc_sharded_edge_batch_all", "upc_shard_bounds", "upc_device_scratch", "upc_memcpy_to_device", "upc_memcpy_to_host",
    "upc_outcome_layout_for", "upc_comm_all_gather_outcomes", "upc_comm_wait", "upc_comm_synchronize", "upc_sharded_edge_batch", "upc_sharded_edge_batch_simulate", "upc_sharded_edge_batch_finish",
    "upc_edge_batch_submit", "upc_edge_batch_collect",
)


def check(code):
    if code != UPC_OK:
        msg = load_library().upc_last_error().decode("utf-8", "replace")
        if code == UPC_ERR_INVALID_ARGUMENT:
            raise ValueError("upc: " + msg)
        raise UpcError(code, msg)


def fill_noise_tape(seed, robot, params, n, first_particle=0):
    """Host-side tape generation through the library (no GPU needed)."""
    lib = load_library()
    tape = np.zeros(tape_shape(params, robot.dof, n), dtype=np.float64)
    check(lib.upc_fill_noise_tape(seed, C.byref(robot.desc), C.byref(params), first_particle, n, n, tape.ctypes.data))
    return tape


def fill_noise_tape_seeded(seed, robot, params, n, first_particle=0):
    """The draws of the seeded entry points (counter-based generator) written out as a tape."""
    lib = load_library()
    tape = np.zeros(tape_shape(params, robot.dof, n), dtype=np.float64)
    check(lib.upc_fill_noise_tape_seeded(seed, C.byref(robot.desc), C.byref(params), first_particle, n, n, tape.ctypes.data))
    return tape


def shard_bounds(total, rank, world):
    lo, hi = C.c_int64(0), C.c_int64(0)
    load_library().upc_shard_bounds(total, rank, world, C.byref(lo), C.byref(hi))
    return int(lo.value), int(hi.value)


def outcome_layout(width, particles, sets):
    lay = OutcomeLayout()
    check(load_library().upc_outcome_layout_for(width, particles, sets, C.byref(lay)))
    return lay

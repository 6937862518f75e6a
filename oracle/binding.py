"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of oracle/_ref/libref_pid.so.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

from uncertainty_planning_core_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when the reference tree is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    if force:
        subprocess.check_call(["make", "-B", "-C", _HERE, "liboracle.so"], stdout=sys.stderr)
    elif not os.path.exists(so) or os.path.exists(os.path.join(_HERE, "upc_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=sys.stderr)
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=sys.stderr)
    return so


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(so):
        build()
    lib = C.CDLL(so)
    vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.oracle_last_error.restype = C.c_char_p
    lib.oracle_num_threads.restype = i32
    lib.oracle_set_num_threads.argtypes = [i32]
    lib.oracle_pid_sequence.argtypes = [dp, dp, dp, dp, vp, vp, i32, vp]
    lib.oracle_pid_sequence.restype = dp
    lib.oracle_math.argtypes = [i32, i64, vp, vp, vp]
    lib.oracle_se3_log.argtypes = [vp, vp]
    lib.oracle_se3_exp.argtypes = [vp, vp]
    lib.oracle_quaternion_from_rotation.argtypes = [vp, vp]
    lib.oracle_quaternion_from_rotation.restype = None
    lib.oracle_forward_simulate.argtypes = [C.POINTER(abi.EnvDesc), C.POINTER(abi.RobotDesc), C.POINTER(abi.SimParams),
                                            i64, vp, i64, vp, vp, vp, vp, C.POINTER(abi.Stats)]
    lib.oracle_check_config_collision.argtypes = [C.POINTER(abi.EnvDesc), C.POINTER(abi.RobotDesc), vp, dp, dp, C.POINTER(i32)]
    lib.oracle_cluster_particles.argtypes = [C.POINTER(abi.EnvDesc), C.POINTER(abi.RobotDesc), C.POINTER(abi.ClusterParams),
                                             i64, i64, vp, vp, vp, C.POINTER(abi.Stats)]
    lib.oracle_pair_distances.argtypes = [C.POINTER(abi.RobotDesc), i64, vp, vp]
    lib.oracle_pair_distances_reference_style.argtypes = [C.POINTER(abi.RobotDesc), i64, vp, vp]
    lib.oracle_region_signatures.argtypes = [C.POINTER(abi.EnvDesc), C.POINTER(abi.RobotDesc), i64, vp, vp]
    lib.oracle_sdf_lookup.argtypes = [C.POINTER(abi.EnvDesc), i64, vp, vp, vp]
    lib.oracle_complete_link.argtypes = [vp, i64, dp, vp, vp]
    lib.oracle_nearest_neighbors.argtypes = [C.POINTER(abi.RobotDesc), i64, vp, vp, vp, vp, i64, vp, dp, vp, vp]
    lib.oracle_nearest_neighbors.restype = i32
    lib.oracle_cluster_statistics.argtypes = [C.POINTER(abi.RobotDesc), i64, vp, vp, vp, i32, dp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.oracle_cluster_statistics.restype = i32
    lib.oracle_forward_simulate_traced.argtypes = [C.POINTER(abi.EnvDesc), C.POINTER(abi.RobotDesc), C.POINTER(abi.SimParams),
                                                   i64, vp, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.oracle_forward_simulate_traced.restype = i32
    lib.oracle_sensor_values.argtypes = [i64, vp, vp, vp]
    lib.oracle_sensor_values.restype = None
    lib.oracle_actuator_values.argtypes = [dp, dp, i64, vp, vp, vp, vp]
    lib.oracle_actuator_values.restype = None
    lib.oracle_philox4x32_10.argtypes = [vp, vp, vp]
    lib.oracle_philox4x32_10.restype = None
    lib.oracle_truncated_normal_quantile.argtypes = [i64, vp, vp]
    lib.oracle_truncated_normal_quantile.restype = None
    lib.oracle_centered_uniform52.argtypes = [i64, vp, vp, vp]
    lib.oracle_centered_uniform52.restype = None
    lib.oracle_log.argtypes = [i64, vp, vp]
    lib.oracle_log.restype = None
    lib.oracle_standard_draws.argtypes = [C.c_uint64, i64, i64, C.c_uint32, C.c_uint32, C.c_uint32, vp]
    lib.oracle_standard_draws.restype = None
    lib.oracle_fill_noise_tape_counter.argtypes = [C.c_uint64, C.POINTER(abi.RobotDesc), C.POINTER(abi.SimParams), i64, i64, i64, vp]
    lib.oracle_fill_noise_tape_counter.restype = i32
    rp = C.POINTER(abi.RobotDesc)
    lib.oracle_rob_sequence.argtypes = [rp, i32, i32, dp, i32, vp, vp, vp, vp, vp, vp]
    lib.oracle_rob_kinematics.argtypes = [rp, i32, vp, vp, vp]
    lib.oracle_rob_distances.argtypes = [rp, i32, vp, vp, vp, vp]
    lib.oracle_rob_average.argtypes = [rp, i32, vp, vp]
    lib.upc_variance_width.argtypes = [i32, i32]
    for name in ("oracle_rob_sequence", "oracle_rob_kinematics", "oracle_rob_distances", "oracle_rob_average", "upc_variance_width"):
        getattr(lib, name).restype = i32
    for name in ("oracle_forward_simulate", "oracle_check_config_collision", "oracle_cluster_particles",
                 "oracle_pair_distances", "oracle_pair_distances_reference_style", "oracle_region_signatures",
                 "oracle_sdf_lookup", "oracle_complete_link"):
        getattr(lib, name).restype = i32
    _LIB = lib
    return lib


def load_ref_pid():
    """The reference's own PID controller compiled unmodified (None when oracle/_ref was never built)."""
    global _REF
    if _REF is not None:
        return _REF
    so = os.path.join(_HERE, "_ref", "libref_pid.so")
    if not os.path.exists(so):
        return None
    lib = C.CDLL(so)
    lib.ref_pid_sequence.argtypes = [C.c_double] * 4 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.ref_pid_sequence.restype = C.c_double
    _REF = lib
    return lib


_REF_UNC = None


def load_ref_unc():
    """The reference's own simple_uncertainty_models.hpp compiled unmodified against the stand-in arc_utilities headers
    (None when oracle/_ref was never built)."""
    global _REF_UNC
    if _REF_UNC is not None:
        return _REF_UNC
    so = os.path.join(_HERE, "_ref", "libref_unc.so")
    if not os.path.exists(so):
        return None
    lib = C.CDLL(so)
    lib.ref_sensor_values.argtypes = [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ref_sensor_values.restype = None
    lib.ref_actuator_values.argtypes = [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ref_actuator_values.restype = None
    lib.ref_actuator_max_velocity_noise.argtypes = [C.c_double, C.c_double]
    lib.ref_actuator_max_velocity_noise.restype = C.c_double
    _REF_UNC = lib
    return lib


_REF_ROB = None


def load_ref_rob():
    """The reference's own simple_robot_models.hpp compiled unmodified against the stand-in Eigen / arc_utilities headers of
    oracle/ref_standins (None when oracle/_ref was never built)."""
    global _REF_ROB
    if _REF_ROB is not None:
        return _REF_ROB
    so = os.path.join(_HERE, "_ref", "libref_rob.so")
    if not os.path.exists(so):
        return None
    lib = C.CDLL(so)
    rp, vp, ci, cd = C.POINTER(abi.RobotDesc), C.c_void_p, C.c_int, C.c_double
    lib.ref_rob_sequence.argtypes = [rp, ci, ci, cd, ci, vp, vp, vp, vp, vp]
    lib.ref_rob_kinematics.argtypes = [rp, ci, vp, vp, vp]
    lib.ref_rob_distances.argtypes = [rp, ci, vp, vp, vp, vp, ci]
    lib.ref_rob_average.argtypes = [rp, ci, vp, vp]
    lib.ref_state_statistics.argtypes = [rp, ci, vp, cd, ci, vp, vp, vp, vp]
    _REF_ROB = lib
    return lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def rob_sequence(robot, starts, targets, dt, n_steps, z=None, reference=False):
    """ResetPosition(start); n_steps x {u = GenerateControlAction(target, dt, draws); ApplyControlInput(u * dt, draws)} for every
    (start, target) row.  z = standard draws [n_seq][n_steps][2][dof] in standard deviations (None = the noise-free forms).
    reference=True runs the reference's own robot classes (oracle/_ref/libref_rob.so), else the oracle's.
    Returns (controls [n_seq][n_steps][dof], configs [n_seq][n_steps][width])."""
    starts, targets = _f64(starts), _f64(targets)
    n_seq, width = starts.shape
    d = robot.desc
    noisy = z is not None
    z = _f64(z if noisy else np.zeros((n_seq, n_steps, 2, d.dof)))
    assert z.shape == (n_seq, n_steps, 2, d.dof)
    controls = np.zeros((n_seq, n_steps, d.dof))
    configs = np.zeros((n_seq, n_steps, width))
    if reference:
        rc = load_ref_rob().ref_rob_sequence(C.byref(d), n_seq, n_steps, dt, int(noisy), starts.ctypes.data, targets.ctypes.data, z.ctypes.data,
                                             controls.ctypes.data, configs.ctypes.data)
        assert rc == 0
    else:
        # the tape's draws: sensor z * max_sensor_noise / 2, actuator z / 2
        bounds = np.array([d.axis[a].max_sensor_noise for a in range(d.dof)])
        sensor = _f64(z[:, :, 0, :] * (np.abs(bounds) * 0.5))
        actuator = _f64(z[:, :, 1, :] * 0.5)
        _check(load().oracle_rob_sequence(C.byref(d), n_seq, n_steps, dt, int(noisy), starts.ctypes.data, targets.ctypes.data, sensor.ctypes.data,
                                          actuator.ctypes.data, controls.ctypes.data, configs.ctypes.data))
    return controls, configs


def rob_kinematics(robot, configs, reference=False):
    """(link transforms [n][links][3][4], translational point Jacobians [n][points][3][dof]) at every configuration"""
    configs = _f64(configs)
    d = robot.desc
    n = configs.shape[0]
    transforms = np.zeros((n, d.num_links, 3, 4))
    jac = np.zeros((n, d.num_points, 3, d.dof))
    if reference:
        assert load_ref_rob().ref_rob_kinematics(C.byref(d), n, configs.ctypes.data, transforms.ctypes.data, jac.ctypes.data) == 0
    else:
        _check(load().oracle_rob_kinematics(C.byref(d), n, configs.ctypes.data, transforms.ctypes.data, jac.ctypes.data))
    return transforms, jac


def rob_distances(robot, a, b, reference=False):
    """(ComputeConfigurationDistance, ComputePerDimensionConfigurationDistance) of the row pairs of a and b"""
    a, b = _f64(a), _f64(b)
    d = robot.desc
    n = a.shape[0]
    dim_width = load().upc_variance_width(d.kind, d.dof)
    dist = np.zeros(n)
    dims = np.zeros((n, dim_width))
    if reference:
        assert load_ref_rob().ref_rob_distances(C.byref(d), n, a.ctypes.data, b.ctypes.data, dist.ctypes.data, dims.ctypes.data, dim_width) == 0
    else:
        _check(load().oracle_rob_distances(C.byref(d), n, a.ctypes.data, b.ctypes.data, dist.ctypes.data, dims.ctypes.data))
    return dist, dims


def rob_average(robot, configs, reference=False):
    configs = _f64(configs)
    out = np.zeros(configs.shape[1])
    if reference:
        assert load_ref_rob().ref_rob_average(C.byref(robot.desc), configs.shape[0], configs.ctypes.data, out.ctypes.data) == 0
    else:
        _check(load().oracle_rob_average(C.byref(robot.desc), configs.shape[0], configs.ctypes.data, out.ctypes.data))
    return out


def ref_state_statistics(robot, configs, step_size):
    """UncertaintyPlannerState(particles)::UpdateStatistics(robot) of the reference (oracle/_ref/libref_rob.so): dict with the keys of one
    cluster of cluster_statistics()"""
    configs = _f64(configs)
    d = robot.desc
    V = load().upc_variance_width(d.kind, d.dof)
    expectation, scalars, variances, si_variances = np.zeros(configs.shape[1]), np.zeros(2), np.zeros(V), np.zeros(V)
    assert load_ref_rob().ref_state_statistics(C.byref(d), configs.shape[0], configs.ctypes.data, float(step_size), V, expectation.ctypes.data, scalars.ctypes.data,
                                               variances.ctypes.data, si_variances.ctypes.data) == 0
    return dict(expectation=expectation, variance=scalars[0], si_variance=scalars[1], variances=variances, si_variances=si_variances)


def sensor_values(process_values, draws):
    p = np.ascontiguousarray(process_values, dtype=np.float64)
    d = np.ascontiguousarray(draws, dtype=np.float64)
    out = np.zeros_like(p)
    load().oracle_sensor_values(p.size, p.ctypes.data, d.ctypes.data, out.ctypes.data)
    return out


def actuator_values(noise_bound, actuator_limit, inputs, draws):
    u = np.ascontiguousarray(inputs, dtype=np.float64)
    d = np.ascontiguousarray(draws, dtype=np.float64)
    out, noiseless = np.zeros_like(u), np.zeros_like(u)
    load().oracle_actuator_values(noise_bound, actuator_limit, u.size, u.ctypes.data, d.ctypes.data, out.ctypes.data, noiseless.ctypes.data)
    return out, noiseless


def _check(code):
    if code != 0:
        raise RuntimeError("oracle error %d: %s" % (code, load().oracle_last_error().decode()))


def philox4x32_10(key, counter):
    k = np.ascontiguousarray(key, dtype=np.uint32)
    c = np.ascontiguousarray(counter, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    load().oracle_philox4x32_10(k.ctypes.data, c.ctypes.data, out.ctypes.data)
    return out


def truncated_normal_quantile(t):
    """z(t) = Phi^-1(1/2 + t (Phi(2) - Phi(-2))) for centred uniforms t in (-1/2, 1/2)"""
    t = np.ascontiguousarray(t, dtype=np.float64)
    out = np.zeros_like(t)
    load().oracle_truncated_normal_quantile(t.size, t.ctypes.data, out.ctypes.data)
    return out


def centered_uniform52(hi, lo):
    hi = np.ascontiguousarray(hi, dtype=np.uint32)
    lo = np.ascontiguousarray(lo, dtype=np.uint32)
    out = np.zeros(hi.size)
    load().oracle_centered_uniform52(hi.size, hi.ctypes.data, lo.ctypes.data, out.ctypes.data)
    return out


def log(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    load().oracle_log(x.size, x.ctypes.data, out.ctypes.data)
    return out


def standard_draws(seed, first, n, step, slot, axis):
    out = np.zeros(n)
    load().oracle_standard_draws(seed, first, n, step, slot, axis, out.ctypes.data)
    return out


def fill_noise_tape_counter(seed, robot, params, n, first_particle=0):
    """The oracle's own restatement of the counter-based noise as a tape for oracle.forward_simulate."""
    tape = np.zeros(abi.tape_shape(params, robot.dof, n), dtype=np.float64)
    _check(load().oracle_fill_noise_tape_counter(seed, C.byref(robot.desc), C.byref(params), first_particle, n, n, tape.ctypes.data))
    return tape


def num_threads():
    return int(load().oracle_num_threads())


def set_num_threads(n):
    load().oracle_set_num_threads(int(n))


def pid_sequence(kp, ki, kd, clamp, errors, dts, lib=None, fn="oracle_pid_sequence"):
    errors = np.ascontiguousarray(errors, dtype=np.float64)
    dts = np.ascontiguousarray(dts, dtype=np.float64)
    out = np.zeros_like(errors)
    f = getattr(lib or load(), fn)
    f(kp, ki, kd, clamp, errors.ctypes.data, dts.ctypes.data, len(errors), out.ctypes.data)
    return out


def ref_pid_sequence(kp, ki, kd, clamp, errors, dts):
    return pid_sequence(kp, ki, kd, clamp, errors, dts, lib=load_ref_pid(), fn="ref_pid_sequence")


def math_fn(which, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(a if b is None else b, dtype=np.float64)
    out = np.zeros_like(a)
    load().oracle_math(which, a.size, a.ctypes.data, b.ctypes.data, out.ctypes.data)
    return out


def se3_log(cfg12):
    cfg12 = np.ascontiguousarray(cfg12, dtype=np.float64)
    out = np.zeros(6)
    load().oracle_se3_log(cfg12.ctypes.data, out.ctypes.data)
    return out


def se3_exp(twist6):
    twist6 = np.ascontiguousarray(twist6, dtype=np.float64)
    out = np.zeros(12)
    load().oracle_se3_exp(twist6.ctypes.data, out.ctypes.data)
    return out


def quaternion_from_rotation(cfg12):
    cfg12 = np.ascontiguousarray(cfg12, dtype=np.float64)
    out = np.zeros(4)
    load().oracle_quaternion_from_rotation(cfg12.ctypes.data, out.ctypes.data)
    return out


def forward_simulate(env, robot, params, starts, targets, tape):
    """starts [n, C], targets [1 or n, C], tape [S, 1+M, dof, n] -> (configs [n, C], collided [n] bool, stats dict)"""
    starts = np.asarray(starts, dtype=np.float64)
    targets = np.asarray(targets, dtype=np.float64)
    n = starts.shape[0]
    s_soa = np.ascontiguousarray(starts.T)
    t_soa = np.ascontiguousarray(targets.T)
    tape = np.ascontiguousarray(tape, dtype=np.float64)
    assert tape.shape == abi.tape_shape(params, robot.dof, n)
    out = np.zeros_like(s_soa)
    coll = np.zeros(n, dtype=np.uint8)
    stats = abi.Stats()
    _check(load().oracle_forward_simulate(C.byref(env.desc), C.byref(robot.desc), C.byref(params), n, s_soa.ctypes.data,
                                          targets.shape[0], t_soa.ctypes.data, tape.ctypes.data, out.ctypes.data,
                                          coll.ctypes.data, C.byref(stats)))
    return np.ascontiguousarray(out.T), coll.astype(bool), stats.as_dict()


def forward_simulate_traced(env, robot, params, starts, targets, tape):
    """forward_simulate plus the per-interval record: dict(configs [S, n, C], controls [S, n, 2*dof], steps [2, n])"""
    starts = np.asarray(starts, dtype=np.float64)
    targets = np.asarray(targets, dtype=np.float64)
    n, width = starts.shape
    s_soa = np.ascontiguousarray(starts.T)
    t_soa = np.ascontiguousarray(targets.T)
    tape = np.ascontiguousarray(tape, dtype=np.float64)
    out = np.zeros_like(s_soa)
    coll = np.zeros(n, dtype=np.uint8)
    S = params.num_steps
    tc, tu, ts = np.zeros((S, width, n)), np.zeros((S, 2 * robot.dof, n)), np.zeros((2, n), dtype=np.int32)
    _check(load().oracle_forward_simulate_traced(C.byref(env.desc), C.byref(robot.desc), C.byref(params), n, s_soa.ctypes.data, targets.shape[0],
                                                 t_soa.ctypes.data, tape.ctypes.data, out.ctypes.data, coll.ctypes.data, tc.ctypes.data,
                                                 tu.ctypes.data, ts.ctypes.data))
    return np.ascontiguousarray(out.T), coll.astype(bool), dict(configs=np.ascontiguousarray(tc.transpose(0, 2, 1)),
                                                                controls=np.ascontiguousarray(tu.transpose(0, 2, 1)), steps=ts)


def check_config_collision(env, robot, config, contact_distance=0.0, inflation=0.0):
    cfg = np.ascontiguousarray(config, dtype=np.float64)
    out = C.c_int32(0)
    _check(load().oracle_check_config_collision(C.byref(env.desc), C.byref(robot.desc), cfg.ctypes.data, contact_distance,
                                                inflation, C.byref(out)))
    return bool(out.value)


def cluster_particles(env, robot, cparams, configs, n_sets=1):
    cfg = np.asarray(configs, dtype=np.float64)
    total = cfg.shape[0]
    set_size = total // n_sets
    soa = np.ascontiguousarray(cfg.T)
    labels = np.zeros(total, dtype=np.int32)
    ncl = np.zeros(n_sets, dtype=np.int32)
    stats = abi.Stats()
    _check(load().oracle_cluster_particles(C.byref(env.desc), C.byref(robot.desc), C.byref(cparams), n_sets, set_size,
                                           soa.ctypes.data, labels.ctypes.data, ncl.ctypes.data, C.byref(stats)))
    return labels, ncl, stats.as_dict()


def pair_distances(robot, configs, reference_style=False):
    cfg = np.asarray(configs, dtype=np.float64)
    n = cfg.shape[0]
    soa = np.ascontiguousarray(cfg.T)
    out = np.zeros((n, n))
    fn = load().oracle_pair_distances_reference_style if reference_style else load().oracle_pair_distances
    _check(fn(C.byref(robot.desc), n, soa.ctypes.data, out.ctypes.data))
    return out


def region_signatures(env, robot, configs):
    cfg = np.asarray(configs, dtype=np.float64)
    n = cfg.shape[0]
    soa = np.ascontiguousarray(cfg.T)
    out = np.zeros((n, robot.num_points), dtype=np.uint32)
    _check(load().oracle_region_signatures(C.byref(env.desc), C.byref(robot.desc), n, soa.ctypes.data, out.ctypes.data))
    return out


def sdf_lookup(env, points):
    pts = np.ascontiguousarray(points, dtype=np.float64)
    n = pts.shape[0]
    d = np.zeros(n)
    g = np.zeros((n, 3))
    _check(load().oracle_sdf_lookup(C.byref(env.desc), n, pts.ctypes.data, d.ctypes.data, g.ctypes.data))
    return d, g


def cluster_statistics(robot, configs, labels, collided, allow_contacts, step_size, n_clusters):
    cfg = np.asarray(configs, dtype=np.float64)
    n = cfg.shape[0]
    K, W, V = int(n_clusters), robot.width, abi.variance_width(robot.kind, robot.dof)
    soa = np.ascontiguousarray(cfg.T)
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    col = np.ascontiguousarray(np.asarray(collided).astype(np.uint8))
    out = dict(counts=np.zeros(K, dtype=np.int64), any_collided=np.zeros(K, dtype=np.uint8), expectations=np.zeros((K, W)),
               variance=np.zeros(K), si_variance=np.zeros(K), variances=np.zeros((K, V)), si_variances=np.zeros((K, V)))
    _check(load().oracle_cluster_statistics(C.byref(robot.desc), n, soa.ctypes.data, lab.ctypes.data, col.ctypes.data, 1 if allow_contacts else 0,
                                            float(step_size), K, out["counts"].ctypes.data, out["any_collided"].ctypes.data,
                                            out["expectations"].ctypes.data, out["variance"].ctypes.data, out["si_variance"].ctypes.data,
                                            out["variances"].ctypes.data, out["si_variances"].ctypes.data))
    out["any_collided"] = out["any_collided"].astype(bool)
    return out


def nearest_neighbors(robot, expectations, feasibility_weight, variance_weight, queries, step_size, use_for_nn=None):
    exp = np.asarray(expectations, dtype=np.float64).reshape(-1, robot.width)
    qs = np.asarray(queries, dtype=np.float64).reshape(-1, robot.width)
    ns, nq = exp.shape[0], qs.shape[0]
    fw = np.ascontiguousarray(feasibility_weight, dtype=np.float64)
    vw = np.ascontiguousarray(variance_weight, dtype=np.float64)
    use = None if use_for_nn is None else np.ascontiguousarray(np.asarray(use_for_nn).astype(np.uint8))
    esoa, qsoa = np.ascontiguousarray(exp.T), np.ascontiguousarray(qs.T)
    idx = np.zeros(nq, dtype=np.int64)
    dist = np.zeros(nq, dtype=np.float64)
    _check(load().oracle_nearest_neighbors(C.byref(robot.desc), ns, esoa.ctypes.data, fw.ctypes.data, vw.ctypes.data,
                                           None if use is None else use.ctypes.data, nq, qsoa.ctypes.data, float(step_size),
                                           idx.ctypes.data, dist.ctypes.data))
    return idx, dist


def complete_link(matrix, threshold):
    m = np.ascontiguousarray(matrix, dtype=np.float64)
    n = m.shape[0]
    labels = np.zeros(n, dtype=np.int32)
    ncl = C.c_int32(0)
    _check(load().oracle_complete_link(m.ctypes.data, n, threshold, labels.ctypes.data, C.byref(ncl)))
    return labels, int(ncl.value)

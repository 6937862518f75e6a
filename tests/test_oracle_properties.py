"""Invariants lifted from the reference's live asserts checked on the oracle, plus
independent numpy restatements of the self-specified pieces (trilinear SDF, complete link)."""
import itertools
import math

import numpy as np
import pytest

import cases
from common import canonical_partition
from uncertainty_planning_core_b200 import abi, scenarios


def _trilinear_numpy(env, p):
    g = np.array(env.desc.grid_from_world[:]).reshape(3, 4)
    u = (g[:, :3] @ p + g[:, 3]) / env.resolution
    shape = np.array(env.sdf.shape)
    if np.any(u < 0) or np.any(u >= shape):
        return float(env.desc.sdf_oob_value)
    c = u - 0.5
    i0 = np.floor(c).astype(int)
    f = c - i0
    lo = np.clip(i0, 0, shape - 1)
    hi = np.clip(i0 + 1, 0, shape - 1)
    acc = 0.0
    for dx, dy, dz in itertools.product((0, 1), repeat=3):
        w = (f[0] if dx else 1 - f[0]) * (f[1] if dy else 1 - f[1]) * (f[2] if dz else 1 - f[2])
        acc += w * float(env.sdf[hi[0] if dx else lo[0], hi[1] if dy else lo[1], hi[2] if dz else lo[2]])
    return acc


def test_sdf_trilinear_matches_numpy_and_gradient_matches_finite_differences(oracle):
    sc = cases.se3_contact()
    env = sc["env"]
    rng = np.random.default_rng(5)
    pts = rng.uniform(-0.2, 4.2, size=(400, 3))
    d, g = oracle.sdf_lookup(env, pts)
    want = np.array([_trilinear_numpy(env, p) for p in pts])
    assert np.max(np.abs(d - want)) < 1e-6
    inside = (d < 100)
    eps = 1e-5
    checked = 0
    for p, gv in zip(pts[inside][:100], g[inside][:100]):
        u = p / env.resolution - 0.5
        if np.min(np.abs(u - np.round(u))) < 0.01:
            continue  # finite differences straddle an interpolation-cell border
        fd = np.array([(oracle.sdf_lookup(env, (p + eps * e)[None])[0][0] - oracle.sdf_lookup(env, (p - eps * e)[None])[0][0]) / (2 * eps)
                       for e in np.eye(3)])
        assert np.max(np.abs(fd - gv)) < 1e-4
        checked += 1
    assert checked > 30
    # out of bounds: oob distance, zero gradient
    d, g = oracle.sdf_lookup(env, np.array([[-1.0, 1.0, 1.0], [1.0, 1.0, 4.1]]))
    assert np.all(d == float(env.desc.sdf_oob_value)) and np.all(g == 0.0)


def test_complete_link_against_brute_force_definition(oracle):
    rng = np.random.default_rng(6)
    for trial in range(30):
        n = int(rng.integers(2, 18))
        x = rng.uniform(0, 1, size=(n, 2))
        D = np.linalg.norm(x[:, None] - x[None], axis=2)
        thr = float(rng.uniform(0.1, 0.6))
        labels, k = oracle.complete_link(D, thr)
        # brute-force agglomeration with the DESIGN.md tie-break
        clusters = [[i] for i in range(n)]
        while len(clusters) > 1:
            best = None
            for a in range(len(clusters)):
                for b in range(a + 1, len(clusters)):
                    link = max(D[i, j] for i in clusters[a] for j in clusters[b])
                    if best is None or link < best[0]:
                        best = (link, a, b)
            if best[0] > thr:
                break
            clusters[best[1]] = sorted(clusters[best[1]] + clusters[best[2]])
            del clusters[best[2]]
        assert canonical_partition(labels) == sorted(tuple(c) for c in clusters)
        assert k == len(clusters)
        # every cluster has diameter <= thr, and labels are numbered by smallest member
        for c in clusters:
            assert max(D[i, j] for i in c for j in c) <= thr
        firsts = [min(np.nonzero(labels == l)[0]) for l in range(k)]
        assert firsts == sorted(firsts)


def test_cluster_particles_edge_cases_and_partition_invariant(oracle):
    c = cases.cl_se2_random()
    env, robot, cp = c["env"], c["robot"], c["cparams"]
    # 0 particles -> no clusters; 1 particle -> itself (uncertainty_contact_planning.hpp:1989-1996)
    labels, ncl, _ = oracle.cluster_particles(env, robot, cp, np.zeros((0, 3)))
    assert ncl[0] == 0
    labels, ncl, _ = oracle.cluster_particles(env, robot, cp, c["configs"][:1])
    assert ncl[0] == 1 and labels[0] == 0
    labels, ncl, stats = oracle.cluster_particles(env, robot, cp, c["configs"])
    # partition: every index exactly once (the assert at :2028)
    assert sorted(i for grp in canonical_partition(labels) for i in grp) == list(range(len(labels)))
    assert set(labels.tolist()) == set(range(int(ncl[0])))
    D = oracle.pair_distances(robot, c["configs"])
    for grp in canonical_partition(labels):
        assert D[np.ix_(grp, grp)].max() <= cp.distance_clustering_threshold
    # max <= threshold => a single cluster and no hierarchical call (:1964-1968)
    s = cases.cl_se2_single_blob()
    labels, ncl, stats = oracle.cluster_particles(s["env"], s["robot"], s["cparams"], s["configs"])
    assert ncl[0] == 1 and stats["cluster_fallback_calls"] == 0


def test_blob_inputs_cluster_into_their_blobs(oracle):
    for kind, case in ((abi.ROBOT_SE2, cases.cl_se2_blobs), (abi.ROBOT_SE3, cases.cl_se3_blobs), (abi.ROBOT_LINKED, cases.cl_linked_blobs)):
        c = case()
        labels, ncl, _ = oracle.cluster_particles(c["env"], c["robot"], c["cparams"], c["configs"])
        D = oracle.pair_distances(c["robot"], c["configs"])
        same = labels[:, None] == labels[None, :]
        assert D[same].max() <= 0.1 and D[~same].min() > 0.1


def test_reference_style_distance_matrix_is_identical(oracle):
    c = cases.cl_se3_random()
    assert np.array_equal(oracle.pair_distances(c["robot"], c["configs"]),
                          oracle.pair_distances(c["robot"], c["configs"], reference_style=True))


def test_actuator_bound_and_angle_range_after_simulation(oracle):
    sc = cases.se2_pid_sensor_noise()
    n = sc["starts"].shape[0]
    tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], n)
    # tape bounds: sensor draws within +-max_sensor_noise, actuator draws within [-1, 1]
    assert np.max(np.abs(tape[:, 0, :2, :])) <= 0.01 and np.max(np.abs(tape[:, 0, 2, :])) <= 0.02
    assert np.max(np.abs(tape[:, 1:, :, :])) <= 1.0
    assert abs(float(np.mean(tape[:, 1, :, :]))) < 0.02 and 0.35 < float(np.std(tape[:, 1, :, :])) < 0.5
    cfg, coll, stats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
    assert np.all(cfg[:, 2] > -math.pi) and np.all(cfg[:, 2] <= math.pi)      # simple_robot_models.hpp:115
    # per controller interval a coordinate moves at most limit*(1+noise_bound)*dt (simple_uncertainty_models.hpp:72-86)
    bound = 0.9 * 1.2 * 0.01 * sc["params"].num_steps
    assert np.max(np.abs(cfg[:, :2] - sc["starts"][:, :2])) <= bound + 0.2     # + contact corrections


def test_zero_noise_free_space_reaches_target(oracle):
    for make in (scenarios.c1_se2, scenarios.c2_se3):
        kw = dict(n=4, num_steps=200)
        kw.update(dict(shape=(64, 64, 1), resolution=0.16) if make is scenarios.c1_se2 else dict(shape=(32, 32, 32), resolution=0.125))
        sc = make(**kw)
        env = abi.Environment(np.full(sc["env"].sdf.shape, 5.0, dtype=np.float32), sc["env"].resolution, sc["env"].desc.grid_from_world[:])
        robot = scenarios.se2_robot(max_actuator_noise=0.0) if make is scenarios.c1_se2 else scenarios.se3_robot(max_actuator_noise=0.0)
        tape = np.zeros(abi.tape_shape(sc["params"], robot.dof, 4))
        cfg, coll, _ = oracle.forward_simulate(env, robot, sc["params"], sc["starts"], sc["targets"], tape)
        assert not coll.any()
        assert np.max(np.abs(cfg - sc["targets"])) < 2e-3
    sc = cases.linked_contact()
    env = abi.Environment(np.full(sc["env"].sdf.shape, 5.0, dtype=np.float32), sc["env"].resolution)
    robot = scenarios.linked_arm(max_actuator_noise=0.0)
    sc["params"].num_steps = 200
    tape = np.zeros(abi.tape_shape(sc["params"], 7, 3))
    cfg, coll, _ = oracle.forward_simulate(env, robot, sc["params"], sc["starts"][:3], sc["targets"], tape)
    assert not coll.any() and np.max(np.abs(cfg - sc["targets"])) < 2e-3


def test_contacts_are_resolved_out_of_collision(oracle):
    for make in (cases.se2_contact, cases.se3_contact, cases.linked_contact):
        sc = make()
        n = sc["starts"].shape[0]
        tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], n)
        cfg, coll, stats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
        assert coll.any() and stats["resolver_iterations"] > 0
        for i in range(n):
            assert not oracle.check_config_collision(sc["env"], sc["robot"], cfg[i], sc["params"].contact_distance)
        # the same edge with contacts disallowed: collided particles stay at a collision-free configuration too
        sc["params"].allow_contacts = 0
        cfg2, coll2, stats2 = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
        assert stats2["resolver_iterations"] == 0 and coll2.any()
        for i in range(0, n, 7):
            assert not oracle.check_config_collision(sc["env"], sc["robot"], cfg2[i], sc["params"].contact_distance)


def test_results_do_not_depend_on_thread_count(oracle):
    sc = cases.se3_contact()
    n = sc["starts"].shape[0]
    tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], n)
    before = oracle.num_threads()
    try:
        oracle.set_num_threads(1)
        a = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
        oracle.set_num_threads(max(2, before))
        b = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
    finally:
        oracle.set_num_threads(before)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # tapes are keyed by particle: a shard of the batch sees the same draws
    shard = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], 10, first_particle=20)
    assert np.array_equal(shard, tape[..., 20:30])


def test_cluster_statistics_properties(oracle):
    """per-cluster statistics (uncertainty_planner_state.hpp:591-714): means, contact filter, single-particle cases"""
    c = cases.cl_se2_blobs()
    cfg = c["configs"]
    labels, ncl, _ = oracle.cluster_particles(c["env"], c["robot"], c["cparams"], cfg)
    K = int(ncl[0])
    assert K >= 2
    collided = np.zeros(len(cfg), dtype=bool)
    collided[labels == 1] = True
    st = oracle.cluster_statistics(c["robot"], cfg, labels, collided, True, 0.25, K)
    assert st["counts"].sum() == len(cfg) and st["any_collided"][1] and st["any_collided"].sum() == 1
    for k in range(K):
        m = cfg[labels == k]
        assert np.allclose(st["expectations"][k][:2], m[:, :2].mean(axis=0), atol=1e-12)
        assert abs(st["expectations"][k][2] - m[:, 2].mean()) < 1e-3           # circular mean of a tight blob
        assert np.isclose(st["si_variance"][k], st["variance"][k] / 0.25 ** 2, rtol=1e-12)
        assert np.allclose(st["variances"][k][:2], ((m[:, :2] - st["expectations"][k][:2]) ** 2).mean(axis=0), rtol=1e-9, atol=1e-15)
    st2 = oracle.cluster_statistics(c["robot"], cfg, labels, collided, False, 0.25, K)
    assert st2["counts"][1] == 0 and not st2["any_collided"].any() and np.all(st2["expectations"][1] == 0.0)
    # one particle per cluster: expectation is the particle, variances vanish
    one = oracle.cluster_statistics(c["robot"], cfg[:4], np.arange(4, dtype=np.int32), np.zeros(4, dtype=bool), True, 0.25, 4)
    assert np.array_equal(one["expectations"], cfg[:4]) and not one["variance"].any() and not one["variances"].any()
    # SE3: expectation of a cluster is a rotation matrix with the mean translation
    c3 = cases.cl_se3_blobs()
    l3, n3, _ = oracle.cluster_particles(c3["env"], c3["robot"], c3["cparams"], c3["configs"])
    s3 = oracle.cluster_statistics(c3["robot"], c3["configs"], l3, np.zeros(len(l3), dtype=bool), True, 0.25, int(n3[0]))
    for k in range(int(n3[0])):
        if s3["counts"][k] < 2:
            continue
        R = s3["expectations"][k][:9].reshape(3, 3)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.linalg.det(R) > 0.999
        assert np.allclose(s3["expectations"][k][9:], c3["configs"][l3 == k][:, 9:].mean(axis=0), atol=1e-12)


def test_nearest_neighbors_against_numpy(oracle):
    """StateDistance / GetNearestNeighbor (uncertainty_contact_planning.hpp:1488-1542) restated in numpy for SE(2)"""
    c = cases.cl_se2_random()
    rng = np.random.default_rng(5)
    states = c["configs"]
    ns = len(states)
    fw, vw = rng.uniform(0.3, 1.0, ns), rng.uniform(0.3, 1.0, ns)
    use = rng.random(ns) < 0.7
    queries = np.stack([rng.uniform(1.0, 1.6, 40), rng.uniform(1.0, 1.5, 40), rng.uniform(-3.0, 3.0, 40)], axis=1)
    idx, dist = oracle.nearest_neighbors(c["robot"], states, fw, vw, queries, 0.125, use)
    for q in range(40):
        d = states - queries[q]
        ang = np.abs((d[:, 2] + np.pi) % (2 * np.pi) - np.pi)
        ed = (np.hypot(d[:, 0], d[:, 1]) + ang) / 0.125
        full = np.where(use, fw * ed * vw, np.inf)
        assert idx[q] == int(np.argmin(full))
        assert np.isclose(dist[q], full[idx[q]], rtol=1e-12)
    none, dn = oracle.nearest_neighbors(c["robot"], states, fw, vw, queries, 0.125, np.zeros(ns, dtype=bool))
    assert (none == -1).all() and np.isinf(dn).all()

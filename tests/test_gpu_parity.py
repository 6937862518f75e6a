"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs and against the committed golden fixtures.

Stated bar: collision flags, solver counters and cluster labels bit-exact; configurations within 0 ulp
(FP64 everywhere, nvcc -fmad=false vs g++ -ffp-contract=off, shared elementary functions)."""
import os

import numpy as np
import pytest

import cases
from common import canonical_partition, max_ulp_diff
from uncertainty_planning_core_b200 import abi, scenarios

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CONFIG_ULP_TOLERANCE = 0


@pytest.fixture(scope="module")
def simulators():
    from uncertainty_planning_core_b200 import B200Simulator
    cache = {}

    def get(key, env, robot):
        if key not in cache:
            cache[key] = B200Simulator(env, robot)
        return cache[key]
    yield get
    for s in cache.values():
        s.close()


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_SIM_CASES))
def test_forward_simulation_matches_oracle_for_every_lane_count(oracle, simulators, name):
    sc = cases.GOLDEN_SIM_CASES[name]()
    n = sc["starts"].shape[0]
    tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], n)
    ocfg, ocoll, ostats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
    g = np.load(os.path.join(HERE, "golden", "sim_%s.npz" % name))
    assert np.array_equal(ocfg, g["configs"]) and np.array_equal(ocoll, g["collided"])
    sim = simulators(name, sc["env"], sc["robot"])
    for lanes in (0, 1, 2, 3, 4, 5, 6, 8, 16, 32):
        sim.SetLanesPerParticle(lanes)
        sim.ResetStatistics()
        cfg, coll = sim.ForwardSimulateRobots(sc["starts"], sc["targets"], tape, sc["params"])
        assert np.array_equal(coll, ocoll), (name, lanes)
        assert max_ulp_diff(cfg, ocfg) <= CONFIG_ULP_TOLERANCE, (name, lanes, np.abs(cfg - ocfg).max())
        stats = sim.GetStatistics()
        for k in cases.SIM_STAT_KEYS:
            assert int(stats[k]) == ostats[k], (name, lanes, k)
        assert int(stats["particles_simulated"]) == n and stats["kernel_launches"] >= 1
    sim.SetLanesPerParticle(0)


@pytest.mark.parametrize("pivot_pass", ["forced", "auto", "off"])
@pytest.mark.parametrize("name", sorted(cases.GOLDEN_CLUSTER_CASES))
def test_clustering_matches_oracle(oracle, simulators, name, pivot_pass, monkeypatch):
    """every golden clustering case through the default pipeline (pivot pass first) and with the pivot pass
    forced off (adjacency / root-check / exact paths only); "forced" tries the pivot pass on these small sets too"""
    from uncertainty_planning_core_b200 import OutcomeClustering
    if pivot_pass == "off":
        monkeypatch.setenv("UPC_B200_NO_PIVOT_PASS", "1")
    elif pivot_pass == "forced":
        monkeypatch.setenv("UPC_B200_FORCE_PIVOT_PASS", "1")
    c = cases.GOLDEN_CLUSTER_CASES[name]()
    n_sets = c.get("n_sets", 1)
    olabels, oncl, ostats = oracle.cluster_particles(c["env"], c["robot"], c["cparams"], c["configs"], n_sets)
    g = np.load(os.path.join(HERE, "golden", "cluster_%s.npz" % name))
    assert np.array_equal(olabels, g["labels"]) and np.array_equal(oncl, g["n_clusters"])
    sim = simulators("cl_" + name, c["env"], c["robot"])
    ptpm = c.get("ptpm")
    oc = OutcomeClustering(sim, c["cparams"].feature_type, c["cparams"].signature_matching_threshold,
                           c["cparams"].distance_clustering_threshold,
                           goal_distance_threshold=c["cparams"].goal_distance_threshold,
                           ptpm_params=ptpm[0] if ptpm else None, ptpm_tape=ptpm[1] if ptpm else None)
    labels, ncl = oc.ClusterIndices(c["configs"], n_sets)
    assert np.array_equal(ncl, oncl), name
    assert np.array_equal(labels, olabels), name


def test_pivot_pass_boundaries(oracle, simulators, monkeypatch):
    """the pivot pass may only settle what the adjacency pass would settle identically - blobs whose radius sits just
    inside / outside threshold/2, pivots just inside / outside 2*threshold, more clusters than pivots, several sets per call"""
    from uncertainty_planning_core_b200 import OutcomeClustering
    monkeypatch.setenv("UPC_B200_FORCE_PIVOT_PASS", "1")
    c = cases.cl_se2_blobs()
    sim = simulators("cl_pivot", c["env"], c["robot"])
    thr = 0.1
    oc = OutcomeClustering(sim, abi.FEATURE_NONE, 0.125, thr)
    cp = abi.ClusterParams()
    cp.feature_type = abi.FEATURE_NONE
    cp.distance_clustering_threshold = thr
    rng = np.random.default_rng(23)

    def ring(centre, radius, m):
        ang = rng.uniform(0, 2 * np.pi, m)
        return np.stack([centre[0] + radius * np.cos(ang), centre[1] + radius * np.sin(ang), np.zeros(m)], axis=1)

    sets = []
    for radius in (0.049, 0.0499995, 0.0500005, 0.051, 0.08):                 # points on a circle around the pivot (first particle)
        sets.append(np.concatenate([[[1.0, 1.0, 0.0]], ring((1.0, 1.0), radius, 199)], axis=0))
    for gap in (0.19, 0.1999995, 0.2000005, 0.21, 0.3):                       # two tight blobs whose pivots are `gap` apart
        a = np.concatenate([[[1.0, 1.0, 0.0]], ring((1.0, 1.0), 0.01, 99)], axis=0)
        b = np.concatenate([[[1.0 + gap, 1.0, 0.0]], ring((1.0 + gap, 1.0), 0.01, 99)], axis=0)
        sets.append(np.concatenate([a, b], axis=0))
    many = np.stack([1.0 + 0.5 * np.arange(200), np.full(200, 1.0), np.zeros(200)], axis=1)   # 200 singleton clusters > 64 pivots
    sets.append(many)
    # wide blobs whose smallest member sits on the rim: a filled disc (a central member exists: settled through the centre search),
    # the same disc with diameter above the threshold, and an empty ring (no member is central)
    for rad, filled in ((0.044, True), (0.047, True), (0.06, True), (0.044, False)):
        m = 199
        ang = rng.uniform(0, 2 * np.pi, m)
        rr = rad * np.sqrt(rng.uniform(0, 1, m)) if filled else np.full(m, rad)
        disc = np.stack([1.0 + rr * np.cos(ang), 1.0 + rr * np.sin(ang), np.zeros(m)], axis=1)
        sets.append(np.concatenate([[[1.0 + rad, 1.0, 0.0]], disc], axis=0))
    for cfg in sets:
        labels, ncl = oc.ClusterIndices(cfg)
        olabels, oncl, _ = oracle.cluster_particles(c["env"], c["robot"], cp, cfg)
        assert np.array_equal(ncl, oncl) and np.array_equal(labels, olabels)
    batch = np.concatenate(sets, axis=0)
    labels, ncl = oc.ClusterIndices(batch, n_sets=len(sets))
    olabels, oncl, _ = oracle.cluster_particles(c["env"], c["robot"], cp, batch, len(sets))
    assert np.array_equal(ncl, oncl) and np.array_equal(labels, olabels)


def test_cluster_particles_assembly_and_edge_cases(oracle, simulators):
    from uncertainty_planning_core_b200 import OutcomeClustering
    c = cases.cl_se2_blobs()
    sim = simulators("cl_edge", c["env"], c["robot"])
    oc = OutcomeClustering(sim, abi.FEATURE_NONE, 0.125, 0.1)
    cfg = c["configs"]
    collided = np.zeros(len(cfg), dtype=bool)
    collided[::3] = True
    assert oc.ClusterParticles((cfg[:0], collided[:0]), True) == []
    one = oc.ClusterParticles((cfg[:1], collided[:1]), False)
    assert len(one) == 1 and len(one[0][0]) == 1            # a single particle is returned as is, collided or not
    with_contacts = oc.ClusterParticles((cfg, collided), True)
    without = oc.ClusterParticles((cfg, collided), False)
    assert len(with_contacts) == len(without) == 5
    assert sum(len(k[0]) for k in with_contacts) == len(cfg)
    assert sum(len(k[0]) for k in without) == int((~collided).sum())
    # all members collided: the cluster is emitted empty (uncertainty_contact_planning.hpp:2025)
    labels, _ = oc.ClusterIndices(cfg)
    coll2 = labels == 2
    emptied = oc.ClusterParticles((cfg, coll2), False)
    assert len(emptied) == 5 and len(emptied[2][0]) == 0
    # PolicyParticleClusteringFn: the current configuration joins as the last index
    idx = oc.PolicyParticleClusteringFn(cfg[:40], cfg[41])
    assert sorted(i for grp in idx for i in grp) == list(range(41))
    # two-particle and size-one sets in a batch
    labels, ncl = oc.ClusterIndices(cfg[:6], n_sets=6)
    assert labels.tolist() == [0] * 6 and ncl.tolist() == [1] * 6
    far = np.array([[1.0, 1.0, 0.0], [3.0, 1.0, 0.0]])
    labels, ncl = oc.ClusterIndices(far)
    assert labels.tolist() == [0, 1] and ncl.tolist() == [2]


def test_plain_grid_path_matches_oracle_too(oracle):
    """Grids too large for the corner-cell table (e.g. 256^3) use the plain SDF + cell-minimum cull: force that path."""
    from uncertainty_planning_core_b200 import B200Simulator
    os.environ["UPC_B200_NO_CORNER_CELLS"] = "1"
    try:
        for make in (cases.se2_contact, cases.se3_contact, cases.linked_contact):
            sc = make()
            n = sc["starts"].shape[0]
            tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], n)
            ocfg, ocoll, ostats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
            sim = B200Simulator(sc["env"], sc["robot"])
            for lanes in (1, 4, 32):
                sim.SetLanesPerParticle(lanes)
                cfg, coll = sim.ForwardSimulateRobots(sc["starts"], sc["targets"], tape, sc["params"])
                assert np.array_equal(coll, ocoll) and max_ulp_diff(cfg, ocfg) <= CONFIG_ULP_TOLERANCE
            sim.close()
    finally:
        del os.environ["UPC_B200_NO_CORNER_CELLS"]


def test_sensor_slot_is_not_read_for_noise_free_sensors(oracle):
    """with every max_sensor_noise zero the sensor slot of the tape (exact zeros by construction) is neither
    uploaded nor read - poisoning it changes nothing, on the oracle and on the GPU (host-pointer path, sliced upload)"""
    from uncertainty_planning_core_b200 import B200Simulator, scenarios
    for sc in (scenarios.c2_se3(seed=5, n=4096, num_steps=32, shape=(64, 64, 64), resolution=0.0625), cases.se2_contact(), cases.linked_contact()):
        assert all(sc["robot"].desc.axis[a].max_sensor_noise == 0.0 for a in range(sc["robot"].dof))
        n = sc["starts"].shape[0]
        tape = abi.fill_noise_tape(3, sc["robot"], sc["params"], n)
        assert not tape[:, 0].any()
        poisoned = tape.copy()
        poisoned[:, 0] = np.nan
        ocfg, ocoll, _ = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
        pcfg, pcoll, _ = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], poisoned)
        assert np.array_equal(ocfg, pcfg) and np.array_equal(ocoll, pcoll)
        sim = B200Simulator(sc["env"], sc["robot"])
        cfg, coll = sim.ForwardSimulateRobots(sc["starts"], sc["targets"], poisoned, sc["params"])
        assert np.array_equal(coll, ocoll) and max_ulp_diff(cfg, ocfg) <= CONFIG_ULP_TOLERANCE
        sim.close()


def test_single_precision_pre_cull_changes_nothing(oracle):
    """the float pre-cull of the collision check and of the resolver's point pass is a shortcut only - results with it (default, used when
    a lane holds >= 8 points of a link) and without it are the oracle's, bit for bit; the C2-shaped case has 128 points"""
    from uncertainty_planning_core_b200 import B200Simulator, scenarios
    sc = scenarios.c2_se3(seed=5, n=96, num_steps=48, shape=(64, 64, 64), resolution=0.0625)
    n = sc["starts"].shape[0]
    tape = abi.fill_noise_tape(9, sc["robot"], sc["params"], n)
    ocfg, ocoll, ostats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
    assert ocoll.any() and ostats["resolver_iterations"] > 0
    for off in (False, True):
        if off:
            os.environ["UPC_B200_NO_FP32_CULL"] = "1"
        try:
            sim = B200Simulator(sc["env"], sc["robot"])
        finally:
            os.environ.pop("UPC_B200_NO_FP32_CULL", None)
        # 128 .. 8 points per lane: all take the pre-cull in the collision check when it is on; 2..8 lanes also in the
        # resolver's point pass (mask-then-evaluate, upc_sim.cuh)
        for lanes in (1, 2, 3, 4, 5, 6, 8, 16):
            sim.SetLanesPerParticle(lanes)
            cfg, coll = sim.ForwardSimulateRobots(sc["starts"], sc["targets"], tape, sc["params"])
            assert np.array_equal(coll, ocoll) and max_ulp_diff(cfg, ocfg) <= CONFIG_ULP_TOLERANCE
            st = sim.GetStatistics()
            sim.ResetStatistics()
            assert st["resolver_iterations"] == ostats["resolver_iterations"]
        sim.close()


def test_link_level_cull_changes_nothing(oracle):
    """the link-level cull (one dilated-minimum lookup per link) is a shortcut only - results with it (default) and
    without it (UPC_B200_NO_LINK_CULL=1) are the oracle's, for every robot kind, both SDF layouts and several lane counts"""
    from uncertainty_planning_core_b200 import B200Simulator
    for make in (cases.se2_contact, cases.se3_contact, cases.linked_contact, cases.linked_mixed_joints):
        sc = make()
        n = sc["starts"].shape[0]
        tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], n)
        ocfg, ocoll, ostats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
        for off in (False, True):
            for plain in (False, True):
                if off:
                    os.environ["UPC_B200_NO_LINK_CULL"] = "1"
                if plain:
                    os.environ["UPC_B200_NO_CORNER_CELLS"] = "1"
                try:
                    sim = B200Simulator(sc["env"], sc["robot"])
                finally:
                    os.environ.pop("UPC_B200_NO_LINK_CULL", None)
                    os.environ.pop("UPC_B200_NO_CORNER_CELLS", None)
                for lanes in (1, 4, 32):
                    sim.SetLanesPerParticle(lanes)
                    sim.ResetStatistics()
                    cfg, coll = sim.ForwardSimulateRobots(sc["starts"], sc["targets"], tape, sc["params"])
                    assert np.array_equal(coll, ocoll) and max_ulp_diff(cfg, ocfg) <= CONFIG_ULP_TOLERANCE, (make.__name__, off, plain, lanes)
                    st = sim.GetStatistics()
                    for k in cases.SIM_STAT_KEYS:
                        assert int(st[k]) == ostats[k], (make.__name__, off, plain, lanes, k)
                sim.close()


def test_group_level_cull_changes_nothing(oracle):
    """the group-level cull of the one-lane programs (one dilated-minimum lookup per point group) is a shortcut only -
    results with it (default) and without it (UPC_B200_NO_GROUP_CULL=1) are the oracle's, for every robot kind, both SDF layouts,
    with and without the link-level cull in front of it.  (The shipped build compiles the group tests out - measured slower on
    B200, upc_device.cuh: UPC_GROUP_CULL - so there the switch changes nothing by construction; -DUPC_GROUP_CULL=1 builds are what
    this test is for, next to the emulation parity suite which always compiles them in.)"""
    from uncertainty_planning_core_b200 import B200Simulator, scenarios
    c2 = scenarios.c2_se3(seed=5, n=96, num_steps=48, shape=(64, 64, 64), resolution=0.0625)     # 128 points: four chunks of four groups
    c2["seed"] = 9
    for sc in (cases.se2_contact(), cases.se3_contact(), cases.linked_contact(), cases.linked_mixed_joints(), c2):
        n = sc["starts"].shape[0]
        tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], n)
        ocfg, ocoll, ostats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
        for off in ("", "GROUP", "LINK"):
            for plain in (False, True):
                if off:
                    os.environ["UPC_B200_NO_%s_CULL" % off] = "1"
                if plain:
                    os.environ["UPC_B200_NO_CORNER_CELLS"] = "1"
                try:
                    sim = B200Simulator(sc["env"], sc["robot"])
                finally:
                    os.environ.pop("UPC_B200_NO_GROUP_CULL", None)
                    os.environ.pop("UPC_B200_NO_LINK_CULL", None)
                    os.environ.pop("UPC_B200_NO_CORNER_CELLS", None)
                sim.SetLanesPerParticle(1)
                sim.ResetStatistics()
                cfg, coll = sim.ForwardSimulateRobots(sc["starts"], sc["targets"], tape, sc["params"])
                assert np.array_equal(coll, ocoll) and max_ulp_diff(cfg, ocfg) <= CONFIG_ULP_TOLERANCE, (sc.get("name"), off, plain)
                st = sim.GetStatistics()
                for k in cases.SIM_STAT_KEYS:
                    assert int(st[k]) == ostats[k], (sc.get("name"), off, plain, k)
                sim.close()


def test_check_config_collision_matches_oracle(oracle, simulators):
    for make in (cases.se2_contact, cases.se3_contact, cases.linked_contact):
        sc = make()
        sim = simulators("chk_" + sc["name"], sc["env"], sc["robot"])
        rng = np.random.default_rng(3)
        hits = 0
        for i in range(40):
            cfg = sc["starts"][0] + (sc["targets"][0] - sc["starts"][0]) * rng.uniform(0, 1.6)
            if sc["robot"].kind == abi.ROBOT_SE3:
                cfg[:9] = sc["starts"][0][:9]
            want = oracle.check_config_collision(sc["env"], sc["robot"], cfg, 0.0, 0.0)
            assert sim.CheckConfigCollision(cfg) == want
            infl = oracle.check_config_collision(sc["env"], sc["robot"], cfg, 0.0, 2.0 * sc["env"].resolution)
            assert sim.CheckConfigCollision(cfg, inflation_ratio=2.0) == infl
            hits += int(want)
        assert 0 < hits < 40


def test_cluster_statistics_match_oracle(oracle, simulators):
    """next row 8f-1: counts / any-collided bit-exact, expectations and variances 0 ulp (sequential sums replayed in order)"""
    from uncertainty_planning_core_b200 import OutcomeClustering
    for case in (cases.cl_se2_blobs, cases.cl_se2_wrap, cases.cl_se3_blobs, cases.cl_se3_random, cases.cl_linked_blobs, cases.cl_linked_random_mixed):
        c = case()
        sim = simulators("st_" + case.__name__, c["env"], c["robot"])
        oc = OutcomeClustering(sim, abi.FEATURE_NONE, 0.125, c["cparams"].distance_clustering_threshold)
        cfg = c["configs"]
        labels, ncl = oc.ClusterIndices(cfg)
        rng = np.random.default_rng(9)
        collided = rng.random(len(cfg)) < 0.3
        K = int(ncl[0])
        if K >= 2:
            collided[labels == 1] = True          # a cluster whose members all collided
        for allow in (True, False):
            got = oc.ClusterStatistics(cfg, labels, collided, allow, 0.25, K)
            want = oracle.cluster_statistics(c["robot"], cfg, labels, collided, allow, 0.25, K)
            assert np.array_equal(got["counts"], want["counts"]) and np.array_equal(got["any_collided"], want["any_collided"])
            for key in ("expectations", "variance", "si_variance", "variances", "si_variances"):
                assert max_ulp_diff(got[key], want[key]) == 0, (case.__name__, allow, key)
            keep = allow | ~collided
            assert int(got["counts"].sum()) == int(keep.sum())
            if not allow:
                assert not got["any_collided"].any()
                if K >= 2:
                    assert got["counts"][1] == 0


def test_nearest_neighbors_match_oracle(oracle, simulators):
    """next row 8f-3: GetNearestNeighbor for a batch of samples - indices equal (ties to the smallest index, disabled and
    NaN-weighted states never selected), distances 0 ulp"""
    from uncertainty_planning_core_b200 import B200Simulator
    rng = np.random.default_rng(17)
    for case in (cases.cl_se2_random, cases.cl_se2_wrap, cases.cl_se3_random, cases.cl_linked_random_mixed):
        c = case()
        sim = simulators("nn_" + case.__name__, c["env"], c["robot"])
        states = np.concatenate([c["configs"], c["configs"][:20]], axis=0)      # 20 duplicated states: exact ties
        ns = states.shape[0]
        pf = rng.uniform(0.0, 1.0, ns)
        siv = rng.uniform(0.0, 0.5, size=(ns, abi.variance_width(c["robot"].kind, c["robot"].dof)))
        fw, vw = B200Simulator.StateDistanceWeights(pf, siv, 0.75, 0.75)
        fw[-20:], vw[-20:] = fw[:20], vw[:20]
        use = rng.random(ns) < 0.8
        queries = states[rng.integers(ns, size=64)] + 0.01 * rng.normal(size=(64, states.shape[1])) * (c["robot"].kind != abi.ROBOT_SE3)
        queries[:8] = states[:8]                                                  # distance 0 to two states each
        for mask in (None, use):
            gi, gd = sim.NearestNeighbors(states, fw, vw, queries, 0.25, mask)
            wi, wd = oracle.nearest_neighbors(c["robot"], states, fw, vw, queries, 0.25, mask)
            assert np.array_equal(gi, wi)
            assert max_ulp_diff(gd, wd) == 0
            assert gi.min() >= 0 and (mask is None or mask[gi].all())
        assert np.array_equal(sim.NearestNeighbors(states, fw, vw, queries[:8], 0.25)[0], np.arange(8))   # ties keep the smallest index
        bad = fw.copy()
        bad[::2] = np.nan
        gi, gd = sim.NearestNeighbors(states, bad, vw, queries, 0.25)
        wi, wd = oracle.nearest_neighbors(c["robot"], states, bad, vw, queries, 0.25)
        assert np.array_equal(gi, wi) and (gi % 2 == 1).all() and max_ulp_diff(gd, wd) == 0
        gi, gd = sim.NearestNeighbors(states, fw, vw, queries, 0.25, np.zeros(ns, dtype=bool))
        assert (gi == -1).all() and np.isinf(gd).all()
        gi, gd = sim.NearestNeighbors(states[:0], fw[:0], vw[:0], queries, 0.25)
        assert (gi == -1).all() and np.isinf(gd).all()


def test_edge_batch_equals_edge_by_edge(oracle, simulators):
    """BASELINE config 5 in small: a batch of edges simulated and clustered in one call each gives, edge for edge, what separate
    calls give (particles and particle sets are independent) - and what the oracle gives"""
    from uncertainty_planning_core_b200 import OutcomeClustering, scenarios
    sc = scenarios.c5_batch(edges=12, particles=96, num_steps=40)
    E, P = 12, 96
    sim = simulators("c5_small", sc["env"], sc["robot"])
    tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], E * P)
    cfg, coll = sim.ForwardSimulateEdgeBatch(sc["starts"][::P], sc["targets"][::P], P, tape, sc["params"])
    assert coll.any() and not coll.all()
    oc = OutcomeClustering(sim, abi.FEATURE_CRS, 0.125, 0.1)
    labels, ncl = oc.ClusterIndices(cfg.reshape(E * P, -1), n_sets=E)
    for e in range(E):
        sl = slice(e * P, (e + 1) * P)
        one_cfg, one_coll = sim.ForwardSimulateRobots(sc["starts"][sl], sc["targets"][sl][:1], np.ascontiguousarray(tape[..., sl]), sc["params"])
        assert np.array_equal(one_cfg, cfg[e]) and np.array_equal(one_coll, coll[e])
        one_labels, one_ncl = oc.ClusterIndices(one_cfg)
        assert np.array_equal(one_labels, labels[sl]) and one_ncl[0] == ncl[e]
        ocfg, ocoll, _ = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"][sl], sc["targets"][sl], np.ascontiguousarray(tape[..., sl]))
        assert np.array_equal(ocfg, cfg[e]) and np.array_equal(ocoll, coll[e])
        ol, on, _ = oracle.cluster_particles(sc["env"], sc["robot"], oc.params, ocfg)
        assert np.array_equal(ol, labels[sl]) and on[0] == ncl[e]


def test_count_within_and_reverse_edge_probability(oracle, simulators):
    """next row 8f-2: counts are bit-exact against the oracle's distances (<= threshold, uncertainty_contact_planning.hpp:2083)"""
    for make in (cases.se2_contact, cases.se3_contact, cases.linked_contact):
        sc = make()
        n = sc["starts"].shape[0]
        tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], n)
        sim = simulators("cnt_" + sc["name"], sc["env"], sc["robot"])
        cfg, _ = sim.ForwardSimulateRobots(sc["starts"], sc["targets"], tape, sc["params"])
        both = np.concatenate([cfg, sc["targets"]], axis=0)
        D = oracle.pair_distances(sc["robot"], both)[:n, n]
        thr = float(np.median(D))
        count, flags = sim.CountWithin(cfg, sc["targets"], thr)
        assert np.array_equal(flags, D <= thr) and count == int((D <= thr).sum()) and 0 < count < n
        # per-particle targets
        count2, flags2 = sim.CountWithin(cfg, np.roll(cfg, 1, axis=0), 1e-3)
        Dp = np.array([oracle.pair_distances(sc["robot"], np.stack([cfg[i], cfg[i - 1]]))[0, 1] for i in range(n)])
        assert np.array_equal(flags2, Dp <= 1e-3) and count2 == int((Dp <= 1e-3).sum())
        attempts, reached = sim.ComputeReverseEdgeProbability(cfg, sc["starts"][0], tape, sc["params"], thr)
        ocfg, _, _ = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], cfg, sc["starts"][:1], tape)
        Dr = oracle.pair_distances(sc["robot"], np.concatenate([ocfg, sc["starts"][:1]], axis=0))[:n, n]
        assert attempts == n and reached == int((Dr <= thr).sum())


def test_device_pointer_entry_points_on_the_torch_stream(oracle, simulators):
    import ctypes as C
    import torch
    sc = cases.se3_contact()
    n = sc["starts"].shape[0]
    tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], n)
    ocfg, ocoll, _ = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
    sim = simulators("dev_ptr", sc["env"], sc["robot"])
    dev = torch.device("cuda:0")
    d_starts = torch.from_numpy(np.ascontiguousarray(sc["starts"].T)).to(dev)
    d_targets = torch.from_numpy(np.ascontiguousarray(sc["targets"].T)).to(dev)
    d_tape = torch.from_numpy(tape).to(dev)
    d_out = torch.zeros_like(d_starts)
    d_coll = torch.zeros(n, dtype=torch.uint8, device=dev)
    lib = abi.load_library()
    stream = torch.cuda.current_stream().cuda_stream
    abi.check(lib.upc_forward_simulate_device(sim.handle, C.byref(sc["params"]), n, d_starts.data_ptr(), 1, d_targets.data_ptr(),
                                              d_tape.data_ptr(), d_out.data_ptr(), d_coll.data_ptr(), stream))
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy().T, ocfg) and np.array_equal(d_coll.cpu().numpy().astype(bool), ocoll)
    cp = cases._cparams(abi.FEATURE_CRS, 0.125, 0.05)
    d_labels = torch.zeros(n, dtype=torch.int32, device=dev)
    d_ncl = torch.zeros(1, dtype=torch.int32, device=dev)
    abi.check(lib.upc_cluster_particles_device(sim.handle, C.byref(cp), 1, n, d_out.data_ptr(), d_labels.data_ptr(), d_ncl.data_ptr(), stream))
    torch.cuda.synchronize()
    olabels, oncl, _ = oracle.cluster_particles(sc["env"], sc["robot"], cp, ocfg)
    assert np.array_equal(d_labels.cpu().numpy(), olabels) and int(d_ncl.item()) == int(oncl[0])


def test_errors_mirror_the_reference_interface(simulators):
    sc = cases.se2_contact()
    sim = simulators("errs", sc["env"], sc["robot"])
    tape = abi.fill_noise_tape(1, sc["robot"], sc["params"], 96)
    with pytest.raises(ValueError):
        sim.ForwardSimulateRobots(sc["starts"], np.zeros((5, 3)), tape, sc["params"])      # targets neither 1 nor n
    with pytest.raises(ValueError):
        sim.ForwardSimulateRobots(sc["starts"][:, :2], sc["targets"], tape, sc["params"])
    with pytest.raises(ValueError):
        sim.ForwardSimulateRobots(sc["starts"], sc["targets"], tape[:10], sc["params"])
    bad = abi.make_sim_params(0, 0.01, 0.08)
    with pytest.raises(ValueError):
        sim.ForwardSimulateRobots(sc["starts"], sc["targets"], tape, bad)
    from uncertainty_planning_core_b200 import OutcomeClustering
    oc = OutcomeClustering(sim, abi.FEATURE_PTPM)          # no simulation parameters / tape given
    with pytest.raises(ValueError):
        oc.ClusterIndices(sc["starts"])
    cfg, coll = sim.ForwardSimulateRobots(sc["starts"][:0], sc["targets"], tape[..., :0], sc["params"])  # empty batch
    assert cfg.shape == (0, 3) and coll.shape == (0,)


# ------------------------------------------------------------------ BASELINE.json sizes: size-independent properties

def test_full_size_c2_properties(oracle, simulators):
    """10,000 SE(3) particles / 128^3 SDF (BASELINE config 2): lane-count invariance, shard invariance,
    oracle agreement on a subsample of the very same run, resolved contacts are collision free."""
    sc = scenarios.c2_se3()
    n = sc["starts"].shape[0]
    tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], n)
    sim = simulators("c2_full", sc["env"], sc["robot"])
    results = {}
    for lanes in (1, 8, 32):
        sim.SetLanesPerParticle(lanes)
        results[lanes] = sim.ForwardSimulateRobots(sc["starts"], sc["targets"], tape, sc["params"])
    sim.SetLanesPerParticle(0)
    for lanes in (8, 32):
        assert np.array_equal(results[1][0], results[lanes][0]) and np.array_equal(results[1][1], results[lanes][1])
    cfg, coll = results[1]
    # particles are independent: any shard with its tape slice reproduces the full run
    lo, hi = 3000, 3700
    part_cfg, part_coll = sim.ForwardSimulateRobots(sc["starts"][lo:hi], sc["targets"], np.ascontiguousarray(tape[..., lo:hi]), sc["params"])
    assert np.array_equal(part_cfg, cfg[lo:hi]) and np.array_equal(part_coll, coll[lo:hi])
    # the oracle on a subsample of the same run
    pick = np.arange(0, n, 97)
    ocfg, ocoll, _ = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"][pick], sc["targets"], np.ascontiguousarray(tape[..., pick]))
    assert np.array_equal(ocoll, coll[pick]) and max_ulp_diff(ocfg, cfg[pick]) <= CONFIG_ULP_TOLERANCE
    assert coll.any()
    for i in pick[:30]:
        assert not sim.CheckConfigCollision(cfg[i])
    R = cfg[:, :9].reshape(-1, 3, 3)
    assert np.max(np.abs(R @ np.transpose(R, (0, 2, 1)) - np.eye(3))) < 1e-9


def test_full_size_clustering_properties(oracle, simulators):
    """16,384 particles: blobs come back as blobs, permuting the input permutes the partition, the single-blob
    case takes the no-fallback path, and a 2,048-particle subsample agrees with the oracle."""
    from uncertainty_planning_core_b200 import OutcomeClustering
    for kind, case in ((abi.ROBOT_SE2, cases.cl_se2_blobs), (abi.ROBOT_SE3, cases.cl_se3_blobs), (abi.ROBOT_LINKED, cases.cl_linked_blobs)):
        c = case()
        sim = simulators("cl_full_%d" % kind, c["env"], c["robot"])
        oc = OutcomeClustering(sim, abi.FEATURE_NONE, 0.125, 0.1)
        n = 16384 if kind != abi.ROBOT_SE3 else 4096
        cfg, owner = scenarios.c4_blobs(kind, n, k=8, seed=77)
        sim.ResetStatistics()
        labels, ncl = oc.ClusterIndices(cfg)
        assert int(ncl[0]) == 8
        assert canonical_partition(labels) == canonical_partition(owner)
        assert sim.GetStatistics()["cluster_fallback_calls"] == 0
        perm = np.random.default_rng(5).permutation(n)
        plabels, _ = oc.ClusterIndices(cfg[perm])
        assert canonical_partition(plabels) == canonical_partition(owner[perm])
        sub = np.arange(0, n, n // 1024)[:1024]
        olabels, oncl, _ = oracle.cluster_particles(c["env"], c["robot"], oc.params, cfg[sub])
        slabels, sncl = oc.ClusterIndices(cfg[sub])
        assert np.array_equal(olabels, slabels)

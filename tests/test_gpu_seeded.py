"""GPU parity tests of the seeded (counter-based noise) entry points, the traced single-particle form, the planner-batch
step and the library's multi-GPU path - all through the C ABI, against the CPU oracle replaying the oracle's own tape of
the same seed.  Bar: flags, counters, labels bit-exact; configurations 0 ulp."""
import ctypes as C
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

    def get(key, env, robot, device=0):
        if (key, device) not in cache:
            cache[(key, device)] = B200Simulator(env, robot, device=device)
        return cache[(key, device)]
    yield get
    for s in cache.values():
        s.close()


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_SIM_CASES))
def test_seeded_simulation_matches_oracle_for_every_lane_count(oracle, simulators, name):
    sc = cases.GOLDEN_SIM_CASES[name]()
    n = sc["starts"].shape[0]
    seed, first = cases.SEEDED_GOLDEN_SEED, cases.SEEDED_GOLDEN_FIRST
    tape = oracle.fill_noise_tape_counter(seed, sc["robot"], sc["params"], n, first_particle=first)
    ocfg, ocoll, ostats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
    g = np.load(os.path.join(HERE, "golden", "sim_seeded_%s.npz" % name))
    assert np.array_equal(ocfg, g["configs"]) and np.array_equal(ocoll, g["collided"])
    sim = simulators(name, sc["env"], sc["robot"])
    for lanes in (0, 1, 2, 3, 4, 5, 6, 8, 16, 32):
        sim.SetLanesPerParticle(lanes)
        sim.ResetStatistics()
        cfg, coll = sim.ForwardSimulateRobotsSeeded(sc["starts"], sc["targets"], seed, sc["params"], first_particle=first)
        assert np.array_equal(coll, ocoll), (name, lanes)
        assert max_ulp_diff(cfg, ocfg) <= CONFIG_ULP_TOLERANCE, (name, lanes, np.abs(cfg - ocfg).max())
        stats = sim.GetStatistics()
        for k in cases.SIM_STAT_KEYS:
            assert int(stats[k]) == ostats[k], (name, lanes, k)
        assert stats["h2d_bytes"] <= 2 * sc["starts"].nbytes        # starts + targets only: no tape crosses the boundary
    sim.SetLanesPerParticle(0)
    # the tape form replaying the product's own tape of the same draws gives the same bits
    ptape = abi.fill_noise_tape_seeded(seed, sc["robot"], sc["params"], n, first_particle=first)
    assert np.array_equal(ptape, tape)
    cfg2, coll2 = sim.ForwardSimulateRobots(sc["starts"], sc["targets"], ptape, sc["params"])
    assert np.array_equal(cfg2, ocfg) and np.array_equal(coll2, ocoll)


def test_seeded_results_do_not_depend_on_the_batch(oracle, simulators):
    """keying is by global particle index: any shard reproduces its slice of the full run"""
    sc = scenarios.c2_se3(seed=5, n=2048, num_steps=32, shape=(64, 64, 64), resolution=0.0625)
    sim = simulators("c2_small", sc["env"], sc["robot"])
    cfg, coll = sim.ForwardSimulateRobotsSeeded(sc["starts"], sc["targets"], 99, sc["params"])
    for lo, hi in ((0, 1), (100, 612), (2040, 2048)):
        pcfg, pcoll = sim.ForwardSimulateRobotsSeeded(sc["starts"][lo:hi], sc["targets"], 99, sc["params"], first_particle=lo)
        assert np.array_equal(pcfg, cfg[lo:hi]) and np.array_equal(pcoll, coll[lo:hi])
    other, _ = sim.ForwardSimulateRobotsSeeded(sc["starts"], sc["targets"], 100, sc["params"])
    assert not np.array_equal(other, cfg)


@pytest.mark.parametrize("make", [cases.se2_contact, cases.se2_microsteps, cases.se3_sensor_noise_spheres, cases.linked_contact])
def test_traced_simulation_matches_oracle(oracle, simulators, make):
    """a2: ForwardSimulateRobot with enable_tracing (simple_simulator_interface.hpp:40-86, 621): one record per controller interval"""
    sc = make()
    n = min(16, sc["starts"].shape[0])
    starts = sc["starts"][:n]
    targets = sc["targets"] if sc["targets"].shape[0] == 1 else sc["targets"][:n]
    sim = simulators(sc["name"] + make.__name__, sc["env"], sc["robot"])
    seed = 31337
    tape = oracle.fill_noise_tape_counter(seed, sc["robot"], sc["params"], n)
    ocfg, ocoll, otrace = oracle.forward_simulate_traced(sc["env"], sc["robot"], sc["params"], starts, targets, tape)
    for kwargs in (dict(seed=seed), dict(noise_tape=tape)):
        cfg, coll, trace = sim.ForwardSimulateRobotTraced(starts, targets, sc["params"], **kwargs)
        assert np.array_equal(cfg, ocfg) and np.array_equal(coll, ocoll)
        assert np.array_equal(trace["steps"], otrace["steps"])
        assert np.array_equal(trace["configs"], otrace["configs"]) and np.array_equal(trace["controls"], otrace["controls"])
    last = otrace["steps"][0] - 1
    assert np.array_equal(trace["configs"][last, np.arange(n)], cfg)
    # the single-particle call the planner makes (n = 1)
    one_cfg, one_coll, one_trace = sim.ForwardSimulateRobotTraced(starts[:1], targets[:1], sc["params"], seed=seed)
    assert np.array_equal(one_cfg[0], ocfg[0]) and one_trace["steps"][0, 0] == otrace["steps"][0, 0]


def test_planner_batch_step_matches_oracle(oracle, simulators):
    """BASELINE config 5 in small through upc_simulate_and_cluster_edges: one call = simulate every edge (seeded) + cluster every
    edge's outcome; equals the oracle edge for edge, and the edge-by-edge seeded calls"""
    from uncertainty_planning_core_b200 import OutcomeClustering
    E, P = 24, 128
    sc = scenarios.c5_batch(edges=E, particles=P, num_steps=40)
    sim = simulators("c5_small", sc["env"], sc["robot"])
    cp = cases._cparams(abi.FEATURE_CRS, 0.125, 0.1)
    es, et = sc["starts"][::P], sc["targets"][::P]
    seed = 2024
    cfg, coll, labels, ncl = sim.SimulateAndClusterEdges(es, et, P, seed, sc["params"], cp)
    assert coll.any() and not coll.all()
    tape = oracle.fill_noise_tape_counter(seed, sc["robot"], sc["params"], E * P)
    ocfg, ocoll, _ = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
    assert np.array_equal(cfg.reshape(E * P, -1), ocfg) and np.array_equal(coll.reshape(-1), ocoll)
    olabels, oncl, _ = oracle.cluster_particles(sc["env"], sc["robot"], cp, ocfg, E)
    assert np.array_equal(labels.reshape(-1), olabels) and np.array_equal(ncl, oncl)
    # edge by edge with the global particle offset
    for e in (0, 7, E - 1):
        one_cfg, one_coll = sim.ForwardSimulateRobotsSeeded(es[e:e + 1], et[e:e + 1], seed, sc["params"], first_particle=e * P, n=P)
        assert np.array_equal(one_cfg, cfg[e]) and np.array_equal(one_coll, coll[e])


def test_ptpm_with_seeded_pair_simulations(oracle, simulators):
    """POINT_TO_POINT_MOVEMENT (uncertainty_contact_planning.hpp:1801-1863) with the pair simulations drawing seeded noise, for the
    SE(2) and the linked robot"""
    from uncertainty_planning_core_b200 import OutcomeClustering
    for make in (cases.cl_se2_ptpm, cases.cl_linked_ptpm):
        c = make()
        sim = simulators("ptpm_" + make.__name__, c["env"], c["robot"])
        cp = c["cparams"]
        seeded = abi.ClusterParams.from_buffer_copy(cp)
        seeded.ptpm_tape = None
        seeded.ptpm_seed = 77
        olabels, oncl, _ = oracle.cluster_particles(c["env"], c["robot"], seeded, c["configs"])
        oc = OutcomeClustering(sim, abi.FEATURE_PTPM, cp.signature_matching_threshold, cp.distance_clustering_threshold,
                               goal_distance_threshold=cp.goal_distance_threshold, ptpm_params=c["ptpm"][0], ptpm_seed=77)
        labels, ncl = oc.ClusterIndices(c["configs"])
        assert np.array_equal(labels, olabels) and np.array_equal(ncl, oncl)


def test_policy_particle_clustering_matches_oracle(oracle, simulators):
    """a18: PolicyParticleClusteringFn (uncertainty_contact_planning.hpp:1558-1578) = the clustering of (particles + current config)"""
    from uncertainty_planning_core_b200 import OutcomeClustering
    for case in (cases.cl_se2_crs, cases.cl_se3_crs, cases.cl_linked_crs, cases.cl_se2_blobs):
        c = case()
        sim = simulators("pol_" + case.__name__, c["env"], c["robot"])
        cp = c["cparams"]
        oc = OutcomeClustering(sim, cp.feature_type, cp.signature_matching_threshold, cp.distance_clustering_threshold)
        particles, current = c["configs"][:-1], c["configs"][-1]
        idx = oc.PolicyParticleClusteringFn(particles, current)
        olabels, oncl, _ = oracle.cluster_particles(c["env"], c["robot"], cp, np.concatenate([particles, current[None, :]], axis=0))
        assert len(idx) == int(oncl[0])
        for k, members in enumerate(idx):
            assert members == np.nonzero(olabels == k)[0].tolist()
        assert any(len(c["configs"]) - 1 in m for m in idx)         # the current configuration is the last index


# ------------------------------------------------------------------ the library's multi-GPU path

def _edge_batch_through_comm(lib, sims, comms, sc, cp, E, P, seed):
    import torch
    W = len(sims)
    width = sims[0].robot.width
    cap = (E + W - 1) // W
    lay = abi.outcome_layout(width, cap * P, cap)
    es = np.ascontiguousarray(sc["starts"][::P].T)
    et = np.ascontiguousarray(sc["targets"][::P].T)
    d_s, d_t, d_b = [], [], []
    for r in range(W):
        dev = torch.device("cuda", r)
        d_s.append(torch.from_numpy(es).to(dev))
        d_t.append(torch.from_numpy(et).to(dev))
        d_b.append(torch.zeros(lay.block_bytes * W, dtype=torch.uint8, device=dev))
    if W == 1:
        abi.check(lib.upc_sharded_edge_batch(comms[0], C.byref(sc["params"]), C.byref(cp), seed, E, P, d_s[0].data_ptr(), d_t[0].data_ptr(),
                                             d_b[0].data_ptr(), None))
    else:
        arr = lambda ptrs: (C.c_void_p * W)(*ptrs)
        abi.check(lib.upc_sharded_edge_batch_all(W, arr([c.value for c in comms]), C.byref(sc["params"]), C.byref(cp), seed, E, P,
                                                 arr([t.data_ptr() for t in d_s]), arr([t.data_ptr() for t in d_t]),
                                                 arr([t.data_ptr() for t in d_b]), None))
    for c in comms:
        abi.check(lib.upc_comm_synchronize(c))
    out = []
    for r in range(W):                      # every rank must hold every rank's block after the all-gather
        host = d_b[r].cpu().numpy()
        cfg, coll, labels, ncl = [], [], [], []
        for q in range(W):
            lo, hi = abi.shard_bounds(E, q, W)
            n = (hi - lo) * P
            blk = host[q * lay.block_bytes:(q + 1) * lay.block_bytes]
            labels.append(blk[lay.labels_offset:lay.labels_offset + 4 * n].view(np.int32))
            ncl.append(blk[lay.n_clusters_offset:lay.n_clusters_offset + 4 * (hi - lo)].view(np.int32))
            coll.append(blk[lay.collided_offset:lay.collided_offset + n].astype(bool))
            cfg.append(blk[lay.configs_offset:lay.configs_offset + 8 * width * n].view(np.float64).reshape(width, n).T)
        out.append((np.concatenate(cfg), np.concatenate(coll), np.concatenate(labels), np.concatenate(ncl)))
    return out


def test_sharded_edge_batch_through_the_library_communicator(oracle, simulators):
    """upc_comm_* + upc_sharded_edge_batch[_all]: edges split over the GPUs present (1 on a single-GPU box), SDF replicated,
    one in-place NCCL all-gather issued by the library; the gathered outcome equals the unsharded call and the oracle"""
    import torch
    lib = abi.load_library()
    E, P = 22, 64            # 22 edges: uneven shards on 4 and 8 GPUs
    sc = scenarios.c5_batch(edges=E, particles=P, num_steps=32)
    cp = cases._cparams(abi.FEATURE_CRS, 0.125, 0.1)
    seed = 606
    base = simulators("c5_comm", sc["env"], sc["robot"])
    cfg, coll, labels, ncl = base.SimulateAndClusterEdges(sc["starts"][::P], sc["targets"][::P], P, seed, sc["params"], cp)
    tape = oracle.fill_noise_tape_counter(seed, sc["robot"], sc["params"], E * P)
    ocfg, ocoll, _ = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
    assert np.array_equal(cfg.reshape(E * P, -1), ocfg) and np.array_equal(coll.reshape(-1), ocoll)
    version = C.c_int32(0)
    abi.check(lib.upc_comm_nccl_version(C.byref(version)))
    assert version.value >= 20000
    worlds = sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()})
    for W in worlds:
        sims = [simulators("c5_comm", sc["env"], sc["robot"], device=r) for r in range(W)]
        comms = (C.c_void_p * W)()
        handles = (C.c_void_p * W)(*[s.handle.value for s in sims])
        abi.check(lib.upc_comm_create_all(W, handles, comms))
        comm_list = [C.c_void_p(comms[r]) for r in range(W)]
        try:
            assert lib.upc_comm_world(comm_list[0]) == W and lib.upc_comm_rank(comm_list[W - 1]) == W - 1
            for rank_view in _edge_batch_through_comm(lib, sims, comm_list, sc, cp, E, P, seed):
                gcfg, gcoll, glabels, gncl = rank_view
                assert np.array_equal(gcfg, cfg.reshape(E * P, -1)) and np.array_equal(gcoll, coll.reshape(-1))
                assert np.array_equal(glabels, labels.reshape(-1)) and np.array_equal(gncl, ncl)
        finally:
            for c in comm_list:
                lib.upc_comm_destroy(c)


def test_pipelined_halves_of_the_sharded_step_equal_the_one_call_form(oracle, simulators):
    """upc_sharded_edge_batch_simulate / _finish: batch k + 1 simulated on stream A while batch k is clustered (and exchanged) on
    stream B, two outcome buffers, one host thread - what bench.py times.  Six batches with different seeds in flight back to back
    must each equal upc_sharded_edge_batch issued alone."""
    import torch
    lib = abi.load_library()
    E, P = 40, 96
    sc = scenarios.c5_batch(edges=E, particles=P, num_steps=32)
    cp = cases._cparams(abi.FEATURE_CRS, 0.125, 0.1)
    sim = simulators("c5_pipe", sc["env"], sc["robot"])
    handles = (C.c_void_p * 1)(sim.handle.value)
    comms = (C.c_void_p * 1)()
    abi.check(lib.upc_comm_create_all(1, handles, comms))
    comm = C.c_void_p(comms[0])
    try:
        width = sim.robot.width
        lay = abi.outcome_layout(width, E * P, E)
        dev = torch.device("cuda", 0)
        d_s = torch.from_numpy(np.ascontiguousarray(sc["starts"][::P].T)).to(dev)
        d_t = torch.from_numpy(np.ascontiguousarray(sc["targets"][::P].T)).to(dev)
        K = 6
        want = []
        ref = torch.zeros(lay.block_bytes, dtype=torch.uint8, device=dev)
        for k in range(K):
            abi.check(lib.upc_sharded_edge_batch(comm, C.byref(sc["params"]), C.byref(cp), 900 + k, E, P, d_s.data_ptr(), d_t.data_ptr(), ref.data_ptr(), None))
            abi.check(lib.upc_comm_synchronize(comm))
            torch.cuda.synchronize()
            want.append(ref.cpu().numpy().copy())
        A, B = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        bufs = [torch.zeros(lay.block_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
        ev_sim = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        got = [None] * K

        def finish(k):
            B.wait_event(ev_sim[k % 2])
            abi.check(lib.upc_sharded_edge_batch_finish(comm, C.byref(cp), E, P, bufs[k % 2].data_ptr(), B.cuda_stream))
            ev_done[k % 2].record(B)

        for k in range(K):
            if k >= 2:
                ev_done[k % 2].synchronize()                      # batch k - 2 is complete: read it back before its buffer is reused
                got[k - 2] = bufs[k % 2].cpu().numpy().copy()
                A.wait_event(ev_done[k % 2])
                abi.check(lib.upc_comm_wait(comm, A.cuda_stream))
            abi.check(lib.upc_sharded_edge_batch_simulate(comm, C.byref(sc["params"]), 900 + k, E, P, d_s.data_ptr(), d_t.data_ptr(), bufs[k % 2].data_ptr(),
                                                          A.cuda_stream))
            ev_sim[k % 2].record(A)
            if k >= 1:
                finish(k - 1)
        finish(K - 1)
        torch.cuda.synchronize()
        abi.check(lib.upc_comm_synchronize(comm))
        got[K - 2] = bufs[(K - 2) % 2].cpu().numpy().copy()
        got[K - 1] = bufs[(K - 1) % 2].cpu().numpy().copy()
        n = E * P
        for k in range(K):
            for off, size in ((lay.configs_offset, 8 * width * n), (lay.collided_offset, n), (lay.labels_offset, 4 * n), (lay.n_clusters_offset, 4 * E)):
                assert np.array_equal(got[k][off:off + size], want[k][off:off + size]), (k, off)
        assert not np.array_equal(want[0][lay.configs_offset:lay.configs_offset + 64], want[1][lay.configs_offset:lay.configs_offset + 64])
    finally:
        lib.upc_comm_destroy(comm)


def test_submit_collect_pairs_equal_the_one_call_edge_batch(oracle, simulators):
    """upc_edge_batch_submit / upc_edge_batch_collect with two batches in flight (submit k + 1 before collect k) return, batch by
    batch, what upc_simulate_and_cluster_edges returns for the same arguments - which test_planner_batch_step_matches_oracle pins to
    the oracle.  Both start forms: one start per edge, one start per particle."""
    lib = abi.load_library()
    E, P = 24, 128
    sc = scenarios.c5_batch(edges=E, particles=P, num_steps=32)
    cp = cases._cparams(abi.FEATURE_CRS, 0.125, 0.1)
    sim = simulators("c5_submit", sc["env"], sc["robot"])
    width, n = sim.robot.width, E * P
    es = np.ascontiguousarray(sc["starts"][::P].T)
    et = np.ascontiguousarray(sc["targets"][::P].T)
    rng = np.random.default_rng(3)
    per_particle = np.ascontiguousarray((sc["starts"] + rng.normal(scale=1e-3, size=sc["starts"].shape)).T)

    def one_call(seed, starts):
        out, coll = np.zeros((width, n)), np.zeros(n, dtype=np.uint8)
        labels, ncl = np.zeros(n, dtype=np.int32), np.zeros(E, dtype=np.int32)
        abi.check(lib.upc_simulate_and_cluster_edges(sim.handle, C.byref(sc["params"]), C.byref(cp), seed, 0, E, P, starts.shape[1], starts.ctypes.data, et.ctypes.data,
                                                     out.ctypes.data, coll.ctypes.data, labels.ctypes.data, ncl.ctypes.data))
        return out, coll, labels, ncl

    K = 5
    starts_of = lambda k: per_particle if k % 2 else es
    want = [one_call(700 + k, starts_of(k)) for k in range(K)]
    got = []
    res = [(np.zeros((width, n)), np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.int32), np.zeros(E, dtype=np.int32)) for _ in range(2)]

    def submit(k):
        st = starts_of(k)
        abi.check(lib.upc_edge_batch_submit(sim.handle, C.byref(sc["params"]), 700 + k, 0, E, P, st.shape[1], st.ctypes.data, et.ctypes.data, k % 2))

    submit(0)
    for k in range(K):
        if k + 1 < K:
            submit(k + 1)
        o, c, l, m = res[k % 2]
        abi.check(lib.upc_edge_batch_collect(sim.handle, C.byref(cp), k % 2, o.ctypes.data, c.ctypes.data, l.ctypes.data, m.ctypes.data))
        got.append((o.copy(), c.copy(), l.copy(), m.copy()))
    for k in range(K):
        for a, b in zip(got[k], want[k]):
            assert np.array_equal(a, b), k
    assert not np.array_equal(want[0][0], want[2][0])            # different seeds really give different batches
    # a slot can only be collected once, and only after a submit
    o, c, l, m = res[0]
    assert lib.upc_edge_batch_collect(sim.handle, C.byref(cp), 0, o.ctypes.data, c.ctypes.data, l.ctypes.data, m.ctypes.data) != 0


# ------------------------------------------------------------------ BASELINE.json sizes

def test_full_size_c3_subsample_matches_oracle(oracle, simulators):
    """BASELINE config 3: 100,000 particles of the 7-DoF arm against the 256^3 SDF (plain-grid path), seeded noise; a 1,024-particle
    subsample of the very same run is recomputed by the oracle; lane-count invariance on a slice"""
    sc = scenarios.c3_linked()
    n = sc["starts"].shape[0]
    assert n == 100000 and sc["env"].sdf.shape == (256, 256, 256)
    sim = simulators("c3_full", sc["env"], sc["robot"])
    seed = 3003
    cfg, coll = sim.ForwardSimulateRobotsSeeded(sc["starts"][:1], sc["targets"], seed, sc["params"], n=n)
    assert coll.any()
    pick = np.arange(0, n, 97)[:1024]
    for i0 in range(0, len(pick), 256):
        idx = pick[i0:i0 + 256]
        tapes = [oracle.fill_noise_tape_counter(seed, sc["robot"], sc["params"], 1, first_particle=int(i)) for i in idx]
        tape = np.concatenate(tapes, axis=3)
        ocfg, ocoll, _ = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"][idx], sc["targets"], tape)
        assert np.array_equal(ocoll, coll[idx]) and max_ulp_diff(ocfg, cfg[idx]) <= CONFIG_ULP_TOLERANCE
    for lanes in (4, 16):
        sim.SetLanesPerParticle(lanes)
        pcfg, pcoll = sim.ForwardSimulateRobotsSeeded(sc["starts"][:1], sc["targets"], seed, sc["params"], first_particle=5000, n=512)
        assert np.array_equal(pcfg, cfg[5000:5512]) and np.array_equal(pcoll, coll[5000:5512])
    sim.SetLanesPerParticle(0)
    lim = 2.9 + 1e-12
    assert np.all(np.abs(cfg) <= lim)                                  # joint limits hold (simple_robot_models.hpp:1077-1100)
    for i in pick[:20]:
        assert not sim.CheckConfigCollision(cfg[i])


@pytest.mark.parametrize("pivot", ["auto", "off"])
def test_clustering_at_65536_particles(oracle, simulators, pivot, monkeypatch):
    """BASELINE config 4 at its largest size, all three robots, with the pivot pass and with the adjacency pass (512 MiB bit
    matrix): the generating blobs come back, and a 4,096-particle subsample agrees with the oracle"""
    from uncertainty_planning_core_b200 import OutcomeClustering
    if pivot == "off":
        monkeypatch.setenv("UPC_B200_NO_PIVOT_PASS", "1")
    n = 65536
    for kind, case in ((abi.ROBOT_SE2, cases.cl_se2_blobs), (abi.ROBOT_SE3, cases.cl_se3_blobs), (abi.ROBOT_LINKED, cases.cl_linked_blobs)):
        c = case()
        sim = simulators("cl_64k_%d" % kind, c["env"], c["robot"])
        oc = OutcomeClustering(sim, abi.FEATURE_NONE, 0.125, 0.1)
        cfg, owner = scenarios.c4_blobs(kind, n, k=8, seed=41)
        labels, ncl = oc.ClusterIndices(cfg)
        assert int(ncl[0]) == 8
        # labels are canonical (clusters numbered by smallest member), so equality with the relabelled owners is exact
        first = {}
        for i, o in enumerate(owner.tolist()):
            first.setdefault(o, len(first))
        assert np.array_equal(labels, np.array([first[o] for o in owner.tolist()], dtype=np.int32))
        sub = np.arange(0, n, 16)[:4096]
        olabels, oncl, _ = oracle.cluster_particles(c["env"], c["robot"], oc.params, cfg[sub])
        slabels, sncl = oc.ClusterIndices(cfg[sub])
        assert np.array_equal(olabels, slabels) and np.array_equal(oncl, sncl)
        assert canonical_partition(labels[sub]) == canonical_partition(olabels)


def test_full_size_c5_properties(oracle, simulators):
    """BASELINE config 5 at full size (4,096 edges x 1,000 particles, one simulation + one clustering pass): sampled edges equal
    the oracle bit for bit; labels are canonical per edge; every edge keeps all its particles"""
    E, P = 4096, 1000
    sc = scenarios.c5_batch(edges=E, particles=P)
    sim = simulators("c5_full", sc["env"], sc["robot"])
    cp = cases._cparams(abi.FEATURE_CRS, 0.125, 0.1)
    seed = 5005
    cfg, coll, labels, ncl = sim.SimulateAndClusterEdges(sc["starts"][::P], sc["targets"][::P], P, seed, sc["params"], cp)
    assert ncl.min() >= 1 and labels.min() == 0 and np.all(labels.max(axis=1) == ncl - 1)
    assert np.all(labels[:, 0] == 0)                                   # canonical numbering: the first particle opens cluster 0
    for e in (0, 1, 1777, 4095):
        sl = slice(e * P, (e + 1) * P)
        tape = oracle.fill_noise_tape_counter(seed, sc["robot"], sc["params"], P, first_particle=e * P)
        ocfg, ocoll, _ = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"][sl], sc["targets"][sl], tape)
        assert np.array_equal(ocfg, cfg[e]) and np.array_equal(ocoll, coll[e])
        ol, on, _ = oracle.cluster_particles(sc["env"], sc["robot"], cp, ocfg)
        assert np.array_equal(ol, labels[e]) and int(on[0]) == int(ncl[e])

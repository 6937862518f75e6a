"""Host-side mirror of the reference's plugin interfaces over the C ABI.

B200Simulator mirrors simple_simulator_interface::SimulatorInterface
(include/uncertainty_planning_core/simple_simulator_interface.hpp:383-624): same method names, argument
meaning and error behaviour (ValueError where the reference throws std::invalid_argument).
OutcomeClustering mirrors UncertaintyPlanningSpace::ClusterParticles / PolicyParticleClusteringFn
(uncertainty_contact_planning.hpp:1986-2034, :1558-1578).

Configurations are numpy arrays [n, C] on this side (one row per particle, C = 3 for SE2 (x, y, theta),
12 for SE3 (R row-major then t), dof for linked robots) and are transposed to the ABI's [C][n] layout here.
The production host shim is the C++ header include/upc_b200_shim.hpp; this Python mirror exists for the
parity tests and the benchmark.
"""
import ctypes as C

import numpy as np

from . import abi


class B200Simulator:
    def __init__(self, environment, robot, debug_level=0, device=0):
        self._lib = abi.load_library()
        self.environment = environment
        self.robot = robot
        self.debug_level = debug_level
        self._h = C.c_void_p()
        abi.check(self._lib.upc_create(C.byref(environment.desc), device, C.byref(self._h)))
        abi.check(self._lib.upc_set_robot(self._h, C.byref(robot.desc)))

    # -- life cycle -------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.upc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def SetLanesPerParticle(self, lanes):
        abi.check(self._lib.upc_set_lanes_per_particle(self._h, lanes))

    # -- SimulatorInterface ---------------------------------------------------------------------------------
    def GetStatistics(self):
        s = abi.Stats()
        abi.check(self._lib.upc_get_stats(self._h, C.byref(s)))
        return {k: float(v) for k, v in s.as_dict().items()}

    def ResetStatistics(self):
        abi.check(self._lib.upc_reset_stats(self._h))

    def GetResolution(self):
        return self.environment.resolution

    def CheckConfigCollision(self, config, inflation_ratio=0.0, contact_distance=0.0):
        cfg = np.ascontiguousarray(config, dtype=np.float64).reshape(-1)
        if cfg.shape[0] != self.robot.width:
            raise ValueError("configuration has %d values, robot needs %d" % (cfg.shape[0], self.robot.width))
        out = C.c_int32(0)
        inflation = inflation_ratio * self.environment.resolution
        abi.check(self._lib.upc_check_config_collision(self._h, cfg.ctypes.data, contact_distance, inflation, C.byref(out)))
        return bool(out.value)

    def ForwardSimulateRobots(self, start_positions, target_positions, noise_tape, params):
        """Batched forward simulation.  Returns (configs [n, C], collided [n] bool).

        start_positions [n, C]; target_positions [1, C] or [n, C]; noise_tape
        [num_steps, 1 + max_microsteps, dof, n] (abi.fill_noise_tape or the caller's own draws);
        params: abi.SimParams carrying forward_simulation_time (as num_steps * controller_interval),
        simulation_shortcut_distance, use_individual_jacobians and allow_contacts.
        """
        starts = np.asarray(start_positions, dtype=np.float64)
        targets = np.asarray(target_positions, dtype=np.float64)
        if starts.ndim != 2 or starts.shape[1] != self.robot.width:
            raise ValueError("start_positions must be [n, %d]" % self.robot.width)
        n = starts.shape[0]
        if targets.ndim != 2 or targets.shape[1] != self.robot.width or targets.shape[0] not in (1, n):
            raise ValueError("target_positions must have size 1 or size start_positions.size()")
        tape = np.ascontiguousarray(noise_tape, dtype=np.float64)
        if tape.shape != abi.tape_shape(params, self.robot.dof, n):
            raise ValueError("noise tape has shape %s, expected %s" % (tape.shape, abi.tape_shape(params, self.robot.dof, n)))
        starts_soa = np.ascontiguousarray(starts.T)
        targets_soa = np.ascontiguousarray(targets.T)
        out = np.zeros_like(starts_soa)
        collided = np.zeros(n, dtype=np.uint8)
        abi.check(self._lib.upc_forward_simulate(self._h, C.byref(params), n, starts_soa.ctypes.data, targets.shape[0],
                                                 targets_soa.ctypes.data, tape.ctypes.data, out.ctypes.data,
                                                 collided.ctypes.data))
        return np.ascontiguousarray(out.T), collided.astype(bool)

    def ForwardSimulateRobotsSeeded(self, start_positions, target_positions, seed, params, first_particle=0, n=None):
        """Batched forward simulation with the counter-based noise of DESIGN.md 4.2 (no tape crosses the boundary).
        start_positions / target_positions hold one row per particle, one row per block of particles (their row counts must
        divide n: edge batches) or a single row; particle i draws from the stream (seed, first_particle + i).
        Returns (configs [n, C], collided [n] bool)."""
        starts = np.asarray(start_positions, dtype=np.float64).reshape(-1, self.robot.width)
        targets = np.asarray(target_positions, dtype=np.float64).reshape(-1, self.robot.width)
        n = starts.shape[0] if n is None else int(n)
        starts_soa = np.ascontiguousarray(starts.T)
        targets_soa = np.ascontiguousarray(targets.T)
        out = np.zeros((self.robot.width, n))
        collided = np.zeros(n, dtype=np.uint8)
        abi.check(self._lib.upc_forward_simulate_seeded(self._h, C.byref(params), int(seed), int(first_particle), n, starts.shape[0],
                                                        starts_soa.ctypes.data, targets.shape[0], targets_soa.ctypes.data,
                                                        out.ctypes.data, collided.ctypes.data))
        return np.ascontiguousarray(out.T), collided.astype(bool)

    def ForwardSimulateRobotTraced(self, start_positions, target_positions, params, noise_tape=None, seed=0, first_particle=0):
        """ForwardSimulateRobot with enable_tracing (simple_simulator_interface.hpp:621; trace types :40-63) for a small batch:
        returns (configs [n, C], collided [n], trace) with trace = dict(configs [S, n, C] - the resolved configuration at the
        end of every controller interval, what ExtractTrajectoryFromTrace (:66-86) reads -, controls [S, n, 2*dof] - control
        input, then the control step per microstep -, steps [2, n] - intervals executed, microsteps of the last one).
        noise_tape=None selects seeded noise."""
        starts = np.asarray(start_positions, dtype=np.float64).reshape(-1, self.robot.width)
        targets = np.asarray(target_positions, dtype=np.float64).reshape(-1, self.robot.width)
        n, S, W, dof = starts.shape[0], params.num_steps, self.robot.width, self.robot.dof
        if targets.shape[0] not in (1, n):
            raise ValueError("target_positions must have size 1 or size start_positions.size()")
        tape = None
        if noise_tape is not None:
            tape = np.ascontiguousarray(noise_tape, dtype=np.float64)
            if tape.shape != abi.tape_shape(params, dof, n):
                raise ValueError("noise tape has shape %s, expected %s" % (tape.shape, abi.tape_shape(params, dof, n)))
        s_soa, t_soa = np.ascontiguousarray(starts.T), np.ascontiguousarray(targets.T)
        out, coll = np.zeros((W, n)), np.zeros(n, dtype=np.uint8)
        tc, tu, ts = np.zeros((S, W, n)), np.zeros((S, 2 * dof, n)), np.zeros((2, n), dtype=np.int32)
        abi.check(self._lib.upc_forward_simulate_traced(self._h, C.byref(params), n, s_soa.ctypes.data, targets.shape[0], t_soa.ctypes.data,
                                                        None if tape is None else tape.ctypes.data, int(seed), int(first_particle),
                                                        out.ctypes.data, coll.ctypes.data, tc.ctypes.data, tu.ctypes.data, ts.ctypes.data))
        trace = dict(configs=np.ascontiguousarray(tc.transpose(0, 2, 1)), controls=np.ascontiguousarray(tu.transpose(0, 2, 1)), steps=ts)
        return np.ascontiguousarray(out.T), coll.astype(bool), trace

    def SimulateAndClusterEdges(self, edge_starts, edge_targets, particles_per_edge, seed, params, cluster_params, first_particle=0):
        """The planner-batch step in one call (BASELINE config 5): E candidate edges x particles_per_edge particles simulated
        with seeded noise (uncertainty_contact_planning.hpp:2039-2071 per edge), each edge's outcome clustered (:1986-2034).
        edge_starts: [E, C] (one start per edge) or [E * P, C] (the parent states' own particles, edge-major).
        Returns (configs [E, P, C], collided [E, P], labels [E, P], n_clusters [E])."""
        es = np.asarray(edge_starts, dtype=np.float64).reshape(-1, self.robot.width)
        et = np.asarray(edge_targets, dtype=np.float64).reshape(-1, self.robot.width)
        E, P, W = et.shape[0], int(particles_per_edge), self.robot.width
        n = E * P
        if es.shape[0] not in (E, n):
            raise ValueError("edge_starts must hold one configuration per edge or one per particle (edge-major)")
        s_soa, t_soa = np.ascontiguousarray(es.T), np.ascontiguousarray(et.T)
        out, coll = np.zeros((W, n)), np.zeros(n, dtype=np.uint8)
        labels, ncl = np.zeros(n, dtype=np.int32), np.zeros(E, dtype=np.int32)
        abi.check(self._lib.upc_simulate_and_cluster_edges(self._h, C.byref(params), C.byref(cluster_params), int(seed), int(first_particle), E, P,
                                                           es.shape[0], s_soa.ctypes.data, t_soa.ctypes.data, out.ctypes.data, coll.ctypes.data,
                                                           labels.ctypes.data, ncl.ctypes.data))
        return np.ascontiguousarray(out.T).reshape(E, P, W), coll.astype(bool).reshape(E, P), labels.reshape(E, P), ncl

    def ForwardSimulateEdgeBatch(self, edge_starts, edge_targets, particles_per_edge, noise_tape, params):
        """A batch of candidate edges in ONE call (BASELINE config 5): edge e contributes `particles_per_edge` particles that
        start at edge_starts[e] and move towards edge_targets[e] - the targets.size() == starts.size() form of
        ForwardSimulateRobots (simple_simulator_interface.hpp:623; used at uncertainty_contact_planning.hpp:1826).
        noise_tape covers E * particles_per_edge particles, edge-major.  Returns (configs [E, P, C], collided [E, P]);
        particles are independent, so the result equals E separate calls with the matching tape columns."""
        es = np.asarray(edge_starts, dtype=np.float64).reshape(-1, self.robot.width)
        et = np.asarray(edge_targets, dtype=np.float64).reshape(-1, self.robot.width)
        if es.shape != et.shape:
            raise ValueError("one target per start")
        E, P = es.shape[0], int(particles_per_edge)
        cfg, coll = self.ForwardSimulateRobots(np.repeat(es, P, axis=0), np.repeat(et, P, axis=0), noise_tape, params)
        return cfg.reshape(E, P, self.robot.width), coll.reshape(E, P)

    def CountWithin(self, configs, targets, threshold):
        """How many configurations lie within `threshold` (<=) of their target, and which: the count behind
        ComputeReverseEdgeProbability / the goal-reached fraction (uncertainty_contact_planning.hpp:2079-2088)."""
        cfg = np.asarray(configs, dtype=np.float64)
        tgt = np.asarray(targets, dtype=np.float64).reshape(-1, self.robot.width)
        n = cfg.shape[0]
        if cfg.ndim != 2 or cfg.shape[1] != self.robot.width or tgt.shape[0] not in (1, n):
            raise ValueError("configs must be [n, %d] and targets [1 or n, %d]" % (self.robot.width, self.robot.width))
        soa, tsoa = np.ascontiguousarray(cfg.T), np.ascontiguousarray(tgt.T)
        flags = np.zeros(n, dtype=np.uint8)
        count = C.c_int64(0)
        abi.check(self._lib.upc_count_within(self._h, n, soa.ctypes.data, tgt.shape[0], tsoa.ctypes.data, float(threshold),
                                             C.byref(count), flags.ctypes.data))
        return int(count.value), flags.astype(bool)

    @staticmethod
    def StateDistanceWeights(motion_pfeasibility, space_independent_variances, feasibility_alpha, variance_alpha):
        """The two per-state factors of StateDistance (uncertainty_contact_planning.hpp:1492-1499):
        feasibility_weight = (1 - Pfeasibility) * alpha + (1 - alpha); variance_weight = erf(|variances|_1) * alpha_v + (1 - alpha_v)."""
        import math
        pf = np.asarray(motion_pfeasibility, dtype=np.float64)
        var = np.abs(np.asarray(space_independent_variances, dtype=np.float64)).reshape(pf.shape[0], -1)
        raw = np.zeros(pf.shape[0])
        for k in range(var.shape[1]):
            raw = raw + var[:, k]
        fw = (1.0 - pf) * feasibility_alpha + (1.0 - feasibility_alpha)
        vw = np.array([math.erf(v) for v in raw]) * variance_alpha + (1.0 - variance_alpha)
        return fw, vw

    def NearestNeighbors(self, expectations, feasibility_weight, variance_weight, queries, step_size, use_for_nn=None):
        """GetNearestNeighbor (uncertainty_contact_planning.hpp:1506-1542) for a batch of sampled configurations:
        returns (index [Q] int64, -1 when no state is enabled; distance [Q])."""
        exp = np.asarray(expectations, dtype=np.float64).reshape(-1, self.robot.width)
        qs = np.asarray(queries, dtype=np.float64).reshape(-1, self.robot.width)
        ns, nq = exp.shape[0], qs.shape[0]
        fw = np.ascontiguousarray(feasibility_weight, dtype=np.float64)
        vw = np.ascontiguousarray(variance_weight, dtype=np.float64)
        if fw.shape != (ns,) or vw.shape != (ns,):
            raise ValueError("one feasibility and one variance weight per state")
        use = None if use_for_nn is None else np.ascontiguousarray(np.asarray(use_for_nn).astype(np.uint8))
        esoa, qsoa = np.ascontiguousarray(exp.T), np.ascontiguousarray(qs.T)
        idx = np.zeros(nq, dtype=np.int64)
        dist = np.zeros(nq, dtype=np.float64)
        abi.check(self._lib.upc_nearest_neighbors(self._h, ns, esoa.ctypes.data, fw.ctypes.data, vw.ctypes.data,
                                                  None if use is None else use.ctypes.data, nq, qsoa.ctypes.data, float(step_size),
                                                  idx.ctypes.data, dist.ctypes.data))
        return idx, dist

    def ComputeReverseEdgeProbability(self, child_particles, parent_expectation, noise_tape, params, goal_distance_threshold):
        """(attempts, reached): simulate the child's particles back towards the parent with contacts allowed and count those
        that end within goal_distance_threshold of it (uncertainty_contact_planning.hpp:2073-2090)."""
        p = abi.SimParams.from_buffer_copy(params)
        p.allow_contacts = 1
        cfg, _ = self.ForwardSimulateRobots(child_particles, np.asarray(parent_expectation, dtype=np.float64).reshape(1, -1), noise_tape, p)
        reached, _ = self.CountWithin(cfg, parent_expectation, goal_distance_threshold)
        return cfg.shape[0], reached

    def ForwardSimulateRobot(self, start_position, target_position, noise_tape, params):
        cfg, coll = self.ForwardSimulateRobots(np.asarray(start_position, dtype=np.float64).reshape(1, -1),
                                               np.asarray(target_position, dtype=np.float64).reshape(1, -1),
                                               noise_tape, params)
        return cfg[0], bool(coll[0])


class OutcomeClustering:
    """ClusterParticles / PolicyParticleClusteringFn over a B200Simulator's handle."""

    def __init__(self, simulator, clustering_type=abi.FEATURE_CRS, signature_matching_threshold=0.125,
                 distance_clustering_threshold=0.1, goal_distance_threshold=0.0, ptpm_params=None, ptpm_tape=None, ptpm_seed=0):
        """ptpm_params / ptpm_tape: simulation parameters and noise tape [S, 1+M, dof, n*(n-1)] for the
        POINT_TO_POINT_MOVEMENT first pass (uncertainty_contact_planning.hpp:1801-1863); without a tape the pair
        simulations draw seeded noise from ptpm_seed."""
        self.sim = simulator
        self._lib = simulator._lib
        self.params = abi.ClusterParams()
        self.params.feature_type = clustering_type
        self.params.signature_matching_threshold = signature_matching_threshold
        self.params.distance_clustering_threshold = distance_clustering_threshold
        self.params.goal_distance_threshold = goal_distance_threshold
        self._keep = (ptpm_params, None if ptpm_tape is None else np.ascontiguousarray(ptpm_tape, dtype=np.float64))
        self.params.ptpm_seed = int(ptpm_seed)
        if ptpm_params is not None:
            self.params.ptpm_sim = C.pointer(ptpm_params)
            self.params.ptpm_tape = None if self._keep[1] is None else self._keep[1].ctypes.data

    def ClusterIndices(self, configs, n_sets=1):
        """Index clustering of n_sets equally sized particle sets stacked in configs [n_sets*set_size, C].
        Returns (labels [total] int32, n_clusters [n_sets] int32)."""
        cfg = np.asarray(configs, dtype=np.float64)
        if cfg.ndim != 2 or cfg.shape[1] != self.sim.robot.width:
            raise ValueError("configs must be [n, %d]" % self.sim.robot.width)
        total = cfg.shape[0]
        if n_sets < 1 or total % n_sets:
            raise ValueError("configs must hold n_sets equally sized sets")
        set_size = total // n_sets
        soa = np.ascontiguousarray(cfg.T)
        labels = np.zeros(total, dtype=np.int32)
        ncl = np.zeros(n_sets, dtype=np.int32)
        abi.check(self._lib.upc_cluster_particles(self.sim.handle, C.byref(self.params), n_sets, set_size,
                                                  soa.ctypes.data, labels.ctypes.data, ncl.ctypes.data))
        return labels, ncl

    def ClusterParticles(self, particles, allow_contacts):
        """particles: (configs [n, C], collided [n] bool).  Returns a list of clusters, each a
        (configs, collided) pair; collided particles are dropped from the clusters when contacts are not
        allowed, and clusters emptied that way are still emitted (uncertainty_contact_planning.hpp:2008-2026)."""
        configs, collided = particles
        configs = np.asarray(configs, dtype=np.float64)
        collided = np.asarray(collided, dtype=bool)
        n = configs.shape[0]
        if n == 0:
            return []
        if n == 1:
            return [(configs.copy(), collided.copy())]
        labels, ncl = self.ClusterIndices(configs)
        clusters = []
        for k in range(int(ncl[0])):
            keep = (labels == k) & (allow_contacts | ~collided)
            clusters.append((configs[keep], collided[keep]))
        return clusters

    def ClusterStatistics(self, configs, labels, collided, allow_contacts, step_size, n_clusters):
        """Per-cluster reduction behind UncertaintyPlannerState::UpdateStatistics (uncertainty_planner_state.hpp:339-351):
        dict of counts [K], any_collided [K], expectations [K, C], variance [K], si_variance [K], variances [K, V],
        si_variances [K, V] over the particles that survive the contact filter."""
        cfg = np.asarray(configs, dtype=np.float64)
        n = cfg.shape[0]
        robot = self.sim.robot
        K, W, V = int(n_clusters), robot.width, abi.variance_width(robot.kind, robot.dof)
        soa = np.ascontiguousarray(cfg.T)
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        col = np.ascontiguousarray(np.asarray(collided).astype(np.uint8))
        out = dict(counts=np.zeros(K, dtype=np.int64), any_collided=np.zeros(K, dtype=np.uint8), expectations=np.zeros((K, W)),
                   variance=np.zeros(K), si_variance=np.zeros(K), variances=np.zeros((K, V)), si_variances=np.zeros((K, V)))
        abi.check(self._lib.upc_cluster_statistics(self.sim.handle, n, soa.ctypes.data, lab.ctypes.data, col.ctypes.data,
                                                   1 if allow_contacts else 0, float(step_size), K, out["counts"].ctypes.data,
                                                   out["any_collided"].ctypes.data, out["expectations"].ctypes.data,
                                                   out["variance"].ctypes.data, out["si_variance"].ctypes.data,
                                                   out["variances"].ctypes.data, out["si_variances"].ctypes.data))
        out["any_collided"] = out["any_collided"].astype(bool)
        return out

    def PolicyParticleClusteringFn(self, particles, current_config):
        """Index clusters of (particles + [current_config]) as lists of indices (:1558-1578)."""
        allp = np.concatenate([np.asarray(particles, dtype=np.float64),
                               np.asarray(current_config, dtype=np.float64).reshape(1, -1)], axis=0)
        if allp.shape[0] == 1:
            return [[0]]
        labels, ncl = self.ClusterIndices(allp)
        return [np.nonzero(labels == k)[0].tolist() for k in range(int(ncl[0]))]

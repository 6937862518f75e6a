/*
 * upc_b200.h - C ABI of the B200-native particle hot path of uncertainty_planning_core.
 *
 * The entry points below are what a reference-side plugin binds:
 *
 *   upc_forward_simulate*      replaces a concrete SimulatorInterface::ForwardSimulateRobots
 *                              (pure virtual at include/uncertainty_planning_core/simple_simulator_interface.hpp:623,
 *                               called at uncertainty_contact_planning.hpp:2065 and :1826)
 *   upc_check_config_collision replaces SimulatorInterface::CheckConfigCollision (simple_simulator_interface.hpp:617)
 *   upc_cluster_particles*     replaces UncertaintyPlanningSpace::ClusterParticles index clustering
 *                              (uncertainty_contact_planning.hpp:1986-2034: PerformSpatialFeatureClustering :1915 +
 *                               RunDistanceClusteringOnInitialClusters :1947) and PolicyParticleClusteringFn (:1558)
 *   upc_get_stats/reset_stats  back SimulatorInterface::GetStatistics / ResetStatistics (simple_simulator_interface.hpp:613-615)
 *
 * Everything is POD: plain pointers and sizes, int32 status codes, no C++ or torch types.
 * All floating-point data is IEEE double unless stated; configurations cross the boundary as
 * structure-of-arrays [C][n] (component-major, particle-minor) with C = upc_config_width(kind, dof).
 */
#ifndef UPC_B200_H
#define UPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UPC_MAX_DOF 16
#define UPC_MAX_JOINTS 16
#define UPC_MAX_LINKS 17

/* status codes */
#define UPC_OK 0
#define UPC_ERR_INVALID_ARGUMENT 1 /* maps to std::invalid_argument in the C++ shim */
#define UPC_ERR_CUDA 2             /* maps to std::runtime_error */
#define UPC_ERR_NO_ROBOT 3
#define UPC_ERR_UNSUPPORTED 4

/* robot kinds: the three instantiations fixed by uncertainty_planning_core.hpp:130-166 */
#define UPC_ROBOT_SE2 0
#define UPC_ROBOT_SE3 1
#define UPC_ROBOT_LINKED 2

/* SimpleJointModel::JOINT_TYPE values (simple_robot_models.hpp:881) */
#define UPC_JOINT_FIXED 0
#define UPC_JOINT_REVOLUTE 1
#define UPC_JOINT_CONTINUOUS 2
#define UPC_JOINT_PRISMATIC 4

/* SPATIAL_FEATURE_CLUSTERING_TYPE (uncertainty_contact_planning.hpp:52-99), plus NONE = distance pass only */
#define UPC_FEATURE_NONE 0
#define UPC_FEATURE_CONVEX_REGION_SIGNATURE 1
#define UPC_FEATURE_ACTUATION_CENTER_CONNECTIVITY 2
#define UPC_FEATURE_POINT_TO_POINT_MOVEMENT 3

typedef struct upc_handle upc_handle;

/*
 * Environment: dense voxel grids in the sdf_tools layout index = (x*ny + y)*nz + z.
 * grid_from_world is the inverse of the grids' origin transform (row-major 3x4), so that
 * p_grid = grid_from_world * p_world and voxel (ix,iy,iz) covers [ix,ix+1)*resolution etc.
 * A world point is in bounds iff 0 <= floor(p_grid/resolution) < n on every axis.
 */
typedef struct upc_env_desc
{
    int32_t nx, ny, nz;
    int32_t reserved0;
    double resolution;
    double grid_from_world[12];
    const float* sdf;               /* nx*ny*nz signed distances (metres), host memory */
    const uint32_t* convex_segment; /* nx*ny*nz convex-segment bitmasks, may be NULL (CRS clustering unavailable) */
    const float* occupancy;         /* nx*ny*nz occupancies, may be NULL (AC clustering unavailable) */
    float sdf_oob_value;            /* distance reported for out-of-bounds points */
    float occupancy_oob_value;      /* occupancy of the default cell returned for out-of-bounds lookups */
} upc_env_desc;

/* Per-axis controller / sensor / actuator parameters (SE2_ROBOT_CONFIG, SE3_ROBOT_CONFIG, LINKED_ROBOT_CONFIG
 * fields at simple_robot_models.hpp:38-90, :453-505, :1183-1214, expanded per axis). */
typedef struct upc_axis_params
{
    double kp, ki, kd, integral_clamp;
    double velocity_limit;
    double max_sensor_noise;
    double max_actuator_noise;
} upc_axis_params;

typedef struct upc_joint_desc
{
    int32_t type;         /* UPC_JOINT_* */
    int32_t parent_link;  /* index into links */
    int32_t child_link;
    int32_t reserved0;
    double axis[3];
    double lower_limit, upper_limit; /* ignored for continuous joints */
    double transform[12];            /* parent link origin -> joint, row-major 3x4 */
} upc_joint_desc;

/*
 * Robot model. Body points are (x, y, z, radius) in their link's frame, grouped by link:
 * link l owns points [link_point_offset[l], link_point_offset[l+1]). radius = 0 reproduces the reference's
 * point model (EigenHelpers::VectorVector4d with w = 1); radius > 0 gives sphere-approximated links.
 * SE2/SE3 robots have one link ("robot"), dof 3 / 6, and ignore the joint table.
 */
typedef struct upc_robot_desc
{
    int32_t kind; /* UPC_ROBOT_* */
    int32_t dof;  /* active degrees of freedom: 3, 6, or number of non-fixed joints */
    int32_t num_links;
    int32_t num_joints; /* linked only, fixed joints included, parents before children (simple_robot_models.hpp:1299-1343) */
    int32_t num_points;
    int32_t reserved0;
    int32_t link_point_offset[UPC_MAX_LINKS + 1];
    const double* points;                 /* num_points * 4 doubles, host memory */
    upc_axis_params axis[UPC_MAX_DOF];    /* SE2: x, y, zr; SE3: x, y, z, xr, yr, zr; linked: active joints in order */
    double distance_weights[UPC_MAX_DOF]; /* linked only (simple_robot_models.hpp:2191-2217) */
    double base_transform[12];            /* linked only */
    upc_joint_desc joints[UPC_MAX_JOINTS];
} upc_robot_desc;

/*
 * Simulation parameters: the ForwardSimulateRobots arguments (simple_simulator_interface.hpp:623) plus the
 * solver settings the concrete simulator owns (normative spec in DESIGN.md section 3).
 */
typedef struct upc_sim_params
{
    int32_t num_steps;        /* controller intervals per edge: max(1, (uint32)(forward_simulation_time * controller_frequency)) */
    int32_t max_microsteps;   /* M_max: actuator-draw slots per controller interval present on the tape (>= 1) */
    int32_t allow_contacts;
    int32_t use_individual_jacobians;
    int32_t max_resolver_iterations;   /* default 25 */
    int32_t resolver_decay_iterations; /* default 5 */
    int32_t reserved0, reserved1;
    double controller_interval;          /* seconds */
    double simulation_shortcut_distance; /* early exit when distance(config, target) < this; <= 0 disables */
    double contact_distance;             /* a body point collides iff sdf(p) - radius < contact_distance */
    double resolve_distance;             /* corrections push points out to this distance (>= contact_distance) */
    double microstep_distance;           /* max workspace motion per microstep (metres) */
    double resolver_initial_scale;       /* default 1.0 */
    double resolver_decay_rate;          /* default 0.5 */
    double resolver_damping;             /* lambda of the damped normal equations, default 1e-6 */
} upc_sim_params;

typedef struct upc_cluster_params
{
    int32_t feature_type; /* UPC_FEATURE_* */
    int32_t reserved0;
    double signature_matching_threshold;
    double distance_clustering_threshold;
    /* POINT_TO_POINT_MOVEMENT only (uncertainty_contact_planning.hpp:1801-1863) */
    double goal_distance_threshold;
    const upc_sim_params* ptpm_sim;
    const double* ptpm_tape; /* tape for the n*(n-1) pair simulations, same layout as upc_forward_simulate; NULL = seeded noise */
    uint64_t ptpm_seed;      /* seeded noise of the pair simulations (pair simulation k is particle k of the stream) */
} upc_cluster_params;

/* counters surfaced through GetStatistics() */
typedef struct upc_stats
{
    int64_t kernel_launches;
    int64_t particles_simulated;
    int64_t particle_steps;         /* controller intervals actually executed (early exits excluded) */
    int64_t collided_particles;
    int64_t resolver_iterations;    /* total contact-resolver iterations */
    int64_t failed_resolves;        /* microsteps reverted because the resolver did not clear the contact */
    int64_t cluster_calls;
    int64_t cluster_fallback_calls; /* particle sets that needed the exact complete-link path */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
} upc_stats;

/* number of doubles per configuration: SE2 3 (x, y, theta); SE3 12 (R row-major, then t); linked dof */
int32_t upc_config_width(int32_t kind, int32_t dof);

/* doubles per particle on the noise tape: num_steps * (1 + max_microsteps) * dof */
int64_t upc_tape_doubles_per_particle(const upc_sim_params* params, int32_t dof);

const char* upc_last_error(void);
const char* upc_version(void);

int32_t upc_create(const upc_env_desc* env, int32_t device, upc_handle** out);
int32_t upc_destroy(upc_handle* h);
int32_t upc_set_robot(upc_handle* h, const upc_robot_desc* robot);
int32_t upc_get_stats(upc_handle* h, upc_stats* out);
int32_t upc_reset_stats(upc_handle* h);
/* tuning: lanes cooperating on one particle (1, 2, 3, 4, 5, 6, 8, 16 or 32; counts that do not divide 32 leave the last lanes of a warp idle);
 * 0 restores the built-in choice. Results do not depend on it. */
int32_t upc_set_lanes_per_particle(upc_handle* h, int32_t lanes);

/* pinned host memory helpers for callers that want the fast H2D/D2H path */
int32_t upc_host_alloc(void** ptr, int64_t bytes);
int32_t upc_host_free(void* ptr);

/*
 * Batched forward simulation. starts: [C][n]; targets: [C][n_targets] with n_targets == 1 or n;
 * (n_targets may also be any divisor of n: one target per block of n / n_targets consecutive particles)
 * tape: [num_steps][1 + max_microsteps][dof][n] pre-truncated draws (slot 0 of each step = sensor draws in
 * [-max_sensor_noise, +max_sensor_noise]; slots 1.. = actuator draws in [-1, 1]); out_configs: [C][n];
 * out_collided: n bytes, 1 iff the particle touched the environment during the edge.
 * Host-pointer form: copies inputs to the device, runs, copies results back, synchronises.
 */
int32_t upc_forward_simulate(upc_handle* h, const upc_sim_params* params, int64_t n, const double* starts,
                             int64_t n_targets, const double* targets, const double* tape, double* out_configs,
                             uint8_t* out_collided);
/* Device-pointer form: all buffers in device memory, asynchronous on `stream` (a cudaStream_t; NULL selects the handle's own
 * non-blocking stream - pass cudaStreamLegacy (0x1) to name the legacy default stream explicitly). */
int32_t upc_forward_simulate_device(upc_handle* h, const upc_sim_params* params, int64_t n, const double* d_starts,
                                    int64_t n_targets, const double* d_targets, const double* d_tape,
                                    double* d_out_configs, uint8_t* d_out_collided, void* stream);

/*
 * Seeded forms: no tape crosses the boundary.  Draw (particle first_particle + i, controller interval,
 * slot, axis) is a pure function of `seed` - Philox-4x32-10 bits through the inverse CDF of the normal distribution truncated
 * to +-2 sigma - evaluated inside the kernel; this replaces the per-thread generators of
 * uncertainty_contact_planning.hpp:343-356 and the draws of simple_uncertainty_models.hpp:38-45, 77-88.  Results do not
 * depend on batch size, lane count or GPU count: keying is by global particle index.
 * starts: [C][n_starts], targets: [C][n_targets]; n_starts and n_targets must each divide n, and particle i uses entry
 * i / (n / n_x) - one entry per particle (n_x == n), one per edge of an edge batch, or a single one (n_x == 1).
 */
int32_t upc_forward_simulate_seeded(upc_handle* h, const upc_sim_params* params, uint64_t seed, int64_t first_particle, int64_t n,
                                    int64_t n_starts, const double* starts, int64_t n_targets, const double* targets, double* out_configs,
                                    uint8_t* out_collided);
int32_t upc_forward_simulate_seeded_device(upc_handle* h, const upc_sim_params* params, uint64_t seed, int64_t first_particle, int64_t n,
                                           int64_t n_starts, const double* d_starts, int64_t n_targets, const double* d_targets,
                                           double* d_out_configs, uint8_t* d_out_collided, void* stream);
/* the same draws written out as a tape (host memory, layout of upc_forward_simulate): replaying it reproduces the seeded results bit for bit */
int32_t upc_fill_noise_tape_seeded(uint64_t seed, const upc_robot_desc* robot, const upc_sim_params* params, int64_t first_particle,
                                   int64_t n, int64_t tape_stride, double* tape);

/*
 * Traced simulation of a small batch: ForwardSimulateRobot / ForwardSimulateMutableRobot with enable_tracing
 * (simple_simulator_interface.hpp:619-621; trace types :40-63; consumer ExtractTrajectoryFromTrace :66-86, used at
 * uncertainty_contact_planning.hpp:510 and :1171).  tape == NULL selects seeded noise (seed, first_particle).
 *   trace_configs  [num_steps][C][n]      resolved configuration at the end of every controller interval
 *   trace_controls [num_steps][2*dof][n]  control input (velocity command) and the control step applied per microstep
 *   trace_steps    [2][n]                 controller intervals executed (early exit: < num_steps), microsteps of the last one
 * Host pointers; synchronises.
 */
int32_t upc_forward_simulate_traced(upc_handle* h, const upc_sim_params* params, int64_t n, const double* starts, int64_t n_targets,
                                    const double* targets, const double* tape, uint64_t seed, int64_t first_particle, double* out_configs,
                                    uint8_t* out_collided, double* trace_configs, double* trace_controls, int32_t* trace_steps);

/*
 * The planner-batch step in ONE call (BASELINE config 5): n_edges candidate edges x particles_per_edge particles simulated with
 * seeded noise (uncertainty_contact_planning.hpp:2039-2071 for every edge) and the outcome of every edge clustered (:1986-2034),
 * outcomes staying in device memory in between.  edge_starts: [C][n_starts] with n_starts == n_edges (every particle of an edge
 * starts at the edge's start: a fresh node, uncertainty_planner_state.hpp:527-529) or n_starts == n (the parent state's own
 * particles, uncertainty_contact_planning.hpp:2046-2055, edge-major); edge_targets: [C][n_edges]; out_configs [C][n],
 * out_collided [n], out_labels [n] (cluster of the particle within its edge), out_n_clusters [n_edges];
 * n = n_edges * particles_per_edge, particles edge-major.  Host pointers; synchronises.
 */
int32_t upc_simulate_and_cluster_edges(upc_handle* h, const upc_sim_params* params, const upc_cluster_params* cparams, uint64_t seed,
                                       int64_t first_particle, int64_t n_edges, int64_t particles_per_edge, int64_t n_starts, const double* edge_starts,
                                       const double* edge_targets, double* out_configs, uint8_t* out_collided, int32_t* out_labels,
                                       int32_t* out_n_clusters);

/*
 * The same step as a pair of calls, for callers that keep two batches in flight (slot = 0 / 1): upc_edge_batch_submit uploads the
 * end points of a batch and queues its simulation - asynchronous, it returns as soon as the work is queued; the host arrays must
 * stay valid until the batch has been collected - and upc_edge_batch_collect clusters the batch in `slot` on the handle's second,
 * high-priority stream and brings configurations, flags, labels and cluster counts to the host (blocks until they are there).
 * Issue submit(k + 1) before collect(k): the simulation of the next batch fills the machine while the clustering and the
 * download of this one run underneath.  Results equal upc_simulate_and_cluster_edges for the same arguments.
 */
int32_t upc_edge_batch_submit(upc_handle* h, const upc_sim_params* params, uint64_t seed, int64_t first_particle, int64_t n_edges,
                              int64_t particles_per_edge, int64_t n_starts, const double* edge_starts, const double* edge_targets, int32_t slot);
int32_t upc_edge_batch_collect(upc_handle* h, const upc_cluster_params* cparams, int32_t slot, double* out_configs, uint8_t* out_collided,
                               int32_t* out_labels, int32_t* out_n_clusters);

/* CheckConfigCollision: 1 iff any body point of `config` has sdf - radius < contact_distance + inflation */
int32_t upc_check_config_collision(upc_handle* h, const double* config, double contact_distance, double inflation,
                                   int32_t* out_in_collision);

/*
 * Outcome clustering of n_sets independent particle sets of set_size particles each
 * (configs: [C][n_sets*set_size], set s owns particles [s*set_size, (s+1)*set_size)).
 * labels[i] = cluster of particle i within its set, clusters numbered by increasing smallest member;
 * n_clusters[s] = cluster count of set s. Collided particles take part in clustering
 * (uncertainty_contact_planning.hpp:2000-2025 drops them afterwards, in the shim).
 */
int32_t upc_cluster_particles(upc_handle* h, const upc_cluster_params* params, int64_t n_sets, int64_t set_size,
                              const double* configs, int32_t* labels, int32_t* n_clusters);
int32_t upc_cluster_particles_device(upc_handle* h, const upc_cluster_params* params, int64_t n_sets, int64_t set_size,
                                     const double* d_configs, int32_t* d_labels, int32_t* d_n_clusters, void* stream);

/*
 * Host-side noise tape generator: fills tape[.][first..first+n) for particles
 * first_particle .. first_particle+n-1 of a batch whose tape has `tape_stride` particles per row
 * (tape_stride >= n; pass tape + offset to fill a shard).  Particle i always draws from its own
 * std::mt19937_64 stream keyed by (seed, i): tapes do not depend on thread count or GPU sharding.
 */
int32_t upc_fill_noise_tape(uint64_t seed, const upc_robot_desc* robot, const upc_sim_params* params, int64_t first_particle,
                            int64_t n, int64_t tape_stride, double* tape);

/*
 * "Next" row: GetNearestNeighbor (uncertainty_contact_planning.hpp:1506-1542) for n_queries sampled
 * configurations at once.  Tree state i is described by what StateDistance (:1488-1504) reads from it: its expectation
 * (expectations [C][n_states]), feasibility_weight[i] = (1 - Pfeasibility) * alpha + (1 - alpha),
 * variance_weight[i] = erf(|space-independent variances|_1) * alpha_v + (1 - alpha_v) - both change only when the state
 * does, so the caller keeps them - and use_for_nn[i] (NULL = all enabled).
 *   distance(i, q) = (feasibility_weight[i] * (ComputeConfigurationDistance(expectation_i, query_q) / step_size)) * variance_weight[i]
 * out_index[q] = the enabled state of smallest distance (strict <, so ties keep the smallest index, as the reference's
 * static-schedule loop does; NaN distances are never selected), -1 if there is none; out_distance[q] (may be NULL) its
 * distance, INFINITY if none.  Host pointers; synchronises.
 */
int32_t upc_nearest_neighbors(upc_handle* h, int64_t n_states, const double* expectations, const double* feasibility_weight,
                              const double* variance_weight, const uint8_t* use_for_nn, int64_t n_queries, const double* queries,
                              double step_size, int64_t* out_index, double* out_distance);

/*
 * "Next" row: the count behind ComputeReverseEdgeProbability (uncertainty_contact_planning.hpp:2073-2090)
 * and the goal-reached fraction: how many of n configurations ([C][n]) lie within `threshold` (<=) of their target
 * (targets [C][n_targets], n_targets == 1 or n) under the robot's ComputeConfigurationDistance.  Also writes the
 * per-particle flags when `within` is not NULL.  Host pointers; synchronises.
 */
int32_t upc_count_within(upc_handle* h, int64_t n, const double* configs, int64_t n_targets, const double* targets, double threshold,
                         int64_t* out_count, uint8_t* within);

/*
 * "Next" row: per-cluster statistics of an outcome, the reduction behind
 * UncertaintyPlannerState::UpdateStatistics (uncertainty_planner_state.hpp:339-351, 591-714) and the counts of
 * uncertainty_contact_planning.hpp:2105-2162.  For each cluster k in [0, n_clusters), over the particles with
 * labels[i] == k that survive the contact filter (collided[i] == 0 || allow_contacts), in ascending particle order:
 *   counts[k], any_collided[k]
 *   expectations[k*C .. )          AverageConfigurations (arithmetic / circular means; SE(3): mean translation + sign-aligned
 *                                  normalised quaternion mean - DESIGN.md section 6)
 *   variance[k], si_variance[k]    sum of d(expectation, p)^2 / n  and  (d / step_size)^2 / n
 *   variances[k*V .. ), si_variances[k*V .. )   per-dimension versions; V = upc_variance_width(kind, dof)
 * A cluster with no surviving particle reports count 0 and zeros.  Host pointers; synchronises.
 */
int32_t upc_variance_width(int32_t kind, int32_t dof); /* SE2 3, SE3 4 (dx, dy, dz, rotation), linked dof */
int32_t upc_cluster_statistics(upc_handle* h, int64_t n, const double* configs, const int32_t* labels, const uint8_t* collided,
                               int32_t allow_contacts, double step_size, int32_t n_clusters, int64_t* counts, uint8_t* any_collided,
                               double* expectations, double* variance, double* si_variance, double* variances, double* si_variances);

/* Roofline denominators measured on the device the handle lives on (diagnostics, not on the hot path):
 * random 32-byte-sector gather bandwidth over a buffer of `buffer_bytes` (L2-resident when it fits), and the FP64 FMA rate. */
int32_t upc_measure_l2_gather_gbs(upc_handle* h, int64_t buffer_bytes, double* out_gbs);
int32_t upc_measure_fp64_tflops(upc_handle* h, double* out_tflops);

/* Device memory for hosts that do not link the CUDA runtime themselves (the C++ shim's multi-GPU path): four grow-only scratch
 * buffers per handle (slot 0..3; contents are lost when a slot grows) and blocking copies ordered on the handle's stream. */
int32_t upc_device_scratch(upc_handle* h, int32_t slot, int64_t bytes, void** out);
int32_t upc_memcpy_to_device(upc_handle* h, void* d_dst, const void* src, int64_t bytes);
int32_t upc_memcpy_to_host(upc_handle* h, void* dst, const void* d_src, int64_t bytes);

/* blocks until everything queued on the handle's stream has finished */
int32_t upc_synchronize(upc_handle* h);

/*
 * Multi-GPU (BASELINE north_star: "the batch of candidate RRT edges or particles is sharded with the SDF
 * replicated on every GPU and a single NCCL gather of outcome clusters").  Every rank owns a complete handle (its own copy of
 * the environment and robot); a upc_comm ties the handles of one box together through NCCL (bound at run time).
 *   one process per GPU:   rank 0 calls upc_comm_get_unique_id, ships the 128 bytes to the other ranks by any means, every
 *                          rank calls upc_comm_create(handle, world, rank, id)
 *   one process, n GPUs:   upc_comm_create_all(n, handles, comms) and the *_all entry point below
 */
typedef struct upc_comm upc_comm;
#define UPC_COMM_ID_BYTES 128
int32_t upc_comm_get_unique_id(uint8_t* id128);
int32_t upc_comm_create(upc_handle* h, int32_t world, int32_t rank, const uint8_t* id128, upc_comm** out);
int32_t upc_comm_create_all(int32_t n, upc_handle* const* handles, upc_comm** out);
int32_t upc_comm_destroy(upc_comm* c);
int32_t upc_comm_world(const upc_comm* c);
int32_t upc_comm_rank(const upc_comm* c);
int32_t upc_comm_nccl_version(int32_t* out);
/* contiguous, balanced [lo, hi) range of `total` items owned by `rank` */
void upc_shard_bounds(int64_t total, int32_t rank, int32_t world, int64_t* lo, int64_t* hi);

/* Outcome block of one rank inside the gather buffer (all offsets in bytes from the block start, 256-byte aligned):
 * labels int32[particles] | n_clusters int32[sets] | collided uint8[particles] | configs double[C][n_local] (stride = the
 * rank's own particle count).  `particles` / `sets` are the capacities, i.e. the largest shard. */
typedef struct upc_outcome_layout
{
    int64_t particles, sets;
    int64_t labels_offset, n_clusters_offset, collided_offset, configs_offset;
    int64_t block_bytes;
} upc_outcome_layout;
int32_t upc_outcome_layout_for(int32_t width, int64_t particles, int64_t sets, upc_outcome_layout* out);

/* In-place all-gather of the outcome blocks (rank r's block at d_blocks + r * block_bytes on every rank), ordered after the
 * work queued on `stream` (NULL = the handle's stream) and executed on the communicator's own stream; `stream` does not wait
 * for it - upc_comm_wait makes a stream wait, upc_comm_synchronize blocks the host.  Work queued on `stream` after the call
 * does wait for the exchange issued by the PREVIOUS call, so a caller that alternates between two gather buffers may issue
 * the next step straight away (the exchange then overlaps the next step's simulation); with a single buffer call
 * upc_comm_wait before the next step. */
int32_t upc_comm_all_gather_outcomes(upc_comm* c, void* d_blocks, int64_t block_bytes, void* stream);
int32_t upc_comm_wait(upc_comm* c, void* stream);
int32_t upc_comm_synchronize(upc_comm* c);

/*
 * One sharded planner-batch step: this rank simulates (seeded, keyed by global particle index) and clusters its contiguous
 * block of the n_edges edges straight into its slot of d_blocks, then the single all-gather runs.  d_edge_starts /
 * d_edge_targets: [C][n_edges] in device memory, replicated; d_blocks: world * block_bytes with the layout of
 * upc_outcome_layout_for(width, ceil(n_edges / world) * particles_per_edge, ceil(n_edges / world)).  Asynchronous.
 */
int32_t upc_sharded_edge_batch(upc_comm* c, const upc_sim_params* params, const upc_cluster_params* cparams, uint64_t seed,
                               int64_t n_edges, int64_t particles_per_edge, const double* d_edge_starts, const double* d_edge_targets,
                               void* d_blocks, void* stream);
/*
 * The two halves of that step as separate calls, for callers that pipeline batches with two streams: the simulation of batch
 * k + 1 (stream A: fills the machine) overlaps the clustering and the exchange of batch k (stream B: small latency-bound
 * launches and the host decisions between them).  _simulate is asynchronous and never waits on the host; _finish clusters this
 * rank's block on `stream` - order it after the simulation of the same batch with an event - and issues the all-gather behind
 * it.  Simulation and clustering of one handle touch disjoint device state, so both may be in flight, issued from one host
 * thread; the d_blocks of batches in flight must be distinct buffers (call upc_comm_wait on stream A before a buffer is reused).
 */
int32_t upc_sharded_edge_batch_simulate(upc_comm* c, const upc_sim_params* params, uint64_t seed, int64_t n_edges, int64_t particles_per_edge,
                                        const double* d_edge_starts, const double* d_edge_targets, void* d_blocks, void* stream);
int32_t upc_sharded_edge_batch_finish(upc_comm* c, const upc_cluster_params* cparams, int64_t n_edges, int64_t particles_per_edge, void* d_blocks,
                                      void* stream);
/* the same step for all n ranks of a single-process communicator set: one entry per rank in every array (device pointers on
 * that rank's GPU; streams may be NULL) */
int32_t upc_sharded_edge_batch_all(int32_t n, upc_comm* const* comms, const upc_sim_params* params, const upc_cluster_params* cparams, uint64_t seed,
                                   int64_t n_edges, int64_t particles_per_edge, const double* const* d_edge_starts,
                                   const double* const* d_edge_targets, void* const* d_blocks, void* const* streams);

#ifdef __cplusplus
}
#endif

#endif /* UPC_B200_H */

/*
 * upc_oracle.hpp - CPU oracle for the particle hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, link or
 * call anything under oracle/.  The product (uncertainty_planning_core_b200/, include/) never does.
 *
 * PARITY STATUS: *partially pinned*.  The reference (UM-ARM-Lab/uncertainty_planning_core) ships no tests and no golden vectors,
 * its simulator is pure virtual (simple_simulator_interface.hpp:613-623) and the heavy arithmetic lives in un-vendored, un-pinned
 * arc_utilities / sdf_tools (package.xml:11-19).  What CAN be compiled from the reference is compiled UNMODIFIED into oracle/_ref
 * (oracle/Makefile) and this oracle is tested against it: simple_pid_controller.hpp (tests/test_oracle_pid.py),
 * simple_uncertainty_models.hpp (tests/test_oracle_unc.py), simple_robot_models.hpp and uncertainty_planner_state.hpp on the stand-in
 * Eigen / arc_utilities headers of oracle/ref_standins (tests/test_oracle_rob.py: control / apply sequences, kinematics, Jacobians,
 * distances, averages, per-cluster statistics of all three robots).  "Parity unpinned" remains for what the reference does not
 * contain or cannot compile: the simulator loop, SDF lookup and contact resolver, complete-link clustering and the clustering member
 * functions of UncertaintyPlanningSpace - restated from the cited lines / specified in DESIGN.md.
 *
 * Style: sequential, one robot object per particle, array-of-structs configurations - the shape of the
 * reference's CPU path (OpenMP parallel-for over particles / pairs as at uncertainty_contact_planning.hpp:1509,
 * 1588, 1770) - deliberately not the GPU's data layout.
 */
#ifndef UPC_ORACLE_HPP
#define UPC_ORACLE_HPP

#include <stdint.h>
#include <vector>
#include <string>
#include <utility>
#include "upc_b200.h"
#include "upc_math.h"

namespace upc_oracle
{
/* rigid transform, rows of [R | t] */
struct Mat34
{
    double m[3][4];
};

Mat34 Identity34();
Mat34 Compose(const Mat34& a, const Mat34& b);
void TransformPoint(const Mat34& a, const double p[3], double out[3]);

/* simple_pid_controller.hpp:53-136 */
struct Pid
{
    double kp, ki, kd, integral_clamp, error_integral, last_error;
    void Initialize(double in_kp, double in_ki, double in_kd, double in_clamp);
    void Zero();
    double ComputeFeedbackTerm(double current_error, double timestep);
};

/* simple_uncertainty_models.hpp:20-46 with the draw supplied by the tape */
inline double SensorValue(double process_value, double draw) { return process_value + draw; }

/* simple_uncertainty_models.hpp:48-106 */
struct Actuator
{
    double actuator_limit, noise_bound;
    void Initialize(double in_noise_bound, double in_limit);
    double GetControlValue(double control_input) const;
    double GetControlValue(double control_input, double draw) const;
};

/* SE(3) helpers (normative closed forms; EigenHelpers un-vendored) */
void Se3Log(const Mat34& t, double twist[6]);
Mat34 Se3Exp(const double twist[6]);
Mat34 Se3Between(const Mat34& a, const Mat34& b); /* a^-1 * b */
void QuaternionFromRotation(const Mat34& t, double q[4]); /* x, y, z, w */

struct Environment
{
    int32_t nx, ny, nz;
    double resolution, inv_resolution;
    Mat34 grid_from_world;
    const float* sdf;
    const uint32_t* convex_segment;
    const float* occupancy;
    float sdf_oob_value, occupancy_oob_value;

    explicit Environment(const upc_env_desc& d);
    /* voxel index of a world point; false when out of bounds */
    bool LocationToGridIndex(const double world[3], int64_t idx[3]) const;
};

struct SimStats
{
    int64_t particle_steps, collided_particles, resolver_iterations, failed_resolves;
};

/* complete-link threshold clustering on a dense symmetric matrix */
std::vector<std::vector<int64_t>> CompleteLinkCluster(const std::vector<double>& matrix, int64_t n, double threshold);
}

extern "C" {
const char* oracle_last_error(void);
int32_t oracle_num_threads(void);
void oracle_set_num_threads(int32_t n);
double oracle_pid_sequence(double kp, double ki, double kd, double clamp, const double* errors, const double* dts,
                           int32_t n, double* out_terms);
/* elementary functions, vectorised, for tests/test_math.py; which: 0 sin 1 cos 2 atan2(a,b) 3 acos 4 wrap 5 signed angle distance(a,b) */
void oracle_math(int32_t which, int64_t n, const double* a, const double* b, double* out);
void oracle_sensor_values(int64_t n, const double* process_values, const double* draws, double* out);
void oracle_actuator_values(double noise_bound, double actuator_limit, int64_t n, const double* inputs, const double* draws, double* out, double* noiseless);
void oracle_se3_log(const double* cfg12, double* twist6);
void oracle_se3_exp(const double* twist6, double* cfg12);
void oracle_quaternion_from_rotation(const double* cfg12, double* q4); /* x, y, z, w */
int32_t oracle_forward_simulate(const upc_env_desc* env, const upc_robot_desc* robot, const upc_sim_params* params,
                                int64_t n, const double* starts, int64_t n_targets, const double* targets,
                                const double* tape, double* out_configs, uint8_t* out_collided, upc_stats* stats);
int32_t oracle_forward_simulate_traced(const upc_env_desc* env, const upc_robot_desc* robot, const upc_sim_params* params, int64_t n,
                                       const double* starts, int64_t n_targets, const double* targets, const double* tape, double* out_configs,
                                       uint8_t* out_collided, double* trace_configs, double* trace_controls, int32_t* trace_steps);
/* counter-based noise (upc_oracle_noise.cpp) */
int32_t oracle_fill_noise_tape_counter(uint64_t seed, const upc_robot_desc* robot, const upc_sim_params* params, int64_t first_particle,
                                       int64_t n, int64_t tape_stride, double* tape);
int32_t oracle_check_config_collision(const upc_env_desc* env, const upc_robot_desc* robot, const double* config,
                                      double contact_distance, double inflation, int32_t* out);
int32_t oracle_cluster_particles(const upc_env_desc* env, const upc_robot_desc* robot, const upc_cluster_params* params,
                                 int64_t n_sets, int64_t set_size, const double* configs, int32_t* labels,
                                 int32_t* n_clusters, upc_stats* stats);
/* pieces exposed for unit tests */
int32_t oracle_pair_distances(const upc_robot_desc* robot, int64_t n, const double* configs, double* matrix);
int32_t oracle_region_signatures(const upc_env_desc* env, const upc_robot_desc* robot, int64_t n, const double* configs,
                                 uint32_t* signatures);
int32_t oracle_sdf_lookup(const upc_env_desc* env, int64_t n, const double* world_points, double* dist, double* grad);
int32_t oracle_complete_link(const double* matrix, int64_t n, double threshold, int32_t* labels, int32_t* n_clusters);
int32_t oracle_cluster_statistics(const upc_robot_desc* robot, int64_t n, const double* configs, const int32_t* labels, const uint8_t* collided,
                                  int32_t allow_contacts, double step_size, int32_t n_clusters, int64_t* counts, uint8_t* any_collided,
                                  double* expectations, double* variance, double* si_variance, double* variances, double* si_variances);
/* GetNearestNeighbor / StateDistance, uncertainty_contact_planning.hpp:1488-1542 (per-state weights supplied by the caller) */
int32_t oracle_nearest_neighbors(const upc_robot_desc* robot, int64_t n_states, const double* expectations, const double* feasibility_weight,
                                 const double* variance_weight, const uint8_t* use_for_nn, int64_t n_queries, const double* queries,
                                 double step_size, int64_t* out_index, double* out_distance);
/* reference-shaped distance pass (std::function per pair + heap-allocated delta vector per pair, as at
 * uncertainty_contact_planning.hpp:1955-1960 and simple_robot_models.hpp:412,2195), for the honest CPU baseline */
int32_t oracle_pair_distances_reference_style(const upc_robot_desc* robot, int64_t n, const double* configs, double* matrix);
}

#endif

/*
 * upc_b200_shim.hpp - header-only C++ host shim over the C ABI (upc_b200.h).
 *
 * B200Simulator<Robot, Configuration, RNG, ConfigAlloc> implements the reference's plugin interface
 *   simple_simulator_interface::SimulatorInterface<Robot, Configuration, RNG, ConfigAlloc>
 *   (include/uncertainty_planning_core/simple_simulator_interface.hpp:383-624; pure virtuals at :613-623)
 * and B200OutcomeClustering<Robot, Configuration, ConfigAlloc> carries the two clustering entry points of
 *   UncertaintyPlanningSpace (uncertainty_contact_planning.hpp:1986 ClusterParticles, :1558 PolicyParticleClusteringFn).
 *
 * Two build modes:
 *   - with -DUPC_SHIM_WITH_REFERENCE the class derives from the real SimulatorInterface (needs the reference's
 *     headers, Eigen, ROS, arc_utilities, sdf_tools) and overrides its virtuals with the exact signatures;
 *   - without it (this container has none of those dependencies) the same code compiles stand-alone: the
 *     methods keep the reference's names, argument order and meaning, with the ros::Publisher argument and the
 *     sdf_tools grids replaced by the POD descriptors of upc_b200.h.  tests/cpp/shim_smoke.cpp exercises it.
 *
 * What a model adapter has to provide (ConfigTraits / RobotTraits below): how a Configuration is flattened to
 * upc_config_width doubles, and the robot description (body points, per-axis gains / limits / noise bounds,
 * joint table).  The reference's robot classes do not expose their sensor noise bounds, so the description is
 * given to the simulator at construction, next to the robot the planner holds.
 *
 * Error convention: status codes from the C ABI become exceptions - std::invalid_argument for
 * UPC_ERR_INVALID_ARGUMENT, std::runtime_error otherwise.  Contract violations the reference asserts on
 * (target_positions.size() not in {1, N}) are reported as std::invalid_argument.
 */
#ifndef UPC_B200_SHIM_HPP
#define UPC_B200_SHIM_HPP

#include <array>
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <limits>
#include <vector>

#include "upc_b200.h"

#ifdef UPC_SHIM_WITH_REFERENCE
#include <uncertainty_planning_core/simple_simulator_interface.hpp>
#include <uncertainty_planning_core/simple_robot_models.hpp>
#include <uncertainty_planning_core/uncertainty_contact_planning.hpp>
#endif

namespace upc_b200
{
inline void check(int32_t code)
{
    if (code == UPC_OK) return;
    const std::string msg = std::string("upc_b200: ") + upc_last_error();
    if (code == UPC_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}

/* ---- reference types mirrored for the stand-alone build ------------------------------------------------------- */

#ifdef UPC_SHIM_WITH_REFERENCE
using uncertainty_contact_planning::SPATIAL_FEATURE_CLUSTERING_TYPE;
using uncertainty_contact_planning::CONVEX_REGION_SIGNATURE;
using uncertainty_contact_planning::ACTUATION_CENTER_CONNECTIVITY;
using uncertainty_contact_planning::POINT_TO_POINT_MOVEMENT;
using uncertainty_contact_planning::COMPARE;
template <typename Configuration, typename ConfigAlloc = std::allocator<Configuration>>
using ForwardSimulationContactResolverStepTrace = simple_simulator_interface::ForwardSimulationContactResolverStepTrace<Configuration, ConfigAlloc>;
template <typename Configuration, typename ConfigAlloc = std::allocator<Configuration>>
using ForwardSimulationResolverTrace = simple_simulator_interface::ForwardSimulationResolverTrace<Configuration, ConfigAlloc>;
template <typename Configuration, typename ConfigAlloc = std::allocator<Configuration>>
using ForwardSimulationStepTrace = simple_simulator_interface::ForwardSimulationStepTrace<Configuration, ConfigAlloc>;
typedef Eigen::VectorXd ControlVector;
inline ControlVector MakeControlVector(const double* v, int n) { ControlVector out(n); for (int i = 0; i < n; i++) out(i) = v[i]; return out; }
#else
/* uncertainty_contact_planning.hpp:52, same enumerators in the same order (CRS = 0, AC = 1, PTPM = 2, COMPARE = 3) */
enum SPATIAL_FEATURE_CLUSTERING_TYPE {CONVEX_REGION_SIGNATURE, ACTUATION_CENTER_CONNECTIVITY, POINT_TO_POINT_MOVEMENT, COMPARE};
/* simple_simulator_interface.hpp:40-63, with std::vector<double> standing in for Eigen::VectorXd */
typedef std::vector<double> ControlVector;
inline ControlVector MakeControlVector(const double* v, int n) { return ControlVector(v, v + n); }
template <typename Configuration, typename ConfigAlloc = std::allocator<Configuration>>
struct ForwardSimulationContactResolverStepTrace
{
    std::vector<Configuration, ConfigAlloc> contact_resolution_steps;
};
template <typename Configuration, typename ConfigAlloc = std::allocator<Configuration>>
struct ForwardSimulationResolverTrace
{
    ControlVector control_input;
    ControlVector control_input_step;
    std::vector<ForwardSimulationContactResolverStepTrace<Configuration, ConfigAlloc>> contact_resolver_steps;
};
template <typename Configuration, typename ConfigAlloc = std::allocator<Configuration>>
struct ForwardSimulationStepTrace
{
    std::vector<ForwardSimulationResolverTrace<Configuration, ConfigAlloc>> resolver_steps;
    inline void Reset() { resolver_steps.clear(); }
};
/* simple_simulator_interface.hpp:66-86 */
template <typename Configuration, typename ConfigAlloc = std::allocator<Configuration>>
inline std::vector<Configuration, ConfigAlloc> ExtractTrajectoryFromTrace(const ForwardSimulationStepTrace<Configuration, ConfigAlloc>& trace)
{
    std::vector<Configuration, ConfigAlloc> execution_trajectory;
    execution_trajectory.reserve(trace.resolver_steps.size());
    for (size_t step_idx = 0; step_idx < trace.resolver_steps.size(); step_idx++)
        execution_trajectory.push_back(trace.resolver_steps[step_idx].contact_resolver_steps.back().contact_resolution_steps.back());
    return execution_trajectory;
}
#endif

/*
 * SPATIAL_FEATURE_CLUSTERING_TYPE (uncertainty_contact_planning.hpp:52: CRS = 0, AC = 1, PTPM = 2, COMPARE = 3) -> UPC_FEATURE_*
 * (NONE = 0, CRS = 1, AC = 2, PTPM = 3).  COMPARE runs all three and prints a comparison (uncertainty_contact_planning.hpp:1915-1945):
 * a diagnostic mode, not offered here.
 */
inline int32_t FromReferenceClusteringType(const SPATIAL_FEATURE_CLUSTERING_TYPE clustering_type)
{
    if (clustering_type == CONVEX_REGION_SIGNATURE) return UPC_FEATURE_CONVEX_REGION_SIGNATURE;
    if (clustering_type == ACTUATION_CENTER_CONNECTIVITY) return UPC_FEATURE_ACTUATION_CENTER_CONNECTIVITY;
    if (clustering_type == POINT_TO_POINT_MOVEMENT) return UPC_FEATURE_POINT_TO_POINT_MOVEMENT;
    throw std::invalid_argument("upc_b200: spatial feature clustering type must be CRS, AC or PTPM (COMPARE is not supported)");
}

/* ---- configuration marshalling ------------------------------------------------------------------------- */

template <typename Configuration>
struct ConfigTraits; /* static int width(const Configuration&); static void flatten(const Configuration&, double*);
                        static Configuration inflate(const double*, const Configuration& like); */

template <size_t N>
struct ConfigTraits<std::array<double, N>>
{
    static int width(const std::array<double, N>&) { return (int)N; }
    static void flatten(const std::array<double, N>& c, double* out) { for (size_t i = 0; i < N; i++) out[i] = c[i]; }
    static std::array<double, N> inflate(const double* in, const std::array<double, N>&)
    {
        std::array<double, N> c;
        for (size_t i = 0; i < N; i++) c[i] = in[i];
        return c;
    }
};

#ifdef UPC_SHIM_WITH_REFERENCE
template <>
struct ConfigTraits<Eigen::Matrix<double, 3, 1>> /* SE2Config, uncertainty_planning_core.hpp:132 */
{
    static int width(const Eigen::Matrix<double, 3, 1>&) { return 3; }
    static void flatten(const Eigen::Matrix<double, 3, 1>& c, double* out) { out[0] = c(0); out[1] = c(1); out[2] = c(2); }
    static Eigen::Matrix<double, 3, 1> inflate(const double* in, const Eigen::Matrix<double, 3, 1>&) { return Eigen::Matrix<double, 3, 1>(in[0], in[1], in[2]); }
};
template <>
struct ConfigTraits<Eigen::Isometry3d> /* SE3Config, uncertainty_planning_core.hpp:144: R row-major then t */
{
    static int width(const Eigen::Isometry3d&) { return 12; }
    static void flatten(const Eigen::Isometry3d& c, double* out)
    {
        for (int r = 0; r < 3; r++) { for (int k = 0; k < 3; k++) out[r * 3 + k] = c(r, k); out[9 + r] = c(r, 3); }
    }
    static Eigen::Isometry3d inflate(const double* in, const Eigen::Isometry3d&)
    {
        Eigen::Isometry3d c = Eigen::Isometry3d::Identity();
        for (int r = 0; r < 3; r++) { for (int k = 0; k < 3; k++) c(r, k) = in[r * 3 + k]; c(r, 3) = in[9 + r]; }
        return c;
    }
};
template <>
struct ConfigTraits<std::vector<simple_robot_models::SimpleJointModel>> /* LinkedConfig, uncertainty_planning_core.hpp:156 */
{
    typedef std::vector<simple_robot_models::SimpleJointModel> C;
    static int width(const C& c) { return (int)c.size(); }
    static void flatten(const C& c, double* out) { for (size_t i = 0; i < c.size(); i++) out[i] = c[i].GetValue(); }
    static C inflate(const double* in, const C& like)
    {
        C c;
        c.reserve(like.size());
        for (size_t i = 0; i < like.size(); i++) c.push_back(like[i].CopyWithNewValue(in[i])); /* limits + type ride along */
        return c;
    }
};
#endif

/* ---- robot / environment descriptions ------------------------------------------------------------------------ */

/* Owns the storage behind a upc_robot_desc. */
struct RobotDescription
{
    upc_robot_desc desc{};
    std::vector<double> points; /* x, y, z, radius per body point, grouped by link */

    void finalize() { desc.points = points.data(); desc.num_points = (int32_t)(points.size() / 4); }
};

/* Owns the storage behind a upc_env_desc. */
struct EnvironmentDescription
{
    upc_env_desc desc{};
    std::vector<float> sdf, occupancy;
    std::vector<uint32_t> convex_segment;

    void finalize()
    {
        desc.sdf = sdf.data();
        desc.convex_segment = convex_segment.empty() ? nullptr : convex_segment.data();
        desc.occupancy = occupancy.empty() ? nullptr : occupancy.data();
    }
};

/* Solver settings the concrete simulator owns; defaults follow the upstream example simulator. */
struct SolverConfig
{
    double controller_frequency = 100.0;
    int32_t max_microsteps = 4;
    double contact_distance = 0.0;
    double resolve_margin_ratio = 0.1; /* resolve_distance = contact_distance + ratio * resolution */
    double microstep_ratio = 0.5;      /* microstep_distance = ratio * resolution */
    int32_t max_resolver_iterations = 25;
    int32_t resolver_decay_iterations = 5;
    double resolver_initial_scale = 1.0;
    double resolver_decay_rate = 0.5;
    double resolver_damping = 1.0e-6;
};

class Handle
{
public:
    Handle(const EnvironmentDescription& env, const RobotDescription& robot, int device = 0) : h_(nullptr)
    {
        check(upc_create(&env.desc, device, &h_));
        const int32_t rc = upc_set_robot(h_, &robot.desc);
        if (rc != UPC_OK)
        {
            upc_destroy(h_);
            check(rc);
        }
    }
    ~Handle() { upc_destroy(h_); }
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;
    upc_handle* get() const { return h_; }

private:
    upc_handle* h_;
};

/* Grow-only page-locked host staging (upc_host_alloc): transfers from pinned memory run at full PCIe speed and need no driver-side
 * staging copy; the buffers are reused from call to call, so nothing is allocated or zero-filled per call. */
class PinnedBuffer
{
public:
    PinnedBuffer() : ptr_(nullptr), bytes_(0) {}
    ~PinnedBuffer() { if (ptr_) upc_host_free(ptr_); }
    PinnedBuffer(const PinnedBuffer&) = delete;
    PinnedBuffer& operator=(const PinnedBuffer&) = delete;
    template <typename T>
    T* get(const size_t count)
    {
        const size_t want = count * sizeof(T);
        if (want > bytes_)
        {
            if (ptr_) upc_host_free(ptr_);
            ptr_ = nullptr;
            bytes_ = 0;
            size_t cap = 4096;
            while (cap < want) cap *= 2;
            check(upc_host_alloc(&ptr_, (int64_t)cap));
            bytes_ = cap;
        }
        return static_cast<T*>(ptr_);
    }

private:
    void* ptr_;
    size_t bytes_;
};

/* staging of one simulator: inputs, outputs, flags, labels, cluster counts; guarded by a mutex because the interface methods are
 * const and may be called from several threads (threading) */
struct Staging
{
    std::mutex mutex;
    PinnedBuffer starts, targets, out, collided, labels, counts;
};

/* ---- SimulatorInterface ---------------------------------------------------------------------------------------------- */

template <typename Robot, typename Configuration, typename RNG, typename ConfigAlloc = std::allocator<Configuration>>
class B200Simulator
#ifdef UPC_SHIM_WITH_REFERENCE
    : public simple_simulator_interface::SimulatorInterface<Robot, Configuration, RNG, ConfigAlloc>
#endif
{
public:
    typedef ConfigTraits<Configuration> Traits;

#ifdef UPC_SHIM_WITH_REFERENCE
    /* ctor forwarding of simple_simulator_interface.hpp:398-405, plus the descriptions the kernels need */
    B200Simulator(const sdf_tools::TaggedObjectCollisionMapGrid& environment, const sdf_tools::SignedDistanceField& environment_sdf,
                  const simple_simulator_interface::SurfaceNormalGrid& surface_normals_grid, const int32_t debug_level,
                  const EnvironmentDescription& env_desc, const RobotDescription& robot_desc, const SolverConfig& solver = SolverConfig(), int device = 0)
        : simple_simulator_interface::SimulatorInterface<Robot, Configuration, RNG, ConfigAlloc>(environment, environment_sdf, surface_normals_grid, debug_level),
          env_(env_desc), robot_(robot_desc), solver_(solver), handle_(new Handle(env_desc, robot_desc, device)) { env_.finalize(); robot_.finalize(); }
#define UPC_SHIM_OVERRIDE override
#define UPC_SHIM_PUBLISHER , ros::Publisher& display_debug_publisher
#define UPC_SHIM_UNUSED_PUBLISHER (void)display_debug_publisher;
#else
    B200Simulator(const EnvironmentDescription& env_desc, const RobotDescription& robot_desc, const int32_t debug_level = 0,
                  const SolverConfig& solver = SolverConfig(), int device = 0)
        : env_(env_desc), robot_(robot_desc), solver_(solver), handle_(new Handle(env_desc, robot_desc, device)), debug_level_(debug_level)
    {
        env_.finalize(); /* the copies point into their own storage, not the caller's */
        robot_.finalize();
    }
#define UPC_SHIM_OVERRIDE
#define UPC_SHIM_PUBLISHER
#define UPC_SHIM_UNUSED_PUBLISHER
#endif
    virtual ~B200Simulator() {}

    upc_handle* handle() const { return handle_->get(); }
    const SolverConfig& solver() const { return solver_; }

    /* simple_simulator_interface.hpp:613 */
    std::map<std::string, double> GetStatistics() const UPC_SHIM_OVERRIDE
    {
        upc_stats s;
        check(upc_get_stats(handle_->get(), &s));
        std::map<std::string, double> m;
        m["kernel_launches"] = (double)s.kernel_launches;
        m["particles_simulated"] = (double)s.particles_simulated;
        m["particle_steps"] = (double)s.particle_steps;
        m["collided_particles"] = (double)s.collided_particles;
        m["resolver_iterations"] = (double)s.resolver_iterations;
        m["unsuccessful_resolves"] = (double)s.failed_resolves;
        m["cluster_calls"] = (double)s.cluster_calls;
        m["cluster_fallback_calls"] = (double)s.cluster_fallback_calls;
        m["h2d_bytes"] = (double)s.h2d_bytes;
        m["d2h_bytes"] = (double)s.d2h_bytes;
        return m;
    }

    /* simple_simulator_interface.hpp:615 */
    void ResetStatistics() UPC_SHIM_OVERRIDE { check(upc_reset_stats(handle_->get())); }

    /* simple_simulator_interface.hpp:617 */
    bool CheckConfigCollision(const Robot& immutable_robot, const Configuration& config, const double inflation_ratio = 0.0) const UPC_SHIM_OVERRIDE
    {
        (void)immutable_robot;
        std::vector<double> flat((size_t)Traits::width(config));
        Traits::flatten(config, flat.data());
        int32_t hit = 0;
        check(upc_check_config_collision(handle_->get(), flat.data(), solver_.contact_distance, inflation_ratio * env_.desc.resolution, &hit));
        return hit != 0;
    }

    /* simple_simulator_interface.hpp:623 - the batched hot path */
    std::vector<std::pair<Configuration, bool>> ForwardSimulateRobots(const Robot& immutable_robot, const std::vector<Configuration, ConfigAlloc>& start_positions,
                                                                      const std::vector<Configuration, ConfigAlloc>& target_positions, std::vector<RNG>& rng,
                                                                      const double forward_simulation_time, const double simulation_shortcut_distance,
                                                                      const bool use_individual_jacobians, const bool allow_contacts UPC_SHIM_PUBLISHER) const UPC_SHIM_OVERRIDE
    {
        UPC_SHIM_UNUSED_PUBLISHER
        (void)immutable_robot; /* borrowed, never mutated: the device works on its own copy of the description */
        const size_t n = start_positions.size();
        if (!(target_positions.size() == 1 || target_positions.size() == n))
            throw std::invalid_argument("target_positions must have size 1 or size start_positions.size()");
        if (rng.empty()) throw std::invalid_argument("at least one PRNG is required");
        std::vector<std::pair<Configuration, bool>> results;
        if (n == 0) return results;
        const int width = Traits::width(start_positions[0]);
        const upc_sim_params params = MakeParams(forward_simulation_time, simulation_shortcut_distance, use_individual_jacobians, allow_contacts);
        /* structure-of-arrays staging */
        const size_t nt = target_positions.size();
        std::lock_guard<std::mutex> guard(staging_->mutex);
        double* starts = staging_->starts.template get<double>((size_t)width * n);
        double* targets = staging_->targets.template get<double>((size_t)width * nt);
        double* out = staging_->out.template get<double>((size_t)width * n);
        uint8_t* collided = staging_->collided.template get<uint8_t>(n);
        FlattenToSoa(start_positions, width, starts);
        FlattenToSoa(target_positions, width, targets);
        /* Noise: the caller's generator is advanced by ONE draw per call (a 64-bit seed); particle i then draws from the
         * counter-based stream (seed, i) inside the kernel - no tape is generated or shipped, and results are
         * independent of OpenMP thread counts, which the reference's per-thread generators
         * (uncertainty_contact_planning.hpp:349-356) are not. */
        const uint64_t seed = NextSeed(rng[0]);
        check(upc_forward_simulate_seeded(handle_->get(), &params, seed, 0, (int64_t)n, (int64_t)n, starts, (int64_t)nt, targets, out, collided));
        results.resize(n, std::make_pair(start_positions[0], false));
        const double* outp = out;
#pragma omp parallel for schedule(static) if (n > 4096)
        for (int64_t i = 0; i < (int64_t)n; i++)
        {
            double tmp[16];
            for (int c = 0; c < width; c++) tmp[c] = outp[(size_t)c * n + (size_t)i];
            results[(size_t)i] = std::make_pair(Traits::inflate(tmp, start_positions[(size_t)i]), collided[(size_t)i] != 0);
        }
        return results;
    }

    /* simple_simulator_interface.hpp:621 without a trace (the form the batched callers use) */
    std::pair<Configuration, bool> ForwardSimulateRobot(const Robot& immutable_robot, const Configuration& start_position, const Configuration& target_position,
                                                        RNG& rng, const double forward_simulation_time, const double simulation_shortcut_distance,
                                                        const bool use_individual_jacobians, const bool allow_contacts UPC_SHIM_PUBLISHER) const
    {
        UPC_SHIM_UNUSED_PUBLISHER
        ForwardSimulationStepTrace<Configuration, ConfigAlloc> unused;
        return SimulateOne(immutable_robot, start_position, target_position, rng, forward_simulation_time, simulation_shortcut_distance,
                           use_individual_jacobians, allow_contacts, unused, false);
    }

    /* simple_simulator_interface.hpp:621 - one particle with an optional trace: one ForwardSimulationResolverTrace per controller
     * interval executed, carrying the control input, the control step per microstep and the resolved configuration of the interval
     * (the entry ExtractTrajectoryFromTrace :66-86 reads; callers uncertainty_contact_planning.hpp:510, :1171) */
    std::pair<Configuration, bool> ForwardSimulateRobot(const Robot& immutable_robot, const Configuration& start_position, const Configuration& target_position,
                                                        RNG& rng, const double forward_simulation_time, const double simulation_shortcut_distance,
                                                        const bool use_individual_jacobians, const bool allow_contacts,
                                                        ForwardSimulationStepTrace<Configuration, ConfigAlloc>& trace,
                                                        const bool enable_tracing UPC_SHIM_PUBLISHER) const UPC_SHIM_OVERRIDE
    {
        UPC_SHIM_UNUSED_PUBLISHER
        return SimulateOne(immutable_robot, start_position, target_position, rng, forward_simulation_time, simulation_shortcut_distance,
                           use_individual_jacobians, allow_contacts, trace, enable_tracing);
    }

#ifdef UPC_SHIM_WITH_REFERENCE
    /* simple_simulator_interface.hpp:619 */
    std::pair<Configuration, bool> ForwardSimulateMutableRobot(Robot& mutable_robot, const Configuration& target_position, RNG& rng,
                                                               const double forward_simulation_time, const double simulation_shortcut_distance,
                                                               const bool use_individual_jacobians, const bool allow_contacts,
                                                               ForwardSimulationStepTrace<Configuration, ConfigAlloc>& trace,
                                                               const bool enable_tracing, ros::Publisher& display_debug_publisher) const override
    {
        (void)display_debug_publisher;
        const std::pair<Configuration, bool> r = SimulateOne(mutable_robot, mutable_robot.GetPosition(), target_position, rng, forward_simulation_time,
                                                             simulation_shortcut_distance, use_individual_jacobians, allow_contacts, trace, enable_tracing);
        mutable_robot.UpdatePosition(r.first);
        return r;
    }
#endif

    /*
     * The planner-batch step (BASELINE config 5; what a speculative multi-edge extension of
     * uncertainty_contact_planning.hpp:2301-2409 would issue): for every candidate edge e, particles_per_edge particles from
     * edge_starts[e] towards edge_targets[e] (ForwardSimulateParticles :2039-2071), then the edge's outcome clustered
     * (ClusterParticles :1986-2034).  ONE simulation launch and ONE clustering pass on the device; result[e] = the clusters of
     * edge e in the shape ClusterParticles returns them.
     */
    template <typename Clustering>
    std::vector<std::vector<std::vector<std::pair<Configuration, bool>>>> SimulateAndClusterEdges(
        const Robot& immutable_robot, const std::vector<Configuration, ConfigAlloc>& edge_starts, const std::vector<Configuration, ConfigAlloc>& edge_targets,
        const size_t particles_per_edge, std::vector<RNG>& rng, const double forward_simulation_time, const double simulation_shortcut_distance,
        const bool use_individual_jacobians, const bool allow_contacts, const Clustering& clustering) const
    {
        (void)immutable_robot;
        const size_t E = edge_starts.size(), P = particles_per_edge, n = E * P;
        if (edge_targets.size() != E) throw std::invalid_argument("one target per edge");
        if (rng.empty()) throw std::invalid_argument("at least one PRNG is required");
        std::vector<std::vector<std::vector<std::pair<Configuration, bool>>>> result(E);
        if (n == 0) return result;
        const int width = Traits::width(edge_starts[0]);
        const upc_sim_params params = MakeParams(forward_simulation_time, simulation_shortcut_distance, use_individual_jacobians, allow_contacts);
        std::lock_guard<std::mutex> guard(staging_->mutex);
        double* starts = staging_->starts.template get<double>((size_t)width * E);
        double* targets = staging_->targets.template get<double>((size_t)width * E);
        double* out = staging_->out.template get<double>((size_t)width * n);
        uint8_t* collided = staging_->collided.template get<uint8_t>(n);
        int32_t* labels = staging_->labels.template get<int32_t>(n);
        int32_t* n_clusters = staging_->counts.template get<int32_t>(E);
        FlattenToSoa(edge_starts, width, starts);
        FlattenToSoa(edge_targets, width, targets);
        const uint64_t seed = NextSeed(rng[0]);
        check(upc_simulate_and_cluster_edges(handle_->get(), &params, &clustering.params(), seed, 0, (int64_t)E, (int64_t)P, (int64_t)E, starts, targets, out, collided,
                                             labels, n_clusters));
        AssembleEdgeClusters(edge_starts, P, n, width, allow_contacts, out, collided, labels, n_clusters, n, 0, E, result);
        return result;
    }

    /*
     * The same step as a pair of calls for planners that keep two batches in flight (upc_edge_batch_submit / _collect): submit batch
     * k + 1, then collect batch k - the next batch is simulated while this one is clustered, downloaded and assembled into clusters
     * on the host.  `slot` is 0 or 1; the edge vectors given to SubmitEdgeBatch must stay alive until the slot has been collected
     * (their start configurations are read again when the clusters are assembled).  CollectEdgeBatch returns what
     * SimulateAndClusterEdges would have returned for the same arguments.
     */
    void SubmitEdgeBatch(const Robot& immutable_robot, const std::vector<Configuration, ConfigAlloc>& edge_starts, const std::vector<Configuration, ConfigAlloc>& edge_targets,
                         const size_t particles_per_edge, std::vector<RNG>& rng, const double forward_simulation_time, const double simulation_shortcut_distance,
                         const bool use_individual_jacobians, const bool allow_contacts, const int slot) const
    {
        (void)immutable_robot;
        if (slot < 0 || slot > 1) throw std::invalid_argument("slot must be 0 or 1");
        const size_t E = edge_starts.size();
        if (E == 0 || particles_per_edge == 0) throw std::invalid_argument("an edge batch needs at least one edge and one particle");
        if (edge_targets.size() != E) throw std::invalid_argument("one target per edge");
        if (rng.empty()) throw std::invalid_argument("at least one PRNG is required");
        const int width = Traits::width(edge_starts[0]);
        const upc_sim_params params = MakeParams(forward_simulation_time, simulation_shortcut_distance, use_individual_jacobians, allow_contacts);
        std::lock_guard<std::mutex> guard(staging_->mutex);
        PendingBatch& pb = pending_[slot];
        double* starts = pb.starts.template get<double>((size_t)width * E);
        double* targets = pb.targets.template get<double>((size_t)width * E);
        FlattenToSoa(edge_starts, width, starts);
        FlattenToSoa(edge_targets, width, targets);
        const uint64_t seed = NextSeed(rng[0]);
        check(upc_edge_batch_submit(handle_->get(), &params, seed, 0, (int64_t)E, (int64_t)particles_per_edge, (int64_t)E, starts, targets, slot));
        pb.edge_starts = &edge_starts;
        pb.particles = particles_per_edge;
        pb.allow_contacts = allow_contacts;
    }

    template <typename Clustering>
    std::vector<std::vector<std::vector<std::pair<Configuration, bool>>>> CollectEdgeBatch(const Clustering& clustering, const int slot) const
    {
        if (slot < 0 || slot > 1) throw std::invalid_argument("slot must be 0 or 1");
        std::lock_guard<std::mutex> guard(staging_->mutex);
        PendingBatch& pb = pending_[slot];
        if (pb.edge_starts == nullptr) throw std::invalid_argument("nothing was submitted to this slot");
        const std::vector<Configuration, ConfigAlloc>& edge_starts = *pb.edge_starts;
        const size_t E = edge_starts.size(), P = pb.particles, n = E * P;
        const int width = Traits::width(edge_starts[0]);
        double* out = pb.out.template get<double>((size_t)width * n);
        uint8_t* collided = pb.collided.template get<uint8_t>(n);
        int32_t* labels = pb.labels.template get<int32_t>(n);
        int32_t* n_clusters = pb.counts.template get<int32_t>(E);
        pb.edge_starts = nullptr;
        check(upc_edge_batch_collect(handle_->get(), &clustering.params(), slot, out, collided, labels, n_clusters));
        std::vector<std::vector<std::vector<std::pair<Configuration, bool>>>> result(E);
        AssembleEdgeClusters(edge_starts, P, n, width, pb.allow_contacts, out, collided, labels, n_clusters, n, 0, E, result);
        return result;
    }

    upc_sim_params MakeParams(const double forward_simulation_time, const double simulation_shortcut_distance, const bool use_individual_jacobians,
                              const bool allow_contacts) const
    {
        upc_sim_params p{};
        const uint32_t steps = (uint32_t)(forward_simulation_time * solver_.controller_frequency);
        p.num_steps = (int32_t)(steps < 1u ? 1u : steps);
        p.max_microsteps = solver_.max_microsteps;
        p.allow_contacts = allow_contacts ? 1 : 0;
        p.use_individual_jacobians = use_individual_jacobians ? 1 : 0;
        p.max_resolver_iterations = solver_.max_resolver_iterations;
        p.resolver_decay_iterations = solver_.resolver_decay_iterations;
        p.controller_interval = 1.0 / solver_.controller_frequency;
        p.simulation_shortcut_distance = simulation_shortcut_distance;
        p.contact_distance = solver_.contact_distance;
        p.resolve_distance = solver_.contact_distance + solver_.resolve_margin_ratio * env_.desc.resolution;
        p.microstep_distance = solver_.microstep_ratio * env_.desc.resolution;
        p.resolver_initial_scale = solver_.resolver_initial_scale;
        p.resolver_decay_rate = solver_.resolver_decay_rate;
        p.resolver_damping = solver_.resolver_damping;
        return p;
    }

    const RobotDescription& robot_description() const { return robot_; }
    const EnvironmentDescription& environment_description() const { return env_; }

    /* one 64-bit draw from the caller's generator: the key of a call's noise streams */
    static uint64_t NextSeed(RNG& generator)
    {
        std::uniform_int_distribution<uint64_t> seed_dist(0, std::numeric_limits<uint64_t>::max());
        return seed_dist(generator);
    }

    /* vector of configurations -> [C][n] doubles */
    static void FlattenToSoa(const std::vector<Configuration, ConfigAlloc>& configs, const int width, double* soa)
    {
        const size_t n = configs.size();
#pragma omp parallel for schedule(static) if (n > 4096)
        for (int64_t i = 0; i < (int64_t)n; i++)
        {
            double tmp[16];
            Traits::flatten(configs[(size_t)i], tmp);
            for (int c = 0; c < width; c++) soa[(size_t)c * n + (size_t)i] = tmp[c];
        }
    }

    /* clusters of edges [first_edge, first_edge + count) from flat outcome arrays (configs [C][stride], edge-major particles starting
     * at column 0 for first_edge): sizes are counted first so that every cluster is allocated once */
    static void AssembleEdgeClusters(const std::vector<Configuration, ConfigAlloc>& edge_starts, const size_t P, const size_t n_columns, const int width,
                                     const bool allow_contacts, const double* configs, const uint8_t* collided, const int32_t* labels,
                                     const int32_t* n_clusters, const size_t stride, const size_t first_edge, const size_t count,
                                     std::vector<std::vector<std::vector<std::pair<Configuration, bool>>>>& result)
    {
        (void)n_columns;
#pragma omp parallel for schedule(static) if (count > 16)
        for (int64_t el = 0; el < (int64_t)count; el++)
        {
            std::vector<std::vector<std::pair<Configuration, bool>>>& clusters = result[first_edge + (size_t)el];
            const size_t K = (size_t)n_clusters[el];
            clusters.resize(K);
            std::vector<size_t> sizes(K, 0);
            for (size_t k = 0; k < P; k++)
            {
                const size_t i = (size_t)el * P + k;
                if (collided[i] == 0 || allow_contacts) sizes[(size_t)labels[i]]++; /* uncertainty_contact_planning.hpp:2019 */
            }
            for (size_t c = 0; c < K; c++) clusters[c].reserve(sizes[c]);
            double tmp[16];
            for (size_t k = 0; k < P; k++)
            {
                const size_t i = (size_t)el * P + k;
                const bool hit = collided[i] != 0;
                if (hit && !allow_contacts) continue;
                for (int c = 0; c < width; c++) tmp[c] = configs[(size_t)c * stride + i];
                clusters[(size_t)labels[i]].push_back(std::make_pair(Traits::inflate(tmp, edge_starts[first_edge + (size_t)el]), hit));
            }
        }
    }

protected:
    std::pair<Configuration, bool> SimulateOne(const Robot& immutable_robot, const Configuration& start_position, const Configuration& target_position, RNG& rng,
                                               const double forward_simulation_time, const double simulation_shortcut_distance,
                                               const bool use_individual_jacobians, const bool allow_contacts,
                                               ForwardSimulationStepTrace<Configuration, ConfigAlloc>& trace, const bool enable_tracing) const
    {
        (void)immutable_robot;
        const int width = Traits::width(start_position), dof = robot_.desc.dof;
        const upc_sim_params params = MakeParams(forward_simulation_time, simulation_shortcut_distance, use_individual_jacobians, allow_contacts);
        std::vector<double> start((size_t)width), target((size_t)width), out((size_t)width);
        Traits::flatten(start_position, start.data());
        Traits::flatten(target_position, target.data());
        const uint64_t seed = NextSeed(rng);
        uint8_t collided = 0;
        if (!enable_tracing)
        {
            check(upc_forward_simulate_seeded(handle_->get(), &params, seed, 0, 1, 1, start.data(), 1, target.data(), out.data(), &collided));
            return std::make_pair(Traits::inflate(out.data(), start_position), collided != 0);
        }
        const size_t S = (size_t)params.num_steps;
        std::vector<double> trace_configs(S * (size_t)width), trace_controls(S * 2 * (size_t)dof);
        int32_t trace_steps[2] = {0, 0};
        check(upc_forward_simulate_traced(handle_->get(), &params, 1, start.data(), 1, target.data(), nullptr, seed, 0, out.data(), &collided,
                                          trace_configs.data(), trace_controls.data(), trace_steps));
        for (int32_t step = 0; step < trace_steps[0]; step++)
        {
            ForwardSimulationResolverTrace<Configuration, ConfigAlloc> step_trace;
            step_trace.control_input = MakeControlVector(&trace_controls[(size_t)step * 2 * (size_t)dof], dof);
            step_trace.control_input_step = MakeControlVector(&trace_controls[(size_t)step * 2 * (size_t)dof + (size_t)dof], dof);
            ForwardSimulationContactResolverStepTrace<Configuration, ConfigAlloc> resolved;
            resolved.contact_resolution_steps.push_back(Traits::inflate(&trace_configs[(size_t)step * (size_t)width], start_position));
            step_trace.contact_resolver_steps.push_back(resolved);
            trace.resolver_steps.push_back(step_trace);
        }
        return std::make_pair(Traits::inflate(out.data(), start_position), collided != 0);
    }

    EnvironmentDescription env_;
    RobotDescription robot_;
    SolverConfig solver_;
    std::shared_ptr<Handle> handle_;
    std::shared_ptr<Staging> staging_ = std::shared_ptr<Staging>(new Staging());
    /* SubmitEdgeBatch / CollectEdgeBatch: page-locked staging and bookkeeping of the two batches in flight */
    struct PendingBatch
    {
        PinnedBuffer starts, targets, out, collided, labels, counts;
        const std::vector<Configuration, ConfigAlloc>* edge_starts = nullptr;
        size_t particles = 0;
        bool allow_contacts = true;
    };
    mutable std::shared_ptr<std::array<PendingBatch, 2>> pending_store_ = std::shared_ptr<std::array<PendingBatch, 2>>(new std::array<PendingBatch, 2>());
    std::array<PendingBatch, 2>& pending_ = *pending_store_;
#ifndef UPC_SHIM_WITH_REFERENCE
    int32_t debug_level_;
#endif
};

/* ---- outcome clustering --------------------------------------------------------------------------------------------- */

template <typename Robot, typename Configuration, typename ConfigAlloc = std::allocator<Configuration>>
class B200OutcomeClustering
{
public:
    typedef ConfigTraits<Configuration> Traits;

    /* clustering_type is the reference's enum (uncertainty_contact_planning.hpp:52, the value the planner holds in
     * spatial_feature_clustering_type_, :144); thresholds as at :150-151 */
    B200OutcomeClustering(upc_handle* handle, const SPATIAL_FEATURE_CLUSTERING_TYPE clustering_type, const double signature_matching_threshold,
                          const double distance_clustering_threshold)
        : handle_(handle)
    {
        params_ = upc_cluster_params{};
        params_.feature_type = FromReferenceClusteringType(clustering_type);
        params_.signature_matching_threshold = signature_matching_threshold;
        params_.distance_clustering_threshold = distance_clustering_threshold;
    }

    /* distance pass only (no spatial-feature first pass): not a mode of the reference, used for the clustering sweep */
    static B200OutcomeClustering DistanceOnly(upc_handle* handle, const double distance_clustering_threshold)
    {
        B200OutcomeClustering c(handle, CONVEX_REGION_SIGNATURE, 0.0, distance_clustering_threshold);
        c.params_.feature_type = UPC_FEATURE_NONE;
        return c;
    }

    /*
     * POINT_TO_POINT_MOVEMENT needs the simulator (uncertainty_contact_planning.hpp:1801-1863: n*(n-1) pair simulations of
     * step_duration_ with shortcut goal_distance_threshold_ / 2, contacts allowed, :1826): pair_params as the simulator's
     * MakeParams(step_duration, goal_distance_threshold * 0.5, use_individual_jacobians, true); every ClusterParticles call draws
     * the pair simulations' seeded noise from `seed`.
     */
    void EnablePointToPointMovement(const upc_sim_params& pair_params, const double goal_distance_threshold, const uint64_t seed)
    {
        ptpm_params_ = pair_params;
        ptpm_params_.allow_contacts = 1;
        params_.goal_distance_threshold = goal_distance_threshold;
        params_.ptpm_seed = seed;
        params_.ptpm_tape = nullptr;
        has_ptpm_ = true;
    }

    B200OutcomeClustering(const B200OutcomeClustering& o) : handle_(o.handle_), params_(o.params_), ptpm_params_(o.ptpm_params_), has_ptpm_(o.has_ptpm_), staging_(o.staging_) {}
    B200OutcomeClustering& operator=(const B200OutcomeClustering& o)
    {
        handle_ = o.handle_; params_ = o.params_; ptpm_params_ = o.ptpm_params_; has_ptpm_ = o.has_ptpm_; staging_ = o.staging_;
        return *this;
    }

    /* the POD parameter block handed to the C ABI (ptpm_sim points into this object) */
    const upc_cluster_params& params() const
    {
        params_.ptpm_sim = has_ptpm_ ? &ptpm_params_ : nullptr;
        return params_;
    }

    /* index clusters of one particle set, numbered by smallest member */
    std::vector<std::vector<size_t>> ClusterIndices(const std::vector<const Configuration*>& configs) const
    {
        const size_t n = configs.size();
        std::vector<std::vector<size_t>> clusters;
        if (n == 0) return clusters;
        const int width = Traits::width(*configs[0]);
        std::lock_guard<std::mutex> guard(staging_->mutex);
        double* soa = staging_->starts.template get<double>((size_t)width * n);
        int32_t* labels = staging_->labels.template get<int32_t>(n);
        int32_t* n_clusters = staging_->counts.template get<int32_t>(1);
#pragma omp parallel for schedule(static) if (n > 4096)
        for (int64_t i = 0; i < (int64_t)n; i++)
        {
            double tmp[16];
            Traits::flatten(*configs[(size_t)i], tmp);
            for (int c = 0; c < width; c++) soa[(size_t)c * n + (size_t)i] = tmp[c];
        }
        check(upc_cluster_particles(handle_, &params(), 1, (int64_t)n, soa, labels, n_clusters));
        clusters.resize((size_t)n_clusters[0]);
        std::vector<size_t> sizes((size_t)n_clusters[0], 0);
        for (size_t i = 0; i < n; i++) sizes[(size_t)labels[i]]++;
        for (size_t k = 0; k < clusters.size(); k++) clusters[k].reserve(sizes[k]);
        for (size_t i = 0; i < n; i++) clusters[(size_t)labels[i]].push_back(i);
        return clusters;
    }

    /* uncertainty_contact_planning.hpp:1986-2034 */
    std::vector<std::vector<std::pair<Configuration, bool>>> ClusterParticles(const std::vector<std::pair<Configuration, bool>>& particles, const bool allow_contacts) const
    {
        if (particles.size() == 0) return std::vector<std::vector<std::pair<Configuration, bool>>>();
        if (particles.size() == 1) return std::vector<std::vector<std::pair<Configuration, bool>>>{particles};
        std::vector<const Configuration*> ptrs;
        ptrs.reserve(particles.size());
        for (size_t i = 0; i < particles.size(); i++) ptrs.push_back(&particles[i].first);
        const std::vector<std::vector<size_t>> index_clusters = ClusterIndices(ptrs);
        std::vector<std::vector<std::pair<Configuration, bool>>> final_clusters;
        final_clusters.reserve(index_clusters.size());
        for (size_t k = 0; k < index_clusters.size(); k++)
        {
            std::vector<std::pair<Configuration, bool>> cluster;
            for (size_t e = 0; e < index_clusters[k].size(); e++)
            {
                const std::pair<Configuration, bool>& particle = particles[index_clusters[k][e]];
                if ((particle.second == false) || allow_contacts) cluster.push_back(particle); /* :2019 */
            }
            final_clusters.push_back(cluster); /* possibly empty, still emitted (:2025) */
        }
        return final_clusters;
    }

    /* uncertainty_contact_planning.hpp:1558-1578 */
    std::vector<std::vector<size_t>> PolicyParticleClusteringFn(const std::vector<Configuration, ConfigAlloc>& particles, const Configuration& current_config) const
    {
        if (particles.size() < 1) throw std::invalid_argument("PolicyParticleClusteringFn needs at least one particle");
        std::vector<const Configuration*> ptrs;
        for (size_t i = 0; i < particles.size(); i++) ptrs.push_back(&particles[i]);
        ptrs.push_back(&current_config);
        return ClusterIndices(ptrs);
    }

    /* Flat per-cluster statistics: what UncertaintyPlannerState::UpdateStatistics (uncertainty_planner_state.hpp:339-351) computes
     * for each child state the clusters become (uncertainty_contact_planning.hpp:2110-2160). */
    struct ClusterStatisticsResult
    {
        int32_t width, variance_width;
        std::vector<int64_t> counts;
        std::vector<uint8_t> any_collided;
        std::vector<double> expectations, variance, si_variance, variances, si_variances; /* [K][width] / [K] / [K][variance_width] */
    };

    ClusterStatisticsResult ClusterStatistics(const std::vector<std::pair<Configuration, bool>>& particles, const std::vector<std::vector<size_t>>& index_clusters,
                                              const bool allow_contacts, const double step_size, const int32_t robot_kind, const int32_t dof) const
    {
        ClusterStatisticsResult r;
        const size_t n = particles.size(), K = index_clusters.size();
        r.width = (n > 0) ? Traits::width(particles[0].first) : 0;
        r.variance_width = upc_variance_width(robot_kind, dof);
        r.counts.assign(K, 0);
        r.any_collided.assign(K, 0);
        r.expectations.assign(K * (size_t)r.width, 0.0);
        r.variance.assign(K, 0.0);
        r.si_variance.assign(K, 0.0);
        r.variances.assign(K * (size_t)r.variance_width, 0.0);
        r.si_variances.assign(K * (size_t)r.variance_width, 0.0);
        if (n == 0 || K == 0) return r;
        std::vector<double> soa((size_t)r.width * n), tmp((size_t)r.width);
        std::vector<int32_t> labels(n, -1);
        std::vector<uint8_t> collided(n);
        for (size_t i = 0; i < n; i++)
        {
            Traits::flatten(particles[i].first, tmp.data());
            for (int c = 0; c < r.width; c++) soa[(size_t)c * n + i] = tmp[(size_t)c];
            collided[i] = particles[i].second ? 1 : 0;
        }
        for (size_t k = 0; k < K; k++)
            for (size_t e = 0; e < index_clusters[k].size(); e++) labels[index_clusters[k][e]] = (int32_t)k;
        check(upc_cluster_statistics(handle_, (int64_t)n, soa.data(), labels.data(), collided.data(), allow_contacts ? 1 : 0, step_size, (int32_t)K,
                                     r.counts.data(), r.any_collided.data(), r.expectations.data(), r.variance.data(), r.si_variance.data(),
                                     r.variances.data(), r.si_variances.data()));
        return r;
    }

    /* ClusterParticles for many equally sized particle sets in ONE device pass (the per-edge outcomes of an edge batch):
     * result[s] = ClusterParticles(sets[s], allow_contacts) */
    std::vector<std::vector<std::vector<std::pair<Configuration, bool>>>> ClusterParticleSets(
        const std::vector<std::vector<std::pair<Configuration, bool>>>& sets, const bool allow_contacts) const
    {
        const size_t S = sets.size();
        std::vector<std::vector<std::vector<std::pair<Configuration, bool>>>> result(S);
        if (S == 0) return result;
        const size_t P = sets[0].size();
        for (size_t k = 0; k < S; k++)
            if (sets[k].size() != P) throw std::invalid_argument("ClusterParticleSets: all particle sets must have the same size");
        if (P == 0) return result;
        if (P == 1)
        {
            for (size_t k = 0; k < S; k++) result[k].push_back(sets[k]); /* :1988-1995: a single particle is its own cluster */
            return result;
        }
        const int width = Traits::width(sets[0][0].first);
        const size_t n = S * P;
        std::vector<double> soa((size_t)width * n);
#pragma omp parallel for schedule(static) if (S > 16)
        for (int64_t k = 0; k < (int64_t)S; k++)
        {
            double tmp[16];
            for (size_t e = 0; e < P; e++)
            {
                Traits::flatten(sets[(size_t)k][e].first, tmp);
                for (int c = 0; c < width; c++) soa[(size_t)c * n + (size_t)k * P + e] = tmp[c];
            }
        }
        std::vector<int32_t> labels(n), n_clusters(S);
        check(upc_cluster_particles(handle_, &params(), (int64_t)S, (int64_t)P, soa.data(), labels.data(), n_clusters.data()));
#pragma omp parallel for schedule(static) if (S > 16)
        for (int64_t k = 0; k < (int64_t)S; k++)
        {
            result[(size_t)k].resize((size_t)n_clusters[(size_t)k]);
            for (size_t e = 0; e < P; e++)
            {
                const std::pair<Configuration, bool>& particle = sets[(size_t)k][e];
                if ((particle.second == false) || allow_contacts) result[(size_t)k][(size_t)labels[(size_t)k * P + e]].push_back(particle);
            }
        }
        return result;
    }

private:
    upc_handle* handle_;
    mutable upc_cluster_params params_;
    upc_sim_params ptpm_params_{};
    bool has_ptpm_ = false;
    std::shared_ptr<Staging> staging_ = std::shared_ptr<Staging>(new Staging());
};

/* ---- one box, several GPUs ------------------------------------------------------------------------------------------ */

/*
 * The planner-batch step sharded over the GPUs of one box: every device holds its own copy of the
 * environment and the robot (a complete simulator), candidate edges are split into contiguous blocks, each device simulates
 * and clusters its block, and ONE NCCL all-gather leaves the outcome records of the whole batch on every device; the host
 * reads them from the first.  Noise is keyed by global particle index, so the result does not depend on the number of devices.
 */
template <typename Robot, typename Configuration, typename RNG, typename ConfigAlloc = std::allocator<Configuration>>
class ShardedSimulator
{
public:
    typedef B200Simulator<Robot, Configuration, RNG, ConfigAlloc> Simulator;
    typedef ConfigTraits<Configuration> Traits;

    ShardedSimulator(const EnvironmentDescription& env_desc, const RobotDescription& robot_desc, const std::vector<int>& devices,
                     const SolverConfig& solver = SolverConfig())
    {
        if (devices.empty()) throw std::invalid_argument("ShardedSimulator needs at least one device");
        for (size_t r = 0; r < devices.size(); r++) simulators_.push_back(std::shared_ptr<Simulator>(new Simulator(env_desc, robot_desc, 0, solver, devices[r])));
        std::vector<upc_handle*> handles;
        for (size_t r = 0; r < devices.size(); r++) handles.push_back(simulators_[r]->handle());
        comms_.assign(devices.size(), nullptr);
        check(upc_comm_create_all((int32_t)devices.size(), handles.data(), comms_.data()));
    }
    ~ShardedSimulator()
    {
        for (size_t r = 0; r < comms_.size(); r++) upc_comm_destroy(comms_[r]);
    }
    ShardedSimulator(const ShardedSimulator&) = delete;
    ShardedSimulator& operator=(const ShardedSimulator&) = delete;

    size_t world() const { return simulators_.size(); }
    const Simulator& simulator(size_t rank) const { return *simulators_[rank]; }

    /* same contract as B200Simulator::SimulateAndClusterEdges */
    template <typename Clustering>
    std::vector<std::vector<std::vector<std::pair<Configuration, bool>>>> SimulateAndClusterEdges(
        const Robot& immutable_robot, const std::vector<Configuration, ConfigAlloc>& edge_starts, const std::vector<Configuration, ConfigAlloc>& edge_targets,
        const size_t particles_per_edge, std::vector<RNG>& rng, const double forward_simulation_time, const double simulation_shortcut_distance,
        const bool use_individual_jacobians, const bool allow_contacts, const Clustering& clustering)
    {
        (void)immutable_robot;
        const size_t E = edge_starts.size(), P = particles_per_edge, W = world();
        if (edge_targets.size() != E) throw std::invalid_argument("one target per edge");
        if (rng.empty()) throw std::invalid_argument("at least one PRNG is required");
        std::vector<std::vector<std::vector<std::pair<Configuration, bool>>>> result(E);
        if (E * P == 0) return result;
        const int width = Traits::width(edge_starts[0]);
        const upc_sim_params params = simulators_[0]->MakeParams(forward_simulation_time, simulation_shortcut_distance, use_individual_jacobians, allow_contacts);
        std::vector<double> starts((size_t)width * E), targets((size_t)width * E);
        Simulator::FlattenToSoa(edge_starts, width, starts.data());
        Simulator::FlattenToSoa(edge_targets, width, targets.data());
        const uint64_t seed = Simulator::NextSeed(rng[0]);
        const size_t cap_edges = (E + W - 1) / W;
        upc_outcome_layout lay;
        check(upc_outcome_layout_for(width, (int64_t)(cap_edges * P), (int64_t)cap_edges, &lay));
        std::vector<const double*> d_starts(W), d_targets(W);
        std::vector<void*> d_blocks(W);
        for (size_t r = 0; r < W; r++)
        {
            double* ds = nullptr;
            double* dt = nullptr;
            void* db = nullptr;
            check(upc_device_scratch(simulators_[r]->handle(), 0, (int64_t)(starts.size() * sizeof(double)), (void**)&ds));
            check(upc_device_scratch(simulators_[r]->handle(), 1, (int64_t)(targets.size() * sizeof(double)), (void**)&dt));
            check(upc_device_scratch(simulators_[r]->handle(), 2, (int64_t)((size_t)lay.block_bytes * W), &db));
            check(upc_memcpy_to_device(simulators_[r]->handle(), ds, starts.data(), (int64_t)(starts.size() * sizeof(double))));
            check(upc_memcpy_to_device(simulators_[r]->handle(), dt, targets.data(), (int64_t)(targets.size() * sizeof(double))));
            d_starts[r] = ds; d_targets[r] = dt; d_blocks[r] = db;
        }
        check(upc_sharded_edge_batch_all((int32_t)W, comms_.data(), &params, &clustering.params(), seed, (int64_t)E, (int64_t)P, d_starts.data(),
                                         d_targets.data(), d_blocks.data(), nullptr));
        for (size_t r = 0; r < W; r++) check(upc_comm_synchronize(comms_[r]));
        /* every device now holds every rank's block: read them all from the first */
        uint8_t* host = host_blocks_.template get<uint8_t>((size_t)lay.block_bytes * W);
        check(upc_memcpy_to_host(simulators_[0]->handle(), host, d_blocks[0], (int64_t)((size_t)lay.block_bytes * W)));
        for (size_t r = 0; r < W; r++)
        {
            int64_t lo = 0, hi = 0;
            upc_shard_bounds((int64_t)E, (int32_t)r, (int32_t)W, &lo, &hi);
            if (hi == lo) continue;
            const uint8_t* block = host + r * (size_t)lay.block_bytes;
            const size_t n_local = (size_t)(hi - lo) * P;
            Simulator::AssembleEdgeClusters(edge_starts, P, n_local, width, allow_contacts, (const double*)(block + lay.configs_offset), block + lay.collided_offset,
                                            (const int32_t*)(block + lay.labels_offset), (const int32_t*)(block + lay.n_clusters_offset), n_local, (size_t)lo,
                                            (size_t)(hi - lo), result);
        }
        return result;
    }

private:
    std::vector<std::shared_ptr<Simulator>> simulators_;
    std::vector<upc_comm*> comms_;
    PinnedBuffer host_blocks_;
};

/*
 * GetNearestNeighbor (uncertainty_contact_planning.hpp:1506-1542) for a batch of sampled configurations.  The caller
 * keeps, per tree state, the expectation and the two StateDistance factors that depend on the state only (:1492-1499).
 * Returns, per sample, the index of the nearest enabled state (-1 if none).
 */
template <typename Configuration>
inline std::vector<int64_t> NearestNeighbors(upc_handle* handle, const std::vector<const Configuration*>& expectations,
                                             const std::vector<double>& feasibility_weights, const std::vector<double>& variance_weights,
                                             const std::vector<uint8_t>& use_for_nearest_neighbors, const std::vector<Configuration>& samples,
                                             const double step_size, std::vector<double>* distances = nullptr)
{
    typedef ConfigTraits<Configuration> Traits;
    const size_t ns = expectations.size(), nq = samples.size();
    std::vector<int64_t> nearest(nq, -1);
    if (distances) distances->assign(nq, std::numeric_limits<double>::infinity());
    if (nq == 0) return nearest;
    if (feasibility_weights.size() != ns || variance_weights.size() != ns || (!use_for_nearest_neighbors.empty() && use_for_nearest_neighbors.size() != ns))
        throw std::invalid_argument("NearestNeighbors: one weight pair (and flag) per state");
    const int width = Traits::width(samples[0]);
    std::vector<double> esoa((size_t)width * ns), qsoa((size_t)width * nq), tmp((size_t)width);
    for (size_t i = 0; i < ns; i++)
    {
        Traits::flatten(*expectations[i], tmp.data());
        for (int c = 0; c < width; c++) esoa[(size_t)c * ns + i] = tmp[(size_t)c];
    }
    for (size_t q = 0; q < nq; q++)
    {
        Traits::flatten(samples[q], tmp.data());
        for (int c = 0; c < width; c++) qsoa[(size_t)c * nq + q] = tmp[(size_t)c];
    }
    check(upc_nearest_neighbors(handle, (int64_t)ns, esoa.data(), feasibility_weights.data(), variance_weights.data(),
                                use_for_nearest_neighbors.empty() ? nullptr : use_for_nearest_neighbors.data(), (int64_t)nq, qsoa.data(), step_size,
                                nearest.data(), distances ? distances->data() : nullptr));
    return nearest;
}

/*
 * Speculative multi-edge extension: the RRT-Extend branch of PerformForwardPropagation (uncertainty_contact_planning.hpp:2320-2340)
 * for a whole batch of sampled configurations in ONE nearest-neighbour launch, ONE simulation launch and ONE clustering pass - the
 * shape that makes planner batches (BASELINE config 5) exist inside a planner.  Per sample q, exactly what the reference does for
 * one: nearest[q] = GetNearestNeighbor(sample) (:1506-1542); target = the sample itself when it lies within step_size of the nearest
 * state's expectation, else InterpolateBetweenConfigurations(expectation, sample, step_size / distance) (:2322-2334, evaluated by
 * the caller's own robot object, so the arithmetic is the reference's); then ForwardSimulateParticles (:2039-2071) from the
 * expectation towards the target and ClusterParticles (:1986-2034) of the outcome.  Which of the speculative edges the planner
 * commits - and in which order their child states enter the tree - stays the planner's decision.  Samples without an enabled
 * state (nearest = -1) get no edge: their entry of `clusters` is empty.
 */
template <typename Configuration, typename ConfigAlloc = std::allocator<Configuration>>
struct ExtensionBatch
{
    std::vector<int64_t> nearest;                                                     /* per sample: index of the state it extends from */
    std::vector<Configuration, ConfigAlloc> targets;                                  /* per sample: the configuration simulated towards */
    std::vector<std::vector<std::vector<std::pair<Configuration, bool>>>> clusters;   /* per sample: what ClusterParticles returns */
};

template <typename Simulator, typename Robot, typename Configuration, typename RNG, typename Clustering, typename ConfigAlloc = std::allocator<Configuration>>
inline ExtensionBatch<Configuration, ConfigAlloc> ExtendTowardsSamples(
    const Simulator& simulator, upc_handle* handle, const Robot& robot, const std::vector<const Configuration*>& expectations,
    const std::vector<double>& feasibility_weights, const std::vector<double>& variance_weights, const std::vector<uint8_t>& use_for_nearest_neighbors,
    const std::vector<Configuration>& samples, const double step_size, const size_t particles_per_edge, std::vector<RNG>& rng,
    const double step_duration, const double simulation_shortcut_distance, const bool use_individual_jacobians, const bool allow_contacts,
    const Clustering& clustering)
{
    ExtensionBatch<Configuration, ConfigAlloc> batch;
    batch.nearest = NearestNeighbors<Configuration>(handle, expectations, feasibility_weights, variance_weights, use_for_nearest_neighbors, samples, step_size);
    batch.targets.assign(samples.begin(), samples.end());
    batch.clusters.resize(samples.size());
    std::vector<Configuration, ConfigAlloc> edge_starts, edge_targets;
    std::vector<size_t> edge_sample;
    for (size_t q = 0; q < samples.size(); q++)
    {
        if (batch.nearest[q] < 0) continue;
        const Configuration& from = *expectations[(size_t)batch.nearest[q]];
        const double target_distance = robot.ComputeConfigurationDistance(from, samples[q]);
        if (target_distance > step_size) batch.targets[q] = robot.InterpolateBetweenConfigurations(from, samples[q], step_size / target_distance);
        edge_starts.push_back(from);
        edge_targets.push_back(batch.targets[q]);
        edge_sample.push_back(q);
    }
    if (edge_starts.empty()) return batch;
    auto edge_clusters = simulator.SimulateAndClusterEdges(robot, edge_starts, edge_targets, particles_per_edge, rng, step_duration, simulation_shortcut_distance,
                                                           use_individual_jacobians, allow_contacts, clustering);
    for (size_t e = 0; e < edge_sample.size(); e++) batch.clusters[edge_sample[e]] = std::move(edge_clusters[e]);
    return batch;
}

} /* namespace upc_b200 */

#endif /* UPC_B200_SHIM_HPP */

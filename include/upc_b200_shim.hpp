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
#include <cstdint>
#include <map>
#include <memory>
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
          env_(env_desc), robot_(robot_desc), solver_(solver), handle_(new Handle(env_desc, robot_desc, device)) {}
#define UPC_SHIM_OVERRIDE override
#define UPC_SHIM_PUBLISHER , ros::Publisher& display_debug_publisher
#define UPC_SHIM_UNUSED_PUBLISHER (void)display_debug_publisher;
#else
    B200Simulator(const EnvironmentDescription& env_desc, const RobotDescription& robot_desc, const int32_t debug_level = 0,
                  const SolverConfig& solver = SolverConfig(), int device = 0)
        : env_(env_desc), robot_(robot_desc), solver_(solver), handle_(new Handle(env_desc, robot_desc, device)), debug_level_(debug_level) {}
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
        std::vector<double> starts((size_t)width * n), targets((size_t)width * target_positions.size()), out((size_t)width * n);
        std::vector<double> tmp((size_t)width);
        for (size_t i = 0; i < n; i++)
        {
            Traits::flatten(start_positions[i], tmp.data());
            for (int c = 0; c < width; c++) starts[(size_t)c * n + i] = tmp[(size_t)c];
        }
        for (size_t i = 0; i < target_positions.size(); i++)
        {
            Traits::flatten(target_positions[i], tmp.data());
            for (int c = 0; c < width; c++) targets[(size_t)c * target_positions.size() + i] = tmp[(size_t)c];
        }
        /* Noise: the caller's generator is advanced by ONE draw per call (a 64-bit tape seed), and particle i then
         * draws from its own stream keyed by (seed, i) - results are independent of OpenMP thread counts, which the
         * reference's per-thread generators (uncertainty_contact_planning.hpp:349-356) are not. */
        std::uniform_int_distribution<uint64_t> seed_dist(0, std::numeric_limits<uint64_t>::max());
        const uint64_t seed = seed_dist(rng[0]);
        std::vector<double> tape((size_t)upc_tape_doubles_per_particle(&params, robot_.desc.dof) * n);
        check(upc_fill_noise_tape(seed, &robot_.desc, &params, 0, (int64_t)n, (int64_t)n, tape.data()));
        std::vector<uint8_t> collided(n);
        check(upc_forward_simulate(handle_->get(), &params, (int64_t)n, starts.data(), (int64_t)target_positions.size(), targets.data(), tape.data(),
                                   out.data(), collided.data()));
        results.reserve(n);
        for (size_t i = 0; i < n; i++)
        {
            for (int c = 0; c < width; c++) tmp[(size_t)c] = out[(size_t)c * n + i];
            results.push_back(std::make_pair(Traits::inflate(tmp.data(), start_positions[i]), collided[i] != 0));
        }
        return results;
    }

    /* simple_simulator_interface.hpp:621 (tracing is a host-side feature of the single-particle path; the batched
     * kernels record no trace, matching the reference which passes none at uncertainty_contact_planning.hpp:2065) */
    std::pair<Configuration, bool> ForwardSimulateRobot(const Robot& immutable_robot, const Configuration& start_position, const Configuration& target_position,
                                                        RNG& rng, const double forward_simulation_time, const double simulation_shortcut_distance,
                                                        const bool use_individual_jacobians, const bool allow_contacts UPC_SHIM_PUBLISHER) const
    {
        UPC_SHIM_UNUSED_PUBLISHER
        std::vector<Configuration, ConfigAlloc> s(1, start_position), t(1, target_position);
        std::vector<RNG> rngs(1, rng);
        std::vector<std::pair<Configuration, bool>> r = ForwardSimulateRobots(immutable_robot, s, t, rngs, forward_simulation_time, simulation_shortcut_distance,
                                                                              use_individual_jacobians, allow_contacts
#ifdef UPC_SHIM_WITH_REFERENCE
                                                                              , display_debug_publisher
#endif
        );
        rng = rngs[0];
        return r[0];
    }

#ifdef UPC_SHIM_WITH_REFERENCE
    std::pair<Configuration, bool> ForwardSimulateRobot(const Robot& immutable_robot, const Configuration& start_position, const Configuration& target_position,
                                                        RNG& rng, const double forward_simulation_time, const double simulation_shortcut_distance,
                                                        const bool use_individual_jacobians, const bool allow_contacts,
                                                        simple_simulator_interface::ForwardSimulationStepTrace<Configuration, ConfigAlloc>& trace,
                                                        const bool enable_tracing, ros::Publisher& display_debug_publisher) const override
    {
        const std::pair<Configuration, bool> r = ForwardSimulateRobot(immutable_robot, start_position, target_position, rng, forward_simulation_time,
                                                                      simulation_shortcut_distance, use_individual_jacobians, allow_contacts, display_debug_publisher);
        if (enable_tracing)
        {
            /* the trace contract (simple_simulator_interface.hpp:66-86) only needs the resolved configuration per step; the
             * device path reports the end of the edge */
            simple_simulator_interface::ForwardSimulationResolverTrace<Configuration, ConfigAlloc> step;
            simple_simulator_interface::ForwardSimulationContactResolverStepTrace<Configuration, ConfigAlloc> contact;
            contact.contact_resolution_steps.push_back(r.first);
            step.contact_resolver_steps.push_back(contact);
            trace.resolver_steps.push_back(step);
        }
        return r;
    }
    std::pair<Configuration, bool> ForwardSimulateMutableRobot(Robot& mutable_robot, const Configuration& target_position, RNG& rng,
                                                               const double forward_simulation_time, const double simulation_shortcut_distance,
                                                               const bool use_individual_jacobians, const bool allow_contacts,
                                                               simple_simulator_interface::ForwardSimulationStepTrace<Configuration, ConfigAlloc>& trace,
                                                               const bool enable_tracing, ros::Publisher& display_debug_publisher) const override
    {
        const std::pair<Configuration, bool> r = ForwardSimulateRobot(mutable_robot, mutable_robot.GetPosition(), target_position, rng, forward_simulation_time,
                                                                      simulation_shortcut_distance, use_individual_jacobians, allow_contacts, trace,
                                                                      enable_tracing, display_debug_publisher);
        mutable_robot.UpdatePosition(r.first);
        return r;
    }
#endif

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

protected:
    EnvironmentDescription env_;
    RobotDescription robot_;
    SolverConfig solver_;
    std::shared_ptr<Handle> handle_;
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

    B200OutcomeClustering(upc_handle* handle, const int32_t clustering_type, const double signature_matching_threshold, const double distance_clustering_threshold)
        : handle_(handle)
    {
        params_ = upc_cluster_params{};
        params_.feature_type = clustering_type;
        params_.signature_matching_threshold = signature_matching_threshold;
        params_.distance_clustering_threshold = distance_clustering_threshold;
    }

    /* index clusters of one particle set, numbered by smallest member */
    std::vector<std::vector<size_t>> ClusterIndices(const std::vector<const Configuration*>& configs) const
    {
        const size_t n = configs.size();
        std::vector<std::vector<size_t>> clusters;
        if (n == 0) return clusters;
        const int width = Traits::width(*configs[0]);
        std::vector<double> soa((size_t)width * n), tmp((size_t)width);
        for (size_t i = 0; i < n; i++)
        {
            Traits::flatten(*configs[i], tmp.data());
            for (int c = 0; c < width; c++) soa[(size_t)c * n + i] = tmp[(size_t)c];
        }
        std::vector<int32_t> labels(n);
        int32_t n_clusters = 0;
        check(upc_cluster_particles(handle_, &params_, 1, (int64_t)n, soa.data(), labels.data(), &n_clusters));
        clusters.resize((size_t)n_clusters);
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

private:
    upc_handle* handle_;
    upc_cluster_params params_;
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

} /* namespace upc_b200 */

#endif /* UPC_B200_SHIM_HPP */

/*
 * shim_smoke.cpp - drives the C++ host shim (include/upc_b200_shim.hpp) the way the reference's planner drives its
 * simulator plugin (uncertainty_contact_planning.hpp:2039-2071 then :1986), stand-alone mode (no ROS / Eigen here),
 * and checks the outcome against the CPU oracle given the very same noise tape.
 * Built and run by tests/test_cpp_shim.py.  Exit code 0 = parity.
 */
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include "upc_b200_shim.hpp"

extern "C" int32_t oracle_forward_simulate(const upc_env_desc*, const upc_robot_desc*, const upc_sim_params*, int64_t, const double*, int64_t,
                                           const double*, const double*, double*, uint8_t*, upc_stats*);
extern "C" int32_t oracle_cluster_statistics(const upc_robot_desc*, int64_t, const double*, const int32_t*, const uint8_t*, int32_t, double, int32_t, int64_t*,
                                             uint8_t*, double*, double*, double*, double*, double*);
extern "C" int32_t oracle_cluster_particles(const upc_env_desc*, const upc_robot_desc*, const upc_cluster_params*, int64_t, int64_t, const double*,
                                            int32_t*, int32_t*, upc_stats*);

typedef std::array<double, 3> SE2Config;
struct StubRobot {}; /* the planner's robot object is only borrowed by the simulator */

int main()
{
    using namespace upc_b200;
    /* 64 x 64 x 1 world, 0.1 m cells, one box obstacle */
    EnvironmentDescription env;
    env.desc.nx = 64; env.desc.ny = 64; env.desc.nz = 1; env.desc.resolution = 0.1;
    const double gfw[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0.05};
    std::memcpy(env.desc.grid_from_world, gfw, sizeof(gfw));
    env.desc.sdf_oob_value = 1000.0f; env.desc.occupancy_oob_value = 0.0f;
    env.sdf.resize(64 * 64); env.convex_segment.resize(64 * 64); env.occupancy.resize(64 * 64);
    for (int x = 0; x < 64; x++)
        for (int y = 0; y < 64; y++)
        {
            const double px = (x + 0.5) * 0.1, py = (y + 0.5) * 0.1;
            const double qx = std::fabs(px - 3.2) - 0.6, qy = std::fabs(py - 3.6) - 0.5;
            const double outside = std::sqrt(std::pow(std::max(qx, 0.0), 2) + std::pow(std::max(qy, 0.0), 2));
            const double inside = std::min(std::max(qx, qy), 0.0);
            env.sdf[(size_t)(x * 64 + y)] = (float)(outside + inside);
            env.occupancy[(size_t)(x * 64 + y)] = (outside + inside) < 0 ? 1.0f : 0.0f;
            env.convex_segment[(size_t)(x * 64 + y)] = (outside + inside) < 0 ? 0u : ((x < 34 ? 1u : 0u) | (x >= 30 ? 2u : 0u));
        }
    env.finalize();
    RobotDescription robot;
    robot.desc.kind = UPC_ROBOT_SE2; robot.desc.dof = 3; robot.desc.num_links = 1;
    for (int ix = 0; ix < 4; ix++)
        for (int iy = 0; iy < 4; iy++) { robot.points.push_back((ix - 1.5) * 0.08); robot.points.push_back((iy - 1.5) * 0.08); robot.points.push_back(0.0); robot.points.push_back(0.0); }
    robot.finalize();
    robot.desc.link_point_offset[0] = 0; robot.desc.link_point_offset[1] = robot.desc.num_points;
    for (int a = 0; a < 3; a++) { upc_axis_params& ax = robot.desc.axis[a]; ax.kp = 5.0; ax.velocity_limit = 1.0; ax.max_sensor_noise = 0.002; ax.max_actuator_noise = 0.15; }

    SolverConfig solver;
    B200Simulator<StubRobot, SE2Config, std::mt19937_64> simulator(env, robot, 0, solver);
    const size_t n = 300;
    std::vector<SE2Config> starts(n, SE2Config{3.0, 2.4, 0.1}), target(1, SE2Config{3.5, 3.05, -0.2});
    std::vector<std::mt19937_64> rngs(4, std::mt19937_64(42));
    std::vector<std::mt19937_64> rngs_copy = rngs;
    StubRobot stub;
    const double sim_time = 0.6, shortcut = 0.01;
    const auto results = simulator.ForwardSimulateRobots(stub, starts, target, rngs, sim_time, shortcut, false, true);
    if (results.size() != n) { std::printf("wrong result count\n"); return 1; }
    if (rngs[0] == rngs_copy[0]) { std::printf("the caller's generator was not advanced\n"); return 1; }

    /* the oracle on the same inputs and the same tape */
    const upc_sim_params params = simulator.MakeParams(sim_time, shortcut, false, true);
    std::uniform_int_distribution<uint64_t> seed_dist(0, std::numeric_limits<uint64_t>::max());
    const uint64_t seed = seed_dist(rngs_copy[0]);
    std::vector<double> tape((size_t)upc_tape_doubles_per_particle(&params, 3) * n), s_soa(3 * n), t_soa(3), out(3 * n);
    check(upc_fill_noise_tape(seed, &robot.desc, &params, 0, (int64_t)n, (int64_t)n, tape.data()));
    for (size_t i = 0; i < n; i++)
        for (int c = 0; c < 3; c++) s_soa[(size_t)c * n + i] = starts[i][(size_t)c];
    for (int c = 0; c < 3; c++) t_soa[(size_t)c] = target[0][(size_t)c];
    std::vector<uint8_t> coll(n);
    if (oracle_forward_simulate(&env.desc, &robot.desc, &params, (int64_t)n, s_soa.data(), 1, t_soa.data(), tape.data(), out.data(), coll.data(), nullptr) != 0) return 2;
    size_t collided = 0;
    for (size_t i = 0; i < n; i++)
    {
        if (results[i].second != (coll[i] != 0)) { std::printf("collision flag %zu differs\n", i); return 3; }
        for (int c = 0; c < 3; c++)
            if (results[i].first[(size_t)c] != out[(size_t)c * n + i]) { std::printf("configuration %zu differs\n", i); return 4; }
        collided += results[i].second ? 1 : 0;
    }

    /* ClusterParticles through the shim vs the oracle's labels */
    B200OutcomeClustering<StubRobot, SE2Config> clustering(simulator.handle(), UPC_FEATURE_CONVEX_REGION_SIGNATURE, 0.125, 0.05);
    const auto clusters = clustering.ClusterParticles(results, true);
    upc_cluster_params cp{};
    cp.feature_type = UPC_FEATURE_CONVEX_REGION_SIGNATURE; cp.signature_matching_threshold = 0.125; cp.distance_clustering_threshold = 0.05;
    std::vector<int32_t> labels(n);
    int32_t ncl = 0;
    if (oracle_cluster_particles(&env.desc, &robot.desc, &cp, 1, (int64_t)n, out.data(), labels.data(), &ncl, nullptr) != 0) return 5;
    if ((size_t)ncl != clusters.size()) { std::printf("cluster count differs: %d vs %zu\n", ncl, clusters.size()); return 6; }
    std::vector<size_t> sizes((size_t)ncl, 0);
    for (size_t i = 0; i < n; i++) sizes[(size_t)labels[i]]++;
    for (size_t k = 0; k < clusters.size(); k++)
        if (clusters[k].size() != sizes[k]) { std::printf("cluster %zu size differs\n", k); return 7; }
    const auto without = clustering.ClusterParticles(results, false);
    size_t kept = 0;
    for (const auto& c : without) kept += c.size();
    if (kept != n - collided) { std::printf("allow_contacts=false kept %zu, expected %zu\n", kept, n - collided); return 8; }
    {
        /* per-cluster statistics through the shim vs the oracle */
        const auto index_clusters = [&]() { std::vector<const SE2Config*> p; for (const auto& r : results) p.push_back(&r.first); return clustering.ClusterIndices(p); }();
        const auto st = clustering.ClusterStatistics(results, index_clusters, true, 0.25, UPC_ROBOT_SE2, 3);
        const size_t K = index_clusters.size();
        std::vector<int64_t> counts(K);
        std::vector<uint8_t> any(K);
        std::vector<double> e(K * 3), v(K), si(K), vs(K * 3), sis(K * 3);
        if (oracle_cluster_statistics(&robot.desc, (int64_t)n, out.data(), labels.data(), coll.data(), 1, 0.25, (int32_t)K, counts.data(), any.data(), e.data(),
                                      v.data(), si.data(), vs.data(), sis.data()) != 0) return 12;
        if (st.counts != counts || st.any_collided != any || st.expectations != e || st.variance != v || st.si_variance != si || st.variances != vs || st.si_variances != sis)
        { std::printf("cluster statistics differ\n"); return 13; }
    }
    const auto idx = clustering.PolicyParticleClusteringFn(std::vector<SE2Config>(starts.begin(), starts.begin() + 10), starts[0]);
    if (idx.size() != 1 || idx[0].size() != 11) { std::printf("policy clustering wrong\n"); return 9; }
    const bool in_collision = simulator.CheckConfigCollision(stub, SE2Config{3.2, 3.6, 0.0});
    const bool free_space = simulator.CheckConfigCollision(stub, SE2Config{1.0, 1.0, 0.0});
    if (!in_collision || free_space) { std::printf("CheckConfigCollision wrong\n"); return 10; }
    bool threw = false;
    try { simulator.ForwardSimulateRobots(stub, starts, std::vector<SE2Config>(2, target[0]), rngs, sim_time, shortcut, false, true); }
    catch (const std::invalid_argument&) { threw = true; }
    if (!threw) { std::printf("bad target count did not throw\n"); return 11; }
    const auto stats = simulator.GetStatistics();
    std::printf("shim ok: %zu particles, %zu collided, %zu clusters, %.0f kernel launches\n", n, collided, clusters.size(), stats.at("kernel_launches"));
    return 0;
}

/*
 * shim_smoke.cpp - drives the C++ host shim (include/upc_b200_shim.hpp) the way the reference's planner drives its
 * simulator plugin (uncertainty_contact_planning.hpp:2039-2071 then :1986), stand-alone mode (no ROS / Eigen here),
 * and checks every outcome against the CPU oracle replaying the same seeded noise.
 * Built and run by tests/test_cpp_shim.py.  Exit code 0 = parity.
 *   shim_smoke            all single-GPU checks
 *   shim_smoke sharded N  the edge batch on N GPUs of this box (ShardedSimulator, NCCL from the library) against one GPU
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "upc_b200_shim.hpp"

extern "C" int32_t oracle_forward_simulate(const upc_env_desc*, const upc_robot_desc*, const upc_sim_params*, int64_t, const double*, int64_t,
                                           const double*, const double*, double*, uint8_t*, upc_stats*);
extern "C" int32_t oracle_forward_simulate_traced(const upc_env_desc*, const upc_robot_desc*, const upc_sim_params*, int64_t, const double*, int64_t,
                                                  const double*, const double*, double*, uint8_t*, double*, double*, int32_t*);
extern "C" int32_t oracle_fill_noise_tape_counter(uint64_t, const upc_robot_desc*, const upc_sim_params*, int64_t, int64_t, int64_t, double*);
extern "C" int32_t oracle_cluster_statistics(const upc_robot_desc*, int64_t, const double*, const int32_t*, const uint8_t*, int32_t, double, int32_t, int64_t*,
                                             uint8_t*, double*, double*, double*, double*, double*);
extern "C" int32_t oracle_cluster_particles(const upc_env_desc*, const upc_robot_desc*, const upc_cluster_params*, int64_t, int64_t, const double*,
                                            int32_t*, int32_t*, upc_stats*);

extern "C" int32_t oracle_nearest_neighbors(const upc_robot_desc*, int64_t, const double*, const double*, const double*, const uint8_t*, int64_t, const double*,
                                            double, int64_t*, double*);

typedef std::array<double, 3> SE2Config;
typedef std::pair<SE2Config, bool> Particle;
typedef std::array<double, 3> SE2ConfigFwd;
/* the planner's robot object is only borrowed by the simulator; the extension-batch helper asks it for SimpleSE2Robot's distance and
 * interpolation (simple_robot_models.hpp:401-430) */
struct StubRobot
{
    static double wrap(double a) { while (a > M_PI) a -= 2.0 * M_PI; while (a <= -M_PI) a += 2.0 * M_PI; return a; }
    double ComputeConfigurationDistance(const SE2ConfigFwd& a, const SE2ConfigFwd& b) const
    {
        return std::sqrt((b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1])) + std::fabs(wrap(b[2] - a[2]));
    }
    SE2ConfigFwd InterpolateBetweenConfigurations(const SE2ConfigFwd& a, const SE2ConfigFwd& b, const double ratio) const
    {
        return SE2ConfigFwd{a[0] + (b[0] - a[0]) * ratio, a[1] + (b[1] - a[1]) * ratio, wrap(a[2] + wrap(b[2] - a[2]) * ratio)};
    }
};

using namespace upc_b200;

static uint64_t next_seed(std::mt19937_64& g)
{
    std::uniform_int_distribution<uint64_t> d(0, std::numeric_limits<uint64_t>::max());
    return d(g);
}

static void make_world(EnvironmentDescription& env, RobotDescription& robot)
{
    /* 64 x 64 x 1 world, 0.1 m cells, one box obstacle */
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
    robot.desc.kind = UPC_ROBOT_SE2; robot.desc.dof = 3; robot.desc.num_links = 1;
    for (int ix = 0; ix < 4; ix++)
        for (int iy = 0; iy < 4; iy++) { robot.points.push_back((ix - 1.5) * 0.08); robot.points.push_back((iy - 1.5) * 0.08); robot.points.push_back(0.0); robot.points.push_back(0.0); }
    robot.finalize();
    robot.desc.link_point_offset[0] = 0; robot.desc.link_point_offset[1] = robot.desc.num_points;
    for (int a = 0; a < 3; a++) { upc_axis_params& ax = robot.desc.axis[a]; ax.kp = 5.0; ax.velocity_limit = 1.0; ax.max_sensor_noise = 0.002; ax.max_actuator_noise = 0.15; }
}

/* E edges around the obstacle: starts below it, targets partly inside it */
static void make_edges(size_t E, std::vector<SE2Config>& starts, std::vector<SE2Config>& targets)
{
    for (size_t e = 0; e < E; e++)
    {
        const double f = (double)e / (double)(E > 1 ? E - 1 : 1);
        starts.push_back(SE2Config{2.6 + 1.2 * f, 2.4 + 0.1 * std::sin(7.0 * f), 0.1 - 0.3 * f});
        targets.push_back(SE2Config{2.9 + 0.8 * f, 2.95 + 0.2 * std::cos(5.0 * f), -0.2 + 0.4 * f});
    }
}

/* oracle clusters of `sets` equally sized outcome sets, in ClusterParticles shape */
static int oracle_edge_clusters(const EnvironmentDescription& env, const RobotDescription& robot, const upc_cluster_params& cp, size_t E, size_t P,
                                const std::vector<double>& out, const std::vector<uint8_t>& coll, bool allow_contacts,
                                std::vector<std::vector<std::vector<Particle>>>& clusters)
{
    const size_t n = E * P;
    std::vector<int32_t> labels(n), ncl(E);
    if (oracle_cluster_particles(&env.desc, &robot.desc, &cp, (int64_t)E, (int64_t)P, out.data(), labels.data(), ncl.data(), nullptr) != 0) return 1;
    clusters.assign(E, std::vector<std::vector<Particle>>());
    for (size_t e = 0; e < E; e++)
    {
        clusters[e].resize((size_t)ncl[e]);
        for (size_t k = 0; k < P; k++)
        {
            const size_t i = e * P + k;
            if (coll[i] && !allow_contacts) continue;
            clusters[e][(size_t)labels[i]].push_back(Particle(SE2Config{out[i], out[n + i], out[2 * n + i]}, coll[i] != 0));
        }
    }
    return 0;
}

static int run_sharded(int devices)
{
    EnvironmentDescription env;
    RobotDescription robot;
    make_world(env, robot);
    SolverConfig solver;
    const size_t E = 37, P = 64; /* 37 edges: uneven shards */
    std::vector<SE2Config> es, et;
    make_edges(E, es, et);
    StubRobot stub;
    B200Simulator<StubRobot, SE2Config, std::mt19937_64> single(env, robot, 0, solver, 0);
    B200OutcomeClustering<StubRobot, SE2Config> clustering(single.handle(), CONVEX_REGION_SIGNATURE, 0.125, 0.05);
    std::vector<std::mt19937_64> r1(1, std::mt19937_64(7)), r2(1, std::mt19937_64(7));
    const auto ref = single.SimulateAndClusterEdges(stub, es, et, P, r1, 0.5, 0.01, false, true, clustering);
    std::vector<int> devs;
    for (int d = 0; d < devices; d++) devs.push_back(d);
    ShardedSimulator<StubRobot, SE2Config, std::mt19937_64> sharded(env, robot, devs, solver);
    for (int rep = 0; rep < 2; rep++)
    {
        std::vector<std::mt19937_64> rr = r2;
        const auto got = sharded.SimulateAndClusterEdges(stub, es, et, P, rr, 0.5, 0.01, false, true, clustering);
        if (got != ref) { std::printf("sharded result differs from the single-GPU result (rep %d)\n", rep); return 20; }
    }
    int32_t version = 0;
    check(upc_comm_nccl_version(&version));
    std::printf("sharded ok: %d GPUs, %zu edges x %zu particles, NCCL %d\n", devices, E, P, version);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc >= 3 && std::strcmp(argv[1], "sharded") == 0) return run_sharded(std::atoi(argv[2]));
    EnvironmentDescription env;
    RobotDescription robot;
    make_world(env, robot);

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
    {
        const auto st = simulator.GetStatistics();
        if (st.at("h2d_bytes") > 4.0e6) { std::printf("a tape crossed the boundary (%.0f bytes uploaded)\n", st.at("h2d_bytes")); return 1; }
    }

    /* the oracle on the same inputs, replaying the oracle's own tape of the same seed */
    const upc_sim_params params = simulator.MakeParams(sim_time, shortcut, false, true);
    const uint64_t seed = next_seed(rngs_copy[0]);
    std::vector<double> tape((size_t)upc_tape_doubles_per_particle(&params, 3) * n), s_soa(3 * n), t_soa(3), out(3 * n);
    if (oracle_fill_noise_tape_counter(seed, &robot.desc, &params, 0, (int64_t)n, (int64_t)n, tape.data()) != 0) return 2;
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

    /* ClusterParticles through the shim, constructed from the reference's enum (CRS = 0), vs the oracle's labels */
    if (FromReferenceClusteringType(CONVEX_REGION_SIGNATURE) != UPC_FEATURE_CONVEX_REGION_SIGNATURE || FromReferenceClusteringType(ACTUATION_CENTER_CONNECTIVITY) != UPC_FEATURE_ACTUATION_CENTER_CONNECTIVITY ||
        FromReferenceClusteringType(POINT_TO_POINT_MOVEMENT) != UPC_FEATURE_POINT_TO_POINT_MOVEMENT) { std::printf("enum mapping wrong\n"); return 5; }
    {
        bool threw = false;
        try { FromReferenceClusteringType(COMPARE); } catch (const std::invalid_argument&) { threw = true; }
        if (!threw) { std::printf("COMPARE was accepted\n"); return 5; }
    }
    B200OutcomeClustering<StubRobot, SE2Config> clustering(simulator.handle(), (SPATIAL_FEATURE_CLUSTERING_TYPE)0 /* CRS as the planner stores it */, 0.125, 0.05);
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
    {
        /* PolicyParticleClusteringFn (uncertainty_contact_planning.hpp:1558-1578) vs the oracle clustering (particles + current) */
        const size_t m = 40;
        std::vector<SE2Config> some;
        for (size_t i = 0; i < m; i++) some.push_back(results[i * 7].first);
        const SE2Config current = results[5].first;
        const auto idx = clustering.PolicyParticleClusteringFn(some, current);
        std::vector<double> soa(3 * (m + 1));
        for (size_t i = 0; i <= m; i++)
            for (int c = 0; c < 3; c++) soa[(size_t)c * (m + 1) + i] = (i < m) ? some[i][(size_t)c] : current[(size_t)c];
        std::vector<int32_t> pl(m + 1);
        int32_t pn = 0;
        if (oracle_cluster_particles(&env.desc, &robot.desc, &cp, 1, (int64_t)(m + 1), soa.data(), pl.data(), &pn, nullptr) != 0) return 9;
        if ((size_t)pn != idx.size()) { std::printf("policy clustering: cluster count differs\n"); return 9; }
        for (size_t k = 0; k < idx.size(); k++)
            for (size_t e = 0; e < idx[k].size(); e++)
                if (pl[idx[k][e]] != (int32_t)k) { std::printf("policy clustering: membership differs\n"); return 9; }
    }
    const bool in_collision = simulator.CheckConfigCollision(stub, SE2Config{3.2, 3.6, 0.0});
    const bool free_space = simulator.CheckConfigCollision(stub, SE2Config{1.0, 1.0, 0.0});
    if (!in_collision || free_space) { std::printf("CheckConfigCollision wrong\n"); return 10; }
    bool threw = false;
    try { simulator.ForwardSimulateRobots(stub, starts, std::vector<SE2Config>(2, target[0]), rngs, sim_time, shortcut, false, true); }
    catch (const std::invalid_argument&) { threw = true; }
    if (!threw) { std::printf("bad target count did not throw\n"); return 11; }

    {
        /* ForwardSimulateRobot with tracing (simple_simulator_interface.hpp:621, 40-86) vs the oracle's per-interval record */
        std::mt19937_64 g(99), g_copy(99);
        ForwardSimulationStepTrace<SE2Config> trace;
        const Particle r = simulator.ForwardSimulateRobot(stub, starts[0], target[0], g, sim_time, shortcut, false, true, trace, true);
        const uint64_t s1 = next_seed(g_copy);
        const size_t S = (size_t)params.num_steps;
        std::vector<double> t1((size_t)upc_tape_doubles_per_particle(&params, 3)), o1(3), tc(S * 3, 0.0), tu(S * 6, 0.0);
        int32_t ts[2] = {0, 0};
        uint8_t c1 = 0;
        if (oracle_fill_noise_tape_counter(s1, &robot.desc, &params, 0, 1, 1, t1.data()) != 0) return 14;
        const std::vector<double> s1soa = {starts[0][0], starts[0][1], starts[0][2]};
        if (oracle_forward_simulate_traced(&env.desc, &robot.desc, &params, 1, s1soa.data(), 1, t_soa.data(), t1.data(), o1.data(), &c1, tc.data(), tu.data(), ts) != 0) return 14;
        if ((int32_t)trace.resolver_steps.size() != ts[0] || ts[0] < 1) { std::printf("trace length %zu, oracle %d\n", trace.resolver_steps.size(), ts[0]); return 15; }
        for (int32_t step = 0; step < ts[0]; step++)
        {
            const auto& st = trace.resolver_steps[(size_t)step];
            const SE2Config& resolved = st.contact_resolver_steps.back().contact_resolution_steps.back();
            for (int c = 0; c < 3; c++)
            {
                if (resolved[(size_t)c] != tc[(size_t)step * 3 + (size_t)c]) { std::printf("trace config differs at step %d\n", step); return 15; }
                if (st.control_input[(size_t)c] != tu[(size_t)step * 6 + (size_t)c] || st.control_input_step[(size_t)c] != tu[(size_t)step * 6 + 3 + (size_t)c])
                { std::printf("trace control differs at step %d\n", step); return 15; }
            }
        }
        const std::vector<SE2Config> trajectory = ExtractTrajectoryFromTrace(trace);
        if (trajectory.empty() || trajectory.back() != r.first || r.first != SE2Config{o1[0], o1[1], o1[2]} || r.second != (c1 != 0)) { std::printf("traced result differs\n"); return 15; }
    }

    {
        /* the planner-batch step: edges simulated + clustered in one call == oracle edge by edge; ClusterParticleSets agrees */
        const size_t E = 12, P = 96;
        std::vector<SE2Config> es, et;
        make_edges(E, es, et);
        std::vector<std::mt19937_64> g(1, std::mt19937_64(1234)), g_copy = g;
        const auto batch = simulator.SimulateAndClusterEdges(stub, es, et, P, g, 0.5, 0.01, false, false, clustering);
        const upc_sim_params bp = simulator.MakeParams(0.5, 0.01, false, false);
        const uint64_t bs = next_seed(g_copy[0]);
        const size_t bn = E * P;
        std::vector<double> bt((size_t)upc_tape_doubles_per_particle(&bp, 3) * bn), bs_soa(3 * bn), bt_soa(3 * bn), bout(3 * bn);
        std::vector<uint8_t> bcoll(bn);
        if (oracle_fill_noise_tape_counter(bs, &robot.desc, &bp, 0, (int64_t)bn, (int64_t)bn, bt.data()) != 0) return 16;
        for (size_t i = 0; i < bn; i++)
            for (int c = 0; c < 3; c++) { bs_soa[(size_t)c * bn + i] = es[i / P][(size_t)c]; bt_soa[(size_t)c * bn + i] = et[i / P][(size_t)c]; }
        if (oracle_forward_simulate(&env.desc, &robot.desc, &bp, (int64_t)bn, bs_soa.data(), (int64_t)bn, bt_soa.data(), bt.data(), bout.data(), bcoll.data(), nullptr) != 0) return 16;
        std::vector<std::vector<std::vector<Particle>>> expect;
        if (oracle_edge_clusters(env, robot, cp, E, P, bout, bcoll, false, expect) != 0) return 16;
        if (batch != expect) { std::printf("edge batch differs from the oracle\n"); return 17; }
        std::vector<std::vector<Particle>> sets(E);
        for (size_t i = 0; i < bn; i++) sets[i / P].push_back(Particle(SE2Config{bout[i], bout[bn + i], bout[2 * bn + i]}, bcoll[i] != 0));
        if (clustering.ClusterParticleSets(sets, false) != expect) { std::printf("ClusterParticleSets differs from the oracle\n"); return 18; }
        /* the same step as submit / collect pairs, two batches in flight: batch by batch what the blocking call returns */
        std::vector<std::mt19937_64> ga(1, std::mt19937_64(99)), gb = ga;
        std::vector<std::vector<std::vector<std::vector<Particle>>>> blocking;
        for (int k = 0; k < 3; k++) blocking.push_back(simulator.SimulateAndClusterEdges(stub, es, et, P, ga, 0.5, 0.01, false, false, clustering));
        simulator.SubmitEdgeBatch(stub, es, et, P, gb, 0.5, 0.01, false, false, 0);
        for (int k = 0; k < 3; k++)
        {
            if (k + 1 < 3) simulator.SubmitEdgeBatch(stub, es, et, P, gb, 0.5, 0.01, false, false, (k + 1) % 2);
            if (simulator.CollectEdgeBatch(clustering, k % 2) != blocking[(size_t)k]) { std::printf("submit / collect batch %d differs from the blocking call\n", k); return 18; }
        }
        if (blocking[0] == blocking[1]) { std::printf("batches with different seeds must differ\n"); return 18; }
        bool empty_slot_throws = false;
        try { simulator.CollectEdgeBatch(clustering, 0); } catch (const std::invalid_argument&) { empty_slot_throws = true; }
        if (!empty_slot_throws) { std::printf("collecting an empty slot did not throw\n"); return 18; }
    }

    {
        /* POINT_TO_POINT_MOVEMENT through the shim (uncertainty_contact_planning.hpp:1801-1863), seeded pair simulations */
        const size_t m = 14;
        std::vector<Particle> few;
        for (size_t i = 0; i < m; i++) few.push_back(results[i * 20]);
        B200OutcomeClustering<StubRobot, SE2Config> ptpm(simulator.handle(), POINT_TO_POINT_MOVEMENT, 0.125, 0.05);
        bool unconfigured_throws = false;
        try { ptpm.ClusterParticles(few, true); } catch (const std::invalid_argument&) { unconfigured_throws = true; }
        if (!unconfigured_throws) { std::printf("PTPM without simulation parameters did not throw\n"); return 19; }
        const double goal_threshold = 0.06;
        const upc_sim_params pair_params = simulator.MakeParams(0.25, goal_threshold * 0.5, false, true);
        ptpm.EnablePointToPointMovement(pair_params, goal_threshold, 555);
        const auto pc = ptpm.ClusterParticles(few, true);
        upc_cluster_params pp = ptpm.params();
        std::vector<double> soa(3 * m);
        for (size_t i = 0; i < m; i++)
            for (int c = 0; c < 3; c++) soa[(size_t)c * m + i] = few[i].first[(size_t)c];
        std::vector<int32_t> pl(m);
        int32_t pn = 0;
        if (oracle_cluster_particles(&env.desc, &robot.desc, &pp, 1, (int64_t)m, soa.data(), pl.data(), &pn, nullptr) != 0) return 19;
        if ((size_t)pn != pc.size()) { std::printf("PTPM cluster count differs: %d vs %zu\n", pn, pc.size()); return 19; }
        std::vector<size_t> psz((size_t)pn, 0);
        for (size_t i = 0; i < m; i++) psz[(size_t)pl[i]]++;
        for (size_t k = 0; k < pc.size(); k++)
            if (pc[k].size() != psz[k]) { std::printf("PTPM cluster %zu size differs\n", k); return 19; }
    }
    {
        /* speculative multi-edge extension (RRT-Extend branch, uncertainty_contact_planning.hpp:2320-2340) for a batch of samples:
         * nearest states == the oracle's, targets == the reference's interpolation rule, clusters == the edge batch issued by hand */
        const StubRobot& planner_robot = stub;
        const std::vector<SE2Config> tree = {{2.6, 2.4, 0.1}, {3.0, 2.5, 0.0}, {3.4, 2.4, -0.2}, {2.8, 2.9, 0.3}, {3.8, 2.6, 0.0}, {2.2, 2.2, 0.0}};
        std::vector<const SE2Config*> expectations;
        for (const SE2Config& c : tree) expectations.push_back(&c);
        const std::vector<double> fw = {1.0, 1.0, 1.2, 1.0, 0.9, 1.0}, vw = {1.0, 1.1, 1.0, 1.0, 1.0, 1.0};
        const std::vector<uint8_t> use = {1, 1, 1, 1, 1, 0};
        std::vector<SE2Config> samples;
        for (int q = 0; q < 10; q++) samples.push_back(SE2Config{2.3 + 0.17 * q, 2.2 + 0.09 * ((q * 7) % 10), -0.4 + 0.09 * q});
        samples[3] = SE2Config{3.02, 2.52, 0.01}; /* closer than the step size: the sample itself is the target */
        const double step_size = 0.3;
        const size_t P = 64;
        std::vector<std::mt19937_64> g(1, std::mt19937_64(4321)), g_copy = g;
        const auto ext = ExtendTowardsSamples(simulator, simulator.handle(), planner_robot, expectations, fw, vw, use, samples, step_size, P, g, 0.5, 0.01, false,
                                              true, clustering);
        std::vector<double> esoa(3 * tree.size()), qsoa(3 * samples.size()), odist(samples.size());
        for (size_t i = 0; i < tree.size(); i++)
            for (int c = 0; c < 3; c++) esoa[(size_t)c * tree.size() + i] = tree[i][(size_t)c];
        for (size_t q = 0; q < samples.size(); q++)
            for (int c = 0; c < 3; c++) qsoa[(size_t)c * samples.size() + q] = samples[q][(size_t)c];
        std::vector<int64_t> onear(samples.size());
        if (oracle_nearest_neighbors(&robot.desc, (int64_t)tree.size(), esoa.data(), fw.data(), vw.data(), use.data(), (int64_t)samples.size(), qsoa.data(), step_size,
                                     onear.data(), odist.data()) != 0) return 20;
        if (ext.nearest != onear) { std::printf("extension batch: nearest states differ from the oracle\n"); return 20; }
        std::vector<SE2Config> es, et;
        bool interpolated = false, direct = false;
        for (size_t q = 0; q < samples.size(); q++)
        {
            if (ext.nearest[q] == 5) { std::printf("extension batch: a disabled state was selected\n"); return 20; }
            const SE2Config& from = tree[(size_t)ext.nearest[q]];
            const double d = planner_robot.ComputeConfigurationDistance(from, samples[q]);
            const SE2Config want = (d > step_size) ? planner_robot.InterpolateBetweenConfigurations(from, samples[q], step_size / d) : samples[q];
            if (ext.targets[q] != want) { std::printf("extension batch: target %zu differs\n", q); return 20; }
            interpolated |= d > step_size;
            direct |= !(d > step_size);
            if (std::fabs(planner_robot.ComputeConfigurationDistance(from, want) - std::min(d, step_size)) > 1e-12) { std::printf("extension batch: step length\n"); return 20; }
            es.push_back(from);
            et.push_back(want);
        }
        if (!interpolated || !direct) { std::printf("extension batch: both target rules must be exercised\n"); return 20; }
        const auto by_hand = simulator.SimulateAndClusterEdges(stub, es, et, P, g_copy, 0.5, 0.01, false, true, clustering);
        if (ext.clusters != by_hand) { std::printf("extension batch: clusters differ from the edge batch issued by hand\n"); return 20; }
        if (g[0] != g_copy[0]) { std::printf("extension batch: generator advanced differently\n"); return 20; }
    }
    const auto stats = simulator.GetStatistics();
    std::printf("shim ok: %zu particles, %zu collided, %zu clusters, %.0f kernel launches\n", n, collided, clusters.size(), stats.at("kernel_launches"));
    return 0;
}

/*
 * plugin_bench.cpp - times the hot path through the reference-facing C++ plugin (include/upc_b200_shim.hpp), from
 * std::vector<Configuration> in to std::vector<std::pair<Configuration, bool>> / clusters out, the way the planner holds its data
 * (uncertainty_contact_planning.hpp:2039-2071 ForwardSimulateParticles, :1986-2034 ClusterParticles).  bench.py writes the
 * workload (the raw POD descriptors of include/upc_b200.h plus their arrays) to a file, runs this driver and reports the result
 * as the `e2e_plugin` key.  Everything a planner would pay for is inside the timed region: AoS -> SoA marshalling, uploads,
 * kernels, downloads, and building the nested result vectors.
 *
 *   plugin_bench <workload file> edges  <steps> <warmup> [gpus]   B200Simulator / ShardedSimulator::SimulateAndClusterEdges
 *   plugin_bench <workload file> single <steps> <warmup>          ForwardSimulateRobots + ClusterParticles on one edge
 */
#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "upc_b200_shim.hpp"

using namespace upc_b200;

struct StubRobot {};

struct Workload
{
    EnvironmentDescription env;
    RobotDescription robot;
    upc_sim_params params;
    double signature_threshold, distance_threshold;
    int64_t n_starts, n_targets, particles_per_start, width;
    std::vector<double> starts, targets; /* row per configuration */
};

static void read_exact(FILE* f, void* dst, size_t bytes)
{
    if (bytes && fread(dst, 1, bytes, f) != bytes)
    {
        std::fprintf(stderr, "workload file truncated\n");
        std::exit(2);
    }
}

static Workload load(const char* path)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path); std::exit(2); }
    Workload w;
    char magic[8];
    read_exact(f, magic, 8);
    if (std::memcmp(magic, "UPCWL01", 8) != 0) { std::fprintf(stderr, "bad workload file\n"); std::exit(2); }
    read_exact(f, &w.env.desc, sizeof(upc_env_desc));
    read_exact(f, &w.robot.desc, sizeof(upc_robot_desc));
    read_exact(f, &w.params, sizeof(upc_sim_params));
    int64_t head[6];
    read_exact(f, head, sizeof(head));
    const int64_t has_seg = head[0], has_occ = head[1];
    w.n_starts = head[2]; w.n_targets = head[3]; w.particles_per_start = head[4]; w.width = head[5];
    read_exact(f, &w.signature_threshold, 8);
    read_exact(f, &w.distance_threshold, 8);
    const size_t cells = (size_t)w.env.desc.nx * (size_t)w.env.desc.ny * (size_t)w.env.desc.nz;
    w.env.sdf.resize(cells);
    read_exact(f, w.env.sdf.data(), cells * 4);
    if (has_seg) { w.env.convex_segment.resize(cells); read_exact(f, w.env.convex_segment.data(), cells * 4); }
    if (has_occ) { w.env.occupancy.resize(cells); read_exact(f, w.env.occupancy.data(), cells * 4); }
    w.env.finalize();
    w.robot.points.resize((size_t)w.robot.desc.num_points * 4);
    read_exact(f, w.robot.points.data(), w.robot.points.size() * 8);
    w.robot.finalize();
    w.starts.resize((size_t)(w.n_starts * w.width));
    w.targets.resize((size_t)(w.n_targets * w.width));
    read_exact(f, w.starts.data(), w.starts.size() * 8);
    read_exact(f, w.targets.data(), w.targets.size() * 8);
    std::fclose(f);
    return w;
}

template <size_t N>
static std::vector<std::array<double, N>> to_configs(const std::vector<double>& rows, int64_t n)
{
    std::vector<std::array<double, N>> out((size_t)n);
    for (int64_t i = 0; i < n; i++)
        for (size_t c = 0; c < N; c++) out[(size_t)i][c] = rows[(size_t)i * N + c];
    return out;
}

template <size_t N>
static int run(const Workload& w, const std::string& mode, int steps, int warmup, int gpus)
{
    typedef std::array<double, N> Config;
    typedef B200Simulator<StubRobot, Config, std::mt19937_64> Simulator;
    SolverConfig solver;
    solver.controller_frequency = 1.0 / w.params.controller_interval;
    solver.max_microsteps = w.params.max_microsteps;
    solver.contact_distance = w.params.contact_distance;
    solver.resolve_margin_ratio = (w.params.resolve_distance - w.params.contact_distance) / w.env.desc.resolution;
    solver.microstep_ratio = w.params.microstep_distance / w.env.desc.resolution;
    solver.max_resolver_iterations = w.params.max_resolver_iterations;
    solver.resolver_decay_iterations = w.params.resolver_decay_iterations;
    solver.resolver_initial_scale = w.params.resolver_initial_scale;
    solver.resolver_decay_rate = w.params.resolver_decay_rate;
    solver.resolver_damping = w.params.resolver_damping;
    const double sim_time = (w.params.num_steps + 0.5) * w.params.controller_interval; /* (uint32)(time * frequency) == num_steps */
    const double shortcut = w.params.simulation_shortcut_distance;
    const std::vector<Config> starts = to_configs<N>(w.starts, w.n_starts), targets = to_configs<N>(w.targets, w.n_targets);
    std::vector<std::mt19937_64> rngs(1, std::mt19937_64(20261020));
    StubRobot stub;
    const size_t P = (size_t)w.particles_per_start;
    double checksum = 0.0;
    size_t clusters_total = 0, particles_out = 0;
    std::chrono::steady_clock::time_point t0;
    upc_stats before{}, after{};
    if (mode == "edges")
    {
        std::string call_name;
        std::shared_ptr<Simulator> one;
        std::shared_ptr<ShardedSimulator<StubRobot, Config, std::mt19937_64>> many;
        std::vector<int> devs;
        for (int d = 0; d < gpus; d++) devs.push_back(d);
        if (gpus > 1) many.reset(new ShardedSimulator<StubRobot, Config, std::mt19937_64>(w.env, w.robot, devs, solver));
        else one.reset(new Simulator(w.env, w.robot, 0, solver, 0));
        upc_handle* h0 = many ? many->simulator(0).handle() : one->handle();
        B200OutcomeClustering<StubRobot, Config> clustering(h0, CONVEX_REGION_SIGNATURE, w.signature_threshold, w.distance_threshold);
        /* one GPU: two batches in flight (SubmitEdgeBatch k + 1, then CollectEdgeBatch k); several GPUs: one blocking call per batch */
        const bool pipelined = !many && std::getenv("UPC_PLUGIN_BENCH_ONE_CALL") == nullptr;
        call_name = pipelined ? "SubmitEdgeBatch / CollectEdgeBatch (two batches in flight)" : "SimulateAndClusterEdges";
        for (int k = 0; k < warmup + steps; k++)
        {
            if (k == warmup)
            {
                check(upc_get_stats(h0, &before));
                t0 = std::chrono::steady_clock::now();
                if (pipelined) one->SubmitEdgeBatch(stub, starts, targets, P, rngs, sim_time, shortcut, false, true, k % 2);
            }
            std::vector<std::vector<std::vector<std::pair<Config, bool>>>> result;
            if (pipelined && k >= warmup)
            {
                if (k + 1 < warmup + steps) one->SubmitEdgeBatch(stub, starts, targets, P, rngs, sim_time, shortcut, false, true, (k + 1) % 2);
                result = one->CollectEdgeBatch(clustering, k % 2);
            }
            else
            {
                result = many ? many->SimulateAndClusterEdges(stub, starts, targets, P, rngs, sim_time, shortcut, false, true, clustering)
                              : one->SimulateAndClusterEdges(stub, starts, targets, P, rngs, sim_time, shortcut, false, true, clustering);
            }
            if (k == warmup + steps - 1)
                for (const auto& edge : result)
                {
                    clusters_total += edge.size();
                    for (const auto& c : edge) { particles_out += c.size(); if (!c.empty()) checksum += c[0].first[0]; }
                }
        }
        const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        check(upc_get_stats(h0, &after));
        const double particle_steps = (double)w.n_starts * (double)P * (double)w.params.num_steps * steps;
        std::printf("{\"mode\": \"edges\", \"call\": \"%s::%s\", \"gpus\": %d, \"ms_per_step\": %.6f, \"value\": %.6e, \"edges\": %lld, \"particles_per_edge\": %zu, "
                    "\"clusters\": %zu, \"particles_returned\": %zu, \"h2d_bytes_per_step\": %.1f, \"d2h_bytes_per_step\": %.1f, \"checksum\": %.17g}\n",
                    many ? "ShardedSimulator" : "B200Simulator", call_name.c_str(), gpus, 1e3 * seconds / steps, particle_steps / seconds, (long long)w.n_starts, P, clusters_total, particles_out,
                    (double)(after.h2d_bytes - before.h2d_bytes) / steps, (double)(after.d2h_bytes - before.d2h_bytes) / steps, checksum);
        return 0;
    }
    /* one edge: the two calls the planner makes per tree extension */
    Simulator simulator(w.env, w.robot, 0, solver, 0);
    B200OutcomeClustering<StubRobot, Config> clustering(simulator.handle(), CONVEX_REGION_SIGNATURE, w.signature_threshold, w.distance_threshold);
    const size_t n = (size_t)w.n_starts * P;
    std::vector<Config> particles;
    particles.reserve(n);
    for (int64_t e = 0; e < w.n_starts; e++)
        for (size_t k = 0; k < P; k++) particles.push_back(starts[(size_t)e]);
    double sim_seconds = 0.0;
    for (int k = 0; k < warmup + steps; k++)
    {
        if (k == warmup)
        {
            check(upc_get_stats(simulator.handle(), &before));
            t0 = std::chrono::steady_clock::now();
        }
        const auto ta = std::chrono::steady_clock::now();
        const auto results = simulator.ForwardSimulateRobots(stub, particles, targets, rngs, sim_time, shortcut, false, true);
        if (k >= warmup) sim_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - ta).count();
        const auto clusters = clustering.ClusterParticles(results, true);
        if (k == warmup + steps - 1)
        {
            clusters_total = clusters.size();
            for (const auto& c : clusters) { particles_out += c.size(); if (!c.empty()) checksum += c[0].first[0]; }
        }
    }
    const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    check(upc_get_stats(simulator.handle(), &after));
    const double particle_steps = (double)n * (double)w.params.num_steps * steps;
    std::printf("{\"mode\": \"single\", \"call\": \"B200Simulator::ForwardSimulateRobots + B200OutcomeClustering::ClusterParticles\", \"gpus\": 1, \"ms_per_step\": %.6f, "
                "\"sim_ms_per_step\": %.6f, \"value\": %.6e, \"particles\": %zu, \"clusters\": %zu, \"particles_returned\": %zu, \"h2d_bytes_per_step\": %.1f, "
                "\"d2h_bytes_per_step\": %.1f, \"checksum\": %.17g}\n",
                1e3 * seconds / steps, 1e3 * sim_seconds / steps, particle_steps / seconds, n, clusters_total, particles_out,
                (double)(after.h2d_bytes - before.h2d_bytes) / steps, (double)(after.d2h_bytes - before.d2h_bytes) / steps, checksum);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 5)
    {
        std::fprintf(stderr, "usage: plugin_bench <workload file> edges|single <steps> <warmup> [gpus]\n");
        return 2;
    }
    const Workload w = load(argv[1]);
    const std::string mode = argv[2];
    const int steps = std::atoi(argv[3]), warmup = std::atoi(argv[4]), gpus = (argc > 5) ? std::atoi(argv[5]) : 1;
    try
    {
        switch (w.width)
        {
            case 3: return run<3>(w, mode, steps, warmup, gpus);
            case 12: return run<12>(w, mode, steps, warmup, gpus);
            case 7: return run<7>(w, mode, steps, warmup, gpus);
            default: std::fprintf(stderr, "unsupported configuration width %lld\n", (long long)w.width); return 2;
        }
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "plugin_bench failed: %s\n", e.what());
        return 1;
    }
}

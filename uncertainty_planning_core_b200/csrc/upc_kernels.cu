/*
 * upc_kernels.cu - dispatch of the simulation kernel (model x lanes x noise source; the kernel itself is
 * upc_sim_launch.cuh, instantiated per model in upc_sim_*.cu) and the small per-configuration kernels:
 * CheckConfigCollision (simple_simulator_interface.hpp:617), distance-to-target counts
 * (uncertainty_contact_planning.hpp:2073-2090) and the nearest tree state (:1488-1542).
 */
#include <cstdlib>
#include <cstring>
#include "upc_internal.h"
#include "upc_sim.cuh"
#include "upc_sim_args.h"

namespace upc
{
/* Built-in choice of lanes per particle.  The per-particle program is a long dependent FP64 chain, so a batch is
 * fastest when all of it is resident at once: take the most lanes (shortest per-lane point loop) for which
 * n * lanes threads still fit in one wave (sm_count * UPC_SIM_MIN_BLOCKS CTAs); big batches get one lane per
 * particle, which is also the most instruction-efficient mapping (no replicated state math). */
int choose_lanes(const upc_handle* h, int64_t n)
{
    int lanes;
    if (h->lanes_override > 0)
    {
        lanes = h->lanes_override;
    }
    else
    {
        const int64_t one_wave_warps = (int64_t)h->sm_count * UPC_SIM_MIN_BLOCKS * (kBlock / 32);
        int max_pts = 1;
        for (int l = 0; l < h->robot.num_links; l++)
        {
            const int c = h->robot.link_off[l + 1] - h->robot.link_off[l];
            if (c > max_pts) max_pts = c;
        }
        /* a warp holds 32 / lanes particles.  3, 5 and 6 lanes (idle lanes at the end of every warp) can be requested through
         * upc_set_lanes_per_particle but are not chosen here: on C2 five lanes fit one wave and shorten the per-lane point loops
         * by a fifth, yet the 33 % extra warps all repeat the replicated state math and the kernel gets slower (2.43 -> 2.56 ms) */
        const int candidates[] = {1, 2, 4, 8, 16, 32};
        lanes = 1;
        for (int c : candidates)
        {
            const int64_t per_warp = 32 / c;
            const int64_t warps = (n + per_warp - 1) / per_warp;
            if (c <= max_pts && warps <= one_wave_warps) lanes = c;
        }
    }
    return lanes;
}

/* tape == nullptr selects the seeded kernels (sp.seed / sp.first_particle) */
/*
 * Heavy-first schedule of an edge batch.  Edges that start or end near an obstacle are the ones whose particles go through contact
 * resolution - several times the work of a free-space edge - and CTAs are scheduled in launch order, so a contact-heavy edge at the
 * end of the launch drains the machine at its own pace (a 512-edge shard of C5: 11.9 ms in natural order, 10.0 ms with the edges
 * sorted by the clearance of their end points; 4,096 edges: 77.5 -> 75 ms).  Key = the smaller SDF value at the body origin of the
 * edge's start and target (nearest voxel), in cells, 64 buckets; one CTA: histogram, scan, scatter.  The order inside a bucket is
 * whatever the atomics give - it only affects WHEN an edge is simulated, never its result.
 */
__global__ void __launch_bounds__(1024) edge_order_kernel(const DevEnv env, int position_offset, int64_t E, const double* __restrict__ starts,
                                                          const double* __restrict__ targets, int32_t* __restrict__ order)
{
    __shared__ int s_count[64];
    __shared__ int s_off[64];
    if (threadIdx.x < 64) s_count[threadIdx.x] = 0;
    __syncthreads();
    auto bucket = [&](int64_t e) {
        float key = INFINITY;
        for (int end = 0; end < 2; end++)
        {
            const double* c = end ? targets : starts;
            const double px = c[(int64_t)(position_offset + 0) * E + e], py = c[(int64_t)(position_offset + 1) * E + e];
            const double pz = (position_offset == 0) ? 0.0 : c[(int64_t)(position_offset + 2) * E + e];
            double u[3];
            xf_point(env.G, px, py, pz, u);
            int ix = (int)floor(u[0] * env.inv_res), iy = (int)floor(u[1] * env.inv_res), iz = (int)floor(u[2] * env.inv_res);
            ix = max(0, min(ix, env.nx - 1)); iy = max(0, min(iy, env.ny - 1)); iz = max(0, min(iz, env.nz - 1));
            key = fminf(key, env.sdf[((int64_t)ix * env.ny + iy) * env.nz + iz]);
        }
        const float cells = key * (float)env.inv_res;
        return (cells > 0.0f) ? min((int)cells, 63) : 0;
    };
    for (int64_t e = threadIdx.x; e < E; e += blockDim.x) atomicAdd(&s_count[bucket(e)], 1);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int run = 0;
        for (int b = 0; b < 64; b++) { s_off[b] = run; run += s_count[b]; }
    }
    __syncthreads();
    for (int64_t e = threadIdx.x; e < E; e += blockDim.x) order[atomicAdd(&s_off[bucket(e)], 1)] = (int32_t)e;
}

cudaError_t launch_forward_simulate(upc_handle* h, const DevSim& sp, int64_t n, const double* starts, int64_t n_targets,
                                    const double* targets, const double* tape, double* out_configs, uint8_t* out_collided,
                                    cudaStream_t stream)
{
    SimArgs a;
    a.env = h->env;
    a.robot = h->robot;
    a.sp = sp;
    a.n = n;
    a.n_targets = n_targets;
    a.starts = starts;
    a.targets = targets;
    a.tape = tape;
    a.out_configs = out_configs;
    a.out_collided = out_collided;
    a.counters = h->d_counters;
    a.edge_order = nullptr;
    a.edge_particles = 1;
    const int lanes = choose_lanes(h, n);
    const bool seeded = (tape == nullptr);
    {
        /* edge batch in one launch (one start and one target per block of particles), free-flying robot: heavy edges first
         * (UPC_B200_NO_EDGE_ORDER=1: natural order) */
        static const bool off = []() { const char* e = getenv("UPC_B200_NO_EDGE_ORDER"); return e != nullptr && e[0] == '1'; }();
        const int64_t E = sp.n_starts;
        if (!off && seeded && h->robot.kind != UPC_ROBOT_LINKED && E >= 8 && E == n_targets && sp.per_start > 1 && sp.per_start == sp.per_target &&
            E * sp.per_start == n && E < ((int64_t)1 << 31) && sp.step_begin == 0 && sp.step_end >= sp.num_steps && sp.trace_configs == nullptr)
        {
            if (h->st_order.reserve((size_t)E * sizeof(int32_t)) != cudaSuccess) return cudaErrorMemoryAllocation;
            edge_order_kernel<<<1, 1024, 0, stream>>>(h->env, (h->robot.kind == UPC_ROBOT_SE3) ? 9 : 0, E, starts, targets, (int32_t*)h->st_order.ptr);
            h->stats.kernel_launches++;
            a.edge_order = (const int32_t*)h->st_order.ptr;
            a.edge_particles = sp.per_start;
        }
    }
    switch (h->robot.kind)
    {
        case UPC_ROBOT_SE2: return launch_sim_se2(a, lanes, seeded, stream);
        case UPC_ROBOT_SE3: return launch_sim_se3(a, lanes, seeded, stream);
        default:
            /* most arms fit the small instantiation (dof <= 8, <= 9 links): half-size normal equations and frame buffers */
            if (h->robot.dof <= LinkedModelSmall::DOF && h->robot.num_links <= LinkedModelSmall::LINKS) return launch_sim_linked_small(a, lanes, seeded, stream);
            return launch_sim_linked(a, lanes, seeded, stream);
    }
}

/* ------------------------------------------------------------------ distance-to-target count (reverse-edge / goal probabilities) */

struct CountArgs
{
    DevRobot robot;
    int64_t n, n_targets;
    const double* configs;
    const double* targets;
    double threshold;
    unsigned long long* count;
    uint8_t* within;
};

/* uncertainty_contact_planning.hpp:2079-2088: particle_distance <= goal_distance_threshold_ */
template <class M>
__global__ void __launch_bounds__(256) count_within_kernel(const __grid_constant__ CountArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (i < a.n)
    {
        const int width = (M::HAS_FRAMES) ? a.robot.dof : M::WIDTH;
        double c[M::WIDTH], t[M::WIDTH];
        const int64_t ti = (a.n_targets == 1) ? 0 : i;
        for (int k = 0; k < width; k++)
        {
            c[k] = a.configs[(int64_t)k * a.n + i];
            t[k] = a.targets[(int64_t)k * a.n_targets + ti];
        }
        ok = M::distance(a.robot, c, t) <= a.threshold;
        if (a.within) a.within[i] = ok ? 1 : 0;
    }
    const unsigned votes = __popc(__ballot_sync(0xffffffffu, ok));
    if ((threadIdx.x & 31) == 0 && votes) atomicAdd(a.count, (unsigned long long)votes);
}

cudaError_t launch_count_within(upc_handle* h, int64_t n, const double* d_configs, int64_t n_targets, const double* d_targets, double threshold,
                                unsigned long long* d_count, uint8_t* d_within, cudaStream_t stream)
{
    CountArgs a;
    a.robot = h->robot; a.n = n; a.n_targets = n_targets; a.configs = d_configs; a.targets = d_targets; a.threshold = threshold;
    a.count = d_count; a.within = d_within;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    switch (h->robot.kind)
    {
        case UPC_ROBOT_SE2: count_within_kernel<Se2Model><<<blocks, 256, 0, stream>>>(a); break;
        case UPC_ROBOT_SE3: count_within_kernel<Se3Model><<<blocks, 256, 0, stream>>>(a); break;
        default: count_within_kernel<LinkedModel><<<blocks, 256, 0, stream>>>(a); break;
    }
    return cudaGetLastError();
}

/* ------------------------------------------------------------------ nearest tree state */

struct NearestArgs
{
    DevRobot robot;
    int64_t n_states, n_queries;
    const double* expectations; /* [C][n_states] */
    const double* fw;
    const double* vw;
    const uint8_t* use;
    const double* queries;      /* [C][n_queries] */
    double step_size;
    long long* out_index;
    double* out_distance;
};

/* uncertainty_contact_planning.hpp:1488-1542: one CTA per query, threads stride over the states in ascending order
 * (strict < keeps the first = smallest index per thread), then a lexicographic (distance, index) minimum over the CTA */
template <class M>
__global__ void __launch_bounds__(256) nearest_state_kernel(const __grid_constant__ NearestArgs a)
{
    __shared__ double s_d[8];
    __shared__ long long s_i[8];
    const int64_t q = blockIdx.x;
    const int width = (M::HAS_FRAMES) ? a.robot.dof : M::WIDTH;
    double qc[M::WIDTH];
    for (int k = 0; k < width; k++) qc[k] = a.queries[(int64_t)k * a.n_queries + q];
    double best = INFINITY;
    long long best_i = -1;
    for (int64_t i = threadIdx.x; i < a.n_states; i += blockDim.x)
    {
        if (a.use != nullptr && a.use[i] == 0) continue;
        double e[M::WIDTH];
        for (int k = 0; k < width; k++) e[k] = a.expectations[(int64_t)k * a.n_states + i];
        const double ed = M::distance(a.robot, e, qc) / a.step_size;
        const double d = (a.fw[i] * ed) * a.vw[i];
        if (d < best)
        {
            best = d;
            best_i = i;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
    {
        const double od = __shfl_xor_sync(0xffffffffu, best, off);
        const long long oi = __shfl_xor_sync(0xffffffffu, best_i, off);
        if (oi >= 0 && (best_i < 0 || od < best || (od == best && oi < best_i)))
        {
            best = od;
            best_i = oi;
        }
    }
    if ((threadIdx.x & 31) == 0)
    {
        s_d[threadIdx.x >> 5] = best;
        s_i[threadIdx.x >> 5] = best_i;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++)
        {
            const double od = s_d[w];
            const long long oi = s_i[w];
            if (oi >= 0 && (best_i < 0 || od < best || (od == best && oi < best_i)))
            {
                best = od;
                best_i = oi;
            }
        }
        a.out_index[q] = best_i;
        if (a.out_distance) a.out_distance[q] = (best_i < 0) ? INFINITY : best;
    }
}

cudaError_t launch_nearest_states(upc_handle* h, int64_t n_states, const double* d_expectations, const double* d_fw, const double* d_vw,
                                  const uint8_t* d_use, int64_t n_queries, const double* d_queries, double step_size, long long* d_index,
                                  double* d_distance, cudaStream_t stream)
{
    NearestArgs a;
    a.robot = h->robot; a.n_states = n_states; a.n_queries = n_queries; a.expectations = d_expectations; a.fw = d_fw; a.vw = d_vw; a.use = d_use;
    a.queries = d_queries; a.step_size = step_size; a.out_index = d_index; a.out_distance = d_distance;
    const unsigned blocks = (unsigned)n_queries;
    switch (h->robot.kind)
    {
        case UPC_ROBOT_SE2: nearest_state_kernel<Se2Model><<<blocks, 256, 0, stream>>>(a); break;
        case UPC_ROBOT_SE3: nearest_state_kernel<Se3Model><<<blocks, 256, 0, stream>>>(a); break;
        default: nearest_state_kernel<LinkedModel><<<blocks, 256, 0, stream>>>(a); break;
    }
    return cudaGetLastError();
}

/* ------------------------------------------------------------------ CheckConfigCollision (single configuration) */

struct CheckArgs
{
    DevEnv env;
    DevRobot robot;
    const double* config;
    double threshold;
    int* out;
};

template <class M>
__global__ void __launch_bounds__(32) check_collision_kernel(const __grid_constant__ CheckArgs a)
{
    extern __shared__ double smem[];
    double* frames = smem;
    Group<32> g;
    g.sub = (int)threadIdx.x;
    g.base = 0;
    g.mask = 0xffffffffu;
    M m;
    if constexpr (M::HAS_FRAMES) m.frames = frames;
    m.load(a.robot, a.config, 1, 0);
    refresh_frames(a.robot, m, g);
    const bool hit = check_collision<M, 32>(a.env, a.robot, m, g, a.robot.pts, a.threshold);
    if (threadIdx.x == 0) *a.out = hit ? 1 : 0;
}

cudaError_t launch_check_collision(upc_handle* h, const double* d_config, double threshold, int* d_out, cudaStream_t stream)
{
    CheckArgs a;
    a.env = h->env;
    a.robot = h->robot;
    a.config = d_config;
    a.threshold = threshold;
    a.out = d_out;
    const size_t smem = (size_t)h->robot.num_links * 12 * sizeof(double);
    switch (h->robot.kind)
    {
        case UPC_ROBOT_SE2: check_collision_kernel<Se2Model><<<1, 32, smem, stream>>>(a); break;
        case UPC_ROBOT_SE3: check_collision_kernel<Se3Model><<<1, 32, smem, stream>>>(a); break;
        default: check_collision_kernel<LinkedModel><<<1, 32, smem, stream>>>(a); break;
    }
    return cudaGetLastError();
}

} /* namespace upc */

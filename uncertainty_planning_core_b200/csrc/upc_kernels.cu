/*
 * upc_kernels.cu - forward-simulation + SDF-collision kernels for sm_100a.
 *
 * One launch advances every particle of a batch along its planner edge: the GPU replacement for the body of
 * SimulatorInterface::ForwardSimulateRobots (simple_simulator_interface.hpp:623).
 *
 * Mapping: G lanes (1..32, a power of two) cooperate on one particle, 128 threads per CTA.  Body points live
 * in shared memory; the SDF is read through the read-only path and stays L2-resident (8 MiB for 128^3,
 * 64 MiB for 256^3 against 126 MB of L2); the noise tape is streamed once from HBM with particle-minor
 * (coalesced) layout.  FP64 throughout, -fmad=false, explicit fma only where DESIGN.md says so.
 */
#include <cstdlib>
#include <cstring>
#include "upc_internal.h"
#include "upc_sim.cuh"

namespace upc
{
#ifndef UPC_SIM_BLOCK
#define UPC_SIM_BLOCK 128
#endif
constexpr int kBlock = UPC_SIM_BLOCK;
/* 3 CTAs of 128 threads per SM (<= 168 registers): measured best on B200 for the latency-bound per-particle program */
#ifndef UPC_SIM_MIN_BLOCKS_SMALL_LINKED
#define UPC_SIM_MIN_BLOCKS_SMALL_LINKED 3
#endif
#ifndef UPC_SIM_MIN_BLOCKS_ONE_LANE
#define UPC_SIM_MIN_BLOCKS_ONE_LANE 2
#endif
#ifndef UPC_SIM_MIN_BLOCKS
#define UPC_SIM_MIN_BLOCKS 3
#endif
#ifndef UPC_SIM_MIN_BLOCKS_SE2_ONE_LANE
#define UPC_SIM_MIN_BLOCKS_SE2_ONE_LANE 4
#endif

struct SimArgs
{
    DevEnv env;
    DevRobot robot;
    DevSim sp;
    int64_t n, n_targets;
    const double* starts;
    const double* targets;
    const double* tape;
    double* out_configs;
    uint8_t* out_collided;
    DevCounters* counters;
};

/* per-thread link-frame storage (two sets) of robots whose frames do not live in shared memory */
template <class M, bool IN_SHARED>
__host__ __device__ constexpr int frame_buffer_doubles()
{
    if constexpr (M::HAS_FRAMES && !IN_SHARED) return 2 * M::LINKS * 12;
    else return 2;
}

/* CTAs per SM the register budget is cut for: three (168 registers) everywhere except the one-lane SE(2) program (four: with
 * more lanes per particle - the latency-bound small batches - SE(2) also prefers three, C1 -6 % with four) and the one-lane program of big linked
 * robots (dof > 8 or > 9 links), whose 8.7 KB of link frames and call frames per thread favour 255 registers and two CTAs
 * (fewer spills, smaller resident local-memory footprint).  The one-lane SE(3) program also ran with two CTAs while its point
 * loops kept 4 - 8 lookups in flight; with the leaner loops (upc_sim.cuh, kCheckInFlight) three win: 10^6 SE(3) particles
 * 2.09e9 -> 2.27e9 particle-steps/s. */
template <class M, int G>
constexpr int sim_min_blocks()
{
    if (G == 1 && M::HAS_FRAMES && M::DOF > 8) return UPC_SIM_MIN_BLOCKS_ONE_LANE;
    if (G == 1 && !M::HAS_FRAMES && M::DOF == 3) return UPC_SIM_MIN_BLOCKS_SE2_ONE_LANE; /* small program: 128 registers, four CTAs (C5 +3 %) */
    if (G == 1 && M::HAS_FRAMES) return UPC_SIM_MIN_BLOCKS_SMALL_LINKED; /* 3.9 KB of stack per thread (C3 28.5 -> 25.9 ms with three) */
    return UPC_SIM_MIN_BLOCKS;
}

template <class M, int G>
__global__ void __launch_bounds__(kBlock, sim_min_blocks<M, G>()) forward_simulate_kernel(const __grid_constant__ SimArgs a)
{
    extern __shared__ double smem[];
    __shared__ unsigned int s_cnt[4];
    double* pts = smem;
    const int npd = a.robot.num_points * 6; /* 4 doubles + 4 floats per point */
    for (int k = threadIdx.x; k < npd; k += kBlock) pts[k] = a.robot.pts[k];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0u;
    __syncthreads();

    /* G lanes per particle, groups never straddle a warp: a warp holds 32 / G particles and, when G does not divide 32, leaves
     * its last 32 % G lanes idle */
    constexpr int kGroupsPerWarp = 32 / G;
    constexpr int kParticlesPerBlock = (kBlock / 32) * kGroupsPerWarp;
    const int lane = (int)(threadIdx.x & 31);
    const int group_in_warp = lane / G;
    const bool lane_used = group_in_warp < kGroupsPerWarp;
    const int particle_in_block = (int)(threadIdx.x >> 5) * kGroupsPerWarp + group_in_warp;
    const int64_t i = lane_used ? ((int64_t)blockIdx.x * kParticlesPerBlock + particle_in_block) : a.n;
    LocalCounters lc = {};
    bool collided = false;
    Group<G> g;
    g.sub = lane % G;
    g.base = lane - g.sub;
    g.mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << g.base);
    if (i < a.n)
    {
        double* frames0 = nullptr;
        double* frames1 = nullptr;
        /* per-particle scratch: [target | previous | PID state] then (linked robots, G >= 4) two sets of link frames.
         * With 1 or 2 lanes per particle - the throughput mapping for big batches - the frames of a linked robot live in
         * per-thread local memory instead: shared memory could not hold them for 64-128 particles per CTA. */
        constexpr bool kFramesInShared = M::HAS_FRAMES && (G >= 4);
        const int per = kFramesInShared ? a.robot.num_links * 12 : 0;
        double* park = smem + npd + (size_t)particle_in_block * (size_t)(2 * M::WIDTH + 2 * M::DOF + 2 * per);
        double local_frames[frame_buffer_doubles<M, kFramesInShared>()];
        if constexpr (kFramesInShared)
        {
            frames0 = park + 2 * M::WIDTH + 2 * M::DOF;
            frames1 = frames0 + per;
        }
        else if constexpr (M::HAS_FRAMES)
        {
            frames0 = local_frames;
            frames1 = local_frames + frame_buffer_doubles<M, kFramesInShared>() / 2;
        }
        simulate_particle<M, G>(a.env, a.robot, a.sp, g, i, a.n, a.starts, a.n_targets, a.targets, a.tape, a.out_configs,
                                a.out_collided, pts, frames0, frames1, park, lc, collided);
        if (g.sub == 0)
        {
            atomicAdd(&s_cnt[0], lc.particle_steps);
            atomicAdd(&s_cnt[1], collided ? 1u : 0u);
            atomicAdd(&s_cnt[2], lc.resolver_iterations);
            atomicAdd(&s_cnt[3], lc.failed_resolves);
#ifdef UPC_PHASE_TIMING
            for (int k = 0; k < 12; k++) atomicAdd((unsigned long long*)&a.counters->phase[k], (unsigned long long)lc.cyc[k]);
#endif
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        atomicAdd(&a.counters->particle_steps, (unsigned long long)s_cnt[0]);
        atomicAdd(&a.counters->collided, (unsigned long long)s_cnt[1]);
        atomicAdd(&a.counters->resolver_iterations, (unsigned long long)s_cnt[2]);
        atomicAdd(&a.counters->failed_resolves, (unsigned long long)s_cnt[3]);
    }
}

template <class M, int G>
static cudaError_t launch_sim_impl(const SimArgs& a, cudaStream_t stream)
{
    constexpr int particles_per_block = (kBlock / 32) * (32 / G);
    const int64_t blocks = (a.n + particles_per_block - 1) / particles_per_block;
    size_t smem = (size_t)a.robot.num_points * 6 * sizeof(double);
    smem += (size_t)particles_per_block * (2 * M::WIDTH + 2 * M::DOF + ((M::HAS_FRAMES && G >= 4) ? 2 * (size_t)a.robot.num_links * 12 : 0)) * sizeof(double);
    if (smem + 256 > 48 * 1024) /* the kernel also has a few bytes of static shared memory */
    {
        cudaError_t e = cudaFuncSetAttribute(forward_simulate_kernel<M, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    /* L1 / shared-memory split: ask for no more shared memory than the resident CTAs use.  Left alone the driver configures
     * 100 KB for this kernel (ncu: launch__shared_mem_config_size) although three CTAs need 50 KB, and the program lives on L1:
     * SDF lookups (94 % hits) and ~ 1.8 KB of local memory per thread (stores hit only 38 %).  UPC_B200_SMEM_CARVEOUT=default
     * switches the hint off. */
    {
        static int configured[64]; /* per device, 0 = not set yet (stored as pct + 2) */
        static const bool use_default = []() { const char* e = std::getenv("UPC_B200_SMEM_CARVEOUT"); return e != nullptr && std::strcmp(e, "default") == 0; }();
        int dev = 0, per_sm = 233472;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
        if (!use_default) cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        const size_t need = (size_t)sim_min_blocks<M, G>() * (smem + 1024 + 64);
        int pct = use_default ? -1 : (int)((need * 100 + (size_t)per_sm - 1) / (size_t)per_sm);
        if (pct > 100) pct = 100;
        if (pct + 2 != configured[dev])
        {
            cudaError_t e = cudaFuncSetAttribute(forward_simulate_kernel<M, G>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            if (e != cudaSuccess) return e;
            configured[dev] = pct + 2;
        }
    }
    forward_simulate_kernel<M, G><<<(unsigned)blocks, kBlock, smem, stream>>>(a);
    return cudaGetLastError();
}

template <class M>
static cudaError_t launch_sim_lanes(const SimArgs& a, int lanes, cudaStream_t stream)
{
    switch (lanes)
    {
        case 1: return launch_sim_impl<M, 1>(a, stream);
        case 2: return launch_sim_impl<M, 2>(a, stream);
        case 3: return launch_sim_impl<M, 3>(a, stream);
        case 4: return launch_sim_impl<M, 4>(a, stream);
        case 5: return launch_sim_impl<M, 5>(a, stream);
        case 6: return launch_sim_impl<M, 6>(a, stream);
        case 8: return launch_sim_impl<M, 8>(a, stream);
        case 16: return launch_sim_impl<M, 16>(a, stream);
        default: return launch_sim_impl<M, 32>(a, stream);
    }
}

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
    const int lanes = choose_lanes(h, n);
    switch (h->robot.kind)
    {
        case UPC_ROBOT_SE2: return launch_sim_lanes<Se2Model>(a, lanes, stream);
        case UPC_ROBOT_SE3: return launch_sim_lanes<Se3Model>(a, lanes, stream);
        default:
            /* most arms fit the small instantiation (dof <= 8, <= 9 links): half-size normal equations and frame buffers */
            if (h->robot.dof <= LinkedModelSmall::DOF && h->robot.num_links <= LinkedModelSmall::LINKS) return launch_sim_lanes<LinkedModelSmall>(a, lanes, stream);
            return launch_sim_lanes<LinkedModel>(a, lanes, stream);
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

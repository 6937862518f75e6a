/*
 * upc_sim_launch.cuh - the forward-simulation + SDF-collision kernel for sm_100a and its launcher, as templates over the
 * robot model, the lanes per particle and the noise source (tape replay / counter-based generator).
 *
 * One launch advances every particle of a batch along its planner edge: the GPU replacement for the body of
 * SimulatorInterface::ForwardSimulateRobots (simple_simulator_interface.hpp:623).
 *
 * Mapping: G lanes (1..32) cooperate on one particle, 128 threads per CTA.  Body points live in shared memory; the SDF is
 * read through the read-only path and stays L2-resident (8 MiB for 128^3, 64 MiB for 256^3 against 126 MB of L2); on tape
 * launches the noise tape is streamed once from HBM with particle-minor (coalesced) layout, on seeded launches the draws
 * are computed in registers (upc_noise.cuh).  FP64 throughout, -fmad=false, explicit fma only where DESIGN.md says so.
 */
#ifndef UPC_SIM_LAUNCH_CUH
#define UPC_SIM_LAUNCH_CUH

#include <cstdlib>
#include <cstring>
#include "upc_sim_args.h"
#include "upc_sim.cuh"
#include "upc_host_prep.h" /* kPointDoubles */

namespace upc
{
#ifndef UPC_SIM_MIN_BLOCKS_SMALL_LINKED
#define UPC_SIM_MIN_BLOCKS_SMALL_LINKED 3
#endif
#ifndef UPC_SIM_MIN_BLOCKS_ONE_LANE
#define UPC_SIM_MIN_BLOCKS_ONE_LANE 2
#endif
#ifndef UPC_SIM_MIN_BLOCKS_SE2_ONE_LANE
#define UPC_SIM_MIN_BLOCKS_SE2_ONE_LANE 2
#endif

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
    if (G == 1 && !M::HAS_FRAMES && M::DOF == 3) return UPC_SIM_MIN_BLOCKS_SE2_ONE_LANE; /* small program: 128 registers, 16 warps per SM (C5 +3 % against 168 / 12) */
    if (G == 1 && M::HAS_FRAMES) return UPC_SIM_MIN_BLOCKS_SMALL_LINKED; /* 3.9 KB of stack per thread (C3 28.5 -> 25.9 ms with three) */
    return UPC_SIM_MIN_BLOCKS;
}

/* threads per CTA: 128, except the one-lane SE(2) program (the planner-batch workhorse): 256 threads x 2 CTAs per SM - the same 16
 * warps and 128 registers as 128 x 4, but with the per-interval barrier (UPC_STEP_BARRIER) eight warps instead of four run through
 * the same stretch of the program together: C5 21.7 -> 20.9 ms (512 x 1: 23.8 ms) */
#ifndef UPC_SIM_BLOCK_SE2_ONE_LANE
#define UPC_SIM_BLOCK_SE2_ONE_LANE 256
#endif
template <class M, int G>
__host__ __device__ constexpr int sim_block()
{
    if (G == 1 && !M::HAS_FRAMES && M::DOF == 3) return UPC_SIM_BLOCK_SE2_ONE_LANE;
    return UPC_SIM_BLOCK;
}

template <class M, int G, bool SEEDED>
__global__ void __launch_bounds__((sim_block<M, G>()), (sim_min_blocks<M, G>())) forward_simulate_kernel(const __grid_constant__ SimArgs a)
{
    constexpr int kBlock = sim_block<M, G>();
    extern __shared__ double smem[];
    __shared__ unsigned int s_cnt[4];
    double* pts = smem;
    const int npd = a.robot.num_points * kPointDoubles; /* 4 doubles + 8 floats per point */
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
    const int particle_slot = (int)(threadIdx.x >> 5) * kGroupsPerWarp + group_in_warp;
    const int64_t i_raw = (int64_t)blockIdx.x * kParticlesPerBlock + particle_slot;
    /* lanes without a particle (tail of the batch, lanes the group size leaves over) walk the program on particle 0 with nothing to
     * do, so that every thread of the CTA meets the same per-interval barrier (simulate_particle: exists); their scratch is the
     * spare slot behind the CTA's particles */
    const bool exists = lane_used && (i_raw < a.n);
    int64_t i = exists ? i_raw : 0;
    if (a.edge_order != nullptr && exists)
    {
        /* heavy-first schedule of an edge batch: this slot's edge is edge_order[slot / P] */
        const int64_t slot_edge = i_raw / a.edge_particles;
        i = (int64_t)a.edge_order[slot_edge] * a.edge_particles + (i_raw - slot_edge * a.edge_particles);
    }
    const int particle_in_block = exists ? particle_slot : kParticlesPerBlock;
    LocalCounters lc = {};
    bool collided = false;
    Group<G> g;
    g.sub = lane % G;
    g.base = lane - g.sub;
    g.mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << g.base);
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
        /* per-thread words of the shared pre-cull masks, behind the scratch of the CTA's particles (and the spare slot) */
        unsigned* masks = nullptr;
        if constexpr (shares_cull_masks<M, G>())
            masks = reinterpret_cast<unsigned*>(smem + npd + (size_t)(kParticlesPerBlock + 1) * (size_t)(2 * M::WIDTH + 2 * M::DOF + 2 * per)) + threadIdx.x;
        simulate_particle<M, G, SEEDED>(a.env, a.robot, a.sp, g, i, a.n, a.starts, a.n_targets, a.targets, a.tape, a.out_configs,
                                a.out_collided, pts, frames0, frames1, park, lc, collided, exists, masks, kBlock);
        if (g.sub == 0 && exists)
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

template <class M, int G, bool SEEDED>
static cudaError_t launch_sim_impl(const SimArgs& a, cudaStream_t stream)
{
    constexpr int kBlock = sim_block<M, G>();
    constexpr int particles_per_block = (kBlock / 32) * (32 / G);
    const int64_t blocks = (a.n + particles_per_block - 1) / particles_per_block;
    size_t smem = (size_t)a.robot.num_points * kPointDoubles * sizeof(double);
    smem += (size_t)(particles_per_block + 1) * (2 * M::WIDTH + 2 * M::DOF + ((M::HAS_FRAMES && G >= 4) ? 2 * (size_t)a.robot.num_links * 12 : 0)) * sizeof(double);
    if (shares_cull_masks<M, G>()) smem += (size_t)kBlock * kMaxMaskChunks * sizeof(unsigned); /* shared pre-cull masks, kMaxMaskChunks words per thread */
    if (smem + 256 > 48 * 1024) /* the kernel also has a few bytes of static shared memory */
    {
        cudaError_t e = cudaFuncSetAttribute(forward_simulate_kernel<M, G, SEEDED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
            cudaError_t e = cudaFuncSetAttribute(forward_simulate_kernel<M, G, SEEDED>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            if (e != cudaSuccess) return e;
            configured[dev] = pct + 2;
        }
    }
    forward_simulate_kernel<M, G, SEEDED><<<(unsigned)blocks, kBlock, smem, stream>>>(a);
    return cudaGetLastError();
}

template <class M, bool SEEDED>
static cudaError_t launch_sim_mode(const SimArgs& a, int lanes, cudaStream_t stream)
{
    switch (lanes)
    {
        case 1: return launch_sim_impl<M, 1, SEEDED>(a, stream);
        case 2: return launch_sim_impl<M, 2, SEEDED>(a, stream);
        case 3: return launch_sim_impl<M, 3, SEEDED>(a, stream);
        case 4: return launch_sim_impl<M, 4, SEEDED>(a, stream);
        case 5: return launch_sim_impl<M, 5, SEEDED>(a, stream);
        case 6: return launch_sim_impl<M, 6, SEEDED>(a, stream);
        case 8: return launch_sim_impl<M, 8, SEEDED>(a, stream);
        case 16: return launch_sim_impl<M, 16, SEEDED>(a, stream);
        default: return launch_sim_impl<M, 32, SEEDED>(a, stream);
    }
}

template <class M>
static cudaError_t launch_sim_lanes(const SimArgs& a, int lanes, bool seeded, cudaStream_t stream)
{
    return seeded ? launch_sim_mode<M, true>(a, lanes, stream) : launch_sim_mode<M, false>(a, lanes, stream);
}

} /* namespace upc */

#endif /* UPC_SIM_LAUNCH_CUH */

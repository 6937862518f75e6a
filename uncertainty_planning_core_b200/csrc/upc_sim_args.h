/*
 * upc_sim_args.h - the argument block of the simulation kernel and the per-model launchers.  The kernel is instantiated
 * per robot model in its own translation unit (upc_sim_se2.cu, upc_sim_se3.cu, upc_sim_linked_small.cu, upc_sim_linked.cu)
 * so that the four compile in parallel; upc_kernels.cu picks the model and the lane count.
 */
#ifndef UPC_SIM_ARGS_H
#define UPC_SIM_ARGS_H

#include <cuda_runtime.h>
#include "upc_device.cuh"

namespace upc
{
#ifndef UPC_SIM_BLOCK
#define UPC_SIM_BLOCK 128
#endif
constexpr int kBlock = UPC_SIM_BLOCK;
/* 3 CTAs of 128 threads per SM (<= 168 registers): measured best on B200 for the latency-bound per-particle program */
#ifndef UPC_SIM_MIN_BLOCKS
#define UPC_SIM_MIN_BLOCKS 3
#endif

struct SimArgs
{
    DevEnv env;
    DevRobot robot;
    DevSim sp;
    int64_t n, n_targets;
    const double* starts;
    const double* targets;
    const double* tape; /* NULL on seeded launches */
    double* out_configs;
    uint8_t* out_collided;
    DevCounters* counters;
    /* edge batches (one start / target per block of edge_particles particles): the order in which the blocks are worked through.
     * Slot j of the launch simulates edge edge_order[j]; results land where they always do.  NULL = natural order. */
    const int32_t* edge_order;
    int64_t edge_particles;
};

/* lanes: 1, 2, 3, 4, 5, 6, 8, 16 or 32 lanes per particle; seeded: draws from the counter-based generator instead of a.tape */
cudaError_t launch_sim_se2(const SimArgs& a, int lanes, bool seeded, cudaStream_t stream);
cudaError_t launch_sim_se3(const SimArgs& a, int lanes, bool seeded, cudaStream_t stream);
cudaError_t launch_sim_linked_small(const SimArgs& a, int lanes, bool seeded, cudaStream_t stream);
cudaError_t launch_sim_linked(const SimArgs& a, int lanes, bool seeded, cudaStream_t stream);
} /* namespace upc */

#endif /* UPC_SIM_ARGS_H */

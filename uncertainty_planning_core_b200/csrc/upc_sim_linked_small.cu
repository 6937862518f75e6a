/* upc_sim_linked_small.cu - forward_simulate_kernel instantiated for LinkedModelSmall (all lane counts, tape and seeded noise) */
#include "upc_sim_launch.cuh"

namespace upc
{
cudaError_t launch_sim_linked_small(const SimArgs& a, int lanes, bool seeded, cudaStream_t stream) { return launch_sim_lanes<LinkedModelSmall>(a, lanes, seeded, stream); }
}

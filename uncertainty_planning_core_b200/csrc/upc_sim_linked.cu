/* upc_sim_linked.cu - forward_simulate_kernel instantiated for LinkedModel (all lane counts, tape and seeded noise) */
#include "upc_sim_launch.cuh"

namespace upc
{
cudaError_t launch_sim_linked(const SimArgs& a, int lanes, bool seeded, cudaStream_t stream) { return launch_sim_lanes<LinkedModel>(a, lanes, seeded, stream); }
}

/* upc_sim_se2.cu - forward_simulate_kernel instantiated for Se2Model (all lane counts, tape and seeded noise) */
#include "upc_sim_launch.cuh"

namespace upc
{
cudaError_t launch_sim_se2(const SimArgs& a, int lanes, bool seeded, cudaStream_t stream) { return launch_sim_lanes<Se2Model>(a, lanes, seeded, stream); }
}

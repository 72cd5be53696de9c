// Instantiates the solve / eval kernels for lane groups of 6 lane(s) per MPC instance (S = 1..4 steps per lane).
#include "kernels.cuh"

namespace neompc {
cudaError_t launch_g6(bool eval, int S, bool ext, const LaunchArgs& a) { return launch_for_g<6>(eval, S, ext, a); }
}  // namespace neompc

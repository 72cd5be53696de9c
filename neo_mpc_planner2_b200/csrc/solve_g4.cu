// Instantiates the solve / eval kernels for lane groups of 4 lane(s) per MPC instance (S = 1..4 steps per lane).
#include "kernels.cuh"

namespace neompc {
cudaError_t launch_g4(bool eval, int S, bool ext, const LaunchArgs& a) { return launch_for_g<4>(eval, S, ext, a); }
}  // namespace neompc

// Instantiates the solve / eval kernels for lane groups of 10 lane(s) per MPC instance (S = 1..4 steps per lane).
#include "kernels.cuh"

namespace neompc {
cudaError_t launch_g10(bool eval, int S, bool ext, const LaunchArgs& a) { return launch_for_g<10>(eval, S, ext, a); }
}  // namespace neompc

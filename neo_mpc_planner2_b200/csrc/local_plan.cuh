// local_plan.cuh — predicted-path output (publishLocalPlan, srv.py:271-310), SURVEY.md §8f row N4.
#pragma once
#include <cuda_runtime.h>

#include "neompc.h"

namespace neompc {

cudaError_t launch_local_plan(const neompc_request* d_reqs, const float* d_plan, unsigned n, int n_steps, double dt,
                              neompc_plan_pose* d_out, cudaStream_t stream);

}  // namespace neompc

// local_plan.cu — the step right after the solve: the predicted path the reference publishes on
// "/mpc_local_plan" (publishLocalPlan, mpc_optimization_server.py:271-310) for a whole batch.  One thread per
// instance, float64, same operation order as the reference (compiled with -fmad=false so products and sums round like
// numpy's): start pose, then per step  yaw += w dt;  x += vx cos(yaw) dt - vy sin(yaw) dt;  y += vx sin(yaw) dt + vy cos(yaw) dt,
// orientation quaternion_from_euler(0, 0, yaw) (srv.py:182-196).
#include "local_plan.cuh"

namespace neompc {

namespace {

__global__ void local_plan_kernel(const neompc_request* __restrict__ reqs, const float* __restrict__ plan, unsigned n,
                                  int n_steps, double dt, neompc_plan_pose* __restrict__ out) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // the reference starts from the TF pose map -> base_link (srv.py:274-286); here: the request's current pose
  double px = (double)reqs[i].pose_x, py = (double)reqs[i].pose_y, yaw = (double)reqs[i].pose_yaw;
  neompc_plan_pose* o = out + (size_t)i * (n_steps + 1);
  const float* x = plan + (size_t)i * 3 * n_steps;
  o[0].x = px; o[0].y = py; o[0].qz = 0.0; o[0].qw = 1.0;            // srv.py:288-291: position only, default orientation
  for (int k = 0; k < n_steps; ++k) {
    const double vx = (double)x[3 * k], vy = (double)x[3 * k + 1], om = (double)x[3 * k + 2];
    yaw += om * dt;                                                  // srv.py:295
    const double c = cos(yaw), s = sin(yaw);
    px += vx * c * dt - vy * s * dt;                                 // srv.py:296
    py += vx * s * dt + vy * c * dt;                                 // srv.py:297
    o[k + 1].x = px; o[k + 1].y = py;
    o[k + 1].qz = sin(yaw * 0.5);                                    // srv.py:183-184,194 (roll = pitch = 0)
    o[k + 1].qw = cos(yaw * 0.5);
  }
}

}  // namespace

cudaError_t launch_local_plan(const neompc_request* d_reqs, const float* d_plan, unsigned n, int n_steps, double dt,
                              neompc_plan_pose* d_out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int block = 128;
  local_plan_kernel<<<(n + block - 1) / block, block, 0, stream>>>(d_reqs, d_plan, n, n_steps, dt, d_out);
  return cudaGetLastError();
}

}  // namespace neompc

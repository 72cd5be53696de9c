// carrot.cu — one warp per robot: closest plan pose, window end, lookahead pick, slow-down hysteresis and the
// Optimizer request, i.e. reference src/NeoMpcPlanner.cpp:66-135 (transformGlobalPlan), :157-171
// (getLookAheadDistance), :173-189 (getLookAheadPoint), :216-246 (hysteresis + request) for a whole fleet.
// Index decisions are made in float64 with the same operation order as the oracle (oracle/carrot_oracle.py);
// this file is compiled with -fmad=false so that products and sums round exactly like numpy's.
#include "carrot.cuh"

namespace neompc {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarpsPerBlock = 4;

__device__ __forceinline__ double dist_to(const double* plan, unsigned i, double rx, double ry) {
  const double dx = plan[3 * (size_t)i] - rx, dy = plan[3 * (size_t)i + 1] - ry;
  return sqrt(dx * dx + dy * dy);
}

// nav2 footprintCostAtPose on raw bytes (declared semantics, oracle/carrot_oracle.py: footprint_raw_cost)
__device__ int footprint_raw_cost(const CarrotConst& C, double x, double y, double yaw, int lane) {
  if (C.cells == nullptr || C.fp_n <= 0) return 0;
  const double c = cos(yaw), s = sin(yaw);
  int worst = 0;
  int mx0 = 0, my0 = 0, mxf = 0, myf = 0;
  for (int v = 0; v <= C.fp_n; ++v) {
    int mx, my;
    if (v < C.fp_n) {
      const double fx = (double)C.fp_x[v], fy = (double)C.fp_y[v];
      const double wx = x + (fx * c - fy * s), wy = y + (fx * s + fy * c);
      if (wx < C.origin_x || wy < C.origin_y) return 254;
      mx = (int)((wx - C.origin_x) / C.resolution);
      my = (int)((wy - C.origin_y) / C.resolution);
      if (mx >= C.W || my >= C.H) return 254;
      if (v == 0) { mxf = mx; myf = my; mx0 = mx; my0 = my; continue; }
    } else {
      mx = mxf; my = myf;
    }
    const int ddx = mx - mx0, ddy = my - my0;
    const int adx = ddx < 0 ? -ddx : ddx, ady = ddy < 0 ? -ddy : ddy;
    const int sxs = ddx >= 0 ? 1 : -1, sys = ddy >= 0 ? 1 : -1;
    const bool xmaj = adx >= ady;
    const int den = xmaj ? adx : ady, numadd = xmaj ? ady : adx;
    for (int k = lane; k <= den; k += 32) {
      const int minor = den > 0 ? (den / 2 + k * numadd) / den : 0;
      const int cx = xmaj ? mx0 + k * sxs : mx0 + minor * sxs;
      const int cy = xmaj ? my0 + minor * sys : my0 + k * sys;
      int val = 254;
      if (cx >= 0 && cy >= 0 && cx < C.W && cy < C.H) val = C.raw_table[__ldg(C.cells + (size_t)cy * C.W + cx)];
      worst = val > worst ? val : worst;
    }
    mx0 = mx; my0 = my;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const int t = __shfl_xor_sync(kFull, worst, o);
    worst = t > worst ? t : worst;
  }
  return worst;
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock)
build_requests_kernel(const __grid_constant__ CarrotConst C, const neompc_robot_tick* __restrict__ ticks, unsigned n,
                      uint32_t first_id, neompc_request* __restrict__ reqs, neompc_carrot_info* __restrict__ info) {
  const unsigned robot = blockIdx.x * kWarpsPerBlock + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (robot >= n) return;                                   // whole warps leave together
  const neompc_robot_tick tk = ticks[robot];
  const double rx = tk.pose_x, ry = tk.pose_y, ryaw = tk.pose_yaw;
  const unsigned L = C.L;
  const unsigned start = tk.plan_start < L ? tk.plan_start : L - 1;

  // closest pose from the pruned start on: first minimum (std::min_element semantics, cpp:81-86)
  double best = 1.0e300;
  unsigned best_i = 0xffffffffu;
  for (unsigned i = start + lane; i < L; i += 32) {
    const double d = dist_to(C.plan, i, rx, ry);
    if (d < best) { best = d; best_i = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(kFull, best, o);
    const unsigned oi = __shfl_xor_sync(kFull, best_i, o);
    if (ob < best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  const unsigned begin = best_i;
  const bool closer = dist_to(C.plan, L - 1, rx, ry) <= C.la_close;                  // cpp:88-96

  // first pose beyond the costmap window (cpp:98-103)
  unsigned end = L;
  for (unsigned base = begin; base < L; base += 32) {
    const unsigned i = base + lane;
    const bool out = i < L && dist_to(C.plan, i, rx, ry) > C.max_transform_dist;
    const unsigned m = __ballot_sync(kFull, out);
    if (m != 0) { end = base + (unsigned)__ffs((int)m) - 1; break; }
  }

  const int fc = footprint_raw_cost(C, rx, ry, ryaw, lane);                         // cpp:218-219
  unsigned status = NEOMPC_CARROT_OK;
  unsigned pick = begin;
  double cxb = 0.0, cyb = 0.0, cyaw = 0.0;
  bool slow = tk.slow_down != 0;
  if (end == begin) {
    status = NEOMPC_CARROT_EMPTY_WINDOW;                                            // cpp:130-132
  } else {
    double lookahead = C.la_min;                                                    // cpp:161-170
    if (!slow || closer) {
      lookahead = C.la_max;
      if (closer) lookahead = C.la_close;
    }
    const double c = cos(ryaw), s = sin(ryaw);
    pick = end - 1;                                                                 // cpp:184-186
    for (unsigned base = begin; base < end; base += 32) {                           // cpp:178-182
      const unsigned i = base + lane;
      bool far = false;
      if (i < end) {
        const double dx = C.plan[3 * (size_t)i] - rx, dy = C.plan[3 * (size_t)i + 1] - ry;
        const double xb = c * dx + s * dy, yb = -s * dx + c * dy;
        far = sqrt(xb * xb + yb * yb) >= lookahead;
      }
      const unsigned m = __ballot_sync(kFull, far);
      if (m != 0) { pick = base + (unsigned)__ffs((int)m) - 1; break; }
    }
    const double dx = C.plan[3 * (size_t)pick] - rx, dy = C.plan[3 * (size_t)pick + 1] - ry;
    const double dyaw = C.plan[3 * (size_t)pick + 2] - ryaw;
    cxb = c * dx + s * dy;
    cyb = -s * dx + c * dy;
    cyaw = atan2(sin(dyaw), cos(dyaw));
    // slow-down hysteresis (cpp:216-232)
    if (fabs(cyaw) < 1.0) slow = false;
    else if (fabs(cyaw) >= 1.0 && fc > 200) slow = true;
    else slow = false;
    if (fc == 255) status = NEOMPC_CARROT_COLLISION;                                // cpp:234-236
  }
  if (lane != 0) return;

  neompc_carrot_info ci;
  ci.status = status;
  ci.plan_start = begin;
  ci.carrot_index = pick;
  ci.flags = (closer ? 1u : 0u) | (slow ? 2u : 0u) | ((unsigned)fc << 8);
  info[robot] = ci;

  // the Optimizer request (cpp:240-246)
  const double gyaw = C.plan[3 * (size_t)(L - 1) + 2];
  const double zc = sin(0.5 * ryaw), wg = cos(0.5 * gyaw);
  neompc_request r;
  r.vel_x = tk.vel_x; r.vel_y = tk.vel_y; r.vel_theta = tk.vel_theta;               // cpp:241
  r.carrot_x = (float)cxb; r.carrot_y = (float)cyb; r.carrot_yaw = (float)cyaw;     // cpp:242
  r.goal_x = (float)C.plan[3 * (size_t)(L - 1)];                                    // cpp:243, :280
  r.goal_y = (float)C.plan[3 * (size_t)(L - 1) + 1];
  r.goal_yaw = (float)gyaw;
  r.pose_x = (float)rx; r.pose_y = (float)ry; r.pose_yaw = (float)ryaw;             // cpp:244
  // yaw(x=0, y=0, z of the current pose, w of the GOAL pose): the srv.py:213 quirk for planar poses
  r.pose_yaw_objective = (float)atan2(2.0 * (wg * zc), 1.0 - 2.0 * (zc * zc));
  r.control_interval = C.control_interval;                                          // cpp:246
  r.delta_t = tk.delta_t;
  r.instance_id = first_id == NEOMPC_STATELESS ? NEOMPC_STATELESS : first_id + robot;
  reqs[robot] = r;
}

}  // namespace

cudaError_t launch_build_requests(const CarrotConst& c, const neompc_robot_tick* d_ticks, unsigned n, uint32_t first_id,
                                  neompc_request* d_reqs, neompc_carrot_info* d_info, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const unsigned grid = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
  build_requests_kernel<<<grid, 32 * kWarpsPerBlock, 0, stream>>>(c, d_ticks, n, first_id, d_reqs, d_info);
  return cudaGetLastError();
}

}  // namespace neompc

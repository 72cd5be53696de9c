// runtime.cu — the C ABI of libneompc (include/neompc.h): handle, device buffers, streams, dispatch.
// Replaces the ROS service hop between NeoMpcPlanner::computeVelocityCommands (reference src/NeoMpcPlanner.cpp:240-252)
// and MpcOptimizationServer.optimizer (reference neo_mpc_planner2/mpc_optimization_server.py:349-403).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>       // types and prototypes only: the library is bound at run time (load_nccl), never linked

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "carrot.cuh"
#include "local_plan.cuh"
#include "kernels.cuh"
#include "mpc_setup.h"

using namespace neompc;

struct neompc_handle {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;   // second stream of the chunked host-buffer path
  neompc_params params{};
  SolverConst c{};
  HostTables tab;
  int encoding = NEOMPC_ENC_OCCUPANCY;
  int G = 1, S = 1;         // lane-group tiling of batches (throughput)
  int Gl = 1, Sl = 1;       // tiling of tiny batches (latency): one step per lane where possible
  bool force_general = false;   // NEOMPC_FORCE_GENERAL=1: always the general kernel build (test knob)
  bool no_full = false;         // NEOMPC_NO_FULL=1: never the full-horizon instantiation of the fast path (test knob)
  bool no_zero_copy = false;    // NEOMPC_NO_ZEROCOPY=1: host batches always go through staged copies (test / A-B knob)
  int last_host_path = 0;       // neompc_last_host_path
  float* d_lut_cost = nullptr;
  uint8_t* d_lut_flag = nullptr;
  uint8_t* d_cells = nullptr;
  size_t cells_cap = 0;
  uint32_t* d_cells4 = nullptr;    // corner-packed copy (mpc_core.cuh: corner_word), what the solve kernel samples
  size_t cells4_cap = 0;
  float* d_state = nullptr;
  unsigned state_rows = 0;
  // staging for the host-buffer entry points
  neompc_request* d_reqs = nullptr;
  neompc_response* d_resp = nullptr;
  float* d_plan = nullptr;
  float* d_twist = nullptr;         // staging of neompc_solve_batch_twists
  size_t cap_twist = 0;
  neompc_optimizer_request* d_msgs = nullptr;
  size_t cap_reqs = 0, cap_plan = 0, cap_msgs = 0;
  uint64_t launches = 0;
  // multi-GPU: one NCCL communicator over the handles of a fleet; the single collective of the path is the
  // all-gather of the solved (vx, vy, omega), enqueued on its own stream behind the solve kernel
  ncclComm_t comm = nullptr;
  int n_ranks = 1, rank = 0;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_solved = nullptr, ev_gathered[2] = {nullptr, nullptr};   // the two most recent gathers, alternating
  unsigned gathers = 0;
  float* d_gather = nullptr;        // staging of the single-process fleet entry: [n_ranks * shard_rows][3]
  size_t cap_gather = 0;
  unsigned* err_word = nullptr;     // mapped pinned word the kernels set when a request names a missing state row
  bool debug_ids = false;           // NEOMPC_DEBUG_IDS=1: host-side uniqueness check of instance ids (slow)
  // carrot selection (row N2): shared global plan, byte -> raw-cost table, staging
  double* d_path = nullptr;
  size_t path_len = 0, path_cap = 0;
  uint8_t* d_raw_table = nullptr;
  double resolution = 0.0;
  neompc_robot_tick* d_ticks = nullptr;
  neompc_carrot_info* d_info = nullptr;
  size_t cap_ticks = 0;
  int sm_count = 0;
  // small-batch mailbox: pinned + mapped host memory the kernels read requests from / write results to directly, so a
  // controller tick (n = 1) costs two launches and one synchronise instead of three staged pageable copies
  neompc_optimizer_request* mb_msgs = nullptr;
  neompc_request* mb_reqs = nullptr;
  neompc_response* mb_resp = nullptr;
  float* mb_plan = nullptr;
  neompc_robot_tick* mb_ticks = nullptr;
  neompc_carrot_info* mb_info = nullptr;
  std::string err;
};

constexpr size_t kMailboxRequests = 64;

static thread_local std::string g_create_error;

namespace {

// makes the handle's device current for the duration of an API call and restores the caller's afterwards
struct DeviceGuard {
  int prev = -1, dev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) : dev(device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};
#define NEOMPC_DEVICE(h)                                                        \
  DeviceGuard guard__((h)->device);                                             \
  if (guard__.err != cudaSuccess) return cuda_fail(h, guard__.err, "cudaSetDevice")

int fail(neompc_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}

int cuda_fail(neompc_handle* h, cudaError_t e, const char* what) {
  return fail(h, NEOMPC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define NEOMPC_CUDA(h, call)                                  \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) return cuda_fail(h, e__, #call);  \
  } while (0)

int upload_tables(neompc_handle* h) {
  build_tables(h->params, h->encoding, h->tab);
  NEOMPC_CUDA(h, cudaMemcpyAsync(h->d_lut_cost, h->tab.cost.data(), kTableSize * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  NEOMPC_CUDA(h, cudaMemcpyAsync(h->d_lut_flag, h->tab.flag.data(), kTableSize, cudaMemcpyHostToDevice, h->stream));
  // costmap byte -> nav2 raw cost, for the plugin's footprintCostAtPose thresholds (cpp:218-236); occupancy grids are
  // mapped back with the inverse of nav2's publisher table (oracle/carrot_oracle.py: raw_byte_table)
  uint8_t raw[256];
  for (int b = 0; b < 256; ++b) {
    if (h->encoding == NEOMPC_ENC_NAV2_RAW) raw[b] = (uint8_t)b;
    else if (b == 0) raw[b] = 0;
    else if (b <= 98) raw[b] = (uint8_t)(1 + (int)std::floor((b - 1) * 251.0 / 97.0 + 0.5));
    else if (b == 99) raw[b] = 253;
    else if (b == 100) raw[b] = 254;
    else raw[b] = 255;
  }
  NEOMPC_CUDA(h, cudaMemcpyAsync(h->d_raw_table, raw, 256, cudaMemcpyHostToDevice, h->stream));
  NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
  return NEOMPC_OK;
}

// everything in SolverConst that does not come from neompc_params survives a parameter update
void rebuild_const(neompc_handle* h) {
  const SolverConst old = h->c;
  build_const(h->params, h->c);
  h->c.cells = old.cells; h->c.cells4 = old.cells4; h->c.pad4 = old.pad4; h->c.pitch4 = old.pitch4; h->c.pad_ok = old.pad_ok;
  h->c.W = old.W; h->c.H = old.H;
  h->c.inv_res = old.inv_res; h->c.inv_res_d = old.inv_res_d;
  h->c.origin_x = old.origin_x; h->c.origin_y = old.origin_y;
  h->c.fp_n = old.fp_n;
  std::memcpy(h->c.fp_x, old.fp_x, sizeof(old.fp_x));
  std::memcpy(h->c.fp_y, old.fp_y, sizeof(old.fp_y));
  h->c.lethal_byte = h->encoding == NEOMPC_ENC_NAV2_RAW ? 254 : 100;
  h->c.cm_scale = 1.0f / (float)h->c.lethal_byte;
  h->c.state = h->d_state;
  h->c.state_rows = h->state_rows;
  h->c.err_word = h->err_word;
  choose_tiling(h->params.control_steps, h->params.lanes_per_instance, &h->G, &h->S);
  h->Gl = h->G; h->Sl = h->S;
  if (h->params.lanes_per_instance <= 0) choose_latency_tiling(h->params.control_steps, &h->Gl, &h->Sl);
}

int ensure_staging(neompc_handle* h, size_t n, bool want_plan, bool want_msgs) {
  if (n > h->cap_reqs) {
    if (h->d_reqs) cudaFree(h->d_reqs);
    if (h->d_resp) cudaFree(h->d_resp);
    h->d_reqs = nullptr; h->d_resp = nullptr; h->cap_reqs = 0;
    NEOMPC_CUDA(h, cudaMalloc(&h->d_reqs, n * sizeof(neompc_request)));
    NEOMPC_CUDA(h, cudaMalloc(&h->d_resp, n * sizeof(neompc_response)));
    h->cap_reqs = n;
  }
  const size_t plan_floats = n * 3 * (size_t)h->params.control_steps;
  if (want_plan && plan_floats > h->cap_plan) {
    if (h->d_plan) cudaFree(h->d_plan);
    h->d_plan = nullptr; h->cap_plan = 0;
    NEOMPC_CUDA(h, cudaMalloc(&h->d_plan, plan_floats * sizeof(float)));
    h->cap_plan = plan_floats;
  }
  if (want_msgs && n > h->cap_msgs) {
    if (h->d_msgs) cudaFree(h->d_msgs);
    h->d_msgs = nullptr; h->cap_msgs = 0;
    NEOMPC_CUDA(h, cudaMalloc(&h->d_msgs, n * sizeof(neompc_optimizer_request)));
    h->cap_msgs = n;
  }
  return NEOMPC_OK;
}

// A batch that cannot fill the device (a controller tick, a fleet of a few thousand robots) is limited by the length
// of the serial instruction stream of one warp, not by throughput; that stream is shortest with one step per lane.
// The latency tiling is used while all instances are resident at once with it (16 warps per SM).  Measured, N = 10:
// one request (4,3) 69 us -> (16,1) 58 us through neompc_solve_msgs; 4096 requests: profiles/latency_n1_r1.txt.
cudaError_t dispatch(neompc_handle* h, bool eval, const LaunchArgs& a) {
  const size_t resident_lanes = (size_t)h->sm_count * 16u * 32u;
  const size_t tn = a.tiling_n ? a.tiling_n : a.n;
  const bool latency = !eval && tn * (size_t)h->Gl <= resident_lanes;
  const int G = latency ? h->Gl : h->G, S = latency ? h->Sl : h->S;
  // general build unless the reference fast path applies (see Forward in mpc_core.cuh)
  const bool ext = h->params.footprint_mode != NEOMPC_FOOTPRINT_STATIC || h->params.costmap_mode != NEOMPC_COSTMAP_NEAREST ||
                   !h->c.disc_only || !h->c.fast_trig || h->c.m != 1 || h->force_general;
  switch (G) {
    case 1: return launch_g1(eval, S, ext, a);
    case 2: return launch_g2(eval, S, ext, a);
    case 3: return launch_g3(eval, S, ext, a);
    case 4: return launch_g4(eval, S, ext, a);
    case 5: return launch_g5(eval, S, ext, a);
    case 6: return launch_g6(eval, S, ext, a);
    case 8: return launch_g8(eval, S, ext, a);
    case 10: return launch_g10(eval, S, ext, a);
    case 16: return launch_g16(eval, S, ext, a);
    case 32: return launch_g32(eval, S, ext, a);
    default: return cudaErrorInvalidValue;
  }
}

// euler_from_quaternion yaw (reference mpc_optimization_server.py:176-178), float64 like the reference
__device__ __forceinline__ double yaw_of(double x, double y, double z, double w) {
  return atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z));
}

// Optimizer.Request (float64, quaternions) -> neompc_request (float32, yaws).  One thread per message.
__global__ void pack_kernel(const neompc_optimizer_request* __restrict__ msgs, unsigned n,
                            neompc_request* __restrict__ reqs) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const neompc_optimizer_request& m = msgs[i];
  neompc_request r;
  r.vel_x = (float)m.current_vel[0];                                  // srv.py:216
  r.vel_y = (float)m.current_vel[1];                                  // srv.py:217
  r.vel_theta = (float)m.current_vel[5];                              // srv.py:218
  r.carrot_x = (float)m.carrot_pose[0];
  r.carrot_y = (float)m.carrot_pose[1];
  r.carrot_yaw = (float)yaw_of(m.carrot_pose[3], m.carrot_pose[4], m.carrot_pose[5], m.carrot_pose[6]);   // srv.py:211
  r.goal_x = (float)m.goal_pose[0];
  r.goal_y = (float)m.goal_pose[1];
  r.goal_yaw = (float)yaw_of(m.goal_pose[3], m.goal_pose[4], m.goal_pose[5], m.goal_pose[6]);             // srv.py:212
  r.pose_x = (float)m.current_pose[0];
  r.pose_y = (float)m.current_pose[1];
  r.pose_yaw = (float)yaw_of(m.current_pose[3], m.current_pose[4], m.current_pose[5], m.current_pose[6]); // srv.py:317
  // the reference takes w from the GOAL pose here (srv.py:213)
  r.pose_yaw_objective = (float)yaw_of(m.current_pose[3], m.current_pose[4], m.current_pose[5], m.goal_pose[6]);
  r.control_interval = (float)m.control_interval;
  r.delta_t = (float)fmin(m.delta_t, 3.0e38);
  r.instance_id = m.instance_id;
  reqs[i] = r;
}

// corner-packed copy of the costmap: one thread per entry of the padded (W + 2 pad) x (H + 2 pad) grid
__global__ void corner_map_kernel(const uint8_t* __restrict__ cells, int W, int H, int lethal_byte, int pad,
                                  uint32_t* __restrict__ out) {
  const int pitch = corner_pitch(W, pad);
  const size_t total = corner_words(W, H, pad);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int iy = (int)(i / pitch) - pad, ix = (int)(i % pitch) - pad;
    out[i] = corner_word(cells, W, H, lethal_byte, ix, iy);
  }
}

__global__ void reset_rows_kernel(float* state, int stride, const uint32_t* ids, unsigned n, unsigned rows) {
  const unsigned i = blockIdx.x;
  if (i >= n) return;
  const uint32_t id = ids[i];
  if (id >= rows) return;
  for (int k = threadIdx.x; k < stride; k += blockDim.x) state[(size_t)id * stride + k] = 0.0f;
}

// (Re)builds the corner-packed copy for the current costmap, encoding and reach (corner_pad_for: parameters and
// resolution); enqueued on the handle's stream, no synchronise.
int rebuild_corner_map(neompc_handle* h) {
  if (h->c.cells == nullptr) { h->c.cells4 = nullptr; h->c.pad4 = 0; h->c.pitch4 = 0; h->c.pad_ok = 0; return NEOMPC_OK; }
  int pad_ok = 0;
  const int pad = corner_pad_for(h->params, h->resolution, &pad_ok);
  const size_t words = corner_words(h->c.W, h->c.H, pad);
  if (words > h->cells4_cap) {
    NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->d_cells4) cudaFree(h->d_cells4);
    h->d_cells4 = nullptr; h->cells4_cap = 0;
    NEOMPC_CUDA(h, cudaMalloc(&h->d_cells4, words * sizeof(uint32_t)));
    h->cells4_cap = words;
  }
  const unsigned blocks = (unsigned)((words + 255) / 256 < 4096 ? (words + 255) / 256 : 4096);
  corner_map_kernel<<<blocks, 256, 0, h->stream>>>(h->d_cells, h->c.W, h->c.H, h->c.lethal_byte, pad, h->d_cells4);
  NEOMPC_CUDA(h, cudaGetLastError());
  h->launches += 1;
  h->c.cells4 = h->d_cells4;
  h->c.pad4 = pad;
  h->c.pad_ok = pad_ok;
  h->c.pitch4 = corner_pitch(h->c.W, pad);
  return NEOMPC_OK;
}

// after a host-buffer solve has been synchronised: did a request name a state row that does not exist?
int check_state_errors(neompc_handle* h) {
  if (*h->err_word == 0u) return NEOMPC_OK;
  *h->err_word = 0u;
  return fail(h, NEOMPC_ERR_STATE,
              "a request's instance_id lies beyond the reserved state rows (neompc_reserve_instances); it was solved as a "
              "cold start and its response carries NEOMPC_FLAG_NO_STATE");
}

// NEOMPC_DEBUG_IDS=1: instance ids of one batch must be unique (two requests with one id race on its state row)
template <class Rec>
int check_unique_ids(neompc_handle* h, const Rec* recs, size_t n) {
  if (!h->debug_ids) return NEOMPC_OK;
  std::vector<uint32_t> ids;
  ids.reserve(n);
  for (size_t i = 0; i < n; ++i)
    if (recs[i].instance_id != NEOMPC_STATELESS) ids.push_back(recs[i].instance_id);
  std::sort(ids.begin(), ids.end());
  if (std::adjacent_find(ids.begin(), ids.end()) != ids.end())
    return fail(h, NEOMPC_ERR_INVALID, "duplicate instance_id within one batch (NEOMPC_DEBUG_IDS check)");
  return NEOMPC_OK;
}


// ---- NCCL, bound at run time --------------------------------------------------------------------------------------
// libneompc.so carries no link-time dependency on NCCL: a controller plugin that drives one GPU never needs it, and a
// Python host has usually loaded its own copy already (torch bundles one under the same SONAME, which dlopen then returns).
struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitRankConfig) CommInitRankConfig = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
};

const NcclApi* load_nccl(std::string& err) {
  static NcclApi api;
  static bool tried = false;
  static std::string load_err;
  if (!tried) {
    tried = true;
    // 1. a copy this process has loaded already (a Python host's torch bundles one under the same SONAME: binding to
    //    another copy first would break a later `import torch`; the Python binding imports torch first for that reason);
    // 2. NEOMPC_NCCL_LIB, an explicit path;  3. the system's libnccl.so.2.
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!api.lib && std::getenv("NEOMPC_NCCL_LIB")) api.lib = dlopen(std::getenv("NEOMPC_NCCL_LIB"), RTLD_NOW | RTLD_LOCAL);
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (api.lib) break;
      api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    }
    if (!api.lib) {
      load_err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
    } else {
#define NEOMPC_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name))
      NEOMPC_SYM(GetUniqueId, "ncclGetUniqueId");
      NEOMPC_SYM(CommInitRank, "ncclCommInitRank");
      NEOMPC_SYM(CommInitRankConfig, "ncclCommInitRankConfig");
      NEOMPC_SYM(CommDestroy, "ncclCommDestroy");
      NEOMPC_SYM(AllGather, "ncclAllGather");
      NEOMPC_SYM(GroupStart, "ncclGroupStart");
      NEOMPC_SYM(GroupEnd, "ncclGroupEnd");
      NEOMPC_SYM(GetErrorString, "ncclGetErrorString");
      NEOMPC_SYM(GetVersion, "ncclGetVersion");
#undef NEOMPC_SYM
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GroupStart || !api.GroupEnd)
        load_err = "libnccl.so.2 lacks a required symbol";
    }
  }
  if (!load_err.empty()) { err = load_err; return nullptr; }
  return &api;
}

int nccl_fail(neompc_handle* h, const NcclApi* api, ncclResult_t r, const char* what) {
  return fail(h, NEOMPC_ERR_NCCL, std::string(what) + ": " + (api && api->GetErrorString ? api->GetErrorString(r) : "NCCL error"));
}

void comm_release(neompc_handle* h) {
  std::string err;
  if (h->comm) {
    const NcclApi* api = load_nccl(err);
    if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
    if (api) api->CommDestroy(h->comm);
    h->comm = nullptr;
  }
  if (h->comm_stream) { cudaStreamDestroy(h->comm_stream); h->comm_stream = nullptr; }
  if (h->ev_solved) { cudaEventDestroy(h->ev_solved); h->ev_solved = nullptr; }
  for (int k = 0; k < 2; ++k)
    if (h->ev_gathered[k]) { cudaEventDestroy(h->ev_gathered[k]); h->ev_gathered[k] = nullptr; }
  h->gathers = 0;
  h->n_ranks = 1; h->rank = 0;
}

// stream + events of the gather; the communicator itself is created by the callers below
int comm_prepare(neompc_handle* h, int n_ranks, int rank) {
  comm_release(h);
  NEOMPC_CUDA(h, cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
  NEOMPC_CUDA(h, cudaEventCreateWithFlags(&h->ev_solved, cudaEventDisableTiming));
  for (int k = 0; k < 2; ++k) NEOMPC_CUDA(h, cudaEventCreateWithFlags(&h->ev_gathered[k], cudaEventDisableTiming));
  h->n_ranks = n_ranks; h->rank = rank;
  return NEOMPC_OK;
}

// The gather moves 12 B per solve (786 KB per rank at C3) while the next batch is being solved on the same SMs.  Holding NCCL
// to very few CTAs (NEOMPC_NCCL_MAX_CTAS) was measured and is NOT the default: with 2 CTAs the gather gets ~10 GB/s next to
// a solve kernel that fills every SM and ends up on the critical path (4 GPUs, C3: step 0.580 ms against 0.566 ms with
// NCCL's own choice; 4 / 8 / 16 CTAs: 0.569 / 0.569 / 0.564 — profiles/nccl_ctas_r2.txt).
ncclResult_t comm_init_rank(const NcclApi* api, ncclComm_t* comm, int n_ranks, const ncclUniqueId& id, int rank) {
  if (api->CommInitRankConfig) {
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    const char* e = std::getenv("NEOMPC_NCCL_MAX_CTAS");
    const int ctas = e ? std::atoi(e) : 0;
    if (ctas > 0) { cfg.minCTAs = 1; cfg.maxCTAs = ctas; }
    return api->CommInitRankConfig(comm, n_ranks, id, rank, &cfg);
  }
  return api->CommInitRank(comm, n_ranks, id, rank);
}

int do_solve_device(neompc_handle* h, const neompc_request* d_reqs, size_t n, neompc_response* d_out,
                    float* d_twist, float* d_plan, cudaStream_t s, size_t tiling_n = 0) {
  if (n == 0) return NEOMPC_OK;
  if (n > 0xFFFFFFF0ull) return fail(h, NEOMPC_ERR_INVALID, "batch too large");
  if ((reinterpret_cast<uintptr_t>(d_reqs) & 15u) != 0)      // the kernel stages request tiles with cp.async.bulk
    return fail(h, NEOMPC_ERR_INVALID, "request array must be 16-byte aligned");
  LaunchArgs a{};
  a.P = h->c;
  a.lut_cost = h->d_lut_cost;
  a.lut_flag = h->d_lut_flag;
  a.reqs = d_reqs;
  a.n = (unsigned)n;
  a.out = d_out;
  a.twist = d_twist;
  a.plan = d_plan;
  a.stream = s;
  a.tiling_n = (unsigned)tiling_n;
  a.no_full = h->no_full;
  cudaError_t e = dispatch(h, false, a);
  if (e != cudaSuccess) return cuda_fail(h, e, "solve kernel launch");
  h->launches += 1;
  return NEOMPC_OK;
}

}  // namespace

extern "C" {

int neompc_version(void) { return NEOMPC_VERSION; }

int neompc_abi_sizes(size_t out[7]) {
  if (!out) return NEOMPC_ERR_INVALID;
  out[0] = sizeof(neompc_request);
  out[1] = sizeof(neompc_response);
  out[2] = sizeof(neompc_params);
  out[3] = sizeof(neompc_optimizer_request);
  out[4] = sizeof(neompc_robot_tick);
  out[5] = sizeof(neompc_carrot_info);
  out[6] = sizeof(neompc_plan_pose);
  return NEOMPC_OK;
}

const char* neompc_last_error(const neompc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int neompc_create(const neompc_params* params, int device, neompc_handle** out) {
  if (!params || !out) return fail(nullptr, NEOMPC_ERR_INVALID, "null argument");
  *out = nullptr;
  std::string err;
  if (!validate_params(*params, err)) return fail(nullptr, NEOMPC_ERR_INVALID, err);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, NEOMPC_ERR_NO_DEVICE,
                std::string("no CUDA device available (libneompc has no CPU fallback): ") +
                    (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (device < 0 || device >= count) return fail(nullptr, NEOMPC_ERR_NO_DEVICE, "device ordinal out of range");
  neompc_handle* h = new (std::nothrow) neompc_handle();
  if (!h) return fail(nullptr, NEOMPC_ERR_INVALID, "out of host memory");
  h->device = device;
  h->params = *params;
  h->force_general = std::getenv("NEOMPC_FORCE_GENERAL") != nullptr;
  h->no_full = std::getenv("NEOMPC_NO_FULL") != nullptr;
  h->no_zero_copy = std::getenv("NEOMPC_NO_ZEROCOPY") != nullptr;
#define CREATE_CUDA(call)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      g_create_error = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
      neompc_destroy(h);                                                                    \
      return NEOMPC_ERR_CUDA;                                                               \
    }                                                                                       \
  } while (0)
  DeviceGuard guard__(device);
  CREATE_CUDA(guard__.err);
  CREATE_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CREATE_CUDA(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  CREATE_CUDA(cudaMalloc(&h->d_raw_table, 256));
  CREATE_CUDA(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
  CREATE_CUDA(cudaMalloc(&h->d_lut_cost, kTableSize * sizeof(float)));
  CREATE_CUDA(cudaMalloc(&h->d_lut_flag, kTableSize + 6));
  CREATE_CUDA(cudaHostAlloc(&h->err_word, sizeof(unsigned), cudaHostAllocMapped));
  *h->err_word = 0u;
  h->debug_ids = std::getenv("NEOMPC_DEBUG_IDS") != nullptr;
  CREATE_CUDA(cudaHostAlloc(&h->mb_msgs, kMailboxRequests * sizeof(neompc_optimizer_request), cudaHostAllocMapped));
  CREATE_CUDA(cudaHostAlloc(&h->mb_reqs, kMailboxRequests * sizeof(neompc_request), cudaHostAllocMapped));
  CREATE_CUDA(cudaHostAlloc(&h->mb_resp, kMailboxRequests * sizeof(neompc_response), cudaHostAllocMapped));
  CREATE_CUDA(cudaHostAlloc(&h->mb_plan, kMailboxRequests * 3 * NEOMPC_MAX_CONTROL_STEPS * sizeof(float), cudaHostAllocMapped));
  CREATE_CUDA(cudaHostAlloc(&h->mb_ticks, kMailboxRequests * sizeof(neompc_robot_tick), cudaHostAllocMapped));
  CREATE_CUDA(cudaHostAlloc(&h->mb_info, kMailboxRequests * sizeof(neompc_carrot_info), cudaHostAllocMapped));
#undef CREATE_CUDA
  build_const(h->params, h->c);
  rebuild_const(h);
  int rc = upload_tables(h);
  if (rc != NEOMPC_OK) { g_create_error = h->err; neompc_destroy(h); return rc; }
  *out = h;
  return NEOMPC_OK;
}

int neompc_destroy(neompc_handle* h) {
  if (!h) return NEOMPC_OK;
  DeviceGuard guard__(h->device >= 0 ? h->device : 0);
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
  comm_release(h);
  cudaFree(h->d_gather);
  cudaFree(h->d_lut_cost); cudaFree(h->d_lut_flag); cudaFree(h->d_cells); cudaFree(h->d_cells4); cudaFree(h->d_state);
  cudaFree(h->d_reqs); cudaFree(h->d_resp); cudaFree(h->d_plan); cudaFree(h->d_msgs); cudaFree(h->d_twist);
  cudaFree(h->d_path); cudaFree(h->d_raw_table); cudaFree(h->d_ticks); cudaFree(h->d_info);
  cudaFreeHost(h->err_word); cudaFreeHost(h->mb_msgs); cudaFreeHost(h->mb_reqs); cudaFreeHost(h->mb_resp); cudaFreeHost(h->mb_plan); cudaFreeHost(h->mb_ticks); cudaFreeHost(h->mb_info);
  delete h;
  return NEOMPC_OK;
}

int neompc_set_params(neompc_handle* h, const neompc_params* params) {
  if (!h || !params) return fail(h, NEOMPC_ERR_INVALID, "null argument");
  std::string err;
  if (!validate_params(*params, err)) return fail(h, NEOMPC_ERR_INVALID, err);
  NEOMPC_DEVICE(h);
  const bool steps_changed = params->control_steps != h->params.control_steps;
  h->params = *params;
  if (steps_changed && h->d_state) {           // rows have a different layout now: start over (srv.py never resizes either)
    const unsigned rows = h->state_rows;
    cudaFree(h->d_state);
    h->d_state = nullptr; h->state_rows = 0;
    rebuild_const(h);
    int rc = neompc_reserve_instances(h, rows);
    if (rc != NEOMPC_OK) return rc;
  }
  rebuild_const(h);
  if (h->c.cells != nullptr && corner_pad_for(h->params, h->resolution) != h->c.pad4) {   // reach changed
    int rc = rebuild_corner_map(h);
    if (rc != NEOMPC_OK) return rc;
  }
  return upload_tables(h);
}

int neompc_get_params(const neompc_handle* h, neompc_params* out) {
  if (!h || !out) return NEOMPC_ERR_INVALID;
  *out = h->params;
  return NEOMPC_OK;
}

static int set_costmap_common(neompc_handle* h, const uint8_t* cells, bool on_device, uint32_t width, uint32_t height,
                              double resolution, double origin_x, double origin_y, int encoding) {
  if (!h) return NEOMPC_ERR_INVALID;
  NEOMPC_DEVICE(h);
  if (encoding != NEOMPC_ENC_OCCUPANCY && encoding != NEOMPC_ENC_NAV2_RAW)
    return fail(h, NEOMPC_ERR_INVALID, "unknown costmap encoding");
  if (cells == nullptr) {                      // free space
    h->c.cells = nullptr; h->c.cells4 = nullptr; h->c.pad4 = h->c.pitch4 = h->c.pad_ok = 0; h->c.W = h->c.H = 0;
    return NEOMPC_OK;
  }
  if (width == 0 || height == 0 || width > 65535u || height > 65535u || !(resolution > 0.0))
    return fail(h, NEOMPC_ERR_INVALID, "bad costmap geometry");
  const size_t bytes = (size_t)width * height;
  if (bytes > h->cells_cap) {
    NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->d_cells) cudaFree(h->d_cells);
    h->d_cells = nullptr; h->cells_cap = 0;
    NEOMPC_CUDA(h, cudaMalloc(&h->d_cells, bytes));
    h->cells_cap = bytes;
  }
  NEOMPC_CUDA(h, cudaMemcpyAsync(h->d_cells, cells, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                 h->stream));
  h->c.cells = h->d_cells;
  h->c.W = (int)width; h->c.H = (int)height;
  h->c.inv_res_d = 1.0 / resolution;
  h->c.inv_res = (float)(1.0 / resolution);
  h->c.origin_x = origin_x; h->c.origin_y = origin_y;
  h->resolution = resolution;
  const bool enc_changed = encoding != h->encoding;
  if (enc_changed) {
    h->encoding = encoding;
    h->c.lethal_byte = encoding == NEOMPC_ENC_NAV2_RAW ? 254 : 100;
    h->c.cm_scale = 1.0f / (float)h->c.lethal_byte;
  }
  int rc = rebuild_corner_map(h);
  if (rc != NEOMPC_OK) return rc;
  if (enc_changed) return upload_tables(h);
  NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
  return NEOMPC_OK;
}

int neompc_set_costmap(neompc_handle* h, const uint8_t* cells, uint32_t width, uint32_t height, double resolution,
                       double origin_x, double origin_y, int encoding) {
  return set_costmap_common(h, cells, false, width, height, resolution, origin_x, origin_y, encoding);
}

int neompc_set_costmap_device(neompc_handle* h, const uint8_t* d_cells, uint32_t width, uint32_t height,
                              double resolution, double origin_x, double origin_y, int encoding) {
  return set_costmap_common(h, d_cells, true, width, height, resolution, origin_x, origin_y, encoding);
}

int neompc_set_footprint(neompc_handle* h, const float* xy, int n_vertices) {
  if (!h) return NEOMPC_ERR_INVALID;
  if (n_vertices < 0 || n_vertices > NEOMPC_MAX_FOOTPRINT_VERTICES || (n_vertices > 0 && !xy))
    return fail(h, NEOMPC_ERR_INVALID, "footprint must have 0..16 vertices");
  h->c.fp_n = n_vertices;
  for (int i = 0; i < n_vertices; ++i) { h->c.fp_x[i] = xy[2 * i]; h->c.fp_y[i] = xy[2 * i + 1]; }
  return NEOMPC_OK;
}

int neompc_reserve_instances(neompc_handle* h, uint32_t n_instances) {
  if (!h) return NEOMPC_ERR_INVALID;
  NEOMPC_DEVICE(h);
  if (n_instances <= h->state_rows) return NEOMPC_OK;
  const int stride = state_stride_for(h->params.control_steps);
  float* fresh = nullptr;
  NEOMPC_CUDA(h, cudaMalloc(&fresh, (size_t)n_instances * stride * sizeof(float)));
  cudaError_t e = cudaMemsetAsync(fresh, 0, (size_t)n_instances * stride * sizeof(float), h->stream);
  if (e == cudaSuccess && h->d_state)
    e = cudaMemcpyAsync(fresh, h->d_state, (size_t)h->state_rows * stride * sizeof(float), cudaMemcpyDeviceToDevice,
                        h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { cudaFree(fresh); return cuda_fail(h, e, "reserve_instances"); }
  if (h->d_state) cudaFree(h->d_state);
  h->d_state = fresh;
  h->state_rows = n_instances;
  h->c.state = h->d_state;
  h->c.state_rows = h->state_rows;
  h->c.state_stride = stride;
  return NEOMPC_OK;
}

int neompc_reset_state(neompc_handle* h, const uint32_t* ids, size_t n) {
  if (!h) return NEOMPC_ERR_INVALID;
  NEOMPC_DEVICE(h);
  if (!h->d_state) return NEOMPC_OK;
  const int stride = h->c.state_stride;
  if (ids == nullptr) {
    NEOMPC_CUDA(h, cudaMemsetAsync(h->d_state, 0, (size_t)h->state_rows * stride * sizeof(float), h->stream));
  } else if (n > 0) {
    uint32_t* d_ids = nullptr;
    NEOMPC_CUDA(h, cudaMalloc(&d_ids, n * sizeof(uint32_t)));
    cudaError_t e = cudaMemcpyAsync(d_ids, ids, n * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
      reset_rows_kernel<<<(unsigned)n, 64, 0, h->stream>>>(h->d_state, stride, d_ids, (unsigned)n, h->state_rows);
      h->launches += 1;
      e = cudaStreamSynchronize(h->stream);
    }
    cudaFree(d_ids);
    if (e != cudaSuccess) return cuda_fail(h, e, "reset_state");
    return NEOMPC_OK;
  }
  NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
  return NEOMPC_OK;
}

int neompc_get_state(neompc_handle* h, uint32_t id, float* initial_guess, float last_control[3], float* waiting_time,
                     uint32_t* flags) {
  if (!h) return NEOMPC_ERR_INVALID;
  if (id >= h->state_rows) return fail(h, NEOMPC_ERR_STATE, "instance id beyond reserved capacity");
  NEOMPC_DEVICE(h);
  const int stride = h->c.state_stride;
  std::vector<float> row(stride);
  NEOMPC_CUDA(h, cudaMemcpyAsync(row.data(), h->d_state + (size_t)id * stride, stride * sizeof(float),
                                 cudaMemcpyDeviceToHost, h->stream));
  NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
  const int n3 = 3 * h->params.control_steps;
  if (initial_guess) std::memcpy(initial_guess, row.data(), n3 * sizeof(float));
  if (last_control) std::memcpy(last_control, row.data() + n3, 3 * sizeof(float));
  if (waiting_time) *waiting_time = row[n3 + 3];
  if (flags) *flags = (row[n3 + 4] != 0.0f ? NEOMPC_FLAG_COLLISION : 0) |
                      (row[n3 + 9] != 0.0f ? NEOMPC_FLAG_COLLISION_FOOTPRINT : 0);
  return NEOMPC_OK;
}

int neompc_solve_batch_device(neompc_handle* h, const neompc_request* d_reqs, size_t n, neompc_response* d_out,
                              float* d_twist_or_null, float* d_plan_or_null, void* stream) {
  if (!h || (n > 0 && (!d_reqs || !d_out))) return fail(h, NEOMPC_ERR_INVALID, "null argument");
  NEOMPC_DEVICE(h);
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  return do_solve_device(h, d_reqs, n, d_out, d_twist_or_null, d_plan_or_null, s);
}

// Host buffers in, host buffers out.  out: full responses (or null), twist_out: packed (vx, vy, omega) (or null).
static int solve_batch_host(neompc_handle* h, const neompc_request* reqs, size_t n, neompc_response* out, float* twist_out,
                            float* plan_or_null) {
  int rc = check_unique_ids(h, reqs, n);
  if (rc != NEOMPC_OK) return rc;
  rc = ensure_staging(h, n, plan_or_null != nullptr, false);
  if (rc != NEOMPC_OK) return rc;
  if (twist_out && n * 3 > h->cap_twist) {
    NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_twist);
    h->d_twist = nullptr; h->cap_twist = 0;
    NEOMPC_CUDA(h, cudaMalloc(&h->d_twist, n * 3 * sizeof(float)));
    h->cap_twist = n * 3;
  }
  if (n <= kMailboxRequests) {                 // small batch: pinned mailbox in, mapped mailbox out (see solve_msgs)
    h->last_host_path = NEOMPC_HOST_PATH_MAILBOX;
    const size_t n3s = n * 3 * (size_t)h->params.control_steps;
    std::memcpy(h->mb_reqs, reqs, n * sizeof(neompc_request));
    NEOMPC_CUDA(h, cudaMemcpyAsync(h->d_reqs, h->mb_reqs, n * sizeof(neompc_request), cudaMemcpyHostToDevice, h->stream));
    rc = do_solve_device(h, h->d_reqs, n, h->mb_resp, nullptr, plan_or_null ? h->mb_plan : nullptr, h->stream);
    if (rc != NEOMPC_OK) return rc;
    NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
    if (out) std::memcpy(out, h->mb_resp, n * sizeof(neompc_response));
    if (twist_out)
      for (size_t i = 0; i < n; ++i) {
        twist_out[3 * i] = h->mb_resp[i].vx; twist_out[3 * i + 1] = h->mb_resp[i].vy; twist_out[3 * i + 2] = h->mb_resp[i].omega;
      }
    if (plan_or_null) std::memcpy(plan_or_null, h->mb_plan, n3s * sizeof(float));
    return check_state_errors(h);
  }
  // Zero-copy: when the caller's buffers are page-locked and device-accessible (cudaHostAlloc / cudaHostRegister /
  // neompc_host_alloc: with unified addressing every pinned allocation is), the kernel moves the data itself — the TMA
  // bulk copy that stages a block's request tile reads the 64-byte records straight from host memory over PCIe, lane 0
  // of a group writes its 12-byte twist (or 32-byte response) straight back — and ONE launch replaces the chunked copy
  // pipeline below: no upload to wait for before the first block starts, no download after the last one finishes.  The
  // transfers (4 MiB in, 768 KiB out at C3) ride under 0.39 ms of compute; what is left of the call is launch +
  // synchronise.  Measured on C3: profiles/host_zero_copy_r2.txt.  Pageable buffers take the chunked path.
  if (!h->no_zero_copy) {
    auto mapped = [](const void* p) -> void* {
      if (p == nullptr) return nullptr;
      cudaPointerAttributes at{};
      if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
    };
    const neompc_request* zr = static_cast<const neompc_request*>(mapped(reqs));
    neompc_response* zo = static_cast<neompc_response*>(mapped(out));
    float* zt = static_cast<float*>(mapped(twist_out));
    float* zp = static_cast<float*>(mapped(plan_or_null));
    const bool all_mapped = zr != nullptr && (out == nullptr || zo != nullptr) && (twist_out == nullptr || zt != nullptr) &&
                            (plan_or_null == nullptr || zp != nullptr) && (reinterpret_cast<uintptr_t>(zr) & 15u) == 0;
    if (all_mapped) {
      h->last_host_path = NEOMPC_HOST_PATH_ZERO_COPY;
      rc = do_solve_device(h, zr, n, zo ? zo : h->d_resp, zt, zp, h->stream, n);
      if (rc != NEOMPC_OK) return rc;
      NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
      return check_state_errors(h);
    }
  }
  h->last_host_path = NEOMPC_HOST_PATH_CHUNKED;
  // Large batches are cut into chunks on two streams, so the H2D copy of one chunk, the solve of another and the D2H
  // copies overlap (the copy engines for the two directions and the SMs are independent).  Problems are independent, so
  // chunking does not change any result.  What stays exposed is the H2D of the first chunk and the D2H of the last one.
  // Three chunks: measured on C3 with the board at its working clocks (profiles/host_chunks_r2.txt), 65536 requests in
  // 0.664 / 0.624 / 0.611 / 0.641 / 0.653 / 0.771 ms with 1 / 2 / 3 / 4 / 6 / 8 chunks (kernel alone 0.55 ms); with the
  // 0.443 ms kernel 0.548 / 0.537 / 0.497 / 0.530 ms with 1 / 2 / 3 / 4 equal chunks and 0.490 ms with sizes 1 : 4 : 6.
  const size_t n3 = 3 * (size_t)h->params.control_steps;
  static const int chunk_override = std::getenv("NEOMPC_CHUNKS") ? std::atoi(std::getenv("NEOMPC_CHUNKS")) : 0;   // tuning knob
  const int chunks = n >= 16384 ? (chunk_override > 0 ? chunk_override : 3) : 1;
  // Chunk sizes: the first chunk is the small one — its H2D copy is the part of the transfer nothing can hide.  Relative
  // weights, default 1 : 4 : 6 for three chunks (NEOMPC_CHUNK_WEIGHTS="a,b,c,..." for tuning; equal sizes otherwise).
  static const std::vector<int> weight_override = [] {
    std::vector<int> w;
    if (const char* e = std::getenv("NEOMPC_CHUNK_WEIGHTS"))
      for (const char* p = e; *p;) { w.push_back(std::atoi(p)); while (*p && *p != ',') ++p; if (*p) ++p; }
    return w;
  }();
  size_t bound[17];
  {
    int w[16], wsum = 0;
    for (int c = 0; c < chunks && c < 16; ++c) {
      w[c] = (int)weight_override.size() == chunks ? std::max(1, weight_override[c])
                                                   : (chunks == 3 && chunk_override == 0 ? (c == 0 ? 1 : c == 1 ? 4 : 6) : 1);
      wsum += w[c];
    }
    size_t acc = 0;
    bound[0] = 0;
    for (int c = 0; c < chunks && c < 16; ++c) {
      acc += (size_t)w[c];
      bound[c + 1] = c + 1 == chunks ? n : std::min(n, (n * acc / (size_t)wsum + 63) / 64 * 64);
    }
  }
  cudaStream_t streams[2] = {h->stream, chunks > 1 ? h->stream2 : h->stream};
  for (int c = 0; c < chunks && c < 16; ++c) {
    const size_t lo = bound[c];
    if (lo >= n) break;
    const size_t cnt = bound[c + 1] - lo;
    if (cnt == 0) continue;
    cudaStream_t s = streams[c & 1];
    NEOMPC_CUDA(h, cudaMemcpyAsync(h->d_reqs + lo, reqs + lo, cnt * sizeof(neompc_request), cudaMemcpyHostToDevice, s));
    rc = do_solve_device(h, h->d_reqs + lo, cnt, h->d_resp + lo, twist_out ? h->d_twist + lo * 3 : nullptr,
                         plan_or_null ? h->d_plan + lo * n3 : nullptr, s, n);
    if (rc != NEOMPC_OK) return rc;
    if (out)
      NEOMPC_CUDA(h, cudaMemcpyAsync(out + lo, h->d_resp + lo, cnt * sizeof(neompc_response), cudaMemcpyDeviceToHost, s));
    if (twist_out)
      NEOMPC_CUDA(h, cudaMemcpyAsync(twist_out + lo * 3, h->d_twist + lo * 3, cnt * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (plan_or_null)
      NEOMPC_CUDA(h, cudaMemcpyAsync(plan_or_null + lo * n3, h->d_plan + lo * n3, cnt * n3 * sizeof(float),
                                     cudaMemcpyDeviceToHost, s));
  }
  NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
  if (chunks > 1) NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream2));
  return check_state_errors(h);
}

int neompc_solve_batch(neompc_handle* h, const neompc_request* reqs, size_t n, neompc_response* out,
                       float* plan_or_null) {
  if (!h || (n > 0 && (!reqs || !out))) return fail(h, NEOMPC_ERR_INVALID, "null argument");
  if (n == 0) return NEOMPC_OK;
  NEOMPC_DEVICE(h);
  return solve_batch_host(h, reqs, n, out, nullptr, plan_or_null);
}

int neompc_solve_batch_twists(neompc_handle* h, const neompc_request* reqs, size_t n, float* twist_out) {
  if (!h || (n > 0 && (!reqs || !twist_out))) return fail(h, NEOMPC_ERR_INVALID, "null argument");
  if (n == 0) return NEOMPC_OK;
  NEOMPC_DEVICE(h);
  return solve_batch_host(h, reqs, n, nullptr, twist_out, nullptr);
}

int neompc_pack_requests(neompc_handle* h, const neompc_optimizer_request* d_msgs, size_t n, neompc_request* d_reqs,
                         void* stream) {
  if (!h || (n > 0 && (!d_msgs || !d_reqs))) return fail(h, NEOMPC_ERR_INVALID, "null argument");
  if (n == 0) return NEOMPC_OK;
  NEOMPC_DEVICE(h);
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_msgs, (unsigned)n, d_reqs);
  NEOMPC_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return NEOMPC_OK;
}

int neompc_solve_msgs(neompc_handle* h, const neompc_optimizer_request* msgs, size_t n, neompc_response* out,
                      float* plan_or_null) {
  if (!h || (n > 0 && (!msgs || !out))) return fail(h, NEOMPC_ERR_INVALID, "null argument");
  if (n == 0) return NEOMPC_OK;
  NEOMPC_DEVICE(h);
  int rc = check_unique_ids(h, msgs, n);
  if (rc != NEOMPC_OK) return rc;
  rc = ensure_staging(h, n, plan_or_null != nullptr, true);
  if (rc != NEOMPC_OK) return rc;
  if (n <= kMailboxRequests) {
    // controller-tick path: messages go through the mapped mailbox (the pack kernel reads them over PCIe), responses
    // and plan are written by the solve kernel straight into mapped host memory
    const size_t n3 = n * 3 * (size_t)h->params.control_steps;
    std::memcpy(h->mb_msgs, msgs, n * sizeof(neompc_optimizer_request));
    rc = neompc_pack_requests(h, h->mb_msgs, n, h->d_reqs, h->stream);
    if (rc != NEOMPC_OK) return rc;
    rc = do_solve_device(h, h->d_reqs, n, h->mb_resp, nullptr, plan_or_null ? h->mb_plan : nullptr, h->stream);
    if (rc != NEOMPC_OK) return rc;
    NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
    std::memcpy(out, h->mb_resp, n * sizeof(neompc_response));
    if (plan_or_null) std::memcpy(plan_or_null, h->mb_plan, n3 * sizeof(float));
    return check_state_errors(h);
  }
  NEOMPC_CUDA(h, cudaMemcpyAsync(h->d_msgs, msgs, n * sizeof(neompc_optimizer_request), cudaMemcpyHostToDevice, h->stream));
  rc = neompc_pack_requests(h, h->d_msgs, n, h->d_reqs, h->stream);
  if (rc != NEOMPC_OK) return rc;
  rc = do_solve_device(h, h->d_reqs, n, h->d_resp, nullptr, plan_or_null ? h->d_plan : nullptr, h->stream);
  if (rc != NEOMPC_OK) return rc;
  NEOMPC_CUDA(h, cudaMemcpyAsync(out, h->d_resp, n * sizeof(neompc_response), cudaMemcpyDeviceToHost, h->stream));
  if (plan_or_null)
    NEOMPC_CUDA(h, cudaMemcpyAsync(plan_or_null, h->d_plan, n * 3 * (size_t)h->params.control_steps * sizeof(float),
                                   cudaMemcpyDeviceToHost, h->stream));
  NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
  return check_state_errors(h);
}

int neompc_eval_objective(neompc_handle* h, const neompc_request* reqs, const float* u, size_t n, float* J,
                          float* grad_or_null) {
  if (!h || (n > 0 && (!reqs || !u || !J))) return fail(h, NEOMPC_ERR_INVALID, "null argument");
  if (n == 0) return NEOMPC_OK;
  NEOMPC_DEVICE(h);
  const size_t n3 = n * 3 * (size_t)h->params.control_steps;
  int rc = ensure_staging(h, n, true, false);
  if (rc != NEOMPC_OK) return rc;
  float *d_u = nullptr, *d_J = nullptr;
  NEOMPC_CUDA(h, cudaMalloc(&d_u, n3 * sizeof(float)));
  cudaError_t e = cudaMalloc(&d_J, n * sizeof(float));
  if (e != cudaSuccess) { cudaFree(d_u); return cuda_fail(h, e, "cudaMalloc"); }
  LaunchArgs a{};
  a.P = h->c; a.lut_cost = h->d_lut_cost; a.lut_flag = h->d_lut_flag;
  a.reqs = h->d_reqs; a.n = (unsigned)n; a.u = d_u; a.J = d_J; a.grad = grad_or_null ? h->d_plan : nullptr;
  a.stream = h->stream;
  e = cudaMemcpyAsync(h->d_reqs, reqs, n * sizeof(neompc_request), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_u, u, n3 * sizeof(float), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) { e = dispatch(h, true, a); h->launches += 1; }
  if (e == cudaSuccess) e = cudaMemcpyAsync(J, d_J, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && grad_or_null)
    e = cudaMemcpyAsync(grad_or_null, h->d_plan, n3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_u); cudaFree(d_J);
  if (e != cudaSuccess) return cuda_fail(h, e, "eval_objective");
  return NEOMPC_OK;
}

int neompc_local_plan_device(neompc_handle* h, const neompc_request* d_reqs, const float* d_plan, size_t n,
                             neompc_plan_pose* d_poses_out, void* stream) {
  if (!h || (n > 0 && (!d_reqs || !d_plan || !d_poses_out))) return fail(h, NEOMPC_ERR_INVALID, "null argument");
  if (n == 0) return NEOMPC_OK;
  if (n > 0xffffffffu) return fail(h, NEOMPC_ERR_INVALID, "batch too large");
  NEOMPC_DEVICE(h);
  const int N = h->params.control_steps;
  // self.dt = prediction_horizon / no_ctrl_steps (srv.py:137), in float64 like the reference
  const double dt = (double)h->params.prediction_horizon / (double)N;
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  cudaError_t e = launch_local_plan(d_reqs, d_plan, (unsigned)n, N, dt, d_poses_out, s);
  if (e != cudaSuccess) return cuda_fail(h, e, "local_plan kernel launch");
  h->launches += 1;
  return NEOMPC_OK;
}

int neompc_local_plan(neompc_handle* h, const neompc_request* reqs, const float* plan, size_t n,
                      neompc_plan_pose* poses_out) {
  if (!h || (n > 0 && (!reqs || !plan || !poses_out))) return fail(h, NEOMPC_ERR_INVALID, "null argument");
  if (n == 0) return NEOMPC_OK;
  NEOMPC_DEVICE(h);
  const size_t N = (size_t)h->params.control_steps;
  int rc = ensure_staging(h, n, true, false);
  if (rc != NEOMPC_OK) return rc;
  neompc_plan_pose* d_poses = nullptr;
  NEOMPC_CUDA(h, cudaMalloc(&d_poses, n * (N + 1) * sizeof(neompc_plan_pose)));
  cudaError_t e = cudaMemcpyAsync(h->d_reqs, reqs, n * sizeof(neompc_request), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_plan, plan, n * 3 * N * sizeof(float), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    rc = neompc_local_plan_device(h, h->d_reqs, h->d_plan, n, d_poses, h->stream);
    if (rc != NEOMPC_OK) { cudaFree(d_poses); return rc; }
    e = cudaMemcpyAsync(poses_out, d_poses, n * (N + 1) * sizeof(neompc_plan_pose), cudaMemcpyDeviceToHost, h->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_poses);
  if (e != cudaSuccess) return cuda_fail(h, e, "local_plan");
  return NEOMPC_OK;
}

int neompc_set_plan(neompc_handle* h, const double* xyyaw, size_t n_poses) {
  if (!h || !xyyaw || n_poses == 0 || n_poses > 0x7fffffffu) return fail(h, NEOMPC_ERR_INVALID, "plan must have >= 1 pose");
  NEOMPC_DEVICE(h);
  if (n_poses > h->path_cap) {
    NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->d_path) cudaFree(h->d_path);
    h->d_path = nullptr; h->path_cap = 0;
    NEOMPC_CUDA(h, cudaMalloc(&h->d_path, n_poses * 3 * sizeof(double)));
    h->path_cap = n_poses;
  }
  NEOMPC_CUDA(h, cudaMemcpyAsync(h->d_path, xyyaw, n_poses * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
  h->path_len = n_poses;
  return NEOMPC_OK;
}

int neompc_build_requests_device(neompc_handle* h, const neompc_carrot_params* cp, const neompc_robot_tick* d_ticks,
                                 size_t n, uint32_t first_instance_id, neompc_request* d_reqs_out,
                                 neompc_carrot_info* d_info_out, void* stream) {
  if (!h || !cp || (n > 0 && (!d_ticks || !d_reqs_out || !d_info_out))) return fail(h, NEOMPC_ERR_INVALID, "null argument");
  if (h->path_len == 0) return fail(h, NEOMPC_ERR_INVALID, "no plan set (neompc_set_plan)");
  if (!(cp->controller_frequency > 0.0f)) return fail(h, NEOMPC_ERR_INVALID, "controller_frequency must be > 0");
  if (n == 0) return NEOMPC_OK;
  NEOMPC_DEVICE(h);
  CarrotConst c{};
  c.plan = h->d_path;
  c.L = (unsigned)h->path_len;
  c.la_min = (double)cp->lookahead_dist_min;
  c.la_max = (double)cp->lookahead_dist_max;
  c.la_close = (double)cp->lookahead_dist_close_to_goal;
  c.control_interval = 1.0f / cp->controller_frequency;
  c.cells = h->c.cells;
  c.W = h->c.W; c.H = h->c.H;
  c.origin_x = h->c.origin_x; c.origin_y = h->c.origin_y;
  c.resolution = h->resolution;
  // max_transform_dist = max(size_x, size_y) * resolution / 2 (cpp:78-79); without a costmap every pose is inside
  c.max_transform_dist = h->c.cells ? (double)(h->c.W > h->c.H ? h->c.W : h->c.H) * h->resolution / 2.0 : 1.0e300;
  c.raw_table = h->d_raw_table;
  c.fp_n = h->c.fp_n;
  std::memcpy(c.fp_x, h->c.fp_x, sizeof(c.fp_x));
  std::memcpy(c.fp_y, h->c.fp_y, sizeof(c.fp_y));
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  cudaError_t e = launch_build_requests(c, d_ticks, (unsigned)n, first_instance_id, d_reqs_out, d_info_out, s);
  if (e != cudaSuccess) return cuda_fail(h, e, "build_requests kernel launch");
  h->launches += 1;
  return NEOMPC_OK;
}

int neompc_build_requests(neompc_handle* h, const neompc_carrot_params* cp, const neompc_robot_tick* ticks, size_t n,
                          uint32_t first_instance_id, neompc_request* reqs_out, neompc_carrot_info* info_out) {
  if (!h || !cp || (n > 0 && (!ticks || !reqs_out || !info_out))) return fail(h, NEOMPC_ERR_INVALID, "null argument");
  if (n == 0) return NEOMPC_OK;
  NEOMPC_DEVICE(h);
  int rc = ensure_staging(h, n, false, false);
  if (rc != NEOMPC_OK) return rc;
  if (n > h->cap_ticks) {
    if (h->d_ticks) cudaFree(h->d_ticks);
    if (h->d_info) cudaFree(h->d_info);
    h->d_ticks = nullptr; h->d_info = nullptr; h->cap_ticks = 0;
    NEOMPC_CUDA(h, cudaMalloc(&h->d_ticks, n * sizeof(neompc_robot_tick)));
    NEOMPC_CUDA(h, cudaMalloc(&h->d_info, n * sizeof(neompc_carrot_info)));
    h->cap_ticks = n;
  }
  NEOMPC_CUDA(h, cudaMemcpyAsync(h->d_ticks, ticks, n * sizeof(neompc_robot_tick), cudaMemcpyHostToDevice, h->stream));
  rc = neompc_build_requests_device(h, cp, h->d_ticks, n, first_instance_id, h->d_reqs, h->d_info, h->stream);
  if (rc != NEOMPC_OK) return rc;
  NEOMPC_CUDA(h, cudaMemcpyAsync(reqs_out, h->d_reqs, n * sizeof(neompc_request), cudaMemcpyDeviceToHost, h->stream));
  NEOMPC_CUDA(h, cudaMemcpyAsync(info_out, h->d_info, n * sizeof(neompc_carrot_info), cudaMemcpyDeviceToHost, h->stream));
  NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
  return NEOMPC_OK;
}

// ---- multi-GPU -------------------------------------------------------------------------------------
int neompc_comm_unique_id(unsigned char id[NEOMPC_COMM_ID_BYTES]) {
  if (!id) return NEOMPC_ERR_INVALID;
  std::string err;
  const NcclApi* api = load_nccl(err);
  if (!api) return fail(nullptr, NEOMPC_ERR_NCCL, err);
  static_assert(sizeof(ncclUniqueId) == NEOMPC_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId uid;
  ncclResult_t r = api->GetUniqueId(&uid);
  if (r != ncclSuccess) return nccl_fail(nullptr, api, r, "ncclGetUniqueId");
  std::memcpy(id, &uid, sizeof(uid));
  return NEOMPC_OK;
}

int neompc_comm_init(neompc_handle* h, const unsigned char id[NEOMPC_COMM_ID_BYTES], int n_ranks, int rank) {
  if (!h || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(h, NEOMPC_ERR_INVALID, "bad communicator arguments");
  std::string err;
  const NcclApi* api = load_nccl(err);
  if (!api) return fail(h, NEOMPC_ERR_NCCL, err);
  NEOMPC_DEVICE(h);
  int rc = comm_prepare(h, n_ranks, rank);
  if (rc != NEOMPC_OK) return rc;
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ncclResult_t r = comm_init_rank(api, &h->comm, n_ranks, uid, rank);
  if (r != ncclSuccess) { comm_release(h); return nccl_fail(h, api, r, "ncclCommInitRank"); }
  return NEOMPC_OK;
}

int neompc_comm_init_all(neompc_handle** handles, int n_handles) {
  if (!handles || n_handles < 1) return NEOMPC_ERR_INVALID;
  for (int i = 0; i < n_handles; ++i) {
    if (!handles[i]) return NEOMPC_ERR_INVALID;
    for (int j = 0; j < i; ++j)
      if (handles[j]->device == handles[i]->device) return fail(handles[0], NEOMPC_ERR_INVALID, "two handles of a fleet share a device");
  }
  neompc_handle* h0 = handles[0];
  std::string err;
  const NcclApi* api = load_nccl(err);
  if (!api) return fail(h0, NEOMPC_ERR_NCCL, err);
  ncclUniqueId uid;
  ncclResult_t r = api->GetUniqueId(&uid);
  if (r != ncclSuccess) return nccl_fail(h0, api, r, "ncclGetUniqueId");
  int prev = -1;
  cudaGetDevice(&prev);
  int rc = NEOMPC_OK;
  for (int i = 0; i < n_handles && rc == NEOMPC_OK; ++i) {
    cudaSetDevice(handles[i]->device);
    rc = comm_prepare(handles[i], n_handles, i);
  }
  if (rc == NEOMPC_OK) {
    // one process, several devices: the ranks are initialised inside one NCCL group (what ncclCommInitAll does)
    r = api->GroupStart();
    for (int i = 0; i < n_handles && r == ncclSuccess; ++i) {
      cudaSetDevice(handles[i]->device);
      r = comm_init_rank(api, &handles[i]->comm, n_handles, uid, i);
    }
    ncclResult_t re = api->GroupEnd();
    if (r == ncclSuccess) r = re;
    if (r != ncclSuccess) {
      for (int i = 0; i < n_handles; ++i) { cudaSetDevice(handles[i]->device); comm_release(handles[i]); }
      rc = nccl_fail(h0, api, r, "ncclCommInitRank (group)");
    }
  }
  if (prev >= 0) cudaSetDevice(prev);
  return rc;
}

int neompc_comm_destroy(neompc_handle* h) {
  if (!h) return NEOMPC_ERR_INVALID;
  NEOMPC_DEVICE(h);
  comm_release(h);
  return NEOMPC_OK;
}

int neompc_comm_info(const neompc_handle* h, int* n_ranks, int* rank) {
  if (!h) return NEOMPC_ERR_INVALID;
  if (n_ranks) *n_ranks = h->n_ranks;
  if (rank) *rank = h->rank;
  return NEOMPC_OK;
}

size_t neompc_shard_rows(size_t n_total, int n_ranks) {
  return n_ranks > 0 ? (n_total + (size_t)n_ranks - 1) / (size_t)n_ranks : 0;
}

// enqueue only (no group call, no synchronise): solve of the local shard into its slot + event + all-gather
static int solve_gather_enqueue(neompc_handle* h, const NcclApi* api, const neompc_request* d_reqs, size_t n_local,
                                size_t shard_rows, neompc_response* d_out, float* d_twist_all, cudaStream_t s, bool gather) {
  if (n_local > shard_rows) return fail(h, NEOMPC_ERR_INVALID, "local shard larger than shard_rows");
  float* slot = d_twist_all + (size_t)h->rank * shard_rows * 3;
  if (n_local < shard_rows)      // rows of the slot past the shard travel with the collective: defined contents
    NEOMPC_CUDA(h, cudaMemsetAsync(slot + n_local * 3, 0, (shard_rows - n_local) * 3 * sizeof(float), s));
  int rc = do_solve_device(h, d_reqs, n_local, d_out, slot, nullptr, s);
  if (rc != NEOMPC_OK) return rc;
  if (h->n_ranks > 1 || h->comm) {
    NEOMPC_CUDA(h, cudaEventRecord(h->ev_solved, s));
    NEOMPC_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_solved, 0));
    if (gather) {
      ncclResult_t r = api->AllGather(slot, d_twist_all, shard_rows * 3, ncclFloat, h->comm, h->comm_stream);
      if (r != ncclSuccess) return nccl_fail(h, api, r, "ncclAllGather");
      NEOMPC_CUDA(h, cudaEventRecord(h->ev_gathered[h->gathers & 1u], h->comm_stream));
      h->gathers += 1;
    }
  }
  return NEOMPC_OK;
}

int neompc_solve_gather_device(neompc_handle* h, const neompc_request* d_reqs, size_t n_local, size_t shard_rows,
                               neompc_response* d_out, float* d_twist_all, void* stream) {
  if (!h || !d_twist_all || (n_local > 0 && (!d_reqs || !d_out))) return fail(h, NEOMPC_ERR_INVALID, "null argument");
  if (!h->comm) return fail(h, NEOMPC_ERR_INVALID, "no communicator (neompc_comm_init / neompc_comm_init_all)");
  std::string err;
  const NcclApi* api = load_nccl(err);
  if (!api) return fail(h, NEOMPC_ERR_NCCL, err);
  NEOMPC_DEVICE(h);
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  return solve_gather_enqueue(h, api, d_reqs, n_local, shard_rows, d_out, d_twist_all, s, true);
}

int neompc_gather_wait(neompc_handle* h, void* stream, int age) {
  if (!h || age < 0 || age > 1) return NEOMPC_ERR_INVALID;
  if (!h->comm || h->gathers <= (unsigned)age) return NEOMPC_OK;          // no such gather yet
  NEOMPC_DEVICE(h);
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  NEOMPC_CUDA(h, cudaStreamWaitEvent(s, h->ev_gathered[(h->gathers - 1u - (unsigned)age) & 1u], 0));
  return NEOMPC_OK;
}

int neompc_fleet_solve(neompc_handle** handles, int n_handles, const neompc_request* reqs, size_t n, float* twist_out,
                       neompc_response* out_or_null) {
  if (!handles || n_handles < 1 || (n > 0 && (!reqs || !twist_out))) return NEOMPC_ERR_INVALID;
  neompc_handle* h0 = handles[0];
  for (int i = 0; i < n_handles; ++i)
    if (!handles[i] || !handles[i]->comm || handles[i]->n_ranks != n_handles || handles[i]->rank != i)
      return fail(h0, NEOMPC_ERR_INVALID, "handles are not the ranks 0..n-1 of one communicator (neompc_comm_init_all)");
  if (n == 0) return NEOMPC_OK;
  std::string err;
  const NcclApi* api = load_nccl(err);
  if (!api) return fail(h0, NEOMPC_ERR_NCCL, err);
  const size_t rows = neompc_shard_rows(n, n_handles);
  int prev = -1;
  cudaGetDevice(&prev);
  int rc = NEOMPC_OK;
  // 1. per device: staging, H2D of the shard, solve into the shard's slot of that device's gather buffer
  for (int i = 0; i < n_handles && rc == NEOMPC_OK; ++i) {
    neompc_handle* h = handles[i];
    cudaSetDevice(h->device);
    const size_t lo = std::min(n, (size_t)i * rows), hi = std::min(n, (size_t)(i + 1) * rows), cnt = hi - lo;
    rc = check_unique_ids(h, reqs + lo, cnt);
    if (rc == NEOMPC_OK) rc = ensure_staging(h, rows, false, false);
    if (rc == NEOMPC_OK && (size_t)n_handles * rows * 3 > h->cap_gather) {
      cudaFree(h->d_gather);
      h->d_gather = nullptr; h->cap_gather = 0;
      cudaError_t e = cudaMalloc(&h->d_gather, (size_t)n_handles * rows * 3 * sizeof(float));
      if (e != cudaSuccess) rc = cuda_fail(h, e, "cudaMalloc(gather)"); else h->cap_gather = (size_t)n_handles * rows * 3;
    }
    if (rc != NEOMPC_OK) break;
    if (cnt > 0) {
      cudaError_t e = cudaMemcpyAsync(h->d_reqs, reqs + lo, cnt * sizeof(neompc_request), cudaMemcpyHostToDevice, h->stream);
      if (e != cudaSuccess) { rc = cuda_fail(h, e, "H2D requests"); break; }
    }
    rc = solve_gather_enqueue(h, api, h->d_reqs, cnt, rows, h->d_resp, h->d_gather, h->stream, false);
    if (rc == NEOMPC_OK && out_or_null && cnt > 0) {
      cudaError_t e = cudaMemcpyAsync(out_or_null + lo, h->d_resp, cnt * sizeof(neompc_response), cudaMemcpyDeviceToHost, h->stream);
      if (e != cudaSuccess) rc = cuda_fail(h, e, "D2H responses");
    }
  }
  // 2. the single collective of the path, all ranks of this process in one NCCL group
  if (rc == NEOMPC_OK) {
    ncclResult_t r = api->GroupStart();
    for (int i = 0; i < n_handles && r == ncclSuccess; ++i) {
      neompc_handle* h = handles[i];
      cudaSetDevice(h->device);
      r = api->AllGather(h->d_gather + (size_t)i * rows * 3, h->d_gather, rows * 3, ncclFloat, h->comm, h->comm_stream);
    }
    ncclResult_t re = api->GroupEnd();
    if (r == ncclSuccess) r = re;
    if (r != ncclSuccess) rc = nccl_fail(h0, api, r, "ncclAllGather (group)");
  }
  // 3. every device now holds all twists; rank 0's copy goes to the host
  if (rc == NEOMPC_OK) {
    cudaSetDevice(h0->device);
    cudaError_t e = cudaMemcpyAsync(twist_out, h0->d_gather, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, h0->comm_stream);
    if (e != cudaSuccess) rc = cuda_fail(h0, e, "D2H twists");
  }
  for (int i = 0; i < n_handles; ++i) {
    neompc_handle* h = handles[i];
    cudaSetDevice(h->device);
    cudaError_t e1 = cudaStreamSynchronize(h->stream), e2 = cudaStreamSynchronize(h->comm_stream);
    if (rc == NEOMPC_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) rc = cuda_fail(h, e1 != cudaSuccess ? e1 : e2, "fleet synchronise");
    if (rc == NEOMPC_OK) rc = check_state_errors(h);
    if (rc != NEOMPC_OK && h != h0) h0->err = h->err;
  }
  if (prev >= 0) cudaSetDevice(prev);
  return rc;
}

int neompc_fleet_get_gathered(neompc_handle* h, size_t n, float* twist_out) {
  if (!h || !twist_out) return NEOMPC_ERR_INVALID;
  if (n * 3 > h->cap_gather) return fail(h, NEOMPC_ERR_INVALID, "no gathered batch of that size on this handle");
  NEOMPC_DEVICE(h);
  NEOMPC_CUDA(h, cudaMemcpyAsync(twist_out, h->d_gather, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->comm_stream));
  NEOMPC_CUDA(h, cudaStreamSynchronize(h->comm_stream));
  return NEOMPC_OK;
}

// One control tick for n robots, fused: carrot selection + request construction (cpp:66-246) feeding the solve
// (srv.py:349-403) on the device, one synchronise.  What NeoMpcPlanner::computeVelocityCommands does per call.
int neompc_control_tick(neompc_handle* h, const neompc_carrot_params* cp, const neompc_robot_tick* ticks, size_t n,
                        uint32_t first_instance_id, neompc_response* out, neompc_carrot_info* info_out,
                        neompc_request* reqs_out_or_null, float* plan_or_null) {
  if (!h || !cp || (n > 0 && (!ticks || !out || !info_out))) return fail(h, NEOMPC_ERR_INVALID, "null argument");
  if (n == 0) return NEOMPC_OK;
  NEOMPC_DEVICE(h);
  int rc = ensure_staging(h, n, plan_or_null != nullptr, false);
  if (rc != NEOMPC_OK) return rc;
  const size_t n3 = n * 3 * (size_t)h->params.control_steps;
  const bool small = n <= kMailboxRequests;
  if (!small && n > h->cap_ticks) {
    if (h->d_ticks) cudaFree(h->d_ticks);
    if (h->d_info) cudaFree(h->d_info);
    h->d_ticks = nullptr; h->d_info = nullptr; h->cap_ticks = 0;
    NEOMPC_CUDA(h, cudaMalloc(&h->d_ticks, n * sizeof(neompc_robot_tick)));
    NEOMPC_CUDA(h, cudaMalloc(&h->d_info, n * sizeof(neompc_carrot_info)));
    h->cap_ticks = n;
  }
  const neompc_robot_tick* d_ticks = h->d_ticks;
  neompc_carrot_info* d_info = h->d_info;
  neompc_response* d_resp = h->d_resp;
  float* d_plan = plan_or_null ? h->d_plan : nullptr;
  if (small) {                                   // controller tick: mapped mailboxes, no staged copies
    std::memcpy(h->mb_ticks, ticks, n * sizeof(neompc_robot_tick));
    d_ticks = h->mb_ticks; d_info = h->mb_info; d_resp = h->mb_resp;
    d_plan = plan_or_null ? h->mb_plan : nullptr;
  } else {
    NEOMPC_CUDA(h, cudaMemcpyAsync(h->d_ticks, ticks, n * sizeof(neompc_robot_tick), cudaMemcpyHostToDevice, h->stream));
  }
  rc = neompc_build_requests_device(h, cp, d_ticks, n, first_instance_id, h->d_reqs, d_info, h->stream);
  if (rc != NEOMPC_OK) return rc;
  rc = do_solve_device(h, h->d_reqs, n, d_resp, nullptr, d_plan, h->stream);
  if (rc != NEOMPC_OK) return rc;
  if (!small) {
    NEOMPC_CUDA(h, cudaMemcpyAsync(out, h->d_resp, n * sizeof(neompc_response), cudaMemcpyDeviceToHost, h->stream));
    NEOMPC_CUDA(h, cudaMemcpyAsync(info_out, h->d_info, n * sizeof(neompc_carrot_info), cudaMemcpyDeviceToHost, h->stream));
    if (plan_or_null) NEOMPC_CUDA(h, cudaMemcpyAsync(plan_or_null, h->d_plan, n3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  }
  if (reqs_out_or_null)
    NEOMPC_CUDA(h, cudaMemcpyAsync(reqs_out_or_null, h->d_reqs, n * sizeof(neompc_request), cudaMemcpyDeviceToHost, h->stream));
  NEOMPC_CUDA(h, cudaStreamSynchronize(h->stream));
  if (small) {
    std::memcpy(out, h->mb_resp, n * sizeof(neompc_response));
    std::memcpy(info_out, h->mb_info, n * sizeof(neompc_carrot_info));
    if (plan_or_null) std::memcpy(plan_or_null, h->mb_plan, n3 * sizeof(float));
  }
  return check_state_errors(h);
}

uint64_t neompc_launch_count(const neompc_handle* h) { return h ? h->launches : 0; }

int neompc_last_host_path(const neompc_handle* h) { return h ? h->last_host_path : NEOMPC_HOST_PATH_NONE; }

int neompc_get_tiling(const neompc_handle* h, int* lanes_per_instance, int* steps_per_lane) {
  if (!h) return NEOMPC_ERR_INVALID;
  if (lanes_per_instance) *lanes_per_instance = h->G;
  if (steps_per_lane) *steps_per_lane = h->S;
  return NEOMPC_OK;
}

int neompc_get_tiling_for(const neompc_handle* h, size_t n, int* lanes_per_instance, int* steps_per_lane) {
  if (!h) return NEOMPC_ERR_INVALID;
  const bool latency = n * (size_t)h->Gl <= (size_t)h->sm_count * 16u * 32u;      // the rule of dispatch()
  if (lanes_per_instance) *lanes_per_instance = latency ? h->Gl : h->G;
  if (steps_per_lane) *steps_per_lane = latency ? h->Sl : h->S;
  return NEOMPC_OK;
}

int neompc_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return NEOMPC_ERR_INVALID;
  return cudaHostAlloc(ptr, bytes, cudaHostAllocDefault) == cudaSuccess ? NEOMPC_OK : NEOMPC_ERR_CUDA;
}

int neompc_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? NEOMPC_OK : NEOMPC_ERR_CUDA; }

}  // extern "C"

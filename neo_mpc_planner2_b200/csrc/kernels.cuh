// kernels.cuh — __global__ wrappers around the lane-group solver core (mpc_core.cuh) and their launchers.
// One translation unit per lanes-per-instance value G (solve_g1.cu ... solve_g32.cu) instantiates S = 1..6.
#pragma once

#include <cuda_runtime.h>

#include "mpc_core.cuh"

namespace neompc {

constexpr int kBlockThreads = 128;      // 4 warps; 128/G instances per block
constexpr int kMaxStepsPerLane = 6;

struct LaunchArgs {
  SolverConst P;
  const float* lut_cost;       // device [kTableSize]
  const uint8_t* lut_flag;     // device [kTableSize]
  const neompc_request* reqs;  // device [n]
  unsigned n;
  neompc_response* out;        // device [n]
  float* twist;                // device [3n] or null
  float* plan;                 // device [n*3N] or null
  // eval only
  const float* u;              // device [n*3N]
  float* J;                    // device [n]
  float* grad;                 // device [n*3N] or null
  cudaStream_t stream;
};

// cost tables staged once per block in shared memory
struct SmemTables {
  float cost[kTableSize];
  uint8_t flag[kTableSize + 2];
};

__device__ __forceinline__ void load_tables(SmemTables& st, const float* lut_cost, const uint8_t* lut_flag) {
  for (int i = threadIdx.x; i < kTableSize; i += blockDim.x) {
    st.cost[i] = __ldg(lut_cost + i);
    st.flag[i] = __ldg(lut_flag + i);
  }
  __syncthreads();
}

// the 64-byte request record as four 16-byte loads (all lanes of a group read the same line: one broadcast)
__device__ __forceinline__ neompc_request load_request(const neompc_request* reqs, unsigned idx, bool valid) {
  neompc_request rq;
  float4* dst = reinterpret_cast<float4*>(&rq);
  if (valid) {
    const float4* src = reinterpret_cast<const float4*>(reqs + idx);
    dst[0] = __ldg(src + 0); dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2); dst[3] = __ldg(src + 3);
  } else {
    dst[0] = dst[1] = dst[2] = dst[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    rq.instance_id = NEOMPC_STATELESS;
  }
  return rq;
}

// resident blocks per SM the register allocator is asked to allow (65536 regs / (128 threads * blocks))
#ifndef NEOMPC_MINBLOCKS_S3
#define NEOMPC_MINBLOCKS_S3 4
#endif
constexpr int min_blocks_for(int S) { return S == 2 ? 5 : S == 3 ? NEOMPC_MINBLOCKS_S3 : S == 1 ? 4 : S == 4 ? 3 : 2; }

template <int G, int S>
__global__ void __launch_bounds__(kBlockThreads, min_blocks_for(S))
solve_kernel(const __grid_constant__ SolverConst P, const float* __restrict__ lut_cost,
             const uint8_t* __restrict__ lut_flag, const neompc_request* __restrict__ reqs, unsigned n,
             neompc_response* __restrict__ out, float* __restrict__ twist, float* __restrict__ plan) {
  extern __shared__ float hist_smem[];
  __shared__ SmemTables st;
  load_tables(st, lut_cost, lut_flag);
  CostTables T{st.cost, st.flag};
  constexpr int kInstPerBlock = kBlockThreads / G;
  const unsigned inst = blockIdx.x * kInstPerBlock + threadIdx.x / G;
  const int lg = threadIdx.x % G;
  const bool valid = inst < n;
  const neompc_request rq = load_request(reqs, inst, valid);
  solve_instance<G, S>(P, T, rq, valid, lg, hist_smem + threadIdx.x, kBlockThreads,
                       valid ? out + inst : nullptr,
                       (valid && twist != nullptr) ? twist + 3 * (size_t)inst : nullptr,
                       (valid && plan != nullptr) ? plan + (size_t)inst * 3 * P.N : nullptr);
}

template <int G, int S>
__global__ void __launch_bounds__(kBlockThreads)
eval_kernel(const __grid_constant__ SolverConst P, const float* __restrict__ lut_cost,
            const uint8_t* __restrict__ lut_flag, const neompc_request* __restrict__ reqs, unsigned n,
            const float* __restrict__ u, float* __restrict__ J, float* __restrict__ grad) {
  __shared__ SmemTables st;
  load_tables(st, lut_cost, lut_flag);
  CostTables T{st.cost, st.flag};
  constexpr int kInstPerBlock = kBlockThreads / G;
  const unsigned inst = blockIdx.x * kInstPerBlock + threadIdx.x / G;
  const int lg = threadIdx.x % G;
  const bool valid = inst < n;
  const neompc_request rq = load_request(reqs, inst, valid);
  eval_instance<G, S>(P, T, rq, valid, lg, valid ? u + (size_t)inst * 3 * P.N : nullptr,
                      valid ? J + inst : nullptr,
                      (valid && grad != nullptr) ? grad + (size_t)inst * 3 * P.N : nullptr);
}

template <int G, int S>
cudaError_t launch_solve_gs(const LaunchArgs& a) {
  const size_t smem = (size_t)kBlockThreads * hist_floats_per_lane<S>(a.P.m) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(solve_kernel<G, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  constexpr int kInstPerBlock = kBlockThreads / G;
  const unsigned grid = (a.n + kInstPerBlock - 1) / kInstPerBlock;
  solve_kernel<G, S><<<grid, kBlockThreads, smem, a.stream>>>(a.P, a.lut_cost, a.lut_flag, a.reqs, a.n, a.out,
                                                               a.twist, a.plan);
  return cudaGetLastError();
}

template <int G, int S>
cudaError_t launch_eval_gs(const LaunchArgs& a) {
  constexpr int kInstPerBlock = kBlockThreads / G;
  const unsigned grid = (a.n + kInstPerBlock - 1) / kInstPerBlock;
  eval_kernel<G, S><<<grid, kBlockThreads, 0, a.stream>>>(a.P, a.lut_cost, a.lut_flag, a.reqs, a.n, a.u, a.J, a.grad);
  return cudaGetLastError();
}

template <int G>
cudaError_t launch_for_g(bool eval, int S, const LaunchArgs& a) {
  switch (S) {
    case 1: return eval ? launch_eval_gs<G, 1>(a) : launch_solve_gs<G, 1>(a);
    case 2: return eval ? launch_eval_gs<G, 2>(a) : launch_solve_gs<G, 2>(a);
    case 3: return eval ? launch_eval_gs<G, 3>(a) : launch_solve_gs<G, 3>(a);
    case 4: return eval ? launch_eval_gs<G, 4>(a) : launch_solve_gs<G, 4>(a);
    case 5: return eval ? launch_eval_gs<G, 5>(a) : launch_solve_gs<G, 5>(a);
    case 6: return eval ? launch_eval_gs<G, 6>(a) : launch_solve_gs<G, 6>(a);
    default: return cudaErrorInvalidValue;
  }
}

// defined in solve_g*.cu
cudaError_t launch_g1(bool eval, int S, const LaunchArgs& a);
cudaError_t launch_g2(bool eval, int S, const LaunchArgs& a);
cudaError_t launch_g4(bool eval, int S, const LaunchArgs& a);
cudaError_t launch_g8(bool eval, int S, const LaunchArgs& a);
cudaError_t launch_g16(bool eval, int S, const LaunchArgs& a);
cudaError_t launch_g32(bool eval, int S, const LaunchArgs& a);

}  // namespace neompc

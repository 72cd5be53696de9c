// kernels.cuh — __global__ wrappers around the lane-group solver core (mpc_core.cuh) and their launchers.
// One translation unit per lanes-per-instance value G (solve_g1.cu ... solve_g32.cu; G = 1,2,3,4,5,6,8,10,16,32)
// instantiates S = 1..4, each with and without the opt-in objective extensions.
#pragma once

#include <cuda_runtime.h>

#include "mpc_core.cuh"

namespace neompc {

#ifndef NEOMPC_BLOCK_THREADS
#define NEOMPC_BLOCK_THREADS 64
#endif
constexpr int kBlockThreads = NEOMPC_BLOCK_THREADS;      // kBlockThreads/G instances per block
constexpr int kMaxStepsPerLane = 4;

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
  unsigned tiling_n;           // batch size the lane tiling is chosen for (0: n); the chunked host path passes the total
  bool no_full;                // test knob NEOMPC_NO_FULL: never the full-horizon instantiation (launch_solve_gs)
};

// thread -> (instance slot of the block, lane inside the group): 32/G groups per warp, leftover lanes idle
template <int G>
struct GroupMap {
  static constexpr int kPerWarp = 32 / G;
  static constexpr int kPerBlock = kPerWarp * (kBlockThreads / 32);
  int slot, lg;
  bool lane_ok;
  __device__ __forceinline__ GroupMap() {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G;
    lg = lane - grp * G;
    lane_ok = grp < kPerWarp;
    slot = warp * kPerWarp + (lane_ok ? grp : kPerWarp - 1);
  }
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// keeps a value in a register: the compiler cannot re-derive it (here: the shared-window address of a static array,
// which it would otherwise rebuild with S2UR + UMOV + ULEA at every use)
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) { asm volatile("" : "+r"(v)); return v; }

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
#ifndef NEOMPC_MINBLOCKS_S2
#define NEOMPC_MINBLOCKS_S2 4
#endif
// NEOMPC_MINBLOCKS_RAW_S3: resident blocks of kBlockThreads (64) per SM the S = 3 kernels are compiled for, unscaled.
// 7 instead of 8: ptxas still allocates 128 registers (8 blocks stay resident) but schedules differently — measured
// C3 0.3965 -> 0.3926 ms, C4 1.488 -> 1.420 ms; 9 blocks (96 registers, 408 B of spills): 0.518 / 1.694 ms
// (profiles/minblocks_sweep_r1.txt).
#ifndef NEOMPC_MINBLOCKS_RAW_S3
#define NEOMPC_MINBLOCKS_RAW_S3 6
#endif
// resident 128-thread-equivalents per SM the register allocator must allow, scaled to the block size
constexpr int min_blocks_for(int S) {
  if (S == 3 && kBlockThreads == 64) return NEOMPC_MINBLOCKS_RAW_S3;
  return (S == 2 ? NEOMPC_MINBLOCKS_S2 : S == 3 ? 4 : S == 1 ? 4 : 3) * (128 / kBlockThreads);
}

// ---- TMA (bulk async copy) staging of the block's request tile ------------------------------------------------
// The 128/G request records of a block are contiguous in HBM (64 B each): one elected thread arms an mbarrier with the
// byte count and issues ONE cp.async.bulk (SASS: UBLKCP) global -> shared; every lane then reads its group's record
// from shared memory instead of issuing four 16-byte global loads per lane.

__device__ __forceinline__ void tma_stage_requests(neompc_request* s_req, unsigned long long* mbar,
                                                   const neompc_request* g_req, unsigned count) {
  const uint32_t bar = smem_addr(mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = count * (uint32_t)sizeof(neompc_request);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(s_req)), "l"(g_req), "r"(bytes), "r"(bar)
                 : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar)
        : "memory");
  }
}

template <int G, int S, bool X, bool F>
__global__ void __launch_bounds__(kBlockThreads, min_blocks_for(S))
solve_kernel(const __grid_constant__ SolverConst P, const float* __restrict__ lut_cost,
             const uint8_t* __restrict__ lut_flag, const neompc_request* __restrict__ reqs, unsigned n,
             neompc_response* __restrict__ out, float* __restrict__ twist, float* __restrict__ plan) {
  extern __shared__ float hist_smem[];
  __shared__ SmemTables st;
  load_tables(st, lut_cost, lut_flag);
  CostTables T{st.cost, st.flag, opaque_u32(smem_addr(st.cost))};
  constexpr int kInstPerBlock = GroupMap<G>::kPerBlock;
  __shared__ alignas(128) neompc_request s_req[kInstPerBlock];
  __shared__ alignas(8) unsigned long long s_mbar;
  const unsigned first = blockIdx.x * kInstPerBlock;
  const unsigned in_block = n - first < (unsigned)kInstPerBlock ? n - first : (unsigned)kInstPerBlock;
  tma_stage_requests(s_req, &s_mbar, reqs + first, in_block);
  // The record stays in shared memory: the solver core reads fields from there when it needs them (prologue and
  // epilogue) instead of holding 16 registers for the whole solve.  Slots past the end of the batch are zeroed.
  if (in_block < (unsigned)kInstPerBlock) {
    for (unsigned sl = in_block + threadIdx.x; sl < (unsigned)kInstPerBlock; sl += blockDim.x) {
      float4* dst = reinterpret_cast<float4*>(&s_req[sl]);
      dst[0] = dst[1] = dst[2] = dst[3] = make_float4(0.f, 0.f, 0.f, 0.f);
      s_req[sl].instance_id = NEOMPC_STATELESS;
    }
    __syncthreads();
  }
  const GroupMap<G> gm;
  const unsigned inst = first + gm.slot;
  const bool valid = gm.lane_ok && inst < n;
  const neompc_request& rq = s_req[gm.slot];
  solve_instance<G, S, X, F>(P, T, rq, valid, gm.lg, hist_smem + threadIdx.x, kBlockThreads,
                       valid ? out + inst : nullptr,
                       (valid && twist != nullptr) ? twist + 3 * (size_t)inst : nullptr,
                       (valid && plan != nullptr) ? plan + (size_t)inst * 3 * P.N : nullptr);
}

template <int G, int S, bool X>
__global__ void __launch_bounds__(kBlockThreads)
eval_kernel(const __grid_constant__ SolverConst P, const float* __restrict__ lut_cost,
            const uint8_t* __restrict__ lut_flag, const neompc_request* __restrict__ reqs, unsigned n,
            const float* __restrict__ u, float* __restrict__ J, float* __restrict__ grad) {
  __shared__ SmemTables st;
  load_tables(st, lut_cost, lut_flag);
  CostTables T{st.cost, st.flag, opaque_u32(smem_addr(st.cost))};
  constexpr int kInstPerBlock = GroupMap<G>::kPerBlock;
  const GroupMap<G> gm;
  const unsigned inst = blockIdx.x * kInstPerBlock + gm.slot;
  const int lg = gm.lg;
  const bool valid = gm.lane_ok && inst < n;
  const neompc_request rq = load_request(reqs, inst, valid);
  eval_instance<G, S, X>(P, T, rq, valid, lg, valid ? u + (size_t)inst * 3 * P.N : nullptr,
                      valid ? J + inst : nullptr,
                      (valid && grad != nullptr) ? grad + (size_t)inst * 3 * P.N : nullptr);
}

template <int G, int S, bool X, bool F>
cudaError_t launch_solve_gsf(const LaunchArgs& a) {
  const size_t smem = (size_t)kBlockThreads * hist_floats_per_lane<S>(a.P.m) * sizeof(float);
  constexpr int kInstPerBlock = GroupMap<G>::kPerBlock;
  const unsigned blocks_needed = (a.n + kInstPerBlock - 1) / kInstPerBlock;
  cudaError_t e = cudaFuncSetAttribute(solve_kernel<G, S, X, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  solve_kernel<G, S, X, F><<<blocks_needed, kBlockThreads, smem, a.stream>>>(a.P, a.lut_cost, a.lut_flag, a.reqs, a.n, a.out,
                                                                          a.twist, a.plan);
  return cudaGetLastError();
}

// The reference fast path has a second instantiation for horizons that fill the lane group exactly (G * S == control_steps:
// C3's (5,2), C4's (10,2)) on a handle with a costmap: no padded steps, so the per-step masks of cost() and backward() fold
// away, and so do the costmap-present test and the bounds-checked sampling path.  Same source-level arithmetic; the compiler
// contracts a few more multiply-adds without the selects, so the two agree to rounding (NEOMPC_NO_FULL forces F = false).
template <int G, int S, bool X>
cudaError_t launch_solve_gs(const LaunchArgs& a) {
  if (!X && G > 1 && a.P.N == G * S && a.P.cells4 != nullptr && a.P.pad_ok && !a.no_full) return launch_solve_gsf<G, S, false, true>(a);
  return launch_solve_gsf<G, S, X, false>(a);
}

template <int G, int S, bool X>
cudaError_t launch_eval_gs(const LaunchArgs& a) {
  constexpr int kInstPerBlock = GroupMap<G>::kPerBlock;
  const unsigned grid = (a.n + kInstPerBlock - 1) / kInstPerBlock;
  eval_kernel<G, S, X><<<grid, kBlockThreads, 0, a.stream>>>(a.P, a.lut_cost, a.lut_flag, a.reqs, a.n, a.u, a.J, a.grad);
  return cudaGetLastError();
}

// ext: the handle's parameters select an opt-in objective extension (moving footprint / bilinear costmap)
template <int G, bool X>
cudaError_t launch_for_gx(bool eval, int S, const LaunchArgs& a) {
  switch (S) {
    case 1: return eval ? launch_eval_gs<G, 1, X>(a) : launch_solve_gs<G, 1, X>(a);
    case 2: return eval ? launch_eval_gs<G, 2, X>(a) : launch_solve_gs<G, 2, X>(a);
    case 3: return eval ? launch_eval_gs<G, 3, X>(a) : launch_solve_gs<G, 3, X>(a);
    case 4: return eval ? launch_eval_gs<G, 4, X>(a) : launch_solve_gs<G, 4, X>(a);
    default: return cudaErrorInvalidValue;
  }
}

template <int G>
cudaError_t launch_for_g(bool eval, int S, bool ext, const LaunchArgs& a) {
  return ext ? launch_for_gx<G, true>(eval, S, a) : launch_for_gx<G, false>(eval, S, a);
}

// defined in solve_g*.cu
cudaError_t launch_g1(bool eval, int S, bool ext, const LaunchArgs& a);
cudaError_t launch_g2(bool eval, int S, bool ext, const LaunchArgs& a);
cudaError_t launch_g3(bool eval, int S, bool ext, const LaunchArgs& a);
cudaError_t launch_g4(bool eval, int S, bool ext, const LaunchArgs& a);
cudaError_t launch_g5(bool eval, int S, bool ext, const LaunchArgs& a);
cudaError_t launch_g6(bool eval, int S, bool ext, const LaunchArgs& a);
cudaError_t launch_g8(bool eval, int S, bool ext, const LaunchArgs& a);
cudaError_t launch_g10(bool eval, int S, bool ext, const LaunchArgs& a);
cudaError_t launch_g16(bool eval, int S, bool ext, const LaunchArgs& a);
cudaError_t launch_g32(bool eval, int S, bool ext, const LaunchArgs& a);

}  // namespace neompc

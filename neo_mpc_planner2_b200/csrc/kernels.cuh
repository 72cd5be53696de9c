// kernels.cuh — __global__ wrappers around the lane-group solver core (mpc_core.cuh) and their launchers.
// One translation unit per lanes-per-instance value G (solve_g1.cu ... solve_g32.cu) instantiates S = 1..6.
#pragma once

#include <cuda_runtime.h>

#include "mpc_core.cuh"

namespace neompc {

#ifndef NEOMPC_BLOCK_THREADS
#define NEOMPC_BLOCK_THREADS 64
#endif
constexpr int kBlockThreads = NEOMPC_BLOCK_THREADS;      // kBlockThreads/G instances per block
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
  unsigned* queue_counter;     // device word for the persistent kernel's work queue; null = one instance per group
  int sm_count;
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
// resident 128-thread-equivalents per SM the register allocator must allow, scaled to the block size
constexpr int min_blocks_for(int S) {
  return (S == 2 ? 5 : S == 3 ? NEOMPC_MINBLOCKS_S3 : S == 1 ? 4 : S == 4 ? 3 : 2) * (128 / kBlockThreads);
}

// ---- TMA (bulk async copy) staging of the block's request tile ------------------------------------------------
// The 128/G request records of a block are contiguous in HBM (64 B each): one elected thread arms an mbarrier with the
// byte count and issues ONE cp.async.bulk (SASS: UBLKCP) global -> shared; every lane then reads its group's record
// from shared memory instead of issuing four 16-byte global loads per lane.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
  __shared__ alignas(128) neompc_request s_req[kInstPerBlock];
  __shared__ alignas(8) unsigned long long s_mbar;
  const unsigned first = blockIdx.x * kInstPerBlock;
  const unsigned in_block = n - first < (unsigned)kInstPerBlock ? n - first : (unsigned)kInstPerBlock;
  tma_stage_requests(s_req, &s_mbar, reqs + first, in_block);
  const unsigned inst = first + threadIdx.x / G;
  const int lg = threadIdx.x % G;
  const bool valid = inst < n;
  // The record stays in shared memory: the solver core reads fields from there when it needs them (prologue and
  // epilogue) instead of holding 16 registers for the whole solve.  Slots past the end of the batch are zeroed.
  if (in_block < (unsigned)kInstPerBlock) {
    if (!valid && lg == 0) {
      float4* dst = reinterpret_cast<float4*>(&s_req[threadIdx.x / G]);
      dst[0] = dst[1] = dst[2] = dst[3] = make_float4(0.f, 0.f, 0.f, 0.f);
      s_req[threadIdx.x / G].instance_id = NEOMPC_STATELESS;
    }
    __syncthreads();
  }
  const neompc_request& rq = s_req[threadIdx.x / G];
  solve_instance<G, S>(P, T, rq, valid, lg, hist_smem + threadIdx.x, kBlockThreads,
                       valid ? out + inst : nullptr,
                       (valid && twist != nullptr) ? twist + 3 * (size_t)inst : nullptr,
                       (valid && plan != nullptr) ? plan + (size_t)inst * 3 * P.N : nullptr);
}

// Persistent variant: every lane group keeps pulling instances from a global counter until the batch is exhausted,
// so a group that converges early does not idle while the slowest instance of its warp finishes (lock-step
// efficiency of the one-instance-per-group kernel is ~0.6 at 8 instances per warp, profiles/lockstep_r1.txt).
// The arithmetic of an instance does not depend on which group runs it or on its neighbours: results are
// bit-identical to solve_kernel's.
template <int G, int S>
__global__ void __launch_bounds__(kBlockThreads, min_blocks_for(S))
solve_queue_kernel(const __grid_constant__ SolverConst P, const float* __restrict__ lut_cost,
                   const uint8_t* __restrict__ lut_flag, const neompc_request* __restrict__ reqs, unsigned n,
                   neompc_response* __restrict__ out, float* __restrict__ twist, float* __restrict__ plan,
                   unsigned* __restrict__ counter) {
  extern __shared__ float hist_smem[];
  __shared__ SmemTables st;
  load_tables(st, lut_cost, lut_flag);
  CostTables T{st.cost, st.flag};
  constexpr int kGroupsPerBlock = kBlockThreads / G;
  const int lg = threadIdx.x % G;
  const unsigned lane = threadIdx.x & 31u;
  float* hist = hist_smem + threadIdx.x;
  const unsigned total_groups = gridDim.x * kGroupsPerBlock;

  Solver<G, S> sv;
  // a defined, inert state for groups that have not been given an instance yet
  sv.prologue(P, T, load_request(reqs, 0, false), false, lg, hist, kBlockThreads);
  unsigned inst = 0;
  bool need = true, exhausted = false, first_round = true;

  constexpr int kGroupsPerWarp = 32 / G;
  constexpr int kRefillAt = kGroupsPerWarp >= 4 ? kGroupsPerWarp / 2 : 1;   // refill once this many groups idle
  while (true) {
    const int idle = __popc(__ballot_sync(kFullMask, need && lg == 0));
    const bool any_active = __any_sync(kFullMask, sv.active);
    if (idle >= kRefillAt || (idle > 0 && !any_active)) {
      // 1. finish the instances of the groups that just converged
      const bool fin = need && sv.has_instance;
      {
        const neompc_request rq_done = load_request(reqs, inst, fin);
        sv.epilogue(P, T, rq_done, fin, lg, fin ? out + inst : nullptr,
                    (fin && twist != nullptr) ? twist + 3 * (size_t)inst : nullptr,
                    (fin && plan != nullptr) ? plan + (size_t)inst * 3 * P.N : nullptr);
      }
      // 2. next instance for every needy group: first round static, then one aggregated atomic per warp
      unsigned nxt;
      if (first_round) {
        nxt = blockIdx.x * kGroupsPerBlock + threadIdx.x / G;
      } else {
        const unsigned want = __ballot_sync(kFullMask, need && lg == 0);
        unsigned base = 0;
        if (lane == 0 && want != 0) base = atomicAdd(counter, (unsigned)__popc(want));
        base = __shfl_sync(kFullMask, base, 0);
        nxt = total_groups + base + (unsigned)__popc(want & ((1u << lane) - 1u));
        nxt = __shfl_sync(kFullMask, nxt, 0, G);          // lane 0 of each group holds the group's ticket
      }
      first_round = false;
      const bool got = need && nxt < n;
      const neompc_request rq = load_request(reqs, nxt, got);
      const bool fp_any = footprint_lethal<G>(P, T, (double)rq.pose_x, (double)rq.pose_y, (double)rq.pose_yaw, lg);
      if (need) {
        sv.init(P, rq, fp_any, got, lg, hist, kBlockThreads);
        inst = nxt;
        exhausted = !got;
      }
      need = false;
    }
    if (!__any_sync(kFullMask, sv.active)) break;
    sv.pass(P, T, hist, kBlockThreads, lg);
    need = need || (!sv.active && !exhausted);
  }
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
  constexpr int kInstPerBlock = kBlockThreads / G;
  const unsigned blocks_needed = (a.n + kInstPerBlock - 1) / kInstPerBlock;
  if (a.queue_counter != nullptr) {
    // persistent launch: as many blocks as fit on the device at once; the rest of the batch is pulled from the queue
    cudaError_t e = cudaFuncSetAttribute(solve_queue_kernel<G, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_queue_kernel<G, S>, kBlockThreads, smem);
    if (e != cudaSuccess) return e;
    const unsigned resident = (unsigned)(per_sm > 0 ? per_sm : 1) * (unsigned)a.sm_count;
    if (blocks_needed > resident) {
      e = cudaMemsetAsync(a.queue_counter, 0, sizeof(unsigned), a.stream);
      if (e != cudaSuccess) return e;
      solve_queue_kernel<G, S><<<resident, kBlockThreads, smem, a.stream>>>(a.P, a.lut_cost, a.lut_flag, a.reqs, a.n,
                                                                            a.out, a.twist, a.plan, a.queue_counter);
      return cudaGetLastError();
    }
  }
  cudaError_t e = cudaFuncSetAttribute(solve_kernel<G, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  solve_kernel<G, S><<<blocks_needed, kBlockThreads, smem, a.stream>>>(a.P, a.lut_cost, a.lut_flag, a.reqs, a.n, a.out,
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

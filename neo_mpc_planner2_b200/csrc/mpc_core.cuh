// mpc_core.cuh — the per-instance MPC solve: rollout, cost, analytic gradient, projection onto box∩disc,
// preconditioned projected L-BFGS, and the optimizer() epilogue.  One *lane group* of G lanes of one warp
// (G = 1, 2, 3, 4, 5, 6, 8, 10, 16 or 32; 32/G groups per warp) owns one MPC instance; each lane holds S <= 4
// consecutive control steps in registers (G*S >= control_steps).  Cross-lane work (prefix/suffix scans of the rollout and its adjoint, dot products)
// is done with warp shuffles; all groups of a warp run in lock step.
//
// What it computes follows /root/reference/neo_mpc_planner2/mpc_optimization_server.py ("srv.py"):
//   objective()            srv.py:204-269      -> Forward::run  (value)  + backward() (analytic gradient)
//   f_constraint + bounds  srv.py:125-134,157  -> project_step  (exact projection onto box ∩ disc)
//   minimize(SLSQP)        srv.py:363-364      -> Solver::pass (preconditioned projected L-BFGS; a different algorithm, same NLP)
//   optimizer() epilogue   srv.py:358-361,366-402 -> solve_instance tail
//   collision_check        srv.py:312-347
//
// The same header compiles for the host (plain g++, G = 1) so the algorithm can be exercised on machines
// without a GPU by tests/hostsim; that build is test tooling and is never linked into libneompc.so.
#pragma once

#include <math.h>
#include <stdint.h>

#include "neompc.h"

#if defined(__CUDACC__)
#define NEOMPC_HD __host__ __device__ __forceinline__
#else
#define NEOMPC_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define NEOMPC_UNROLL _Pragma("unroll")
#else
#define NEOMPC_UNROLL
#endif

// host-only tracing hook for tests/hostsim (compiled out everywhere else)
#if !defined(NEOMPC_TRACE)
#define NEOMPC_TRACE(...) ((void)0)
#endif

namespace neompc {

constexpr int kMaxMemory = 8;          // compile-time cap of L-BFGS pairs
constexpr int kMaxBacktracks = 8;      // arc-search trials per iteration (each shrinks the step by 0.1 .. 0.5)
constexpr float kPinnedAlpha = 0.01f;  // two accepted steps in a row this short: the arc is pinned at a costmap cell edge
constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kCellOob = 256;          // cost-table index of "outside the map" (cost 1.0, lethal)
constexpr int kCellFree = 257;         // cost-table index of "no costmap loaded" (cost 0)
constexpr int kTableSize = 258;

// ---------------------------------------------------------------------------------------------------------
// Per-handle constants, precomputed on the host (runtime.cu: build_const) and passed by value to the kernels.
// ---------------------------------------------------------------------------------------------------------
struct SolverConst {
  int N;               // control_steps
  int m;               // L-BFGS memory
  int max_iter;
  int disc_only;       // 1: disc lies inside the box -> projection is a radial scaling (README parameters)
  int fast_trig;       // 1: max|omega| * prediction_horizon <= pi -> MUFU sin/cos inside the rollout
  float dt;            // prediction_horizon / N                      (srv.py:137)
  float a_trans;       // w_trans / N                                 (srv.py:252)
  float b_orient;      // w_orient / N
  float w_ctrl;        // w_control / N                               (srv.py:253-254)
  float bt_term;       // w_orient * w_terminal                       (srv.py:268)
  float wt_term;       // w_trans * w_terminal
  float w_fp;          // w_footprint: N * (1.0^2 * w_footprint / N)  (srv.py:263); 0 in the moving-footprint mode
  float w_fp_step;     // moving-footprint mode (NEOMPC_FOOTPRINT_MOVING): w_footprint / N per lethal step, else 0
  int fp_mode;         // neompc_params.footprint_mode
  int cm_mode;         // neompc_params.costmap_mode
  float cm_scale;      // 1 / lethal_byte: cell byte -> normalised cost (bilinear mode)
  float cm_w;          // w_costmap / N                               (bilinear mode)
  float cm_wl;         // (1000 - w_costmap) / N                      (bilinear mode)
  int lethal_byte;     // the cell byte whose cost is 1.0 in the current encoding (100 or 254)
  float eps2;          // control_smoothing^2
  float lo[3], hi[3];  // box (srv.py:127-129)
  float R;             // max_vel_trans (srv.py:158)
  float acc[3];        // acceleration limits (srv.py:385-391)
  float lp_gain;       // low_pass_gain (srv.py:366-367)
  float tol_pg;        // projected-gradient tolerance derived from opt_tolerance
  float tol_f;         // relative objective-decrease tolerance derived from opt_tolerance
  float pair_eps;      // curvature pairs with s.y <= pair_eps * y.y are skipped
  float pin_alpha;     // an accepted arc parameter this small counts likewise (kPinnedAlpha)
  float tol_x;         // an accepted step shorter than this (sup norm) counts as "the objective stopped moving"
  // costmap (Costmap2d, srv.py:118)
  const uint8_t* cells;   // device pointer or nullptr (free space)
  const uint32_t* cells4; // corner-packed copy of the costmap (corner_word() below) or nullptr; the hot loop reads this one
  int pad4, pitch4;       // its padding and row pitch
  int pad_ok;             // 1: the padding covers twice a plan's reach -> the solver samples without bounds tests
  float k_lethal;         // lut_cost entry of the lethal byte (unique to it: build_tables)
  int polish_max;         // guidance: iteration cap of the second phase
  float alpha_warm;       // guidance, second phase: first trial at min(1, alpha_warm x the last accepted arc parameter)
  float sur_tol;          // guidance: the first phase stops at sur_tol x the tolerances of the second
  float cm_curv;          // curvature floor of the guided costmap term in the preconditioner (Solver::init)
  int guided;             // 1: the solve starts on the interpolated costmap term (costmap guidance, Solver::sur)
  int W, H;
  float inv_res;
  double origin_x, origin_y, inv_res_d;
  // footprint polygon, robot frame
  int fp_n;
  float fp_x[NEOMPC_MAX_FOOTPRINT_VERTICES], fp_y[NEOMPC_MAX_FOOTPRINT_VERTICES];
  // per-instance state rows: [3*N guess][3 last_control][waiting_time][flags(bits)][goal x,y,yaw][valid]
  float* state;
  int state_stride;       // floats per row
  unsigned state_rows;
  unsigned* err_word;     // set to 1 by any instance whose state row does not exist (nullptr: not reported)
};

// tables derived from (encoding, w_costmap, N): lut_cost[b] = w_costmap * c^2 / N  (srv.py:247,260); a lethal cell
// (c == 1.0) adds SolverConst::cm_wl = (1000 - w_costmap) / N on top (srv.py:257-258)
// lut_flag[b]: bit0 c == 1.0 (lethal), bit1 c >= 0.99 (srv.py:338)
struct CostTables {
  const float* cost;      // [kTableSize]  entry 256 = out-of-bounds (c = 1.0), entry 257 = no costmap (0)
  const uint8_t* flag;    // [kTableSize]
  uint32_t cost_s;        // device: shared-space byte address of `cost` (the kernels stage the tables in shared memory)
  // cost[byte_off / 4]: the hot lookups of Forward::cost address the shared window directly (ld.shared off a register
  // base) — through the generic pointer the compiler rematerialises the window base (S2UR + UMOV + ULEA) at every site
  NEOMPC_HD float cost_at_byte(uint32_t byte_off) const {
#if defined(__CUDA_ARCH__)
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(cost_s + byte_off));
    return v;
#else
    return cost[byte_off >> 2];
#endif
  }
};

// Corner-packed costmap.  Entry (ix, iy), ix in [-pad, W + pad), iy in [-pad, H + pad), holds the bytes of the four
// cells (ix,iy) (ix+1,iy) (ix,iy+1) (ix+1,iy+1) in bits 0-7, 8-15, 16-23, 24-31; a cell outside the map is stored as
// the lethal byte (cost 1.0, DESIGN.md "Costmap semantics").  One aligned 32-bit load then serves both the reference's
// nearest-cell term (srv.py:246-247: the cell containing the point is one of the four) and the interpolated term the
// solver uses for guidance.  `pad` is twice what a feasible plan can travel (max_vel_trans * prediction_horizon /
// resolution cells) plus slack, corner_pad_for(): any start within reach of the map indexes exactly, without bounds tests
// (make_instance).
// Row pitch W + 2 pad; index (iy + pad) * pitch + (ix + pad).
NEOMPC_HD int corner_pitch(int W, int pad) { return W + 2 * pad; }
NEOMPC_HD size_t corner_words(int W, int H, int pad) { return (size_t)(W + 2 * pad) * (size_t)(H + 2 * pad); }
NEOMPC_HD uint32_t corner_word(const uint8_t* cells, int W, int H, int lethal_byte, int ix, int iy) {
  uint32_t w = 0;
  NEOMPC_UNROLL
  for (int q = 0; q < 4; ++q) {
    const int cx = ix + (q & 1), cy = iy + (q >> 1);
    const bool inb = (unsigned)cx < (unsigned)W && (unsigned)cy < (unsigned)H;
    const uint32_t b = inb ? (uint32_t)cells[(size_t)cy * W + cx] : (uint32_t)lethal_byte;
    w |= b << (8 * q);
  }
  return w;
}

constexpr int kStateExtra = 12;   // floats after the 3N guess in a state row
NEOMPC_HD int state_stride_for(int n_steps) { return 3 * n_steps + kStateExtra; }

// ---------------------------------------------------------------------------------------------------------
// lane-group collectives
// ---------------------------------------------------------------------------------------------------------
// G lanes of one warp form a group, 32/G groups per warp (lanes past the last full group idle when G does not
// divide 32, e.g. G = 5: six groups, lanes 30-31 unused).  Powers of two use butterfly exchanges; other sizes run a
// Hillis-Steele scan inside the group and broadcast the total from its last lane.  Either way every lane of a group
// ends up with bit-identical scalars, so the lanes of a group always take the same decisions.
// Groups whose size is not a power of two (G = 3, 5, 6, 10 — the production tilings (5,2) and (10,2) among them)
// exchange through shared memory instead (NEOMPC_XCH, default on): every lane stores its value into the group's
// 16-byte-aligned slot, the warp synchronises, and every lane reads ALL G values of its group back with vector loads
// (G = 5: one LDS.128 + one LDS.32) and combines them locally in index order.  That is one store + two loads + four
// adds where the shuffle scan needs four DEPENDENT shuffle round trips (~25 cycles each) — the dependency chain of a
// collective drops from ~110 to ~45 cycles, which is what this issue/latency-bound kernel is short of.
#ifndef NEOMPC_XCH
#define NEOMPC_XCH 1
#endif
// power-of-two groups of at least this many lanes exchange through shared memory too (64: none — butterflies)
#ifndef NEOMPC_XCH_POW2_MIN
#define NEOMPC_XCH_POW2_MIN 64
#endif
#ifndef NEOMPC_BLOCK_THREADS
#define NEOMPC_BLOCK_THREADS 64
#endif
constexpr int kXchMaxWarps = NEOMPC_BLOCK_THREADS / 32;   // warps per block (kernels.cuh launches exactly this block size)
// first exchange array of the collectives that own their arrays (see Grp::gather)
enum XchSite { kXsAny = 0, kXsZ = 4, kXsXY = 5, kXsJ = 7, kXsSuf2 = 9, kXsSufG = 11, kXsDot1 = 12, kXsDot2 = 13, kXsDesc = 14,
               kXsPair = 15, kXchArrays = 19 };

template <int G>
struct Grp {
  static constexpr bool kPow2 = (G & (G - 1)) == 0;
  static constexpr int kPerWarp = 32 / G;          // groups (= instances) per warp
  static constexpr bool kXch = NEOMPC_XCH && G > 1 && (!kPow2 || G >= NEOMPC_XCH_POW2_MIN);
  static constexpr int kSlot = (G + 3) & ~3;       // floats per group slot of the exchange buffer (16-byte aligned)
  static constexpr int kSlots = (32 + G - 1) / G;  // incl. the partial group of leftover lanes
#if defined(__CUDA_ARCH__)
  static __device__ __forceinline__ int lane() { return (int)(threadIdx.x & 31u); }

  // gather(): all[k][i] = value v[k] of lane i of this lane's group, for K = 1, 2 or 4 scalars at once.
  // The exchange buffer of a warp is kXchArrays arrays of one float per lane slot.  A collective names the first array
  // it uses (`Site`) and owns K consecutive ones; inside them the K scalars of a lane lie side by side, so a lane stores
  // with one STS.32/.64/.128 and reads its group's K*G floats with vector loads.  One warp barrier separates the stores
  // from the loads.  Site 0 (the default, arrays 0..3) ends with a second barrier — loads before the next collective's
  // stores — so that it can be used anywhere; the hot collectives of Solver::pass own their arrays (XchSite) and skip it:
  // between two executions of the same site the warp always passes the barrier of another collective.
  static __device__ __forceinline__ float (*xch_warp())[kSlots * kSlot] {
    __shared__ alignas(16) float xbuf[kXchMaxWarps][kXchArrays][kSlots * kSlot];
    return xbuf[threadIdx.x >> 5];
  }
  template <int K, int Site>
  static __device__ __forceinline__ void gather(const float (&v)[K], float (&all)[K][G]) {
    static_assert(Site + K <= kXchArrays && (Site != 0 || K <= 4), "exchange buffer too small");
    static_assert(K == 1 || K == 2 || K == 4, "one, two or four scalars per exchange");
    const int ln = lane(), grp = ln / G, lg = ln - grp * G;
    // K scalars of a lane are stored side by side (one STS.32/.64/.128); a group's K*G floats start 16-byte aligned
    constexpr int kSpan = (K * G + 3) & ~3;
    static_assert(kSpan * kSlots <= K * kSlots * kSlot, "packed layout must fit the K arrays of the site");
    float* base = &xch_warp()[Site][0] + grp * kSpan;
    if (K == 1) base[lg] = v[0];
    else if (K == 2) *reinterpret_cast<float2*>(base + 2 * lg) = make_float2(v[0], v[K > 1 ? 1 : 0]);
    else *reinterpret_cast<float4*>(base + 4 * lg) = make_float4(v[0], v[K > 1 ? 1 : 0], v[K > 2 ? 2 : 0], v[K > 3 ? 3 : 0]);
    __syncwarp();
    float flat[kSpan];
    NEOMPC_UNROLL
    for (int q = 0; q < kSpan / 4; ++q) {
      if (4 * q + 4 <= K * G) {
        const float4 t = *reinterpret_cast<const float4*>(base + 4 * q);
        flat[4 * q] = t.x; flat[4 * q + 1] = t.y; flat[4 * q + 2] = t.z; flat[4 * q + 3] = t.w;
      } else if (4 * q + 2 == K * G) {
        const float2 t = *reinterpret_cast<const float2*>(base + 4 * q);
        flat[4 * q] = t.x; flat[4 * q + 1] = t.y;
      } else {
        NEOMPC_UNROLL
        for (int e = 4 * q; e < K * G; ++e) flat[e] = base[e];
      }
    }
    NEOMPC_UNROLL
    for (int i = 0; i < G; ++i) {
      NEOMPC_UNROLL
      for (int k = 0; k < K; ++k) all[k][i] = flat[K * i + k];
    }
    if (Site == 0) __syncwarp();
  }
#endif
  template <int Site = 0, class Op>
  static NEOMPC_HD float reduce(float v, Op op) {
#if defined(__CUDA_ARCH__)
    if (kXch) {
      const float in[1] = {v};
      float all[1][G];
      gather<1, Site>(in, all);
      if (kPow2) {                        // pairwise tree (same order in every lane; short dependency chain)
        NEOMPC_UNROLL
        for (int w = G / 2; w > 0; w >>= 1) {
          NEOMPC_UNROLL
          for (int i = 0; i < w; ++i) all[0][i] = op(all[0][2 * i], all[0][2 * i + 1]);
        }
        v = all[0][0];
      } else {
        v = all[0][0];
        NEOMPC_UNROLL
        for (int i = 1; i < G; ++i) v = op(v, all[0][i]);
      }
    } else if (kPow2) {
      NEOMPC_UNROLL
      for (int o = G / 2; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(kFullMask, v, o));
    } else {
      const int ln = lane(), lg = ln % G;
      NEOMPC_UNROLL
      for (int o = 1; o < G; o <<= 1) {
        const float t = __shfl_sync(kFullMask, v, ln - o);
        if (lg >= o) v = op(v, t);
      }
      v = __shfl_sync(kFullMask, v, ln - lg + (G - 1));
    }
#else
    (void)op;
#endif
    return v;
  }
  struct Add { NEOMPC_HD float operator()(float a, float b) const { return a + b; } };
  struct Max { NEOMPC_HD float operator()(float a, float b) const { return fmaxf(a, b); } };
  template <int Site = 0>
  static NEOMPC_HD float sum(float v) { return reduce<Site>(v, Add()); }
  template <int Site = 0>
  static NEOMPC_HD float max(float v) { return reduce<Site>(v, Max()); }
  static NEOMPC_HD int imax(int v) {       // small non-negative flags/counters: exact in float
    return (int)reduce((float)v, Max());
  }
  // two sums at once (one exchange for groups that use the shared-memory path)
  template <int Site = 0>
  static NEOMPC_HD void sum2(float& a, float& b) {
#if defined(__CUDA_ARCH__)
    if (kXch) {
      const float in[2] = {a, b};
      float all[2][G];
      gather<2, Site>(in, all);
      a = all[0][0]; b = all[1][0];
      NEOMPC_UNROLL
      for (int i = 1; i < G; ++i) { a += all[0][i]; b += all[1][i]; }
      return;
    }
#endif
    a = sum(a); b = sum(b);
  }
  // two sums and two maxima at once
  template <int Site = 0>
  static NEOMPC_HD void sum2_max2(float& a, float& b, float& c, float& d) {
#if defined(__CUDA_ARCH__)
    if (kXch) {
      const float in[4] = {a, b, c, d};
      float all[4][G];
      gather<4, Site>(in, all);
      a = all[0][0]; b = all[1][0]; c = all[2][0]; d = all[3][0];
      NEOMPC_UNROLL
      for (int i = 1; i < G; ++i) { a += all[0][i]; b += all[1][i]; c = fmaxf(c, all[2][i]); d = fmaxf(d, all[3][i]); }
      return;
    }
#endif
    a = sum(a); b = sum(b); c = max(c); d = max(d);
  }
  // the value lane 0 of the group holds
  static NEOMPC_HD float bcast0(float v, int lg) {
#if defined(__CUDA_ARCH__)
    if (G > 1) return kPow2 ? __shfl_sync(kFullMask, v, 0, G) : __shfl_sync(kFullMask, v, lane() - lg);
#endif
    (void)lg;
    return v;
  }
  // sum of v over the lanes of the group that come BEFORE this lane
  template <int Site = 0>
  static NEOMPC_HD float excl_prefix(float v, int lg) {
#if defined(__CUDA_ARCH__)
    if (kXch) {
      const float in[1] = {v};
      float all[1][G];
      gather<1, Site>(in, all);
      float acc = 0 < lg ? all[0][0] : 0.0f;
      NEOMPC_UNROLL
      for (int i = 1; i + 1 < G; ++i) { if (i < lg) acc += all[0][i]; }
      return acc;
    }
    if (G > 1) {
      const int ln = lane();
      float incl = v;
      NEOMPC_UNROLL
      for (int o = 1; o < G; o <<= 1) {
        const float t = kPow2 ? __shfl_up_sync(kFullMask, incl, o, G) : __shfl_sync(kFullMask, incl, ln - o);
        if (lg >= o) incl += t;
      }
      const float ex = kPow2 ? __shfl_up_sync(kFullMask, incl, 1, G) : __shfl_sync(kFullMask, incl, ln - 1);
      return lg > 0 ? ex : 0.0f;
    }
#endif
    (void)v; (void)lg;
    return 0.0f;
  }
  // two exclusive prefix sums at once
  template <int Site = 0>
  static NEOMPC_HD void excl_prefix2(float a, float b, int lg, float* pa, float* pb) {
#if defined(__CUDA_ARCH__)
    if (kXch) {
      const float in[2] = {a, b};
      float all[2][G];
      gather<2, Site>(in, all);
      float sa = 0 < lg ? all[0][0] : 0.0f, sb = 0 < lg ? all[1][0] : 0.0f;
      NEOMPC_UNROLL
      for (int i = 1; i + 1 < G; ++i) { if (i < lg) { sa += all[0][i]; sb += all[1][i]; } }
      *pa = sa; *pb = sb;
      return;
    }
#endif
    *pa = excl_prefix(a, lg); *pb = excl_prefix(b, lg);
  }
  // sum of v over the lanes of the group that come AFTER this lane
  template <int Site = 0>
  static NEOMPC_HD float excl_suffix(float v, int lg) {
#if defined(__CUDA_ARCH__)
    if (kXch) {
      const float in[1] = {v};
      float all[1][G];
      gather<1, Site>(in, all);
      float acc = G - 1 > lg ? all[0][G - 1] : 0.0f;
      NEOMPC_UNROLL
      for (int i = G - 2; i > 0; --i) { if (i > lg) acc += all[0][i]; }
      return acc;
    }
    if (G > 1) {
      const int ln = lane();
      float incl = v;
      NEOMPC_UNROLL
      for (int o = 1; o < G; o <<= 1) {
        const float t = kPow2 ? __shfl_down_sync(kFullMask, incl, o, G) : __shfl_sync(kFullMask, incl, ln + o);
        if (lg + o < G) incl += t;
      }
      const float ex = kPow2 ? __shfl_down_sync(kFullMask, incl, 1, G) : __shfl_sync(kFullMask, incl, ln + 1);
      return lg + 1 < G ? ex : 0.0f;
    }
#endif
    (void)v; (void)lg;
    return 0.0f;
  }
  // two exclusive suffix sums at once
  template <int Site = 0>
  static NEOMPC_HD void excl_suffix2(float a, float b, int lg, float* pa, float* pb) {
#if defined(__CUDA_ARCH__)
    if (kXch) {
      const float in[2] = {a, b};
      float all[2][G];
      gather<2, Site>(in, all);
      float sa = G - 1 > lg ? all[0][G - 1] : 0.0f, sb = G - 1 > lg ? all[1][G - 1] : 0.0f;
      NEOMPC_UNROLL
      for (int i = G - 2; i > 0; --i) { if (i > lg) { sa += all[0][i]; sb += all[1][i]; } }
      *pa = sa; *pb = sb;
      return;
    }
#endif
    *pa = excl_suffix(a, lg); *pb = excl_suffix(b, lg);
  }
  // true if the predicate holds for any lane of the WARP (all groups of a warp iterate in lock step)
  static NEOMPC_HD bool warp_any(bool p) {
#if defined(__CUDA_ARCH__)
    return __any_sync(kFullMask, p) != 0;
#else
    return p;
#endif
  }
};

NEOMPC_HD void sincos_f(float a, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  sincosf(a, s, c);
#else
  *s = sinf(a);
  *c = cosf(a);
#endif
}

// sin/cos of the accumulated heading inside the rollout.  |a| <= max|omega| * prediction_horizon; when that bound is
// <= pi (SolverConst::fast_trig) the MUFU path is used: absolute error <= 2^-21.4 on [-pi, pi] (CUDA C Programming
// Guide, intrinsic table), which moves a rollout position by < 1e-8 m.
NEOMPC_HD void sincos_heading(bool fast, float a, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  if (fast) __sincosf(a, s, c); else sincosf(a, s, c);
#else
  (void)fast;
  *s = sinf(a);
  *c = cosf(a);
#endif
}

NEOMPC_HD float rsqrt_f(float v) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));   // one MUFU.RSQ; callers keep v in the normal range
  return r;
#else
  return 1.0f / sqrtf(v);
#endif
}

// a / b where a few ulp do not matter (step lengths, scalings): MUFU.RCP, no denormal slow path
NEOMPC_HD float rcp_approx(float b) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));     // one MUFU.RCP
  return r;
#else
  return 1.0f / b;
#endif
}
NEOMPC_HD float div_approx(float a, float b) { return a * rcp_approx(b); }

NEOMPC_HD float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// ---------------------------------------------------------------------------------------------------------
// Euclidean projection of one control step onto  [lo,hi]^3 ∩ { vx^2 + vy^2 <= R^2 }
// (bounds srv.py:127-133; disc constraint srv.py:157-158).  omega only sees its interval.
// ---------------------------------------------------------------------------------------------------------
// General case (the disc is not contained in the box); kept out of line: the README parameters never take it.
// Takes and returns values so that callers' register arrays never have their address taken.
struct Vec2 { float x, y; };
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
inline Vec2 project_general(const SolverConst& P, float vx, float vy) {
  const float R2 = P.R * P.R;
  const float bx = clampf(vx, P.lo[0], P.hi[0]);
  const float by = clampf(vy, P.lo[1], P.hi[1]);
  if (bx * bx + by * by <= R2) return Vec2{bx, by};   // box projection already inside the disc
  const float n2 = vx * vx + vy * vy;
  if (n2 > 0.0f) {
    const float sc = P.R / sqrtf(n2);
    const float rx = vx * sc, ry = vy * sc;
    const float t = 1e-6f;
    if (rx >= P.lo[0] - t && rx <= P.hi[0] + t && ry >= P.lo[1] - t && ry <= P.hi[1] + t)
      return Vec2{clampf(rx, P.lo[0], P.hi[0]), clampf(ry, P.lo[1], P.hi[1])};   // disc projection inside the box
  }
  // Otherwise the projection is a vertex of the feasible region: a box-edge line meeting the circle.
  float best = 3.4e38f, ox = bx, oy = by;
  bool found = false;
  for (int e = 0; e < 4; ++e) {
    const bool xedge = e < 2;                          // x fixed at a bound, y on the circle
    const float fixed = xedge ? (e == 0 ? P.lo[0] : P.hi[0]) : (e == 2 ? P.lo[1] : P.hi[1]);
    const float rem = R2 - fixed * fixed;
    if (rem < 0.0f) continue;
    const float root = sqrtf(rem);
    for (int sgn = 0; sgn < 2; ++sgn) {
      const float other = sgn ? root : -root;
      const float lo_o = xedge ? P.lo[1] : P.lo[0];
      const float hi_o = xedge ? P.hi[1] : P.hi[0];
      if (other < lo_o || other > hi_o) continue;
      const float cx = xedge ? fixed : other;
      const float cy = xedge ? other : fixed;
      const float d2 = (cx - vx) * (cx - vx) + (cy - vy) * (cy - vy);
      if (d2 < best) { best = d2; ox = cx; oy = cy; found = true; }
    }
  }
  if (!found) {                                        // degenerate parameters: stay safe
    const float n = sqrtf(bx * bx + by * by);
    ox = bx * P.R / n;
    oy = by * P.R / n;
  }
  return Vec2{ox, oy};
}

// X = false: the caller guarantees P.disc_only (the general case analysis is not even compiled in)
template <bool X>
NEOMPC_HD void project_step(const SolverConst& P, float& vx, float& vy, float& om) {
  om = clampf(om, P.lo[2], P.hi[2]);
  if (!X || P.disc_only) {                             // radial scaling, branch-free
    const float n2 = vx * vx + vy * vy;
    const float sc = fminf(1.0f, P.R * rsqrt_f(fmaxf(n2, 1e-30f)));
    vx *= sc;
    vy *= sc;
  } else {
    const Vec2 p = project_general(P, vx, vy);
    vx = p.x;
    vy = p.y;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Per-instance constants (hoisted out of objective(): srv.py:207-221)
// ---------------------------------------------------------------------------------------------------------
struct Instance {
  float cx, cy;          // carrot position                    (srv.py:219)
  float tyaw, fyaw;      // target_yaw, final_yaw              (srv.py:211-212)
  float v0x, v0y, v0z;   // current velocity                   (srv.py:216-218)
  float cq, sq;          // cos/sin of pose_yaw_objective / resolution  (srv.py:213, hoisted from :234-236)
  int bx, by;            // cell containing the current position
  int base4;             // entry of that cell (clamped to the map's one-cell surround) in the corner-packed map
  float fx, fy;          // fractional position inside that cell, in cells
};

// srv.py:207-221 hoisted: everything objective() derives from the request alone
NEOMPC_HD Instance make_instance(const SolverConst& P, const neompc_request& rq) {
  Instance I;
  I.cx = rq.carrot_x; I.cy = rq.carrot_y;
  I.tyaw = rq.carrot_yaw; I.fyaw = rq.goal_yaw;
  I.v0x = rq.vel_x; I.v0y = rq.vel_y; I.v0z = rq.vel_theta;
  sincos_f(rq.pose_yaw_objective, &I.sq, &I.cq);
  I.sq *= P.inv_res; I.cq *= P.inv_res;
  I.bx = I.by = 0; I.fx = I.fy = 0.0f; I.base4 = 0;
  if (P.cells != nullptr) {
    // nav2 worldToMap in float64, then a float32 offset inside the cell keeps sub-cell precision on big maps
    const double gx = ((double)rq.pose_x - P.origin_x) * P.inv_res_d;
    const double gy = ((double)rq.pose_y - P.origin_y) * P.inv_res_d;
    const double bxd = fmin(fmax(floor(gx), -1.0e9), 1.0e9), byd = fmin(fmax(floor(gy), -1.0e9), 1.0e9);
    I.bx = (int)bxd; I.by = (int)byd;
    I.fx = (float)(gx - floor(gx)); I.fy = (float)(gy - floor(gy));
    // Unchecked sampling indexes relative to the start cell.  A start more than a plan's reach outside the map sees lethal
    // cells whatever it does, so it may stand in for any other such start: clamping it there keeps every index a
    // feasible plan can produce inside the padded array (pad = 2 reach + 6, corner_pad_for).
    const int reach = P.pad_ok ? (P.pad4 - 6) / 2 : 0;
    const int bxc = I.bx < -(reach + 2) ? -(reach + 2) : (I.bx > P.W + reach + 1 ? P.W + reach + 1 : I.bx);
    const int byc = I.by < -(reach + 2) ? -(reach + 2) : (I.by > P.H + reach + 1 ? P.H + reach + 1 : I.by);
    I.base4 = (byc + P.pad4) * P.pitch4 + (bxc + P.pad4);
  }
  return I;
}

// terms of J that do not depend on u: terminal distance (srv.py:266) + footprint (srv.py:262-263)
NEOMPC_HD float constant_cost(const SolverConst& P, const neompc_request& rq, bool fp_hit) {
  const float ddx = rq.carrot_x - rq.goal_x, ddy = rq.carrot_y - rq.goal_y;
  return P.wt_term * (ddx * ddx + ddy * ddy) + (fp_hit ? P.w_fp : 0.0f);
}

// What optimizer() remembers about an instance between calls (srv.py:115-117,138,146-149), read from its state row.
struct Carry {
  float last[3];         // last_control                       (srv.py:393-395)
  float waiting;         // waiting_time                       (srv.py:378-382)
  bool latched;          // self.collision                     (srv.py:338-339)
  bool new_goal;         // goal_pose != old_goal              (srv.py:358-361)
  bool stateful;
  bool no_state;         // a state row was asked for that does not exist (id beyond neompc_reserve_instances)
};

NEOMPC_HD Carry load_carry(const SolverConst& P, const neompc_request& rq, bool valid) {
  Carry c;
  c.stateful = valid && rq.instance_id != NEOMPC_STATELESS && P.state != nullptr && rq.instance_id < P.state_rows;
  c.no_state = valid && rq.instance_id != NEOMPC_STATELESS && !c.stateful;
  c.last[0] = c.last[1] = c.last[2] = 0.0f;
  c.waiting = 0.0f;
  c.latched = false;
  c.new_goal = true;
  if (c.stateful) {
    const float* tail = P.state + (size_t)rq.instance_id * P.state_stride + 3 * P.N;
    c.new_goal = !(tail[8] != 0.0f && tail[5] == rq.goal_x && tail[6] == rq.goal_y && tail[7] == rq.goal_yaw);
    c.latched = tail[4] != 0.0f;
    if (!c.new_goal) { c.last[0] = tail[0]; c.last[1] = tail[1]; c.last[2] = tail[2]; c.waiting = tail[3]; }
  }
  return c;
}

// X = true is the general build: opt-in objective extensions (moving footprint, bilinear costmap), the general
// box-and-disc projection and the accurate sincosf.  X = false is the reference fast path — no extension, disc inside the
// box (P.disc_only), heading range within MUFU accuracy (P.fast_trig); the dispatcher picks it only when all three hold —
// so that its hot loop carries none of the other code (merely being present cost 7-20 % there).
// F ("full"): G * S == control_steps, i.e. no padded step, on a handle with a corner-packed costmap whose padding covers a
// plan's reach — the per-step masks, the costmap-present test and the bounds-checked sampling path fold away (launch_solve_gs)
template <int G, int S, bool X, bool F = false>
struct Forward {
  float c[S], s[S], dx[S], dy[S], x[S], y[S], z[S], rinv[S];

  // Table index of the cell under a base-frame offset (x, y) rotated by (cr, sr) from the current position:
  // 0..255 = the cell byte, kCellOob = outside the map, kCellFree = no costmap loaded.  Branch-free.
  static NEOMPC_HD int cell_of(const SolverConst& P, const Instance& I, float cr, float sr, float x, float y) {
    if (P.cells == nullptr) return kCellFree;                  // uniform
    // (cr, sr) = cos/sin of the start yaw pre-multiplied by 1/resolution
    const float gx = I.fx + (cr * x - sr * y);
    const float gy = I.fy + (sr * x + cr * y);
#if defined(__CUDA_ARCH__)
    const int mx = I.bx + __float2int_rd(gx);
    const int my = I.by + __float2int_rd(gy);
#else
    const int mx = I.bx + (int)floorf(gx);
    const int my = I.by + (int)floorf(gy);
#endif
    const bool inb = (unsigned)mx < (unsigned)P.W && (unsigned)my < (unsigned)P.H;
    const unsigned idx = inb ? (unsigned)my * (unsigned)P.W + (unsigned)mx : 0u;      // W*H < 2^32 (set_costmap checks)
#if defined(__CUDA_ARCH__)
    const int cell = (int)__ldg(P.cells + idx);
#else
    const int cell = (int)P.cells[idx];
#endif
    return inb ? cell : kCellOob;
  }

  // Bilinear costmap mode (NEOMPC_COSTMAP_BILINEAR).  Normalised cost c and lethal indicator l are
  // interpolated between the four cell centres around the predicted position; term = cm_w c^2 + cm_wl l^2.
  // Returns the term; (dx, dy) receive its derivative w.r.t. the BASE-frame rollout offset (x, y) in metres
  // (chain rule through the rotation by the start yaw; I.cq / I.sq carry the factor 1/resolution).
  static NEOMPC_HD float bilinear_term(const SolverConst& P, const Instance& I, float x, float y, float* dx, float* dy) {
    const float gx = I.fx + (I.cq * x - I.sq * y) - 0.5f;          // cell-centre coordinates relative to the base cell
    const float gy = I.fy + (I.sq * x + I.cq * y) - 0.5f;
    const float fx0 = floorf(gx), fy0 = floorf(gy);
    const float tx = gx - fx0, ty = gy - fy0;
    const int mx = I.bx + (int)fx0, my = I.by + (int)fy0;
    float c[4], l[4];
    NEOMPC_UNROLL
    for (int q = 0; q < 4; ++q) {
      const int cx = mx + (q & 1), cy = my + (q >> 1);
      const bool inb = (unsigned)cx < (unsigned)P.W && (unsigned)cy < (unsigned)P.H;
      const unsigned idx = inb ? (unsigned)cy * (unsigned)P.W + (unsigned)cx : 0u;
#if defined(__CUDA_ARCH__)
      const int b = (int)__ldg(P.cells + idx);
#else
      const int b = (int)P.cells[idx];
#endif
      c[q] = !inb ? 1.0f : (b <= P.lethal_byte ? (float)b * P.cm_scale : 0.0f);    // outside the map: lethal
      l[q] = (!inb || b == P.lethal_byte) ? 1.0f : 0.0f;
    }
    const float c0 = c[0] + tx * (c[1] - c[0]), c1 = c[2] + tx * (c[3] - c[2]);
    const float l0 = l[0] + tx * (l[1] - l[0]), l1 = l[2] + tx * (l[3] - l[2]);
    const float cv = c0 + ty * (c1 - c0), lv = l0 + ty * (l1 - l0);
    const float dcx = (c[1] - c[0]) + ty * ((c[3] - c[2]) - (c[1] - c[0])), dcy = c1 - c0;   // per cell
    const float dlx = (l[1] - l[0]) + ty * ((l[3] - l[2]) - (l[1] - l[0])), dly = l1 - l0;
    const float ggx = 2.0f * (P.cm_w * cv * dcx + P.cm_wl * lv * dlx);                       // d term / d gx
    const float ggy = 2.0f * (P.cm_w * cv * dcy + P.cm_wl * lv * dly);
    *dx = ggx * I.cq + ggy * I.sq;                                                           // gx = .. + cq x - sq y
    *dy = -ggx * I.sq + ggy * I.cq;                                                          // gy = .. + sq x + cq y
    return P.cm_w * cv * cv + P.cm_wl * lv * lv;
  }

  // Moving-footprint mode (NEOMPC_FOOTPRINT_MOVING): is the robot-frame polygon, placed at the
  // predicted pose (base-frame offset (x, y), heading change with cos/sin (c, s)) of the costmap rollout
  // (srv.py:234-236), in collision?  nav2 FootprintCollisionChecker::footprintCost == 1.0 restated: vertices -> cells,
  // every edge incl. last -> first walked with nav2's LineIterator (incremental Bresenham, both ends included); a
  // vertex outside the map counts as lethal.  Run by ONE lane for one of its steps: no collectives inside.
  static NEOMPC_HD bool footprint_lethal_at(const SolverConst& P, const Instance& I, float x, float y, float c, float s) {
    // heading of the pose = start yaw + z, by angle addition; I.cq / I.sq carry the factor 1/resolution (cells)
    const float cw = I.cq * c - I.sq * s, sw = I.sq * c + I.cq * s;
    const float ox = I.fx + (I.cq * x - I.sq * y), oy = I.fy + (I.sq * x + I.cq * y);
    int lethal = 0;
    int mx0 = 0, my0 = 0, mxf = 0, myf = 0;
    for (int v = 0; v <= P.fp_n; ++v) {
      int mx, my;
      if (v < P.fp_n) {
        const float gx = ox + (cw * P.fp_x[v] - sw * P.fp_y[v]);
        const float gy = oy + (sw * P.fp_x[v] + cw * P.fp_y[v]);
#if defined(__CUDA_ARCH__)
        mx = I.bx + __float2int_rd(gx);
        my = I.by + __float2int_rd(gy);
#else
        mx = I.bx + (int)floorf(gx);
        my = I.by + (int)floorf(gy);
#endif
        if ((unsigned)mx >= (unsigned)P.W || (unsigned)my >= (unsigned)P.H) {   // vertex off the map: cost 1.0
          lethal = 1;
          mx = mx < 0 ? 0 : (mx >= P.W ? P.W - 1 : mx);                         // keep the walk inside the grid
          my = my < 0 ? 0 : (my >= P.H ? P.H - 1 : my);
        }
        if (v == 0) { mxf = mx; myf = my; mx0 = mx; my0 = my; continue; }
      } else {
        mx = mxf; my = myf;                       // closing edge last -> first
      }
      // all vertices are inside the grid here, so every pixel of the edge is too: no per-pixel bounds test
      const int ddx = mx - mx0, ddy = my - my0;
      const int adx = ddx < 0 ? -ddx : ddx, ady = ddy < 0 ? -ddy : ddy;
      const int sxs = ddx >= 0 ? 1 : -1, sys = ddy >= 0 ? 1 : -1;
      const bool xmaj = adx >= ady;
      const int den = xmaj ? adx : ady, numadd = xmaj ? ady : adx;
      const int step_major = xmaj ? sxs : sys * P.W;          // index increments along / across the major axis
      const int step_minor = xmaj ? sys * P.W : sxs;
      int num = den >> 1;
      int idx = my0 * P.W + mx0;
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
      for (int k = 0; k <= den; ++k) {
#if defined(__CUDA_ARCH__)
        const int cell = (int)__ldg(P.cells + idx);
#else
        const int cell = (int)P.cells[idx];
#endif
        lethal |= (cell == P.lethal_byte) ? 1 : 0;
        num += numadd;
        const bool wrap = num >= den;
        num -= wrap ? den : 0;
        idx += step_major + (wrap ? step_minor : 0);
      }
      mx0 = mx; my0 = my;
    }
    return lethal != 0;
  }

  // Rolls the omni-drive model over the horizon (srv.py:230-232): z, cos/sin, per-step displacement, position.
  // u[j][0..2] = (vx, vy, omega) of step lg*S + j.
  NEOMPC_HD void rollout(const SolverConst& P, const float (*u)[3], int lg) {
    const float dt = P.dt;
    // z_i = dt * sum_{k<=i} omega_k                                               (srv.py:230)
    float acc = 0.0f;
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) { acc += u[j][2] * dt; z[j] = acc; }
    const float zoff = Grp<G>::template excl_prefix<kXsZ>(acc, lg);
    float ax = 0.0f, ay = 0.0f;
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) {
      z[j] += zoff;
      sincos_heading(!X || P.fast_trig != 0, z[j], &s[j], &c[j]);
      dx[j] = (u[j][0] * c[j] - u[j][1] * s[j]) * dt;                              // srv.py:231
      dy[j] = (u[j][0] * s[j] + u[j][1] * c[j]) * dt;                              // srv.py:232
      ax += dx[j]; x[j] = ax;
      ay += dy[j]; y[j] = ay;
    }
    float xoff, yoff;
    Grp<G>::template excl_prefix2<kXsXY>(ax, ay, lg, &xoff, &yoff);
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) { x[j] += xoff; y[j] += yoff; }
  }

  // Costmap sample at base-frame offset (x, y): the corner-packed word around the point and the interpolation
  // weights (cell-centre coordinates relative to the start cell).  Checked = false: the caller guarantees a feasible
  // plan, which cannot leave the padded map (corner_pad_for); Checked = true (test hook, arbitrary controls): a point
  // outside the padded map sees four lethal corners.
  template <bool Checked>
  static NEOMPC_HD uint32_t corner_at(const SolverConst& P, const Instance& I, float x, float y, float* tx, float* ty) {
    const float gx = I.fx + (I.cq * x - I.sq * y) - 0.5f;
    const float gy = I.fy + (I.sq * x + I.cq * y) - 0.5f;
    const float flx = floorf(gx), fly = floorf(gy);
    *tx = gx - flx;
    *ty = gy - fly;
    const int ox = (int)flx, oy = (int)fly;
    bool inside = true;
    int idx = I.base4 + oy * P.pitch4 + ox;
    if (Checked || (!F && !P.pad_ok)) {          // absolute entry; anything outside the padded array is outside the map
      const float lim = 1.0e9f;
      const int ix = I.bx + (int)fminf(fmaxf(flx, -lim), lim), iy = I.by + (int)fminf(fmaxf(fly, -lim), lim);
      inside = ix >= -P.pad4 && ix < P.W + P.pad4 && iy >= -P.pad4 && iy < P.H + P.pad4;
      idx = inside ? (iy + P.pad4) * P.pitch4 + (ix + P.pad4) : 0;
    }
#if defined(__CUDA_ARCH__)
    const uint32_t w = __ldg(P.cells4 + idx);
#else
    const uint32_t w = P.cells4[idx];
#endif
    return inside ? w : (uint32_t)P.lethal_byte * 0x01010101u;
  }

  // This lane's share of J (srv.py:246-268) for the rollout held in the struct.  The control term uses
  // sqrt(r^2 + eps^2); its reciprocal is kept for backward().  On return x[j], y[j] no longer hold the positions
  // but the adjoint seeds dJ/dx_j, dJ/dy_j (tracking term + the costmap term where it has a gradient), which is all
  // backward() needs of them.
  // `sur` (per lane group): costmap guidance — the costmap term is the bilinear interpolation of the per-cell term
  // table[cell] between the four surrounding cell centres (equal to the reference's term at every cell centre) and
  // its gradient enters the seeds; otherwise the reference's term table[cell under the point] (gradient 0).
  // A lethal cell (cost == 1.0) weighs 1000 instead of w_costmap (srv.py:257-258): that part stays piecewise constant
  // in both cases — the guidance follows the inflation slope, the wall stays where the reference has it.
  template <bool Checked>
  NEOMPC_HD float cost(const SolverConst& P, const CostTables& T, const Instance& I, const float (*u)[3], int lg,
                       bool sur) {
    const bool bilinear = X && P.cm_mode == NEOMPC_COSTMAP_BILINEAR && P.cells != nullptr;   // uniform
    const bool sampled = F || (!bilinear && P.cells4 != nullptr);                             // uniform (F: the launcher checked)
    uint32_t word[S];
    float tx[S], ty[S];
    if (sampled) {
      NEOMPC_UNROLL
      for (int j = 0; j < S; ++j) word[j] = corner_at<Checked>(P, I, x[j], y[j], &tx[j], &ty[j]);   // (loads issued together)
    }
    // the interpolation is skipped by warps none of whose groups is (still) guided
    const bool any_sur = sampled && Grp<G>::warp_any(sur);                                    // warp-uniform
    const float cqs = sur ? I.cq : 0.0f, sqs = sur ? I.sq : 0.0f, sf = sur ? 1.0f : 0.0f;
    float J = 0.0f;
    float cmx[S], cmy[S];
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) { cmx[j] = 0.0f; cmy[j] = 0.0f; }
    if (bilinear) {
      for (int j = 0; j < S; ++j) {
        const float t = bilinear_term(P, I, x[j], y[j], &cmx[j], &cmy[j]);
        J += (lg * S + j < P.N) ? t : 0.0f;
      }
    }
    if (X && P.fp_mode == NEOMPC_FOOTPRINT_MOVING && P.cells != nullptr && P.fp_n > 0) {      // uniform branch
      for (int j = 0; j < S; ++j) {
        const bool hit = footprint_lethal_at(P, I, x[j], y[j], c[j], s[j]);
        J += (hit && lg * S + j < P.N) ? P.w_fp_step : 0.0f;                        // srv.py:262-263 per step
      }
    }
    // costmap term of each step and its lethal add-on (srv.py:246-247, 257-260), ONE warp-uniform branch for all steps
    float cmv[S], lwv[S];
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) { cmv[j] = 0.0f; lwv[j] = 0.0f; }
    if (sampled) {
      if (any_sur) {
        NEOMPC_UNROLL
        for (int j = 0; j < S; ++j) {
          const uint32_t w = word[j];
          const bool hx = tx[j] >= 0.5f, hy = ty[j] >= 0.5f;                        // the cell containing the point
          const float k00 = T.cost_at_byte((w << 2) & 0x3fcu), k10 = T.cost_at_byte((w >> 6) & 0x3fcu);   // w_costmap c^2 / N per corner
          const float k01 = T.cost_at_byte((w >> 14) & 0x3fcu), k11 = T.cost_at_byte((w >> 22) & 0x3fcu);
          const float near = hy ? (hx ? k11 : k01) : (hx ? k10 : k00);
          const float ax = k10 - k00, bx = k11 - k01;
          const float k0 = k00 + tx[j] * ax, k1 = k01 + tx[j] * bx;
          const float val = k0 + ty[j] * (k1 - k0);
          const float ggx = ax + ty[j] * (bx - ax), ggy = k1 - k0;                 // per cell
          cmv[j] = near + sf * (val - near);
          cmx[j] = ggx * cqs + ggy * sqs;                                           // gx = .. + cq x - sq y
          cmy[j] = ggy * cqs - ggx * sqs;                                           // gy = .. + sq x + cq y
          lwv[j] = near == P.k_lethal ? P.cm_wl : 0.0f;                             // (the lethal entry is unique)
        }
      } else {
        NEOMPC_UNROLL
        for (int j = 0; j < S; ++j) {
          const uint32_t w = word[j];
          const bool hx = tx[j] >= 0.5f, hy = ty[j] >= 0.5f;
          const float near = T.cost_at_byte(((w >> ((hx ? 8u : 0u) + (hy ? 16u : 0u))) & 0xffu) << 2);
          cmv[j] = near;
          lwv[j] = near == P.k_lethal ? P.cm_wl : 0.0f;
        }
      }
    }
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) {
      const int i = lg * S + j;
      const bool on = F || i < P.N;
      const float ex = I.cx - x[j], ey = I.cy - y[j], eo = I.tyaw - z[j];
      float st = P.a_trans * (ex * ex + ey * ey) + P.b_orient * (eo * eo);          // srv.py:250-252
      const float rx = u[j][0] - I.v0x, ry = u[j][1] - I.v0y, rz = u[j][2] - I.v0z;
      const float r2 = fmaxf(rx * rx + ry * ry + rz * rz + P.eps2, 1e-24f);
      rinv[j] = rsqrt_f(r2);
      st += P.w_ctrl * (r2 * rinv[j]);                                              // srv.py:253-254 (smoothed)
      if (sampled) {
        st += cmv[j];
        st += lwv[j];
      }
      J += on ? st : 0.0f;
      x[j] = on ? -2.0f * P.a_trans * ex + cmx[j] : 0.0f;                           // adjoint seeds
      y[j] = on ? -2.0f * P.a_trans * ey + cmy[j] : 0.0f;
    }
    {                                                                               // terminal yaw term, srv.py:267-268
      if (F) {                                       // full horizon: the last step is step S-1 of the last lane
        const float ef = I.fyaw - z[S - 1];
        J += lg == G - 1 ? P.bt_term * (ef * ef) : 0.0f;
      } else {
        const int jl = P.N - 1 - lg * S;             // local index of the last step, if this lane holds it
        float zl = z[0];
        NEOMPC_UNROLL
        for (int j = 1; j < S; ++j) zl = jl == j ? z[j] : zl;
        const float ef = I.fyaw - zl;
        J += (jl >= 0 && jl < S) ? P.bt_term * (ef * ef) : 0.0f;
      }
    }
    return J;
  }

  template <bool Checked>
  NEOMPC_HD float run(const SolverConst& P, const CostTables& T, const Instance& I, const float (*u)[3], int lg,
                      bool sur) {
    rollout(P, u, lg);
    return cost<Checked>(P, T, I, u, lg, sur);
  }

  // Adjoint of run(): gradient of J w.r.t. this lane's controls (of the reference's objective the smooth part: its
  // costmap / footprint terms are piecewise constant; with guidance the interpolated costmap term contributes
  // through the seeds).  Derivation in DESIGN.md ("Analytic gradient").
  NEOMPC_HD void backward(const SolverConst& P, const Instance& I, const float (*u)[3], int lg,
                          float (*g)[3]) const {
    const float dt = P.dt;
    float gx[S], gy[S], gz[S];
    float sx = 0.0f, sy = 0.0f;
    NEOMPC_UNROLL
    for (int j = S - 1; j >= 0; --j) {
      const int i = lg * S + j;
      const bool on = F || i < P.N;
      gz[j] = on ? -2.0f * P.b_orient * (I.tyaw - z[j]) : 0.0f;
      gz[j] += (F ? (j == S - 1 && lg == G - 1) : (i == P.N - 1)) ? -2.0f * P.bt_term * (I.fyaw - z[j]) : 0.0f;
      sx += x[j]; gx[j] = sx;                   // local inclusive suffix sums of the seeds left by cost()
      sy += y[j]; gy[j] = sy;
    }
    float sxoff, syoff;
    Grp<G>::template excl_suffix2<kXsSuf2>(sx, sy, lg, &sxoff, &syoff);
    float sg = 0.0f;
    NEOMPC_UNROLL
    for (int j = S - 1; j >= 0; --j) {
      gx[j] += sxoff;
      gy[j] += syoff;
      sg += gz[j] - gx[j] * dy[j] + gy[j] * dx[j];
      gz[j] = sg;
    }
    const float sgoff = Grp<G>::template excl_suffix<kXsSufG>(sg, lg);
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) {
      const int i = lg * S + j;
      const bool on = F || i < P.N;
      const float rx = u[j][0] - I.v0x, ry = u[j][1] - I.v0y, rz = u[j][2] - I.v0z;
      const float kk = P.w_ctrl * rinv[j];
      g[j][0] = on ? dt * (c[j] * gx[j] + s[j] * gy[j]) + kk * rx : 0.0f;
      g[j][1] = on ? dt * (-s[j] * gx[j] + c[j] * gy[j]) + kk * ry : 0.0f;
      g[j][2] = on ? dt * (gz[j] + sgoff) + kk * rz : 0.0f;
    }
  }
};

// value the reference's objective has at u, given the smoothed value: replaces sqrt(r^2+eps^2) by r
template <int G, int S>
NEOMPC_HD float unsmooth_correction(const SolverConst& P, const Instance& I, const float (*u)[3], int lg) {
  float d = 0.0f;
  NEOMPC_UNROLL
  for (int j = 0; j < S; ++j) {
    const float rx = u[j][0] - I.v0x, ry = u[j][1] - I.v0y, rz = u[j][2] - I.v0z;
    const float r2 = rx * rx + ry * ry + rz * rz;
    const float v = P.w_ctrl * (sqrtf(r2) - sqrtf(r2 + P.eps2));
    d += (lg * S + j < P.N) ? v : 0.0f;
  }
  return Grp<G>::sum(d);
}

// ---------------------------------------------------------------------------------------------------------
// Footprint cost at the current pose: nav2 FootprintCollisionChecker::footprintCost restated
// (vertices -> cells, every edge rasterised with the closed form of nav2's LineIterator, max cell cost).
// Returns true if the footprint touches a lethal cell or leaves the map (reference test "== 1.0",
// srv.py:262, :343).  Vertex placement is done in float64 so the cell indices equal the oracle's.
// ---------------------------------------------------------------------------------------------------------
template <int G>
NEOMPC_HD bool footprint_lethal(const SolverConst& P, const CostTables& T, double px, double py, double yaw, int lg) {
  if (P.cells == nullptr || P.fp_n <= 0) return false;
  const double cyaw = cos(yaw), syaw = sin(yaw);
  int lethal = 0;
  int mx0 = 0, my0 = 0, mxf = 0, myf = 0;
  for (int v = 0; v <= P.fp_n; ++v) {
    int mx, my;
    if (v < P.fp_n) {
      const double wx = px + ((double)P.fp_x[v] * cyaw - (double)P.fp_y[v] * syaw);
      const double wy = py + ((double)P.fp_x[v] * syaw + (double)P.fp_y[v] * cyaw);
      if (wx < P.origin_x || wy < P.origin_y) { lethal = 1; break; }
      mx = (int)((wx - P.origin_x) * P.inv_res_d);
      my = (int)((wy - P.origin_y) * P.inv_res_d);
      if (mx >= P.W || my >= P.H) { lethal = 1; break; }
      if (v == 0) { mxf = mx; myf = my; mx0 = mx; my0 = my; continue; }
    } else {
      mx = mxf; my = myf;                       // closing edge last -> first
    }
    // edge (mx0,my0) -> (mx,my): pixel k of nav2's LineIterator
    const int ddx = mx - mx0, ddy = my - my0;
    const int adx = ddx < 0 ? -ddx : ddx, ady = ddy < 0 ? -ddy : ddy;
    const int sxs = ddx >= 0 ? 1 : -1, sys = ddy >= 0 ? 1 : -1;
    const bool xmaj = adx >= ady;
    const int den = xmaj ? adx : ady, numadd = xmaj ? ady : adx;
    // floor((den/2 + k numadd) / den) without an integer division per pixel: (v + 0.5) / den is never within
    // 0.5/den of an integer, far more than the float error for v < 2^20, so truncating the float product is exact
    const float inv_den = den > 0 ? 1.0f / (float)den : 0.0f;
    const bool small = den < 1024;
    for (int k = lg; k <= den; k += G) {
      const int v = den / 2 + k * numadd;
      const int minor = den <= 0 ? 0 : small ? (int)(((float)v + 0.5f) * inv_den) : v / den;
      const int cxp = xmaj ? mx0 + k * sxs : mx0 + minor * sxs;
      const int cyp = xmaj ? my0 + minor * sys : my0 + k * sys;
      int cell = kCellOob;
      if (cxp >= 0 && cyp >= 0 && cxp < P.W && cyp < P.H) {
#if defined(__CUDA_ARCH__)
        cell = (int)__ldg(P.cells + (size_t)cyp * P.W + cxp);
#else
        cell = (int)P.cells[(size_t)cyp * P.W + cxp];
#endif
      }
      lethal |= (T.flag[cell] & 1);
    }
    mx0 = mx; my0 = my;
  }
  return Grp<G>::imax(lethal) != 0;
}

// ---------------------------------------------------------------------------------------------------------
// History storage for L-BFGS: element e of this lane lives at hist[e * stride] (stride = threads per
// block in shared memory -> conflict-free; 1 in the host build).
// Per pair p (0..m-1): [3S floats s][3S floats y][rho][alpha scratch of the two-loop recursion]
// ---------------------------------------------------------------------------------------------------------
template <int S>
NEOMPC_HD int hist_floats_per_lane(int m) { return m * (6 * S + 2) + 2 * S + 6 * S; }   // pairs + (av, aw) of precondition() + pg + g

template <int S, bool X>
NEOMPC_HD float projected_gradient(const SolverConst& P, const float (*u)[3], const float (*g)[3], float (*pg)[3]) {
  float pgmax = 0.0f;
  NEOMPC_UNROLL
  for (int j = 0; j < S; ++j) {
    float a = u[j][0] - g[j][0], b = u[j][1] - g[j][1], w = u[j][2] - g[j][2];
    project_step<X>(P, a, b, w);
    pg[j][0] = u[j][0] - a; pg[j][1] = u[j][1] - b; pg[j][2] = u[j][2] - w;
    pgmax = fmaxf(pgmax, fmaxf(fabsf(pg[j][0]), fmaxf(fabsf(pg[j][1]), fabsf(pg[j][2]))));
  }
  return pgmax;
}

// ---------------------------------------------------------------------------------------------------------
// One lane group's solve of one optimizer() call (srv.py:349-403), split into
//   prologue()  request -> per-instance constants, footprint cost, state row, start point   (srv.py:350-361)
//   pass()      one iteration of the projected L-BFGS                                        (srv.py:363-364)
//   epilogue()  low-pass, collision check, accel clamp, state, response                      (srv.py:366-402)
// (one instance per group per launch: solve_instance below).
//
// Projected L-BFGS on the smoothed objective over the feasible set  prod_i (box ∩ disc):
//   x_{k+1} = Proj(x_k + alpha d_k),  d_k = -H_k pg_k  (two-loop recursion whose initial matrix is the block-diagonal
//   preconditioner of precondition(); pg = x - Proj(x - g) is the projected gradient; the secant pairs are those of
//   the map pg, so an active disc constraint contributes its curvature), binding constraints frozen along the step,
//   Armijo backtracking along the projection arc on the true objective incl. the piecewise-constant costmap term;
//   when an arc fails: history dropped -> preconditioned gradient -> plain projected gradient -> stop.
//   The first pass of an instance only evaluates its start point.
// Every collective (shuffle / vote) in here is executed by all 32 lanes of the warp; per-group decisions are
// predicates, never branches around a collective.
// ---------------------------------------------------------------------------------------------------------
template <int G, int S, bool X, bool F = false>
struct Solver {
  static constexpr int PAIR = 6 * S + 2;
  // per-instance constants
  Instance I;
  bool fp_hit, has_instance;
  // iterate
  float u[S][3];                 // (gradient and projected gradient of the iterate live in shared memory: g_at(), pg_at())
  float f, pgmax;
  unsigned iters, evals, status;
  unsigned iters_sw;     // guidance: iteration count at the switch to the second phase
  float last_alpha;      // arc parameter of the last accepted step
  int hist_len, head, small_steps;
  bool active, force_pg, plain, first;
  bool sur;              // costmap guidance: this solve is still on the interpolated costmap term (phase 1 of 2)

  // projected gradient at the current iterate, element e of this lane (kept in shared memory behind the history and the
  // preconditioner constants: 3S registers fewer across the whole pass)
  static NEOMPC_HD float* pg_at(const SolverConst& P, float* hist, int stride) {
    return hist + (size_t)((X ? P.m : 1) * PAIR + 2 * S) * stride;
  }

  static NEOMPC_HD float* g_at(const SolverConst& P, float* hist, int stride) {
    return hist + (size_t)((X ? P.m : 1) * PAIR + 5 * S) * stride;
  }

  NEOMPC_HD void prologue(const SolverConst& P, const CostTables& T, const neompc_request& rq, bool valid, int lg,
                          float* hist, int stride) {
    // (every lane of the warp runs the collective inside footprint_lethal, also for groups without an instance)
    const bool fp_any = footprint_lethal<G>(P, T, (double)rq.pose_x, (double)rq.pose_y, (double)rq.pose_yaw, lg);
    init(P, rq, fp_any, valid, lg, hist, stride);
  }

  // collective-free part of the prologue: may be executed by a subset of the groups of a warp
  NEOMPC_HD void init(const SolverConst& P, const neompc_request& rq, bool fp_any, bool valid, int lg,
                      float* hist, int stride) {
    has_instance = valid;
    I = make_instance(P, rq);
    fp_hit = valid && fp_any;

    // per-instance state and the new-goal reset (srv.py:358-361); the epilogue reads the row again
    const Carry cr = load_carry(P, rq, valid);
    const bool stateful = cr.stateful, new_goal = cr.new_goal;
    const float* row = stateful ? P.state + (size_t)rq.instance_id * P.state_stride : nullptr;
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) {
      const int i = lg * S + j;
      const bool ld = stateful && !new_goal && i < P.N;
      u[j][0] = ld ? row[3 * i + 0] : 0.0f;          // warm start (srv.py:397-400) or zeros (srv.py:136,359)
      u[j][1] = ld ? row[3 * i + 1] : 0.0f;
      u[j][2] = ld ? row[3 * i + 2] : 0.0f;
      project_step<X>(P, u[j][0], u[j][1], u[j][2]);
    }
    sur = P.guided != 0 && P.cells4 != nullptr && !(X && P.cm_mode == NEOMPC_COSTMAP_BILINEAR);
    const int m = X ? P.m : 1;                     // the fast path is dispatched only for one history pair
    for (int e = 0; e < m * PAIR; ++e) hist[(size_t)e * stride] = 0.0f;
    for (int e = 0; e < 6 * S; ++e) pg_at(P, hist, stride)[(size_t)e * stride] = 0.0f;       // pg and g
    {
      float* tab = hist + (size_t)(m * PAIR) * stride;               // tracking-term majorants of precondition()
      const float dt2 = 2.0f * P.dt * P.dt;
      NEOMPC_UNROLL
      for (int j = 0; j < S; ++j) {
        const int i = lg * S + j;
        const int rem = P.N - i > 0 ? P.N - i : 1;                   // padded steps: any positive value
        const float rowsum = (float)(rem * (i + 1) + ((rem - 1) * rem) / 2);
        // Floors: with w_control = 0 and w_trans = 0 (or w_orient = 0) a block would be singular.  Without an
        // orientation weight the heading still has curvature through the positions it swings (~ av * lever^2, lever
        // up to max_vel_trans * horizon): 5 % of av stands in for that.
        // With guidance the interpolated costmap term bends over one cell by about its own size: a fraction of
        // (w_costmap / N) / resolution^2 keeps the position block from vanishing when w_trans is (nearly) zero.
        const float acm = sur ? P.cm_curv * P.cm_w * P.inv_res * P.inv_res : 0.0f;
        const float av = fmaxf(dt2 * fmaxf(P.a_trans, acm) * rowsum, 1e-12f);
        tab[(size_t)(2 * j) * stride] = av;
        tab[(size_t)(2 * j + 1) * stride] = fmaxf(dt2 * (P.b_orient * rowsum + P.bt_term * (float)P.N), 0.05f * av);
      }
    }
    f = 0.0f; pgmax = 0.0f;
    iters = 0; evals = 0; status = NEOMPC_STATUS_MAXITER; iters_sw = 0xffffffffu; last_alpha = 1.0f;
    hist_len = 0; head = 0; small_steps = 0;
    active = valid; force_pg = true; plain = false; first = true;
  }

  // r <- H0 r where H0 is the inverse of a block-diagonal majorant of the model Hessian at u.  Per step i:
  //   diag(av, av, aw) + k (I - q rr^T),   r = u_i - v0,  q = 1/(|r|^2 + eps^2),  k = (w_control/N) sqrt(q)
  // k(...) is the exact Hessian of the smoothed control term (srv.py:253-254), which dominates at the README weights.
  // av, aw bound the tracking terms (srv.py:250-252, 267-268): their Hessian in the step velocities is
  // 2 a dt^2 M (x) R_k^T R_l with M_kl = N - max(k,l) (positions are cumulative sums of the steps); M has positive
  // entries, so diag(row sums of M) - M is diagonally dominant, i.e. the row sums majorise it:
  //   rowsum_i = (N-i)(i+1) + (N-i-1)(N-i)/2,  av = 2 (w_trans/N) dt^2 rowsum_i,
  //   aw = 2 (w_orient/N) dt^2 rowsum_i + 2 w_orient w_terminal dt^2 N.
  // The 3x3 block is inverted in closed form (Sherman-Morrison).  `on` = false leaves r unchanged (predicated).
  NEOMPC_HD void precondition(const SolverConst& P, const float* hist, int stride, bool on, float (*r)[3]) const {
    const float* tab = hist + (size_t)((X ? P.m : 1) * PAIR) * stride;   // (av_j, aw_j) written by init()
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) {
      const float rx = u[j][0] - I.v0x, ry = u[j][1] - I.v0y, rz = u[j][2] - I.v0z;
      const float ri = rsqrt_f(rx * rx + ry * ry + rz * rz + P.eps2);
      const float q = ri * ri;
      const float k = P.w_ctrl * ri;
      const float iv = rcp_approx(tab[(size_t)(2 * j) * stride] + k);
      const float iw = rcp_approx(tab[(size_t)(2 * j + 1) * stride] + k);
      const float tx = iv * rx, ty = iv * ry, tz = iw * rz;
      const float kq = k * q;
      const float den = 1.0f - kq * (rx * tx + ry * ty + rz * tz);
      const float c = div_approx(kq * (tx * r[j][0] + ty * r[j][1] + tz * r[j][2]), den);
      r[j][0] = on ? iv * r[j][0] + c * tx : r[j][0];
      r[j][1] = on ? iv * r[j][1] + c * ty : r[j][1];
      r[j][2] = on ? iw * r[j][2] + c * tz : r[j][2];
    }
  }

  // one iteration for every group of the warp (inactive groups compute and discard)
  NEOMPC_HD void pass(const SolverConst& P, const CostTables& T, float* hist, int stride, int lg) {
    const int m = X ? P.m : 1;                     // compile-time 1 on the fast path: the two loops below unroll away
    Forward<G, S, X, F> fw;
    float d[S][3], xt[S][3], r[S][3];

    // ---- direction: two-loop recursion on the projected gradient (loops rolled: small code)
    const bool use_qn = !first && !force_pg && hist_len > 0;
    float* const pgs = pg_at(P, hist, stride);
    float* const gsm = g_at(P, hist, stride);
    NEOMPC_UNROLL
    for (int e = 0; e < 3 * S; ++e) r[e / 3][e % 3] = pgs[(size_t)e * stride];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int k = 0; k < m; ++k) {
      const bool on = use_qn && k < hist_len;
      int p = head - 1 - k; if (p < 0) p += m;
      float* sp = hist + (size_t)(p * PAIR) * stride;
      const float* yp = sp + (size_t)(3 * S) * stride;
      float dot = 0.0f;
      NEOMPC_UNROLL
      for (int e = 0; e < 3 * S; ++e) dot += sp[(size_t)e * stride] * r[e / 3][e % 3];
      dot = Grp<G>::template sum<X ? kXsAny : kXsDot1>(dot);     // (m > 1: the site repeats back to back -> barrier-closed site 0)
      const float a = on ? sp[(size_t)(6 * S) * stride] * dot : 0.0f;
      sp[(size_t)(6 * S + 1) * stride] = a;
      NEOMPC_UNROLL
      for (int e = 0; e < 3 * S; ++e) r[e / 3][e % 3] -= a * yp[(size_t)e * stride];
    }
    const bool use_pc = !plain;                    // initial matrix of the recursion: the block-diagonal preconditioner
    precondition(P, hist, stride, use_pc, r);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int k = m - 1; k >= 0; --k) {
      const bool on = use_qn && k < hist_len;
      int p = head - 1 - k; if (p < 0) p += m;
      const float* sp = hist + (size_t)(p * PAIR) * stride;
      const float* yp = sp + (size_t)(3 * S) * stride;
      float dot = 0.0f;
      NEOMPC_UNROLL
      for (int e = 0; e < 3 * S; ++e) dot += yp[(size_t)e * stride] * r[e / 3][e % 3];
      dot = Grp<G>::template sum<X ? kXsAny : kXsDot2>(dot);
      const float b = on ? sp[(size_t)(6 * S + 1) * stride] - sp[(size_t)(6 * S) * stride] * dot : 0.0f;
      NEOMPC_UNROLL
      for (int e = 0; e < 3 * S; ++e) r[e / 3][e % 3] += b * sp[(size_t)e * stride];
    }
    // d = -r; it must be a descent direction for the projected gradient, else restart from -pg
    float gd = 0.0f, pgn2 = 0.0f;
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) {
      NEOMPC_UNROLL
      for (int q = 0; q < 3; ++q) {
        const float pv = pgs[(size_t)(3 * j + q) * stride];
        d[j][q] = -r[j][q];
        r[j][q] = pv;                            // (r is free from here on: keeps pg for the fallback direction below)
        gd += pv * d[j][q];
        pgn2 += pv * pv;
      }
    }
    // descent test  pg.d < -1e-4 |pg|^2  as ONE group sum
    const float desc = Grp<G>::template sum<kXsDesc>(gd + 1e-4f * pgn2);
    const bool qn_dir = (use_qn || use_pc) && desc < 0.0f;
    // Binding constraints (at the boundary with the gradient pushing outward) stay fixed along the step
    // (two-metric projection): omega at a bound -> no omega step; (vx, vy) on the circle -> tangential step only.
    // Their projected gradient is zero, so pg . d — and with it the descent property — is unchanged.
    if (!X || P.disc_only) {
      NEOMPC_UNROLL
      for (int j = 0; j < S; ++j) {
        const float n2 = u[j][0] * u[j][0] + u[j][1] * u[j][1];
        const float gr = gsm[(size_t)(3 * j) * stride] * u[j][0] + gsm[(size_t)(3 * j + 1) * stride] * u[j][1];
        const bool bind = n2 >= P.R * P.R * (1.0f - 2e-6f) && gr < 0.0f;
        const float dr = div_approx(d[j][0] * u[j][0] + d[j][1] * u[j][1], fmaxf(n2, 1e-30f));
        d[j][0] -= bind ? dr * u[j][0] : 0.0f;
        d[j][1] -= bind ? dr * u[j][1] : 0.0f;
      }
    }
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) {
      const float gw = gsm[(size_t)(3 * j + 2) * stride];
      const bool at_lo = u[j][2] <= P.lo[2] && gw > 0.0f;
      const bool at_hi = u[j][2] >= P.hi[2] && gw < 0.0f;
      d[j][2] = (at_lo || at_hi) ? 0.0f : d[j][2];
    }
    float alpha = 1.0f;
    if (!qn_dir) {
      NEOMPC_UNROLL
      for (int j = 0; j < S; ++j) { d[j][0] = -r[j][0]; d[j][1] = -r[j][1]; d[j][2] = -r[j][2]; }
      // first trial moves the largest component by about the velocity range
      alpha = fmaxf(1.0f, div_approx(P.R, fmaxf(pgmax, 1e-12f)));
    }
    // Second phase of a guided solve: arcs end at cost steps, and the arc parameter that was accepted last time predicts
    // the next one far better than the unit step does — the first trial is bounded by alpha_warm x it (evaluation slots per
    // warp of C3 51 -> 44; applied to the first phase too it costs the C2 tail: profiles/solver_tuning_r2.txt)
    if (iters_sw != 0xffffffffu && qn_dir) alpha = fminf(alpha, P.alpha_warm * last_alpha);
    if (first) alpha = 0.0f;                     // evaluate the start point itself

    // ---- Armijo backtracking along the projection arc (all groups of the warp in lock step)
    bool ls_done = !active, accepted = false;
    float ft = f;
    for (int bt = 0; bt < kMaxBacktracks; ++bt) {
      float gs = 0.0f;
      NEOMPC_UNROLL
      for (int j = 0; j < S; ++j) {
        xt[j][0] = u[j][0] + alpha * d[j][0];
        xt[j][1] = u[j][1] + alpha * d[j][1];
        xt[j][2] = u[j][2] + alpha * d[j][2];
        project_step<X>(P, xt[j][0], xt[j][1], xt[j][2]);
        gs += gsm[(size_t)(3 * j) * stride] * (xt[j][0] - u[j][0]) + gsm[(size_t)(3 * j + 1) * stride] * (xt[j][1] - u[j][1]) +
              gsm[(size_t)(3 * j + 2) * stride] * (xt[j][2] - u[j][2]);
      }
      float ftrial = fw.template run<false>(P, T, I, xt, lg, sur);
      Grp<G>::template sum2<kXsJ>(gs, ftrial);
      NEOMPC_TRACE("   trial bt %d alpha %.4g gs %.4e df %.4e gd %.4e\n", bt, alpha, gs, ftrial - f, gd);
      // (bookkeeping as selects: groups of a warp differ here, and two small divergent branches per trial cost more than
      //  the few instructions they would skip)
      {
        const bool run = !ls_done;
        evals += run ? 1u : 0u;
        ft = run ? ftrial : ft;
        const bool ok = run && (first || (ftrial <= f + 1e-4f * gs && gs < 0.0f));
        ls_done = ls_done || ok;
        accepted = accepted || ok;
      }
      if (!Grp<G>::warp_any(!ls_done)) break;
      {
        const bool still = !ls_done;
        // safeguarded quadratic interpolation of the step length
        const float denom = 2.0f * (ftrial - f - gs);
        const float aq = denom > 0.0f ? div_approx(-gs, denom) : 0.5f;
        alpha = still ? alpha * fminf(0.5f, fmaxf(0.1f, aq)) : alpha;
        // not a descent arc at all: give up on this direction at once (the fallback below takes over)
        ls_done = ls_done || (still && gs >= 0.0f && bt == 0 && qn_dir);
      }
    }
    NEOMPC_TRACE("it %u f %.7f pgmax %.3e qn %d alpha %.4g acc %d ft %.7f evals %u hist %d\n", iters, f, pgmax,
                 (int)qn_dir, alpha, (int)accepted, ft, evals, hist_len);
    for (int j = 0; j < S; ++j)
      NEOMPC_TRACE("      u %+.4f %+.4f %+.4f   pg %+.2e %+.2e %+.2e  g %+.2e %+.2e %+.2e  d %+.2e %+.2e %+.2e\n", u[j][0],
                   u[j][1], u[j][2], pgs[(size_t)(3 * j) * stride], pgs[(size_t)(3 * j + 1) * stride], pgs[(size_t)(3 * j + 2) * stride],
                   gsm[(size_t)(3 * j) * stride], gsm[(size_t)(3 * j + 1) * stride], gsm[(size_t)(3 * j + 2) * stride], d[j][0], d[j][1], d[j][2]);

    // ---- gradient and projected gradient at the last trial point (uniform work for the whole warp)
    float gn[S][3], pgn[S][3];
    fw.backward(P, I, xt, lg, gn);
    float pgmax_n = projected_gradient<S, X>(P, xt, gn, pgn);
    // secant pair of the projected-gradient map: s = x+ - x, y = pg(x+) - pg(x).  On an active disc
    // constraint y carries the curvature of the constraint, which a pair of plain gradients would miss.
    float sy = 0.0f, yy = 0.0f, smax = 0.0f;
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) {
      NEOMPC_UNROLL
      for (int q = 0; q < 3; ++q) {
        const float sv = xt[j][q] - u[j][q], yv = pgn[j][q] - pgs[(size_t)(3 * j + q) * stride];
        sy += sv * yv; yy += yv * yv;
        smax = fmaxf(smax, fabsf(sv));
        d[j][q] = sv; r[j][q] = yv;            // reuse as (s, y)
      }
    }
    Grp<G>::template sum2_max2<kXsPair>(sy, yy, smax, pgmax_n);
    if (active) {
      if (accepted) {
        if (!first && sy > P.pair_eps * yy && yy > 0.0f) {
          float* sp = hist + (size_t)(head * PAIR) * stride;
          NEOMPC_UNROLL
          for (int e = 0; e < 3 * S; ++e) {
            sp[(size_t)e * stride] = d[e / 3][e % 3];
            sp[(size_t)(3 * S + e) * stride] = r[e / 3][e % 3];
          }
          sp[(size_t)(6 * S) * stride] = div_approx(1.0f, sy);
          head = head + 1 == m ? 0 : head + 1;
          hist_len = hist_len < m ? hist_len + 1 : m;
          force_pg = false;
          plain = false;
        }
        const float df = f - ft;
        NEOMPC_UNROLL
        for (int j = 0; j < S; ++j) {
          NEOMPC_UNROLL
          for (int q = 0; q < 3; ++q) {
            u[j][q] = xt[j][q];
            gsm[(size_t)(3 * j + q) * stride] = gn[j][q];
            pgs[(size_t)(3 * j + q) * stride] = pgn[j][q];
          }
        }
        f = ft;
        pgmax = pgmax_n;
        if (!first) last_alpha = alpha;
        if (!first) {
          ++iters;
          // secondary stop: the objective stopped moving (relative) for two accepted steps in a row
          // (a step that short means the arc search is pinned at a costmap cell edge)
          const float ts = sur ? P.sur_tol : 1.0f;        // the guided phase is followed by a second one: looser stop
          if (df <= ts * P.tol_f * fmaxf(1.0f, fabsf(f)) || smax <= ts * P.tol_x || alpha <= P.pin_alpha) ++small_steps;
          else small_steps = 0;
          if (small_steps >= 2) { active = false; status = NEOMPC_STATUS_CONVERGED; }
        }
      } else if (qn_dir && use_qn) {
        hist_len = 0; head = 0; force_pg = true;    // quasi-Newton arc failed: restart without history
        ++iters;
      } else if (qn_dir) {
        plain = true;                                // the preconditioned arc failed as well: plain projected gradient
        ++iters;
      } else {
        active = false;                              // projected-gradient arc failed too
        status = NEOMPC_STATUS_LINESEARCH;
      }
      // ---- convergence test on the projected gradient  pg = x - Proj(x - g)
      if (active && pgmax <= (sur ? P.sur_tol : 1.0f) * P.tol_pg) { active = false; status = NEOMPC_STATUS_CONVERGED; }
      if (active && (int)iters >= P.max_iter) { active = false; status = NEOMPC_STATUS_MAXITER; }
      // the second phase of a guided solve refines a point that is already within a cell of its optimum: bounded
      if (active && !first && iters_sw != 0xffffffffu && iters - iters_sw >= (unsigned)P.polish_max) {
        active = false; status = NEOMPC_STATUS_CONVERGED;
      }
    }
    first = false;
    // Costmap guidance, end of the first phase: the guided solve has settled on the interpolated costmap term; the solve
    // continues from that point on the reference's objective (the next pass re-evaluates the point there).
    if (sur && !active && has_instance && status != NEOMPC_STATUS_MAXITER) {
      sur = false;
      iters_sw = iters; last_alpha = 1.0f;
      active = true; first = true; force_pg = true; plain = false;
      hist_len = 0; head = 0; small_steps = 0;
      status = NEOMPC_STATUS_MAXITER;
    }
  }

  // post-solve part of optimizer() (srv.py:365-402).  Executed by every lane of the warp (it contains collectives);
  // only groups with `fin` set (a finished instance) write response / twist / plan / state.  Solver state is not
  // modified, so groups that are still iterating pass through unharmed.
  NEOMPC_HD void epilogue(const SolverConst& P, const CostTables& T, const neompc_request& rq, bool fin, int lg,
                          neompc_response* resp, float* twist, float* plan) const {
    const bool valid = has_instance && fin;
    const float j_true = f + unsmooth_correction<G, S>(P, I, u, lg) + constant_cost(P, rq, fp_hit);
    const Carry cr = load_carry(P, rq, valid);
    const float* last = cr.last;
    const bool stateful = cr.stateful, new_goal = cr.new_goal;
    float ct, st;                                      // cos/sin of the true pose yaw (srv.py:317)
    sincos_f(rq.pose_yaw, &st, &ct);
    st *= P.inv_res; ct *= P.inv_res;
    if (valid && plan != nullptr) {                   // the raw solution x.x (what publishLocalPlan gets, srv.py:365)
      NEOMPC_UNROLL
      for (int j = 0; j < S; ++j) {
        const int i = lg * S + j;
        if (i < P.N) { plan[3 * i + 0] = u[j][0]; plan[3 * i + 1] = u[j][1]; plan[3 * i + 2] = u[j][2]; }
      }
    }
    // ---- low-pass on the first control (srv.py:366-367; the reference does it in place on x.x)
    float ue[S][3];
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) { ue[j][0] = u[j][0]; ue[j][1] = u[j][1]; ue[j][2] = u[j][2]; }
    if (lg == 0) {
      NEOMPC_UNROLL
      for (int q = 0; q < 3; ++q) ue[0][q] = ue[0][q] * P.lp_gain + last[q] * (1.0f - P.lp_gain);
    }
    // ---- collision_check: re-roll with the TRUE yaw (srv.py:312-347)
    Forward<G, S, X> fw;
    fw.rollout(P, ue, lg);
    int hit = 0;
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) {
      const int cell = Forward<G, S, X>::cell_of(P, I, ct, st, fw.x[j], fw.y[j]);
      hit |= (lg * S + j < P.N) ? (T.flag[cell] >> 1) & 1 : 0;                          // col >= 0.99, srv.py:338
    }
    hit = Grp<G>::imax(hit);
    bool collision = cr.latched || hit != 0;
    float wait = cr.waiting;
    float o[3];
    o[0] = Grp<G>::bcast0(ue[0][0], lg);
    o[1] = Grp<G>::bcast0(ue[0][1], lg);
    o[2] = Grp<G>::bcast0(ue[0][2], lg);
    unsigned flags = 0;
    if (collision || fp_hit) {                                                          // srv.py:374-382
      o[0] = o[1] = o[2] = 0.0f;
      flags |= NEOMPC_FLAG_STOPPED;
      wait += rq.delta_t;
      if (wait >= 3.0f) { collision = false; wait = 0.0f; }
    } else {                                                                            // srv.py:385-391
      NEOMPC_UNROLL
      for (int q = 0; q < 3; ++q) {
        const float lim = P.acc[q] * rq.control_interval;
        o[q] = fmaxf(fminf(o[q], last[q] + lim), last[q] - lim);
      }
    }
    if (collision) flags |= NEOMPC_FLAG_COLLISION;
    if (fp_hit) flags |= NEOMPC_FLAG_COLLISION_FOOTPRINT;
    if (new_goal) flags |= NEOMPC_FLAG_NEW_GOAL;
    if (cr.no_state) flags |= NEOMPC_FLAG_NO_STATE;          // solved as a cold start; the host entry points return NEOMPC_ERR_STATE
    if (!valid) return;
    // ---- state: last_control (srv.py:393-395), warm start (srv.py:397-400), old_goal (srv.py:402)
    if (stateful) {
      float* row = P.state + (size_t)rq.instance_id * P.state_stride;
      float* tail = row + 3 * P.N;
      const bool success = status != NEOMPC_STATUS_MAXITER;
      NEOMPC_UNROLL
      for (int j = 0; j < S; ++j) {
        const int i = lg * S + j;
        if (i < P.N) {
          // success: plan shifted left by one step, tail = low-passed first control (srv.py:198-202)
          const int dst = success ? (i == 0 ? P.N - 1 : i - 1) : i;
          row[3 * dst + 0] = ue[j][0]; row[3 * dst + 1] = ue[j][1]; row[3 * dst + 2] = ue[j][2];
        }
      }
      if (lg == 0) {
        tail[0] = o[0]; tail[1] = o[1]; tail[2] = o[2];
        tail[3] = wait;
        tail[4] = collision ? 1.0f : 0.0f;
        tail[5] = rq.goal_x; tail[6] = rq.goal_y; tail[7] = rq.goal_yaw;
        tail[8] = 1.0f;
        tail[9] = fp_hit ? 1.0f : 0.0f;
      }
    }
    if (lg == 0) {
      neompc_response rs;
      rs.vx = o[0]; rs.vy = o[1]; rs.omega = o[2];
      rs.cost = j_true;
      rs.iters = iters; rs.evals = evals; rs.status = status; rs.flags = flags;
      *resp = rs;
      if (cr.no_state && P.err_word != nullptr) *P.err_word = 1u;
      if (twist != nullptr) { twist[0] = o[0]; twist[1] = o[1]; twist[2] = o[2]; }
    }
  }
};

// One optimizer() call for the instance owned by this lane group (one instance per group per launch).
template <int G, int S, bool X, bool F = false>
NEOMPC_HD void solve_instance(const SolverConst& P, const CostTables& T, const neompc_request& rq, bool valid,
                              int lg, float* hist, int stride, neompc_response* resp, float* twist, float* plan) {
  Solver<G, S, X, F> sv;
  sv.prologue(P, T, rq, valid, lg, hist, stride);
  while (Grp<G>::warp_any(sv.active)) sv.pass(P, T, hist, stride, lg);
  sv.epilogue(P, T, rq, true, lg, resp, twist, plan);
}

// objective value (reference J, unsmoothed) and gradient (smoothed objective) at a given u — test hook
template <int G, int S, bool X>
NEOMPC_HD void eval_instance(const SolverConst& P, const CostTables& T, const neompc_request& rq, bool valid, int lg,
                             const float* uin, float* Jout, float* gout) {
  const Instance I = make_instance(P, rq);
  const bool fp_any = footprint_lethal<G>(P, T, (double)rq.pose_x, (double)rq.pose_y, (double)rq.pose_yaw, lg);
  const bool fp_hit = valid && fp_any;
  float u[S][3], g[S][3];
  NEOMPC_UNROLL
  for (int j = 0; j < S; ++j) {
    const int i = lg * S + j;
    const bool ld = valid && i < P.N;
    u[j][0] = ld ? uin[3 * i + 0] : 0.0f;
    u[j][1] = ld ? uin[3 * i + 1] : 0.0f;
    u[j][2] = ld ? uin[3 * i + 2] : 0.0f;
  }
  Forward<G, S, X> fw;
  const float f = Grp<G>::sum(fw.template run<true>(P, T, I, u, lg, false));
  const float jt = f + unsmooth_correction<G, S>(P, I, u, lg) + constant_cost(P, rq, fp_hit);
  fw.backward(P, I, u, lg, g);
  if (!valid) return;
  if (lg == 0) *Jout = jt;
  if (gout != nullptr) {
    NEOMPC_UNROLL
    for (int j = 0; j < S; ++j) {
      const int i = lg * S + j;
      if (i < P.N) { gout[3 * i + 0] = g[j][0]; gout[3 * i + 1] = g[j][1]; gout[3 * i + 2] = g[j][2]; }
    }
  }
}

}  // namespace neompc

// mpc_setup.h — host-side derivation of the solver constants and cost tables from neompc_params
// (plain C++; shared by the CUDA runtime and by the tests/hostsim tooling).
#pragma once

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "mpc_core.cuh"

namespace neompc {

struct HostTables {
  std::vector<float> cost;     // kTableSize
  std::vector<uint8_t> flag;   // kTableSize
};

inline double cell_cost(int encoding, int byte) {
  // declared costmap semantics (oracle/costmap.py: cost_lut)
  if (encoding == NEOMPC_ENC_OCCUPANCY) return (byte >= 0 && byte <= 100) ? byte / 100.0 : 0.0;
  return (byte >= 0 && byte <= 254) ? byte / 254.0 : 0.0;
}

// lut_cost[b] = w_costmap * c^2 / N (srv.py:247, 260); lethal cells add (1000 - w_costmap) / N in the kernel (srv.py:257-258),
// which recognises them by their table value: lethal_entry() is unique to c == 1.0 (1e-30 stands in when w_costmap = 0)
inline float lethal_entry(const neompc_params& p) {
  return fmaxf((float)((double)p.w_costmap / p.control_steps), 1e-30f);
}
inline void build_tables(const neompc_params& p, int encoding, HostTables& t) {
  t.cost.assign(kTableSize, 0.0f);          // entry kCellFree (no costmap) stays 0
  t.flag.assign(kTableSize, 0);
  for (int b = 0; b <= 256; ++b) {
    const double c = b == 256 ? 1.0 : cell_cost(encoding, b);
    t.cost[b] = c == 1.0 ? lethal_entry(p) : (float)((double)p.w_costmap * c * c / p.control_steps);
    t.flag[b] = (uint8_t)((c == 1.0 ? 1 : 0) | (c >= 0.99 ? 2 : 0));
  }
}

// Padding of the corner-packed map: twice what a feasible plan can travel (the disc constraint bounds every step's speed
// by max_vel_trans, srv.py:157-158) in cells — once for starts that lie outside the map but within reach of it, once for
// the plan itself — plus the interpolation neighbour and slack (mpc_core.cuh: make_instance).
// A reach of more than kMaxCornerReach cells (a very fine grid under a long horizon) would make the padded copy large:
// the padding is capped and the solver samples with bounds tests instead (*pad_ok = 0).
constexpr int kMaxCornerReach = 125;
inline int corner_pad_for(const neompc_params& p, double resolution, int* pad_ok = nullptr) {
  const double reach = std::ceil((double)p.max_vel_trans * (double)p.prediction_horizon / resolution);
  const bool ok = reach <= (double)kMaxCornerReach;
  if (pad_ok) *pad_ok = ok ? 1 : 0;
  return ok ? 2 * (int)reach + 6 : 8;
}

// corner-packed copy of a host costmap (mpc_core.cuh: corner_word); the CUDA runtime builds the same on the device
inline void build_corner_map(const uint8_t* cells, int W, int H, int lethal_byte, int pad, std::vector<uint32_t>& out) {
  out.resize(corner_words(W, H, pad));
  const int pitch = corner_pitch(W, pad);
  for (int iy = -pad; iy < H + pad; ++iy)
    for (int ix = -pad; ix < W + pad; ++ix)
      out[(size_t)(iy + pad) * pitch + (ix + pad)] = corner_word(cells, W, H, lethal_byte, ix, iy);
}

inline bool validate_params(const neompc_params& p, std::string& err) {
  if (p.control_steps < 1 || p.control_steps > NEOMPC_MAX_CONTROL_STEPS) {
    err = "control_steps must be in 1.." + std::to_string(NEOMPC_MAX_CONTROL_STEPS);
    return false;
  }
  if (!(p.prediction_horizon > 0.0f)) { err = "prediction_horizon must be > 0"; return false; }
  if (!(p.max_vel_trans > 0.0f)) { err = "max_vel_trans must be > 0"; return false; }
  if (!(p.min_vel_x <= p.max_vel_x && p.min_vel_y <= p.max_vel_y && p.min_vel_theta <= p.max_vel_theta)) {
    err = "min_vel_* must not exceed max_vel_*";
    return false;
  }
  // the feasible set box ∩ disc must be non-empty: the box point closest to the origin lies in the disc
  const float nx = fminf(fmaxf(0.0f, p.min_vel_x), p.max_vel_x);
  const float ny = fminf(fmaxf(0.0f, p.min_vel_y), p.max_vel_y);
  if (nx * nx + ny * ny > p.max_vel_trans * p.max_vel_trans) {
    err = "velocity box and max_vel_trans disc do not intersect";
    return false;
  }
  if (p.lbfgs_memory < 0 || p.lbfgs_memory > kMaxMemory) { err = "lbfgs_memory must be 0..8"; return false; }
  if (p.max_iterations < 0) { err = "max_iterations must be >= 0"; return false; }
  if (p.costmap_mode != NEOMPC_COSTMAP_NEAREST && p.costmap_mode != NEOMPC_COSTMAP_BILINEAR) {
    err = "costmap_mode must be NEOMPC_COSTMAP_NEAREST or NEOMPC_COSTMAP_BILINEAR";
    return false;
  }
  if (p.footprint_mode != NEOMPC_FOOTPRINT_STATIC && p.footprint_mode != NEOMPC_FOOTPRINT_MOVING) {
    err = "footprint_mode must be NEOMPC_FOOTPRINT_STATIC or NEOMPC_FOOTPRINT_MOVING";
    return false;
  }
  if (p.costmap_guidance != NEOMPC_GUIDANCE_ON && p.costmap_guidance != NEOMPC_GUIDANCE_OFF) {
    err = "costmap_guidance must be NEOMPC_GUIDANCE_ON or NEOMPC_GUIDANCE_OFF";
    return false;
  }
  if (!(p.opt_tolerance > 0.0f)) { err = "opt_tolerance must be > 0"; return false; }
  const int g = p.lanes_per_instance;
  if (!(g == 0 || g == 1 || g == 2 || g == 3 || g == 4 || g == 5 || g == 6 || g == 8 || g == 10 || g == 16 || g == 32)) {
    err = "lanes_per_instance must be 0 (auto) or one of 1, 2, 3, 4, 5, 6, 8, 10, 16, 32";
    return false;
  }
  return true;
}

// Solver tolerances derived from the reference's opt_tolerance (SLSQP's ftol there).  See DESIGN.md
// "Meaning of opt_tolerance": the projected-gradient sup-norm must fall below kPgScale * opt_tolerance, or the
// objective must stop decreasing by more than kFScale * opt_tolerance (relative) twice in a row.
constexpr float kPgScale = 0.2f;
constexpr float kFScale = 4e-3f;
constexpr float kXScale = 0.2f;
constexpr float kCmCurvature = 0.02f;   // SolverConst::cm_curv
constexpr float kGuidedTolScale = 2.0f;  // SolverConst::sur_tol
constexpr int kPolishMaxIterations = 1000; // SolverConst::polish_max
constexpr float kAlphaWarm = 4.0f;       // SolverConst::alpha_warm
      // SolverConst::skip_polish in units of opt_tolerance

inline void build_const(const neompc_params& p, SolverConst& c) {
  std::memset(&c, 0, sizeof(c));
  const int N = p.control_steps;
  c.N = N;
  // with the block-diagonal preconditioner one pair does as well as 3 or 6 (profiles/solver_tuning_r1.txt)
  c.m = p.lbfgs_memory > 0 ? p.lbfgs_memory : 1;
  c.max_iter = p.max_iterations > 0 ? p.max_iterations : 100;
  c.dt = p.prediction_horizon / (float)N;
  c.a_trans = p.w_trans / (float)N;
  c.b_orient = p.w_orient / (float)N;
  c.w_ctrl = p.w_control / (float)N;
  c.bt_term = p.w_orient * p.w_terminal;
  c.wt_term = p.w_trans * p.w_terminal;
  c.fp_mode = p.footprint_mode;
  c.w_fp = p.footprint_mode == NEOMPC_FOOTPRINT_MOVING ? 0.0f : p.w_footprint;
  c.w_fp_step = p.footprint_mode == NEOMPC_FOOTPRINT_MOVING ? p.w_footprint / (float)N : 0.0f;
  c.lethal_byte = 100;                            // NEOMPC_ENC_OCCUPANCY; set_costmap updates it with the encoding
  c.cm_mode = p.costmap_mode;
  c.cm_scale = 1.0f / 100.0f;                     // likewise updated with the encoding
  c.cm_w = p.w_costmap / (float)N;
  c.cm_wl = (1000.0f - p.w_costmap) / (float)N;
  // default smoothing length of the control-term kink: 10 x opt_tolerance within [1e-4, 1e-2] m/s — 1e-2 at the README's
  // opt_tolerance = 1e-3; 1e-4 at the code default 1e-5, where w_control = 0.5 makes the bias (<= w_control eps / N per
  // kinked step) visible against a tightly converged scipy: 1e-3 left 8 % of free-space problems worse than scipy by
  // more than 1e-4, 1e-4 none (profiles/solver_tuning_r2.txt)
  const float eps = p.control_smoothing > 0.0f ? fmaxf(p.control_smoothing, 1e-6f)
                                               : fminf(1e-2f, fmaxf(1e-4f, 10.0f * p.opt_tolerance));
  c.eps2 = eps * eps;
  c.lo[0] = p.min_vel_x; c.lo[1] = p.min_vel_y; c.lo[2] = p.min_vel_theta;
  c.hi[0] = p.max_vel_x; c.hi[1] = p.max_vel_y; c.hi[2] = p.max_vel_theta;
  c.R = p.max_vel_trans;
  c.disc_only = (p.max_vel_trans <= fminf(fminf(-p.min_vel_x, p.max_vel_x), fminf(-p.min_vel_y, p.max_vel_y))) ? 1 : 0;
  c.fast_trig = (fmaxf(fabsf(p.min_vel_theta), fabsf(p.max_vel_theta)) * p.prediction_horizon <= 3.14159265f) ? 1 : 0;
  c.acc[0] = p.acc_x_limit; c.acc[1] = p.acc_y_limit; c.acc[2] = p.acc_theta_limit;
  c.lp_gain = p.low_pass_gain;
  c.tol_pg = kPgScale * p.opt_tolerance;
  c.tol_f = kFScale * p.opt_tolerance;
  c.tol_x = kXScale * p.opt_tolerance;
  // the pinned-arc stop trades accuracy for evaluations at a costmap cell edge: only at tolerances that ask for no more
  // (at the code default 1e-5 it ended solves with a projected gradient of 0.1 still standing)
  c.pin_alpha = p.opt_tolerance > 1e-4f ? kPinnedAlpha : 0.0f;
  c.pair_eps = 1e-10f;
  c.cells = nullptr;
  c.cells4 = nullptr;
  c.pad4 = 0; c.pitch4 = 0; c.pad_ok = 0;
  c.k_lethal = lethal_entry(p);
  c.cm_curv = kCmCurvature;
  c.sur_tol = kGuidedTolScale;
  c.polish_max = kPolishMaxIterations;
  c.alpha_warm = kAlphaWarm;
  c.guided = (p.costmap_guidance == NEOMPC_GUIDANCE_ON && p.costmap_mode == NEOMPC_COSTMAP_NEAREST) ? 1 : 0;
  c.state = nullptr;
  c.state_stride = state_stride_for(N);
  c.state_rows = 0;
}

// lanes-per-instance G and steps-per-lane S for a horizon of n steps (G*S >= n, S <= 4).
// Auto mode minimises a cost model fitted to sweeps on B200 (profiles/tiling_sweep_r1b.txt, tiling_sweep_r1c.txt):
//   time per instance  ~  c(S) * shuffles(G) * lockstep(32/G) / (32/G)
// c(S): warp instructions of one solver pass with S steps per lane (kernels exist for S <= 4; more steps per lane spill registers); shuffles(G): scans and
// reductions take ceil(log2 G) exchange steps, and a group size that is not a power of two pays ~20 % on top (scan +
// broadcast instead of butterflies, computed source lanes); lockstep: the groups of a warp wait for the slowest one,
// which costs more the more groups there are.  Measured in round 2 (guided solve, profiles/tiling_sweep_r2.txt):
// N=10 (5,2) 0.542 ms, (4,3) 0.567, (8,2) 0.625; N=20 (10,2) 1.840 ms, (5,4) 1.842, (8,3) 1.856, (16,2) 2.131;
// round 1, N=3 at throughput sizes: (1,3) 0.168 ms, (2,2) 0.182, (3,1) 0.200.
constexpr int kGroupSizes[] = {1, 2, 3, 4, 5, 6, 8, 10, 16, 32};
inline double tiling_cost(int g, int s) {
  static const double c[5] = {0.0, 0.35, 0.59, 1.0, 1.55};   // (S = 2 kernels run without spills at 120 registers since round 2)
  int lg2 = 0;
  while ((1 << lg2) < g) ++lg2;
  const bool pow2 = (g & (g - 1)) == 0;
  const int per_warp = 32 / g;
  const double shuffles = (1.0 + 0.04 * lg2) * (pow2 ? 1.0 : 1.2);
  const double lockstep = 1.0 + 0.05 * std::log2((double)per_warp);
  return c[s] * shuffles * lockstep / per_warp;
}

inline void choose_tiling(int n_steps, int lanes_override, int* G, int* S) {
  int g = lanes_override;
  if (g <= 0) {
    double best = 1e300;
    for (int cand : kGroupSizes) {
      const int s = (n_steps + cand - 1) / cand;
      if (s > 4) continue;
      const double t = tiling_cost(cand, s);
      if (t < best) { best = t; g = cand; }
    }
  } else {
    // a requested group size that would need more than 4 steps per lane: next larger size that fits
    for (int cand : kGroupSizes)
      if (cand >= g && (n_steps + cand - 1) / cand <= 4) { g = cand; break; }
  }
  *G = g;
  *S = (n_steps + g - 1) / g;
}

// tiling for tiny batches: the fewest steps per lane the supported group sizes allow (see runtime.cu: dispatch)
inline void choose_latency_tiling(int n_steps, int* G, int* S) {
  int g = 1;
  while (g < n_steps && g < 32) g <<= 1;            // powers of two: butterfly reductions are the shortest
  *G = g;
  *S = (n_steps + g - 1) / g;
}

}  // namespace neompc

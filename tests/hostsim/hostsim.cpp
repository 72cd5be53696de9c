// hostsim — TEST TOOLING ONLY.  Compiles the device solver core (neo_mpc_planner2_b200/csrc/mpc_core.cuh) for
// the host with one "lane" per instance (G = 1, S = control_steps) so that the algorithm (projection, analytic
// gradient, projected L-BFGS, optimizer() epilogue) can be checked against the oracle on machines without a
// GPU.  It is NOT part of libneompc.so, is not importable from the neo_mpc_planner2_b200 package and is not a
// fallback: the product path fails loudly without CUDA.
#include <cstdio>
#include <cstdlib>
#include <vector>

static int g_trace = 0;
#define NEOMPC_TRACE(...) do { if (g_trace) fprintf(stderr, __VA_ARGS__); } while (0)
#include "../../neo_mpc_planner2_b200/csrc/mpc_setup.h"

using namespace neompc;

namespace {

struct Env {
  SolverConst c;
  HostTables tab;
  CostTables T;
  std::vector<uint32_t> cells4;
};

void make_env(Env& e, const neompc_params* p, const uint8_t* cells, int W, int H, double res, double ox, double oy,
              int enc, const float* fp_xy, int fp_n, float* state, unsigned state_rows, float tol_pg, float tol_f) {
  build_const(*p, e.c);
  build_tables(*p, enc, e.tab);
  e.T.cost = e.tab.cost.data();
  e.T.flag = e.tab.flag.data();
  e.c.cells = cells;
  e.c.lethal_byte = enc == NEOMPC_ENC_NAV2_RAW ? 254 : 100;
  e.c.cm_scale = 1.0f / (float)e.c.lethal_byte;
  e.c.W = W; e.c.H = H;
  if (cells != nullptr) {
    e.c.pad4 = corner_pad_for(*p, res, &e.c.pad_ok);
    e.c.pitch4 = corner_pitch(W, e.c.pad4);
    build_corner_map(cells, W, H, e.c.lethal_byte, e.c.pad4, e.cells4);
    e.c.cells4 = e.cells4.data();
  }
  e.c.inv_res = (float)(1.0 / res);
  e.c.inv_res_d = 1.0 / res;
  e.c.origin_x = ox; e.c.origin_y = oy;
  e.c.fp_n = fp_n;
  for (int i = 0; i < fp_n && i < NEOMPC_MAX_FOOTPRINT_VERTICES; ++i) { e.c.fp_x[i] = fp_xy[2 * i]; e.c.fp_y[i] = fp_xy[2 * i + 1]; }
  e.c.state = state;
  e.c.state_rows = state_rows;
  if (getenv("HOSTSIM_PINALPHA")) e.c.pin_alpha = atof(getenv("HOSTSIM_PINALPHA"));
  if (getenv("HOSTSIM_CMCURV")) e.c.cm_curv = atof(getenv("HOSTSIM_CMCURV"));
  if (getenv("HOSTSIM_POLISHMAX")) e.c.polish_max = atoi(getenv("HOSTSIM_POLISHMAX"));
  if (getenv("HOSTSIM_ALPHAWARM")) e.c.alpha_warm = atof(getenv("HOSTSIM_ALPHAWARM"));
  if (getenv("HOSTSIM_SURTOL")) e.c.sur_tol = atof(getenv("HOSTSIM_SURTOL"));
  if (getenv("HOSTSIM_PAIREPS")) e.c.pair_eps = atof(getenv("HOSTSIM_PAIREPS"));
  if (getenv("HOSTSIM_TOLX")) { e.c.tol_x = atof(getenv("HOSTSIM_TOLX")); e.c.pin_alpha = 0.0f; }
  if (getenv("HOSTSIM_TOLSCALE")) { const float k = atof(getenv("HOSTSIM_TOLSCALE")); e.c.tol_pg *= k; e.c.tol_f *= k; e.c.tol_x *= k; }
  if (getenv("HOSTSIM_TOLF")) e.c.tol_f *= atof(getenv("HOSTSIM_TOLF"));
  if (getenv("HOSTSIM_TOLPG")) e.c.tol_pg *= atof(getenv("HOSTSIM_TOLPG"));
  if (tol_pg > 0) e.c.tol_pg = tol_pg;
  if (tol_f >= 0) e.c.tol_f = tol_f;
}

// The same rule as the library's dispatcher (runtime.cu: dispatch): the reference fast path (X = false) when no
// objective extension is on, the disc lies inside the box, headings stay in MUFU range and there is one history pair;
// HOSTSIM_GENERAL=1 forces the general build so that both instantiations can be compared.
bool fast_path(const Env& e) {
  if (getenv("HOSTSIM_GENERAL")) return false;
  return e.c.fp_mode == NEOMPC_FOOTPRINT_STATIC && e.c.cm_mode == NEOMPC_COSTMAP_NEAREST && e.c.disc_only &&
         e.c.fast_trig && e.c.m == 1;
}

template <int S>
void solve_all(const Env& e, const neompc_request* reqs, size_t n, neompc_response* out, float* plan) {
  std::vector<float> hist((size_t)hist_floats_per_lane<S>(e.c.m));
  const bool fast = fast_path(e);
  for (size_t i = 0; i < n; ++i) {
    float* pl = plan ? plan + i * 3 * e.c.N : nullptr;
    if (fast) solve_instance<1, S, false>(e.c, e.T, reqs[i], true, 0, hist.data(), 1, &out[i], nullptr, pl);
    else solve_instance<1, S, true>(e.c, e.T, reqs[i], true, 0, hist.data(), 1, &out[i], nullptr, pl);
  }
}

template <int S>
void eval_all(const Env& e, const neompc_request* reqs, const float* u, size_t n, float* J, float* g) {
  const bool fast = fast_path(e);
  for (size_t i = 0; i < n; ++i) {
    float* gi = g ? g + i * 3 * e.c.N : nullptr;
    if (fast) eval_instance<1, S, false>(e.c, e.T, reqs[i], true, 0, u + i * 3 * e.c.N, &J[i], gi);
    else eval_instance<1, S, true>(e.c, e.T, reqs[i], true, 0, u + i * 3 * e.c.N, &J[i], gi);
  }
}

template <int S>
struct Dispatch {
  static void solve(int N, const Env& e, const neompc_request* r, size_t n, neompc_response* o, float* plan) {
    if (N == S) solve_all<S>(e, r, n, o, plan); else Dispatch<S - 1>::solve(N, e, r, n, o, plan);
  }
  static void eval(int N, const Env& e, const neompc_request* r, const float* u, size_t n, float* J, float* g) {
    if (N == S) eval_all<S>(e, r, u, n, J, g); else Dispatch<S - 1>::eval(N, e, r, u, n, J, g);
  }
};
template <>
struct Dispatch<0> {
  static void solve(int, const Env&, const neompc_request*, size_t, neompc_response*, float*) {}
  static void eval(int, const Env&, const neompc_request*, const float*, size_t, float*, float*) {}
};

}  // namespace

extern "C" {

int hostsim_solve(const neompc_params* p, const uint8_t* cells, int W, int H, double res, double ox, double oy, int enc,
                  const float* fp_xy, int fp_n, const neompc_request* reqs, size_t n, neompc_response* out,
                  float* plan, float* state, unsigned state_rows, float tol_pg, float tol_f) {
  std::string err;
  if (!validate_params(*p, err) || p->control_steps > 32) { fprintf(stderr, "hostsim: %s\n", err.c_str()); return -1; }
  Env e;
  make_env(e, p, cells, W, H, res, ox, oy, enc, fp_xy, fp_n, state, state_rows, tol_pg, tol_f);
  Dispatch<32>::solve(p->control_steps, e, reqs, n, out, plan);
  return 0;
}

int hostsim_eval(const neompc_params* p, const uint8_t* cells, int W, int H, double res, double ox, double oy, int enc,
                 const float* fp_xy, int fp_n, const neompc_request* reqs, const float* u, size_t n, float* J, float* g) {
  std::string err;
  if (!validate_params(*p, err) || p->control_steps > 32) { fprintf(stderr, "hostsim: %s\n", err.c_str()); return -1; }
  Env e;
  make_env(e, p, cells, W, H, res, ox, oy, enc, fp_xy, fp_n, nullptr, 0, -1.f, -1.f);
  Dispatch<32>::eval(p->control_steps, e, reqs, u, n, J, g);
  return 0;
}

void hostsim_trace(int on) { g_trace = on; }

int hostsim_state_stride(int n_steps) { return state_stride_for(n_steps); }

// host logic of the dispatcher (mpc_setup.h): lane tiling for a horizon, throughput (latency = 0) or latency form
void hostsim_choose_tiling(int n_steps, int lanes_override, int latency, int* G, int* S) {
  if (latency) choose_latency_tiling(n_steps, G, S); else choose_tiling(n_steps, lanes_override, G, S);
}

// projection onto box ∩ disc, for property tests
void hostsim_project(const neompc_params* p, float* v, size_t n) {
  SolverConst c;
  build_const(*p, c);
  for (size_t i = 0; i < n; ++i) project_step<true>(c, v[3 * i], v[3 * i + 1], v[3 * i + 2]);
}

}  // extern "C"

/*
 * neompc.h — C ABI of libneompc: the batched receding-horizon MPC solve of neo_mpc_planner2 on B200.
 *
 * This is the drop-in boundary for the reference's hot path.  Today that path is reached through the
 * ROS 2 service "optimizer" (neo_srvs2/srv/Optimizer): client created at src/NeoMpcPlanner.cpp:308,
 * request built at :240-246, blocking call at :248-250, twist read at :252; server side registered at
 * neo_mpc_planner2/mpc_optimization_server.py:105 and handled by MpcOptimizationServer.optimizer
 * (:349-403).  libneompc replaces the service hop by an in-process call: plain pointers and sizes, no
 * exceptions, no ROS / torch types.  Every function returns NEOMPC_OK (0) or a negative error code;
 * neompc_last_error() gives the message.  There is NO CPU fallback: creation fails without a CUDA device.
 *
 * Reference paths below are relative to /root/reference ("srv.py" = neo_mpc_planner2/mpc_optimization_server.py,
 * "cpp" = src/NeoMpcPlanner.cpp).
 */
#ifndef NEOMPC_H_
#define NEOMPC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEOMPC_VERSION 100            /* 0.1.0 */
#define NEOMPC_MAX_CONTROL_STEPS 64
#define NEOMPC_MAX_FOOTPRINT_VERTICES 16
#define NEOMPC_STATELESS 0xFFFFFFFFu  /* neompc_request.instance_id: cold start, no per-instance state */

/* neompc_params.footprint_mode: how the footprint term of the objective (srv.py:238-244, 262-263) is evaluated.
 * STATIC reproduces the reference: `update_footprint.points` aliases `self.footprint.points` (srv.py:227) and every
 * vertex is restored right after it is written (:241-244), so the polygon tested at every step is the CURRENT footprint;
 * the term is w_footprint when that polygon touches a lethal cell and 0 otherwise, independent of the controls.
 * MOVING is what the loop at srv.py:238-244 sets out to do (it changes results, hence
 * opt-in): the robot-frame polygon is placed at each predicted pose of the costmap rollout (srv.py:234-236) and
 * rasterised there; every step whose polygon touches a lethal cell (or leaves the map) adds w_footprint / control_steps. */
enum {
  NEOMPC_FOOTPRINT_STATIC = 0,
  NEOMPC_FOOTPRINT_MOVING = 1
};

/* neompc_params.costmap_mode: how the costmap term of the objective (srv.py:246-247, 257-260) samples the grid.
 * NEAREST reproduces the reference: the cost of the cell containing the predicted position; the term is piecewise
 * constant, so neither scipy's finite differences nor the analytic gradient here see the obstacles.
 * BILINEAR (changes results, hence opt-in) interpolates between the four nearest cell
 * centres: c = bilerp(cell cost), l = bilerp(cell cost == 1.0), term = (w_costmap c^2 + (1000 - w_costmap) l^2) / N.
 * At a cell centre this equals the reference's term; in between it is smooth and its gradient enters the solver, so
 * plans bend away from inflated obstacles.  collision_check (srv.py:312-347) stays on the nearest cell in both modes. */
enum {
  NEOMPC_COSTMAP_NEAREST = 0,
  NEOMPC_COSTMAP_BILINEAR = 1
};

/* neompc_params.costmap_guidance: the reference's costmap term is piecewise constant (srv.py:246-247, 257-260), so a
 * gradient-based solver cannot see an inflation slope and stops wherever a cost step blocks its path.  With guidance ON a
 * solve first minimises J with the costmap term interpolated between cell centres (whose gradient bends the plan down the
 * inflation slope), then continues from that point on the reference's objective until it converges there.  OFF solves on
 * the reference's objective from the start (the round-1 behaviour; kept for A/B measurements). */
enum {
  NEOMPC_GUIDANCE_ON = 0,
  NEOMPC_GUIDANCE_OFF = 1
};

/* error codes */
enum {
  NEOMPC_OK = 0,
  NEOMPC_ERR_INVALID = -1,     /* bad argument / parameter */
  NEOMPC_ERR_NO_DEVICE = -2,   /* no usable CUDA device (there is no CPU fallback) */
  NEOMPC_ERR_CUDA = -3,        /* a CUDA runtime call failed */
  NEOMPC_ERR_STATE = -4,       /* instance id beyond the reserved capacity */
  NEOMPC_ERR_NCCL = -5         /* NCCL missing (it is bound at run time, libnccl.so.2) or an NCCL call failed */
};

/* costmap cell encodings (semantics declared in DESIGN.md "Costmap semantics"; the reference reads its
 * costmap through the un-vendored neo_nav2_py_costmap2D package, call sites srv.py:246-247,257,262-263,332-333,343) */
enum {
  NEOMPC_ENC_OCCUPANCY = 0,    /* nav_msgs/OccupancyGrid: 0..100 -> v/100, anything else (unknown) -> 0 */
  NEOMPC_ENC_NAV2_RAW = 1      /* nav2 costmap_raw: 0..254 -> v/254, 255 (no information) -> 0 */
};

/*
 * Solver parameters: the reference's ROS parameters (declared srv.py:49-75, read :78-103) as one POD,
 * plus a few solver knobs that have no reference counterpart (0 = library default).
 * 128 bytes.
 */
typedef struct neompc_params {
  float acc_x_limit, acc_y_limit, acc_theta_limit;                 /* srv.py:49-51, used :385-391 */
  float min_vel_x, min_vel_y, min_vel_trans, min_vel_theta;        /* srv.py:53-56 (min_vel_trans unused, as in the reference) */
  float max_vel_x, max_vel_y, max_vel_trans, max_vel_theta;        /* srv.py:58-61; bounds :127-133, disc constraint :157-158 */
  float w_trans, w_orient, w_control, w_terminal;                  /* srv.py:63-66 */
  float w_costmap, w_footprint;                                    /* srv.py:67-68 */
  float waiting_time;                                              /* srv.py:70 (initial value only; threshold is 3.0 s, :380) */
  float low_pass_gain;                                             /* srv.py:71, used :366-367 */
  float opt_tolerance;                                             /* srv.py:72 (SLSQP ftol there; see DESIGN.md for its meaning here) */
  float prediction_horizon;                                        /* srv.py:73 */
  int32_t control_steps;                                           /* srv.py:75; 1..NEOMPC_MAX_CONTROL_STEPS */
  int32_t max_iterations;      /* L-BFGS iteration cap (default 100 = SLSQP's maxiter default) */
  int32_t lbfgs_memory;        /* history pairs, 1..8 (default 1: with the block-diagonal preconditioner one pair does as well as 3 or 6;
                                  other values take the general kernel build) */
  float control_smoothing;     /* epsilon of sqrt(r^2+eps^2) used for the control-term kink (srv.py:253-254);
                                  default: 10 * opt_tolerance clamped to [1e-4, 1e-2] */
  int32_t lanes_per_instance;  /* 1,2,3,4,5,6,8,10,16 or 32 lanes of a warp cooperate on one instance (a size that would need more than
                                  4 steps per lane is raised to the next that fits); 0 = auto by control_steps */
  int32_t footprint_mode;      /* NEOMPC_FOOTPRINT_* ; 0 = the reference's behaviour */
  int32_t costmap_mode;        /* NEOMPC_COSTMAP_* ; 0 = the reference's behaviour */
  int32_t costmap_guidance;    /* NEOMPC_GUIDANCE_* ; 0 = on.  Solver strategy only: the objective, and with it every reported
                                  cost, stays the reference's (see DESIGN.md "Costmap guidance") */
  int32_t reserved[3];
} neompc_params;

/*
 * One MPC problem = one neo_srvs2/srv/Optimizer request (fields: cpp:241-246; consumed srv.py:350-355),
 * planar form, float32, 64 bytes.  Orientations are yaw angles extracted with the reference's
 * euler_from_quaternion (srv.py:160-180) by whoever marshals the request (neompc_pack_requests does it
 * on the device from full quaternions).
 */
typedef struct neompc_request {
  float vel_x, vel_y, vel_theta;          /* current_vel.linear.x / .y, .angular.z           (srv.py:216-218) */
  float carrot_x, carrot_y, carrot_yaw;   /* carrot_pose, robot base frame                    (srv.py:211,219) */
  float goal_x, goal_y, goal_yaw;         /* goal_pose, plan frame                            (srv.py:212,266) */
  float pose_x, pose_y, pose_yaw;         /* current_pose, costmap global frame; true yaw     (srv.py:315-317) */
  float pose_yaw_objective;               /* yaw the objective's costmap rollout starts from: the reference
                                             mixes goal_pose.orientation.w into it (srv.py:213).  Set it to
                                             pose_yaw to switch the quirk off. */
  float control_interval;                 /* 1 / controller_frequency                         (cpp:246, srv.py:385-391) */
  float delta_t;                          /* wall-clock time since this instance's previous call (srv.py:369-371) */
  uint32_t instance_id;                   /* row of the per-instance state (warm start, last_control, collision
                                             latch ...), or NEOMPC_STATELESS */
} neompc_request;

/* neompc_response.status: how the solver stopped */
enum {
  NEOMPC_STATUS_CONVERGED = 0,   /* projected-gradient tolerance met  -> reference "x.success" path (srv.py:397-398) */
  NEOMPC_STATUS_MAXITER = 1,     /* iteration cap                     -> reference failure path      (srv.py:399-400) */
  NEOMPC_STATUS_LINESEARCH = 2   /* no further decrease found (e.g. at a costmap cell edge); treated as converged */
};
/* neompc_response.flags */
enum {
  NEOMPC_FLAG_COLLISION = 1,            /* self.collision after this call        (srv.py:338-339,380-382) */
  NEOMPC_FLAG_COLLISION_FOOTPRINT = 2,  /* self.collision_footprint              (srv.py:343-347) */
  NEOMPC_FLAG_NEW_GOAL = 4,             /* the new-goal reset ran                (srv.py:358-361) */
  NEOMPC_FLAG_STOPPED = 8,              /* zero twist returned                   (srv.py:374-377) */
  NEOMPC_FLAG_NO_STATE = 16             /* instance_id beyond the reserved rows: solved as a cold start (see neompc_reserve_instances) */
};

/* Optimizer response (output_vel.twist.linear.x/.y, .angular.z; cpp:252, srv.py:375-377,389-391) + diagnostics. 32 bytes. */
typedef struct neompc_response {
  float vx, vy, omega;
  float cost;          /* reference objective J (srv.py:204-269) at the solver's solution, float32 */
  uint32_t iters;      /* L-BFGS iterations */
  uint32_t evals;      /* objective evaluations */
  uint32_t status;
  uint32_t flags;
} neompc_response;

/* Full-fidelity mirror of neo_srvs2/srv/Optimizer.Request in float64 with quaternions, for callers that hold
 * ROS messages (the plugin).  neompc_pack_requests converts these to neompc_request on the device. */
typedef struct neompc_optimizer_request {
  double current_vel[6];        /* Twist: linear xyz, angular xyz */
  double carrot_pose[7];        /* position xyz, orientation xyzw */
  double goal_pose[7];
  double current_pose[7];
  double control_interval;
  double delta_t;
  uint32_t instance_id;
  uint32_t switch_opt;          /* carried, unused — as in the reference (srv.py:354) */
} neompc_optimizer_request;

typedef struct neompc_handle neompc_handle;

/* ---- lifecycle ------------------------------------------------------------------------------------------- */
/* Replaces client creation + wait-for-service in NeoMpcPlanner::configure (cpp:308, :325-330) and the server's
 * constructor (srv.py:44-152).  device: CUDA ordinal (>= 0). */
int neompc_create(const neompc_params* params, int device, neompc_handle** out);
int neompc_destroy(neompc_handle* h);
/* Replaces the dynamic-parameter callback (srv.py:405-439); takes effect from the next solve. */
int neompc_set_params(neompc_handle* h, const neompc_params* params);
int neompc_get_params(const neompc_handle* h, neompc_params* out);
const char* neompc_last_error(const neompc_handle* h);   /* h may be NULL: last error of neompc_create on this thread */
int neompc_version(void);
/* sizeof() of the POD records as compiled, for binding self-checks: out[0..6] = request, response, params, optimizer_request,
 * robot_tick, carrot_info, plan_pose */
int neompc_abi_sizes(size_t out[7]);

/* ---- environment ------------------------------------------------------------------------------------------ */
/* Costmap2d(self) (srv.py:118).  cells: row-major uint8 [height][width], copied to the device (host pointer).
 * cells == NULL removes the costmap (free space everywhere: BASELINE config C1). */
int neompc_set_costmap(neompc_handle* h, const uint8_t* cells, uint32_t width, uint32_t height,
                       double resolution, double origin_x, double origin_y, int encoding);
/* Same, cells already on the device (copied device-to-device on the handle's stream). */
int neompc_set_costmap_device(neompc_handle* h, const uint8_t* d_cells, uint32_t width, uint32_t height,
                              double resolution, double origin_x, double origin_y, int encoding);
/* Robot-frame footprint polygon (x0,y0,x1,y1,...).  The reference receives the world-frame polygon from
 * /local_costmap/published_footprint (srv.py:140-144,154-155); here it is placed at each request's current pose. */
int neompc_set_footprint(neompc_handle* h, const float* xy, int n_vertices);

/* ---- per-instance state (srv.py:115-117,136,138,146-149: initial_guess, last_control, waiting_time, collision,
 *      collision_footprint, old_goal) ----------------------------------------------------------------------
 * neompc_request.instance_id names the state row of a robot; rows exist after neompc_reserve_instances(n) for ids 0..n-1.
 *  - An id beyond the reserved rows is NOT silently accepted: the request is solved as a cold start, its response carries
 *    NEOMPC_FLAG_NO_STATE, and the host-buffer entry points (neompc_solve_batch, neompc_solve_msgs) return NEOMPC_ERR_STATE
 *    after writing the responses.
 *  - Ids must be UNIQUE within one batch: two requests with the same id read and write the same row concurrently (the later
 *    writer wins, nothing is reported).  With the environment variable NEOMPC_DEBUG_IDS=1 the host-buffer entry points
 *    check this on the host (slow) and return NEOMPC_ERR_INVALID.
 * Every call makes the handle's device current for its duration and restores the caller's current device on return. */
int neompc_reserve_instances(neompc_handle* h, uint32_t n_instances);
int neompc_reset_state(neompc_handle* h, const uint32_t* ids, size_t n);     /* ids == NULL: all */
/* Test/inspection hook: copies one instance's state to the host.  initial_guess: 3*control_steps floats. */
int neompc_get_state(neompc_handle* h, uint32_t id, float* initial_guess, float last_control[3],
                     float* waiting_time, uint32_t* flags);

/* ---- the hot path: MpcOptimizationServer.optimizer (srv.py:349-403) for n requests --------------------------- */
/* Host buffers (what NeoMpcPlanner::computeVelocityCommands calls with n = 1, replacing cpp:240-252).
 * Synchronous: H2D copy, solve, D2H copy.  plan_or_null: n * 3*control_steps floats, the solver's solution
 * (the "x.x" of srv.py:363 before the low-pass), e.g. to publish local_plan (srv.py:271-310). */
int neompc_solve_batch(neompc_handle* h, const neompc_request* reqs, size_t n,
                       neompc_response* out, float* plan_or_null);
/* Same, but only the answer a controller needs comes back: twist_out = n * 3 floats (vx, vy, omega), i.e. the
 * Optimizer response's output_vel (cpp:252) without the diagnostics — 12 instead of 32 bytes per problem over PCIe.  A stopped robot (srv.py:374-377) is the zero twist. */
int neompc_solve_batch_twists(neompc_handle* h, const neompc_request* reqs, size_t n, float* twist_out);
/* Device buffers, asynchronous on `stream` (a cudaStream_t; NULL = the handle's own stream).
 * twist_or_null: n*3 floats (vx,vy,omega) packed — the payload of the multi-GPU gather. */
int neompc_solve_batch_device(neompc_handle* h, const neompc_request* d_reqs, size_t n,
                              neompc_response* d_out, float* d_twist_or_null, float* d_plan_or_null,
                              void* stream);
/* Message-level entry: float64 quaternion requests on the host -> device pack (euler_from_quaternion incl. the
 * goal-w quirk, srv.py:160-180, :211-213) -> solve -> responses on the host.  This is the per-tick call of the
 * controller plugin: small batches take a pinned mailbox path and a latency-oriented lane tiling. */
int neompc_solve_msgs(neompc_handle* h, const neompc_optimizer_request* msgs, size_t n,
                      neompc_response* out, float* plan_or_null);
/* Device-side packing only (d_msgs, d_reqs device pointers), asynchronous on `stream`. */
int neompc_pack_requests(neompc_handle* h, const neompc_optimizer_request* d_msgs, size_t n,
                         neompc_request* d_reqs, void* stream);

/* ---- multi-GPU: contiguous shards, ONE collective ------------------------------------------
 * Problems are independent (the reference solves them one at a time, srv.py:349-403), so a batch of n requests is cut
 * into n_ranks contiguous shards of neompc_shard_rows(n, n_ranks) = ceil(n / n_ranks) rows (the last one may be shorter);
 * costmap, footprint and parameters are set on every handle (replicated, <= 16 MB); per-instance state lives on the
 * handle that solves the request (a robot keeps its state as long as it keeps its position in the batch).  The only
 * exchange is an NCCL all-gather of the solved (vx, vy, omega), 12 B per problem, after which every rank holds all of
 * them.  NCCL is bound at run time (dlopen libnccl.so.2): a single-GPU user of this library does not need it.
 * One handle = one GPU = one rank.  Two ways to form the communicator: */
#define NEOMPC_COMM_ID_BYTES 128
/* (a) one process per GPU (MPI / torchrun style): rank 0 creates an id, the caller distributes it, every rank joins. */
int neompc_comm_unique_id(unsigned char id[NEOMPC_COMM_ID_BYTES]);
int neompc_comm_init(neompc_handle* h, const unsigned char id[NEOMPC_COMM_ID_BYTES], int n_ranks, int rank);
/* (b) one process driving several GPUs (a C++ fleet host): handles[i] becomes rank i. */
int neompc_comm_init_all(neompc_handle** handles, int n_handles);
int neompc_comm_destroy(neompc_handle* h);                 /* also done by neompc_destroy */
int neompc_comm_info(const neompc_handle* h, int* n_ranks, int* rank);
size_t neompc_shard_rows(size_t n_total, int n_ranks);
/* Device buffers, asynchronous: solves this rank's shard (n_local <= shard_rows requests) on `stream`, writing its twists
 * into rows [rank * shard_rows, ...) of d_twist_all ([n_ranks * shard_rows][3] floats, rows past a short shard zeroed),
 * then all-gathers d_twist_all in place on the handle's communication stream behind the solve — so the caller may already
 * enqueue the next batch's solve on `stream` (with another d_twist_all) while the gather runs.  neompc_gather_wait makes
 * `stream` wait for the most recent gather (age 0) or for the one before it (age 1: the software-pipelined loop
 * "solve k; wait for gather k-1"). */
int neompc_solve_gather_device(neompc_handle* h, const neompc_request* d_reqs, size_t n_local, size_t shard_rows,
                               neompc_response* d_out, float* d_twist_all, void* stream);
int neompc_gather_wait(neompc_handle* h, void* stream, int age);
/* Host buffers, synchronous, for communicators made with neompc_comm_init_all: shards `reqs`, copies each shard to its
 * GPU, solves, all-gathers, and returns all n twists ([n][3] floats) from rank 0's copy; out_or_null receives the n
 * responses.  neompc_fleet_get_gathered reads another rank's copy of the last gather (they are identical: test hook). */
int neompc_fleet_solve(neompc_handle** handles, int n_handles, const neompc_request* reqs, size_t n, float* twist_out,
                       neompc_response* out_or_null);
int neompc_fleet_get_gathered(neompc_handle* h, size_t n, float* twist_out);

/* ---- the step before the solve: carrot selection of the plugin ------------------- */
/* One robot's inputs for one control tick: what computeVelocityCommands receives (cpp:202-205) plus the plugin state
 * the reference keeps per instance (pruned plan position cpp:127, slow_down_ h:162).  48 bytes. */
typedef struct neompc_robot_tick {
  double pose_x, pose_y, pose_yaw;     /* robot pose in the plan / costmap global frame */
  float vel_x, vel_y, vel_theta;       /* current speed (cpp:204) */
  uint32_t plan_start;                 /* first plan pose still kept (the reference erases the ones before, cpp:127) */
  uint32_t slow_down;                  /* slow_down_ (h:162), updated by cpp:216-232 */
  float delta_t;                       /* forwarded to neompc_request.delta_t */
} neompc_robot_tick;

enum { NEOMPC_CARROT_OK = 0, NEOMPC_CARROT_EMPTY_WINDOW = 1 /* cpp:130-132 */, NEOMPC_CARROT_COLLISION = 2 /* cpp:234-236 */ };

/* Per-robot result of the carrot selection.  16 bytes. */
typedef struct neompc_carrot_info {
  uint32_t status;            /* NEOMPC_CARROT_* */
  uint32_t plan_start;        /* closest plan pose = new pruning position (cpp:81-86,127) */
  uint32_t carrot_index;      /* plan pose chosen as carrot (cpp:173-189) */
  uint32_t flags;             /* bit0 closer_to_goal (cpp:88-96), bit1 new slow_down_ (cpp:216-232), bits 8..15 footprint raw cost */
} neompc_carrot_info;

typedef struct neompc_carrot_params {
  float lookahead_dist_min, lookahead_dist_max, lookahead_dist_close_to_goal;   /* cpp:311-323 */
  float controller_frequency;                                                   /* cpp:323; control_interval = 1/f (cpp:246) */
} neompc_carrot_params;

/* setPlan (cpp:274-281): a global plan shared by all robots, poses (x, y, yaw) in the costmap's global frame; the
 * last pose is the goal_pose of every request (cpp:280, :243).  Host pointer, copied. */
int neompc_set_plan(neompc_handle* h, const double* xyyaw, size_t n_poses);
/* transformGlobalPlan + getLookAheadDistance + getLookAheadPoint + slow-down hysteresis + request construction
 * (cpp:66-135, 157-189, 216-246) for n robots: writes one neompc_request per robot (instance_id = first_instance_id + i,
 * or NEOMPC_STATELESS when first_instance_id == NEOMPC_STATELESS) and one neompc_carrot_info.  Host buffers. */
int neompc_build_requests(neompc_handle* h, const neompc_carrot_params* cp, const neompc_robot_tick* ticks, size_t n,
                          uint32_t first_instance_id, neompc_request* reqs_out, neompc_carrot_info* info_out);
/* Same on device buffers, asynchronous on `stream`; reqs_out can be passed straight to neompc_solve_batch_device. */
int neompc_build_requests_device(neompc_handle* h, const neompc_carrot_params* cp, const neompc_robot_tick* d_ticks,
                                 size_t n, uint32_t first_instance_id, neompc_request* d_reqs_out,
                                 neompc_carrot_info* d_info_out, void* stream);

/* One control tick for n robots in one call — carrot selection + request construction feeding the solve on the device,
 * one synchronise: what NeoMpcPlanner::computeVelocityCommands (cpp:202-255) does per call, without the service hop.
 * Every robot is solved, also those whose info.status is not NEOMPC_CARROT_OK (the caller decides what to do with them: the
 * reference throws ControllerException, cpp:131, :235).  reqs_out_or_null receives the requests that were solved. */
int neompc_control_tick(neompc_handle* h, const neompc_carrot_params* cp, const neompc_robot_tick* ticks, size_t n,
                        uint32_t first_instance_id, neompc_response* out, neompc_carrot_info* info_out,
                        neompc_request* reqs_out_or_null, float* plan_or_null);

/* ---- the step after the solve: the predicted path -------------------------------- */
/* One pose of the nav_msgs/Path the reference publishes on "local_plan" (srv.py:107; publishLocalPlan, srv.py:271-310):
 * position x, y and the orientation quaternion_from_euler(0, 0, yaw) (srv.py:182-196; x = y = 0).  32 bytes. */
typedef struct neompc_plan_pose {
  double x, y;
  double qz, qw;
} neompc_plan_pose;
/* publishLocalPlan(x.x) (srv.py:365) for n solved problems: plan = n * 3*control_steps floats (the plan output of
 * neompc_solve_batch*), poses_out = n * (control_steps + 1) poses; pose 0 is the start pose (position only, identity
 * orientation, srv.py:288-291).  The reference rolls from the TF pose map -> base_link (srv.py:274-286); here the
 * request's current_pose is used (same thing when the costmap's global frame is "map").  Host buffers, synchronous. */
int neompc_local_plan(neompc_handle* h, const neompc_request* reqs, const float* plan, size_t n,
                      neompc_plan_pose* poses_out);
/* Same on device buffers, asynchronous on `stream` (NULL = the handle's stream). */
int neompc_local_plan_device(neompc_handle* h, const neompc_request* d_reqs, const float* d_plan, size_t n,
                             neompc_plan_pose* d_poses_out, void* stream);

/* ---- test hooks ------------------------------------------------------------------------------------------- */
/* objective(cmd_vel) (srv.py:204-269) and its analytic gradient for n (request, u) pairs; host buffers.
 * u: n * 3*control_steps.  J: n.  grad_or_null: n * 3*control_steps (gradient of the smoothed objective the
 * solver uses; costmap/footprint terms are piecewise constant and contribute 0). */
int neompc_eval_objective(neompc_handle* h, const neompc_request* reqs, const float* u, size_t n,
                          float* J, float* grad_or_null);
/* Number of kernels this handle has launched so far. */
uint64_t neompc_launch_count(const neompc_handle* h);
/* How the last neompc_solve_batch / neompc_solve_batch_twists call moved its host buffers (measurement aid):
 * MAILBOX  small batches through the handle's pinned mailbox;
 * ZERO_COPY  the caller's buffers are page-locked and device-accessible (neompc_host_alloc, cudaHostAlloc,
 *            cudaHostRegister): the solve kernel reads the requests and writes the results over PCIe itself, one launch;
 * CHUNKED  pageable buffers: staged copies pipelined with the solve in chunks on two streams. */
enum { NEOMPC_HOST_PATH_NONE = 0, NEOMPC_HOST_PATH_MAILBOX = 1, NEOMPC_HOST_PATH_CHUNKED = 2, NEOMPC_HOST_PATH_ZERO_COPY = 3 };
int neompc_last_host_path(const neompc_handle* h);
/* Lanes-per-instance / steps-per-lane the dispatcher uses for the current parameters. */
int neompc_get_tiling(const neompc_handle* h, int* lanes_per_instance, int* steps_per_lane);
/* ... and for a batch of n requests: batches small enough to be resident at once take the latency tiling (more lanes per
 * instance, fewer steps per lane), larger ones the throughput tiling neompc_get_tiling reports. */
int neompc_get_tiling_for(const neompc_handle* h, size_t n, int* lanes_per_instance, int* steps_per_lane);

/* pinned host memory for the host-buffer entry points (optional; any host memory works, pinned memory takes the
 * zero-copy path of neompc_last_host_path) */
int neompc_host_alloc(void** ptr, size_t bytes);
int neompc_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* NEOMPC_H_ */

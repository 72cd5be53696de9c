/* A plain C99 caller of libneompc, the way a cgo / FFI binding would use it: only include/neompc.h, POD records,
 * int status codes.  Without a CUDA device neompc_create must fail with NEOMPC_ERR_NO_DEVICE and a message (there is
 * no CPU fallback); with one, a single cold-start request is solved.  Built and run by tests/test_abi.py. */
#include <stdio.h>
#include <string.h>

#include "neompc.h"

int main(void) {
  size_t sz[7];
  neompc_params p;
  neompc_handle* h = NULL;
  int rc;
  if (neompc_version() != NEOMPC_VERSION) { printf("version mismatch\n"); return 1; }
  if (neompc_abi_sizes(sz) != NEOMPC_OK || sz[0] != sizeof(neompc_request) || sz[1] != sizeof(neompc_response) ||
      sz[2] != sizeof(neompc_params) || sz[6] != sizeof(neompc_plan_pose)) { printf("record sizes differ\n"); return 1; }
  memset(&p, 0, sizeof p);
  p.acc_x_limit = 2.5f; p.acc_y_limit = 2.5f; p.acc_theta_limit = 3.0f;
  p.min_vel_x = -0.7f; p.min_vel_y = -0.7f; p.min_vel_trans = -0.7f; p.min_vel_theta = -0.7f;
  p.max_vel_x = 0.7f; p.max_vel_y = 0.7f; p.max_vel_trans = 0.7f; p.max_vel_theta = 0.7f;
  p.w_trans = 0.82f; p.w_orient = 0.5f; p.w_control = 0.05f; p.w_terminal = 0.05f; p.w_costmap = 0.05f;
  p.waiting_time = 3.0f; p.low_pass_gain = 0.5f; p.opt_tolerance = 1e-3f; p.prediction_horizon = 0.8f;
  p.control_steps = 3;
  rc = neompc_create(&p, 0, &h);
  if (rc == NEOMPC_ERR_NO_DEVICE) {
    if (h != NULL || strlen(neompc_last_error(NULL)) == 0) { printf("bad failure contract\n"); return 1; }
    printf("abi ok (no CUDA device: %s)\n", neompc_last_error(NULL));
    return 0;
  }
  if (rc != NEOMPC_OK) { printf("neompc_create: %d %s\n", rc, neompc_last_error(NULL)); return 1; }
  {
    neompc_request rq;
    neompc_response rs;
    float plan[9];
    neompc_plan_pose poses[4];
    memset(&rq, 0, sizeof rq);
    rq.carrot_x = 0.4f; rq.carrot_y = 0.1f; rq.carrot_yaw = 0.3f;            /* the known-answer problem of SURVEY 8c */
    rq.goal_x = 3.0f; rq.goal_y = 1.0f; rq.goal_yaw = 0.5f;
    rq.pose_x = 1.0f; rq.pose_y = 2.0f; rq.pose_yaw = 0.2f; rq.pose_yaw_objective = 0.2f;
    rq.control_interval = 1.0f / 30.0f; rq.delta_t = 1.0f / 30.0f; rq.instance_id = NEOMPC_STATELESS;
    if (neompc_solve_batch(h, &rq, 1, &rs, plan) != NEOMPC_OK) { printf("solve: %s\n", neompc_last_error(h)); return 1; }
    if (neompc_local_plan(h, &rq, plan, 1, poses) != NEOMPC_OK) { printf("local_plan: %s\n", neompc_last_error(h)); return 1; }
    /* first call: accel clamp from rest, 2.5/30 and 3.0/30 (srv.py:385-391) */
    if (rs.vx < 0.0832f || rs.vx > 0.0835f || rs.omega < 0.0999f || rs.omega > 0.1001f) {
      printf("unexpected twist %f %f %f\n", rs.vx, rs.vy, rs.omega);
      return 1;
    }
    printf("abi ok (twist %.4f %.4f %.4f, cost %.5f, %u iterations; path ends at %.3f %.3f)\n", rs.vx, rs.vy, rs.omega,
           rs.cost, rs.iters, poses[3].x, poses[3].y);
  }
  return neompc_destroy(h) == NEOMPC_OK ? 0 : 1;
}

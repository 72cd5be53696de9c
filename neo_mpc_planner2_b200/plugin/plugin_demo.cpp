// plugin_demo — drives the plugin through configure / setPlan / computeVelocityCommands on a synthetic scene.
//   plugin_demo                 closed loop of 12 control ticks next to an inflated obstacle; one JSON object per run with,
//                               per tick, the twist, the Optimizer request the plugin built (cpp:240-246), the carrot
//                               bookkeeping, the raw plan and the solver's diagnostics (tests/test_plugin.py replays it
//                               against the oracle's carrot selection and optimizer() state machine on the GPU box)
//   plugin_demo latency W H     full-tick latency of computeVelocityCommands on a W x H costmap, once with the costmap
//                               unchanged between ticks (no upload) and once with a new costmap every tick
//   plugin_demo frames          a plan in another frame than the controller's pose: the reference's exception (cpp:76)
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "NeoMpcPlanner.h"

extern "C" nav2_core::Controller * neompc_plugin_create();

namespace {

std::shared_ptr<rclcpp_lifecycle::LifecycleNode> make_node(int control_steps) {
  auto node = std::make_shared<rclcpp_lifecycle::LifecycleNode>();
  node->params = {{"FollowPath.lookahead_dist_min", 0.3}, {"FollowPath.lookahead_dist_max", 0.45},
                  {"FollowPath.lookahead_dist_close_to_goal", 0.25}, {"controller_frequency", 30.0},
                  {"FollowPath.acc_x_limit", 2.5}, {"FollowPath.acc_y_limit", 2.5}, {"FollowPath.acc_theta_limit", 3.0},
                  {"FollowPath.min_vel_x", -0.7}, {"FollowPath.min_vel_y", -0.7}, {"FollowPath.min_vel_theta", -0.7},
                  {"FollowPath.max_vel_x", 0.7}, {"FollowPath.max_vel_y", 0.7}, {"FollowPath.max_vel_trans", 0.7},
                  {"FollowPath.max_vel_theta", 0.7}, {"FollowPath.w_trans", 0.82}, {"FollowPath.w_orient", 0.5},
                  {"FollowPath.w_control", 0.05}, {"FollowPath.w_terminal", 0.05}, {"FollowPath.w_footprint", 0.0},
                  {"FollowPath.w_costmap", 0.05}, {"FollowPath.opt_tolerance", 1e-3},
                  {"FollowPath.prediction_horizon", 0.8}, {"FollowPath.control_steps", (double)control_steps}};
  return node;
}

// nav2 raw costs, integer arithmetic only (tests/test_plugin.py builds the same grid): a lethal box with an inscribed ring
// and a linear inflation ramp around it (Chebyshev distance d to the box: d <= 2 -> 253, else 252 - 18 (d - 2) down to 0)
void paint_scene(nav2_costmap_2d::Costmap2D & grid, int bx0, int bx1, int by0, int by1, int shift) {
  const int W = (int)grid.getSizeInCellsX(), H = (int)grid.getSizeInCellsY();
  unsigned char * c = grid.getCharMap();
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int dx = x < bx0 + shift ? bx0 + shift - x : (x >= bx1 + shift ? x - (bx1 + shift) + 1 : 0);
      const int dy = y < by0 ? by0 - y : (y >= by1 ? y - by1 + 1 : 0);
      const int d = std::max(dx, dy);
      int v = d == 0 ? 254 : (d <= 2 ? 253 : 252 - 18 * (d - 2));
      c[(size_t)y * W + x] = (unsigned char)std::max(v, 0);
    }
}

std::vector<geometry_msgs::msg::Point> footprint() {
  std::vector<geometry_msgs::msg::Point> fp(4);
  fp[0].x = 0.4; fp[0].y = 0.3; fp[1].x = -0.4; fp[1].y = 0.3; fp[2].x = -0.4; fp[2].y = -0.3; fp[3].x = 0.4; fp[3].y = -0.3;
  return fp;
}

nav_msgs::msg::Path make_plan(int n, const std::string & frame) {
  nav_msgs::msg::Path plan;
  plan.header.frame_id = frame;
  for (int i = 0; i <= n; ++i) {                       // a gentle left curve; the pose yaw follows the tangent, then overshoots
    geometry_msgs::msg::PoseStamped ps;
    const double t = 0.05 * i;
    ps.pose.position.x = -1.0 + t;
    ps.pose.position.y = 0.15 * t * t;
    const double yaw = std::atan2(0.3 * t, 1.0) + (i > 30 ? 1.2 : 0.0);   // |carrot yaw| >= 1 later on: exercises slow_down_
    ps.pose.orientation.z = std::sin(0.5 * yaw);
    ps.pose.orientation.w = std::cos(0.5 * yaw);
    plan.poses.push_back(ps);
  }
  return plan;
}

int run_ticks() {
  auto node = make_node(3);
  auto grid = std::make_shared<nav2_costmap_2d::Costmap2D>(200, 200, 0.05, -5.0, -5.0);
  paint_scene(*grid, 92, 100, 112, 120, 0);            // box around (-0.2, 0.8): its ramp reaches the footprint on the way
  auto costmap = std::make_shared<nav2_costmap_2d::Costmap2DROS>(grid, footprint());
  std::unique_ptr<nav2_core::Controller> ctrl(neompc_plugin_create());
  auto * mpc = dynamic_cast<neo_mpc_planner::NeoMpcPlanner *>(ctrl.get());
  ctrl->configure(node, "FollowPath", std::make_shared<tf2_ros::Buffer>(), costmap);
  ctrl->activate();
  ctrl->setPlan(make_plan(60, costmap->getGlobalFrameID()));
  geometry_msgs::msg::PoseStamped pose;
  pose.header.frame_id = costmap->getGlobalFrameID();
  pose.pose.position.x = -1.02; pose.pose.position.y = 0.03;
  double yaw = 0.1;
  geometry_msgs::msg::Twist vel;
  std::printf("{\"ticks\": [");
  for (int k = 0; k < 12; ++k) {
    pose.pose.orientation.z = std::sin(0.5 * yaw);
    pose.pose.orientation.w = std::cos(0.5 * yaw);
    auto cmd = ctrl->computeVelocityCommands(pose, vel, nullptr);
    const auto & rq = mpc->lastRequest();
    const auto & info = mpc->lastCarrotInfo();
    const auto & rs = mpc->lastResponse();
    std::printf("%s{\"twist\": [%.9g, %.9g, %.9g], \"pose\": [%.17g, %.17g, %.17g], \"vel\": [%.9g, %.9g, %.9g], "
                "\"info\": [%u, %u, %u, %u], \"uploaded\": %d, "
                "\"request\": {\"vel_x\": %.9g, \"vel_y\": %.9g, \"vel_theta\": %.9g, \"carrot_x\": %.9g, \"carrot_y\": %.9g, "
                "\"carrot_yaw\": %.9g, \"goal_x\": %.9g, \"goal_y\": %.9g, \"goal_yaw\": %.9g, \"pose_x\": %.9g, \"pose_y\": %.9g, "
                "\"pose_yaw\": %.9g, \"pose_yaw_objective\": %.9g, \"control_interval\": %.9g, \"delta_t\": %.9g, "
                "\"instance_id\": %u}, \"response\": {\"cost\": %.9g, \"iters\": %u, \"status\": %u, \"flags\": %u}, \"plan\": [",
                k ? ", " : "", cmd.twist.linear.x, cmd.twist.linear.y, cmd.twist.angular.z, pose.pose.position.x,
                pose.pose.position.y, yaw, vel.linear.x, vel.linear.y, vel.angular.z, info.status, info.plan_start,
                info.carrot_index, info.flags, mpc->lastTickUploadedCostmap() ? 1 : 0, rq.vel_x, rq.vel_y, rq.vel_theta,
                rq.carrot_x, rq.carrot_y, rq.carrot_yaw, rq.goal_x, rq.goal_y, rq.goal_yaw, rq.pose_x, rq.pose_y, rq.pose_yaw,
                rq.pose_yaw_objective, rq.control_interval, rq.delta_t, rq.instance_id, rs.cost, rs.iters, rs.status, rs.flags);
    for (size_t i = 0; i < mpc->lastPlan().size(); ++i) std::printf("%s%.9g", i ? ", " : "", mpc->lastPlan()[i]);
    std::printf("]}");
    // the robot moves with the commanded twist for a (long) control period, so that the plan gets pruned along the way
    const double dt = 0.25;
    vel = cmd.twist;
    pose.pose.position.x += (cmd.twist.linear.x * std::cos(yaw) - cmd.twist.linear.y * std::sin(yaw)) * dt;
    pose.pose.position.y += (cmd.twist.linear.x * std::sin(yaw) + cmd.twist.linear.y * std::cos(yaw)) * dt;
    yaw += cmd.twist.angular.z * dt;
    if (k == 7) paint_scene(*grid, 92, 100, 112, 120, 2);      // the obstacle moves: the next tick must upload again
  }
  const auto & lp = mpc->lastLocalPlan();
  std::printf("], \"local_plan\": [");
  for (size_t i = 0; i < lp.size(); ++i) std::printf("%s[%.17g, %.17g, %.17g, %.17g]", i ? ", " : "", lp[i].x, lp[i].y, lp[i].qz, lp[i].qw);
  std::printf("]}\n");
  ctrl->cleanup();
  return 0;
}

int run_latency(int W, int H) {
  auto node = make_node(10);
  auto grid = std::make_shared<nav2_costmap_2d::Costmap2D>(W, H, 0.05, -0.025 * W, -0.025 * H);
  paint_scene(*grid, W / 2 + 12, W / 2 + 20, H / 2 + 12, H / 2 + 20, 0);
  auto costmap = std::make_shared<nav2_costmap_2d::Costmap2DROS>(grid, footprint());
  std::unique_ptr<nav2_core::Controller> ctrl(neompc_plugin_create());
  ctrl->configure(node, "FollowPath", std::make_shared<tf2_ros::Buffer>(), costmap);
  ctrl->setPlan(make_plan(60, costmap->getGlobalFrameID()));
  geometry_msgs::msg::PoseStamped pose;
  pose.pose.position.x = -1.0;
  geometry_msgs::msg::Twist vel;
  auto run = [&](bool repaint, std::vector<double> & us) {
    for (int k = 0; k < 220; ++k) {
      if (repaint) paint_scene(*grid, W / 2 + 12, W / 2 + 20, H / 2 + 12, H / 2 + 20, k % 5);
      pose.pose.position.x = -1.0 + 0.002 * (k % 50);
      const auto t0 = std::chrono::steady_clock::now();
      auto cmd = ctrl->computeVelocityCommands(pose, vel, nullptr);
      const auto t1 = std::chrono::steady_clock::now();
      (void)cmd;
      if (k >= 20) us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
    }
    std::sort(us.begin(), us.end());
  };
  std::vector<double> same, fresh;
  run(false, same);
  run(true, fresh);
  auto pct = [](const std::vector<double> & v, double p) { return v[(size_t)(p * (v.size() - 1))]; };
  std::printf("{\"costmap\": [%d, %d], \"control_steps\": 10, \"ticks\": %zu, "
              "\"costmap_unchanged_us\": {\"median\": %.1f, \"p99\": %.1f}, "
              "\"costmap_new_every_tick_us\": {\"median\": %.1f, \"p99\": %.1f}}\n",
              W, H, same.size(), pct(same, 0.5), pct(same, 0.99), pct(fresh, 0.5), pct(fresh, 0.99));
  ctrl->cleanup();
  return 0;
}

int run_frames() {
  auto node = make_node(3);
  auto grid = std::make_shared<nav2_costmap_2d::Costmap2D>(60, 60, 0.05, -1.5, -1.5);
  auto costmap = std::make_shared<nav2_costmap_2d::Costmap2DROS>(grid, footprint());
  std::unique_ptr<nav2_core::Controller> ctrl(neompc_plugin_create());
  ctrl->configure(node, "FollowPath", std::make_shared<tf2_ros::Buffer>(), costmap);
  ctrl->setPlan(make_plan(20, "map"));                 // the controller's pose is in "odom"
  geometry_msgs::msg::PoseStamped pose;
  pose.header.frame_id = "odom";
  geometry_msgs::msg::Twist vel;
  ctrl->computeVelocityCommands(pose, vel, nullptr);   // must throw
  return 0;
}

}  // namespace

int main(int argc, char ** argv) {
  try {
    if (argc >= 4 && std::strcmp(argv[1], "latency") == 0) return run_latency(std::atoi(argv[2]), std::atoi(argv[3]));
    if (argc >= 2 && std::strcmp(argv[1], "frames") == 0) return run_frames();
    return run_ticks();
  } catch (const nav2_core::ControllerException & e) {
    std::fprintf(stderr, "ControllerException: %s\n", e.what());
    return 2;
  }
}

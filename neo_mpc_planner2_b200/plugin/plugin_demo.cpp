// plugin_demo — drives the plugin through configure / setPlan / computeVelocityCommands on a small synthetic scene
// and prints the twist of each control tick as JSON (used by tests/test_plugin.py on the GPU box).
#include <cmath>
#include <cstdio>

#include "NeoMpcPlanner.h"

extern "C" nav2_core::Controller * neompc_plugin_create();

int main() {
  auto node = std::make_shared<rclcpp_lifecycle::LifecycleNode>();
  node->params = {{"FollowPath.lookahead_dist_min", 0.4}, {"FollowPath.lookahead_dist_max", 0.4},
                  {"FollowPath.lookahead_dist_close_to_goal", 0.4}, {"controller_frequency", 30.0},
                  {"FollowPath.acc_x_limit", 2.5}, {"FollowPath.acc_y_limit", 2.5}, {"FollowPath.acc_theta_limit", 3.0},
                  {"FollowPath.min_vel_x", -0.7}, {"FollowPath.min_vel_y", -0.7}, {"FollowPath.min_vel_theta", -0.7},
                  {"FollowPath.max_vel_x", 0.7}, {"FollowPath.max_vel_y", 0.7}, {"FollowPath.max_vel_trans", 0.7},
                  {"FollowPath.max_vel_theta", 0.7}, {"FollowPath.w_trans", 0.82}, {"FollowPath.w_orient", 0.5},
                  {"FollowPath.w_control", 0.05}, {"FollowPath.w_terminal", 0.05}, {"FollowPath.w_footprint", 0.0},
                  {"FollowPath.w_costmap", 0.05}, {"FollowPath.opt_tolerance", 1e-3},
                  {"FollowPath.prediction_horizon", 0.8}, {"FollowPath.control_steps", 3}};
  auto grid = std::make_shared<nav2_costmap_2d::Costmap2D>(200, 200, 0.05, -5.0, -5.0);
  std::vector<geometry_msgs::msg::Point> fp(4);
  fp[0].x = 0.4; fp[0].y = 0.3; fp[1].x = -0.4; fp[1].y = 0.3; fp[2].x = -0.4; fp[2].y = -0.3; fp[3].x = 0.4; fp[3].y = -0.3;
  auto costmap = std::make_shared<nav2_costmap_2d::Costmap2DROS>(grid, fp);

  std::unique_ptr<nav2_core::Controller> ctrl(neompc_plugin_create());
  try {
    ctrl->configure(node, "FollowPath", std::make_shared<tf2_ros::Buffer>(), costmap);
    ctrl->activate();
    nav_msgs::msg::Path plan;
    for (int i = 0; i <= 40; ++i) {
      geometry_msgs::msg::PoseStamped ps;
      ps.pose.position.x = 0.1 * i;
      ps.pose.position.y = 0.02 * i;
      ps.pose.orientation.z = std::sin(0.1);
      ps.pose.orientation.w = std::cos(0.1);
      plan.poses.push_back(ps);
    }
    ctrl->setPlan(plan);
    geometry_msgs::msg::PoseStamped pose;
    geometry_msgs::msg::Twist vel;
    std::printf("[");
    for (int k = 0; k < 5; ++k) {
      auto cmd = ctrl->computeVelocityCommands(pose, vel, nullptr);
      std::printf("%s[%.9g, %.9g, %.9g]", k ? ", " : "", cmd.twist.linear.x, cmd.twist.linear.y, cmd.twist.angular.z);
      vel = cmd.twist;
      pose.pose.position.x += cmd.twist.linear.x / 30.0;
      pose.pose.position.y += cmd.twist.linear.y / 30.0;
    }
    std::printf("]\n");
    ctrl->cleanup();
  } catch (const nav2_core::ControllerException & e) {
    std::fprintf(stderr, "ControllerException: %s\n", e.what());
    return 2;
  }
  return 0;
}

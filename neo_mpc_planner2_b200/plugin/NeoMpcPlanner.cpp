// NeoMpcPlanner.cpp — the reference's controller plugin (src/NeoMpcPlanner.cpp) with the ROS service hop to the Python
// optimisation server (cpp:240-252) replaced by in-process libneompc calls.  One control tick is ONE library call,
// neompc_control_tick: the plan-following front half — closest plan pose, pruning, costmap window, lookahead distance under
// slow_down_, lookahead point, slow-down hysteresis, the footprint_cost == 255 test (cpp:66-135, 157-189, 216-236) — and the
// request construction (cpp:240-246) run on the device and feed the solve (srv.py:349-403) without leaving it.  The plugin
// keeps what the reference keeps per instance (pruned plan position cpp:127, slow_down_ h:162, goal_pose cpp:280) and throws
// the reference's ControllerExceptions (cpp:70, :76, :131, :235).  See INTEGRATION.md.
// TF: the library works in ONE frame.  A plan whose frame_id differs from the frame of the controller's pose is refused
// with the reference's "Unable to transform ..." exception (cpp:76) instead of being followed in the wrong frame.
#include "NeoMpcPlanner.h"

#include <cmath>
#include <cstring>
#include <limits>

#include "pluginlib/class_list_macros.hpp"

namespace neo_mpc_planner {

namespace {
double yawOf(const geometry_msgs::msg::Quaternion & q) {
  return std::atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));
}
}  // namespace

NeoMpcPlanner::~NeoMpcPlanner() { cleanup(); }

void NeoMpcPlanner::configure(const rclcpp_lifecycle::LifecycleNode::WeakPtr & parent, std::string name,
                              std::shared_ptr<tf2_ros::Buffer> tf,
                              std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros) {
  node_ = parent;
  auto node = node_.lock();
  if (!node) throw nav2_core::ControllerException("Unable to lock node!");          // reference cpp:298-300
  costmap_ros_ = costmap_ros;
  tf_ = tf;
  plugin_name_ = name;
  logger_ = node->get_logger();
  clock_ = node->get_clock();

  node->get_parameter_or(name + ".lookahead_dist_min", lookahead_dist_min_, 0.5);   // reference cpp:311-323
  node->get_parameter_or(name + ".lookahead_dist_max", lookahead_dist_max_, 0.5);
  node->get_parameter_or(name + ".lookahead_dist_close_to_goal", lookahead_dist_close_to_goal_, 0.5);
  node->get_parameter_or(std::string("controller_frequency"), control_frequency_, 20.0);

  // the optimisation server's parameters (reference mpc_optimization_server.py:49-75, same names and defaults),
  // now owned by the plugin
  neompc_params & p = params_;
  auto f = [&](const char * key, float & out, float dflt) { node->get_parameter_or(name + "." + key, out, dflt); };
  f("acc_x_limit", p.acc_x_limit, 0.5f); f("acc_y_limit", p.acc_y_limit, 0.5f); f("acc_theta_limit", p.acc_theta_limit, 0.5f);
  f("min_vel_x", p.min_vel_x, -0.5f); f("min_vel_y", p.min_vel_y, -0.5f); f("min_vel_trans", p.min_vel_trans, 0.5f);
  f("min_vel_theta", p.min_vel_theta, -0.5f);
  f("max_vel_x", p.max_vel_x, 0.5f); f("max_vel_y", p.max_vel_y, 0.5f); f("max_vel_trans", p.max_vel_trans, 0.5f);
  f("max_vel_theta", p.max_vel_theta, 0.5f);
  f("w_trans", p.w_trans, 0.5f); f("w_orient", p.w_orient, 0.5f); f("w_control", p.w_control, 0.5f);
  f("w_terminal", p.w_terminal, 0.5f); f("w_costmap", p.w_costmap, 0.5f); f("w_footprint", p.w_footprint, 2000.0f);
  f("waiting_time", p.waiting_time, 3.0f); f("low_pass_gain", p.low_pass_gain, 0.5f);
  f("opt_tolerance", p.opt_tolerance, 1e-5f); f("prediction_horizon", p.prediction_horizon, 0.5f);
  int steps = 3;
  node->get_parameter_or(name + ".control_steps", steps, 3);
  p.control_steps = steps;
  int device = 0;
  node->get_parameter_or(name + ".cuda_device", device, 0);

  // replaces create_client + the wait-for-service loop (reference cpp:308, :325-330)
  if (neompc_create(&p, device, &mpc_) != NEOMPC_OK)
    throw nav2_core::ControllerException(std::string("neompc_create failed: ") + neompc_last_error(nullptr));
  if (neompc_reserve_instances(mpc_, 1) != NEOMPC_OK)
    throw nav2_core::ControllerException(std::string("neompc_reserve_instances failed: ") + neompc_last_error(mpc_));
  std::vector<float> xy;
  for (const auto & pt : costmap_ros_->getRobotFootprint()) { xy.push_back((float)pt.x); xy.push_back((float)pt.y); }
  if (neompc_set_footprint(mpc_, xy.data(), (int)(xy.size() / 2)) != NEOMPC_OK)
    throw nav2_core::ControllerException(std::string("neompc_set_footprint failed: ") + neompc_last_error(mpc_));
  last_plan_.assign(3 * (size_t)p.control_steps, 0.0f);
  RCLCPP_INFO(logger_, "neompc ready: control_steps=%d on CUDA device %d", p.control_steps, device);
}

void NeoMpcPlanner::cleanup() {
  if (mpc_) { neompc_destroy(mpc_); mpc_ = nullptr; }
}
void NeoMpcPlanner::activate() {}
void NeoMpcPlanner::deactivate() {}

void NeoMpcPlanner::setPlan(const nav_msgs::msg::Path & plan) {                      // reference cpp:274-281
  std::lock_guard<std::mutex> lock(mutex_);
  global_plan_ = plan;
  plan_start_ = 0;                                    // a new plan is un-pruned (the reference erases from its own copy, cpp:127)
  if (plan.poses.empty()) return;
  if (goal_pose_ != plan.poses.back().pose) slow_down_ = true;                       // cpp:277-279
  goal_pose_ = plan.poses.back().pose;
  if (!mpc_) return;
  std::vector<double> xyyaw;
  xyyaw.reserve(3 * plan.poses.size());
  for (const auto & ps : plan.poses) {
    xyyaw.push_back(ps.pose.position.x);
    xyyaw.push_back(ps.pose.position.y);
    xyyaw.push_back(yawOf(ps.pose.orientation));
  }
  if (neompc_set_plan(mpc_, xyyaw.data(), plan.poses.size()) != NEOMPC_OK)
    throw nav2_core::ControllerException(std::string("neompc_set_plan failed: ") + neompc_last_error(mpc_));
}

void NeoMpcPlanner::setSpeedLimit(const double &, const bool &) {}                   // empty in the reference too (cpp:283-288)

// The costmap goes to the device only when it changed: geometry compare + a 64-bit checksum of the cells, taken under the
// costmap's own mutex (controller_server updates the local costmap on another thread).  Returns true if it uploaded.
bool NeoMpcPlanner::uploadCostmapIfChanged() {
  auto * cm = costmap_ros_->getCostmap();
  std::unique_lock<nav2_costmap_2d::Costmap2D::mutex_t> lock(*cm->getMutex());
  const unsigned w = cm->getSizeInCellsX(), h = cm->getSizeInCellsY();
  const unsigned char * cells = cm->getCharMap();
  const size_t bytes = (size_t)w * h;
  // eight independent lanes, rotate-xor-add mixing (no multiply on the critical path): memory-bound on a 1 MB costmap
  uint64_t acc[8] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull, 0x27D4EB2F165667C5ull,
                     0x85EBCA77C2B2AE63ull, 0xD6E8FEB86659FD93ull, 0xA0761D6478BD642Full, 0xE7037ED1A0B428DBull};
  size_t i = 0;
  for (; i + 64 <= bytes; i += 64) {
    uint64_t v[8];
    std::memcpy(v, cells + i, 64);
    for (int k = 0; k < 8; ++k) { acc[k] ^= v[k]; acc[k] = ((acc[k] << 23) | (acc[k] >> 41)) + (acc[k] >> 17) + 0x9E3779B97F4A7C15ull; }
  }
  for (; i < bytes; ++i) acc[i & 7] = (((acc[i & 7] ^ cells[i]) << 23) | ((acc[i & 7] ^ cells[i]) >> 41)) + cells[i];
  uint64_t sum = 0;
  for (int k = 0; k < 8; ++k) sum = ((sum << 7) | (sum >> 57)) ^ (acc[k] * 0x100000001B3ull);
  const bool same = costmap_loaded_ && sum == costmap_sum_ && w == costmap_w_ && h == costmap_h_ &&
                    cm->getResolution() == costmap_res_ && cm->getOriginX() == costmap_ox_ && cm->getOriginY() == costmap_oy_;
  if (same) return false;
  if (neompc_set_costmap(mpc_, cells, w, h, cm->getResolution(), cm->getOriginX(), cm->getOriginY(), NEOMPC_ENC_NAV2_RAW) !=
      NEOMPC_OK)
    throw nav2_core::ControllerException(std::string("neompc_set_costmap failed: ") + neompc_last_error(mpc_));
  costmap_loaded_ = true; costmap_sum_ = sum; costmap_w_ = w; costmap_h_ = h;
  costmap_res_ = cm->getResolution(); costmap_ox_ = cm->getOriginX(); costmap_oy_ = cm->getOriginY();
  return true;
}

geometry_msgs::msg::TwistStamped NeoMpcPlanner::computeVelocityCommands(
    const geometry_msgs::msg::PoseStamped & position, const geometry_msgs::msg::Twist & speed,
    nav2_core::GoalChecker *) {
  std::lock_guard<std::mutex> lock(mutex_);                                           // reference cpp:207
  if (global_plan_.poses.empty()) throw nav2_core::ControllerException("Received plan with zero length");   // cpp:69-71
  if (!mpc_) throw nav2_core::ControllerException("controller not configured");
  // transformPose (cpp:73-77): the robot pose must be expressible in the plan's frame.  Same frame -> nothing to do;
  // anything else would need TF, which this port does not carry: refuse instead of following the plan in the wrong frame.
  const std::string & plan_frame = global_plan_.header.frame_id;
  const std::string pose_frame = position.header.frame_id.empty() ? costmap_ros_->getGlobalFrameID() : position.header.frame_id;
  if (!plan_frame.empty() && plan_frame != pose_frame)
    throw nav2_core::ControllerException("Unable to transform robot pose into global plan's frame");

  last_upload_ = uploadCostmapIfChanged();

  // one tick: front half + request + solve (cpp:209-252)
  neompc_robot_tick tick{};
  tick.pose_x = position.pose.position.x;
  tick.pose_y = position.pose.position.y;
  tick.pose_yaw = yawOf(position.pose.orientation);
  tick.vel_x = (float)speed.linear.x; tick.vel_y = (float)speed.linear.y; tick.vel_theta = (float)speed.angular.z;
  tick.plan_start = plan_start_;
  tick.slow_down = slow_down_ ? 1u : 0u;
  const double now = clock_ ? clock_->now().seconds() : 0.0;
  tick.delta_t = (float)std::fmin(now - last_call_time_, 3.0e38);                     // srv.py:369-371
  last_call_time_ = now;
  neompc_carrot_params cp{};
  cp.lookahead_dist_min = (float)lookahead_dist_min_;
  cp.lookahead_dist_max = (float)lookahead_dist_max_;
  cp.lookahead_dist_close_to_goal = (float)lookahead_dist_close_to_goal_;
  cp.controller_frequency = (float)control_frequency_;
  const int rc = neompc_control_tick(mpc_, &cp, &tick, 1, 0u, &last_response_, &last_info_, &last_request_, last_plan_.data());
  if (rc != NEOMPC_OK)
    throw nav2_core::ControllerException(std::string("neompc_control_tick failed: ") + neompc_last_error(mpc_));
  local_plan_valid_ = false;

  // the state the reference carries from tick to tick, updated before its exceptions can fire (cpp:127, :216-232)
  plan_start_ = last_info_.plan_start;
  closer_to_goal_ = (last_info_.flags & 1u) != 0;
  if (last_info_.status == NEOMPC_CARROT_EMPTY_WINDOW)
    throw nav2_core::ControllerException("Resulting plan has 0 poses in it.");        // cpp:130-132
  slow_down_ = (last_info_.flags & 2u) != 0;
  if (last_info_.status == NEOMPC_CARROT_COLLISION)
    throw nav2_core::ControllerException("MPC detected collision!");                  // cpp:234-236

  geometry_msgs::msg::TwistStamped cmd_vel_final;                                     // cpp:250-254
  cmd_vel_final.header.frame_id = costmap_ros_->getBaseFrameID();
  cmd_vel_final.twist.linear.x = last_response_.vx;
  cmd_vel_final.twist.linear.y = last_response_.vy;
  cmd_vel_final.twist.angular.z = last_response_.omega;
  return cmd_vel_final;
}

// The predicted path the Python server published on 'local_plan' (publishLocalPlan, srv.py:271-310, called :365): computed
// on demand from the last solve.
const std::vector<neompc_plan_pose> & NeoMpcPlanner::lastLocalPlan() {
  std::lock_guard<std::mutex> lock(mutex_);
  if (!local_plan_valid_ && mpc_) {
    last_local_plan_.resize((size_t)params_.control_steps + 1);
    if (neompc_local_plan(mpc_, &last_request_, last_plan_.data(), 1, last_local_plan_.data()) != NEOMPC_OK)
      throw nav2_core::ControllerException(std::string("neompc_local_plan failed: ") + neompc_last_error(mpc_));
    local_plan_valid_ = true;
  }
  return last_local_plan_;
}

}  // namespace neo_mpc_planner

PLUGINLIB_EXPORT_CLASS(neo_mpc_planner::NeoMpcPlanner, nav2_core::Controller)

// NeoMpcPlanner.cpp — controller plugin body with the optimizer service call (reference src/NeoMpcPlanner.cpp:240-252)
// replaced by an in-process libneompc call.  See INTEGRATION.md.  The plan-following front half of the reference
// (TF transforms, pruning, slow-down hysteresis; cpp:66-232) is row N2 of SURVEY.md §8f and is reduced here to the
// carrot pick computeVelocityCommands needs, for plans already expressed in the costmap's global frame.
#include "NeoMpcPlanner.h"

#include <cmath>
#include <limits>

#include "pluginlib/class_list_macros.hpp"

namespace neo_mpc_planner {

namespace {
double yawOf(const geometry_msgs::msg::Quaternion & q) {
  return std::atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));
}
void putPose(double * dst, const geometry_msgs::msg::Pose & p) {
  dst[0] = p.position.x; dst[1] = p.position.y; dst[2] = p.position.z;
  dst[3] = p.orientation.x; dst[4] = p.orientation.y; dst[5] = p.orientation.z; dst[6] = p.orientation.w;
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
  global_plan_ = plan;
  if (!plan.poses.empty()) goal_pose_ = plan.poses.back().pose;
}

void NeoMpcPlanner::setSpeedLimit(const double &, const bool &) {}                   // empty in the reference too (cpp:283-288)

void NeoMpcPlanner::uploadCostmap() {
  auto * cm = costmap_ros_->getCostmap();
  if (neompc_set_costmap(mpc_, cm->getCharMap(), cm->getSizeInCellsX(), cm->getSizeInCellsY(), cm->getResolution(),
                         cm->getOriginX(), cm->getOriginY(), NEOMPC_ENC_NAV2_RAW) != NEOMPC_OK)
    throw nav2_core::ControllerException(std::string("neompc_set_costmap failed: ") + neompc_last_error(mpc_));
}

// Carrot = first plan pose, from the pose closest to the robot onwards, at least `lookahead` away; expressed in the
// robot base frame (what the reference sends as carrot_pose, cpp:114,124,173-189).
geometry_msgs::msg::PoseStamped NeoMpcPlanner::pickCarrot(const geometry_msgs::msg::PoseStamped & robot,
                                                          double lookahead) {
  const auto & poses = global_plan_.poses;
  size_t start = 0;
  double best = std::numeric_limits<double>::max();
  for (size_t i = 0; i < poses.size(); ++i) {
    const double d = std::hypot(poses[i].pose.position.x - robot.pose.position.x,
                                poses[i].pose.position.y - robot.pose.position.y);
    if (d < best) { best = d; start = i; }
  }
  size_t pick = poses.size() - 1;
  for (size_t i = start; i < poses.size(); ++i) {
    if (std::hypot(poses[i].pose.position.x - robot.pose.position.x,
                   poses[i].pose.position.y - robot.pose.position.y) >= lookahead) { pick = i; break; }
  }
  const double yaw = yawOf(robot.pose.orientation), c = std::cos(yaw), s = std::sin(yaw);
  const double dx = poses[pick].pose.position.x - robot.pose.position.x;
  const double dy = poses[pick].pose.position.y - robot.pose.position.y;
  geometry_msgs::msg::PoseStamped carrot;
  carrot.header.frame_id = costmap_ros_->getBaseFrameID();
  carrot.pose.position.x = c * dx + s * dy;
  carrot.pose.position.y = -s * dx + c * dy;
  const double rel = yawOf(poses[pick].pose.orientation) - yaw;
  carrot.pose.orientation.z = std::sin(0.5 * rel);
  carrot.pose.orientation.w = std::cos(0.5 * rel);
  return carrot;
}

geometry_msgs::msg::TwistStamped NeoMpcPlanner::computeVelocityCommands(
    const geometry_msgs::msg::PoseStamped & position, const geometry_msgs::msg::Twist & speed,
    nav2_core::GoalChecker *) {
  std::lock_guard<std::mutex> lock(mutex_);                                           // reference cpp:207
  if (global_plan_.poses.empty()) throw nav2_core::ControllerException("Received plan with zero length");
  if (!mpc_) throw nav2_core::ControllerException("controller not configured");

  const double to_goal = std::hypot(goal_pose_.position.x - position.pose.position.x,
                                    goal_pose_.position.y - position.pose.position.y);
  closer_to_goal_ = to_goal <= lookahead_dist_close_to_goal_;
  const double lookahead = closer_to_goal_ ? lookahead_dist_close_to_goal_ : lookahead_dist_max_;
  const auto carrot_pose = pickCarrot(position, lookahead);

  uploadCostmap();

  // the Optimizer request (reference cpp:240-246), marshalled for the C ABI
  neompc_optimizer_request m{};
  m.current_vel[0] = speed.linear.x; m.current_vel[1] = speed.linear.y; m.current_vel[2] = speed.linear.z;
  m.current_vel[3] = speed.angular.x; m.current_vel[4] = speed.angular.y; m.current_vel[5] = speed.angular.z;
  putPose(m.carrot_pose, carrot_pose.pose);
  putPose(m.goal_pose, goal_pose_);
  putPose(m.current_pose, position.pose);
  m.switch_opt = closer_to_goal_ ? 1u : 0u;
  m.control_interval = 1.0 / control_frequency_;
  const double now = clock_ ? clock_->now().seconds() : 0.0;
  m.delta_t = now - last_call_time_;                                                  // srv.py:369-371
  last_call_time_ = now;
  m.instance_id = 0;

  // the blocking service call and the response read (reference cpp:248-252)
  if (neompc_solve_msgs(mpc_, &m, 1, &last_response_, last_plan_.data()) != NEOMPC_OK)
    throw nav2_core::ControllerException(std::string("neompc_solve_msgs failed: ") + neompc_last_error(mpc_));

  // the predicted path the Python server published on /mpc_local_plan (publishLocalPlan, srv.py:271-310, called :365)
  {
    neompc_request rq{};
    rq.pose_x = (float)position.pose.position.x;
    rq.pose_y = (float)position.pose.position.y;
    const auto & q = position.pose.orientation;
    rq.pose_yaw = (float)std::atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));   // srv.py:176-178
    last_local_plan_.resize((size_t)params_.control_steps + 1);
    if (neompc_local_plan(mpc_, &rq, last_plan_.data(), 1, last_local_plan_.data()) != NEOMPC_OK)
      throw nav2_core::ControllerException(std::string("neompc_local_plan failed: ") + neompc_last_error(mpc_));
  }

  geometry_msgs::msg::TwistStamped cmd_vel_final;
  cmd_vel_final.header.frame_id = costmap_ros_->getBaseFrameID();
  cmd_vel_final.twist.linear.x = last_response_.vx;
  cmd_vel_final.twist.linear.y = last_response_.vy;
  cmd_vel_final.twist.angular.z = last_response_.omega;
  return cmd_vel_final;
}

}  // namespace neo_mpc_planner

PLUGINLIB_EXPORT_CLASS(neo_mpc_planner::NeoMpcPlanner, nav2_core::Controller)

// NeoMpcPlanner — nav2_core::Controller plugin surface of the reference (include/NeoMpcPlanner.h:52-127: configure,
// cleanup, activate, deactivate, computeVelocityCommands, setPlan, setSpeedLimit; exported as
// neo_mpc_planner::NeoMpcPlanner, neo_mpc_planner_plugin.xml:2) with the ROS service client replaced by a libneompc
// handle.  Written against the real Nav2 / rclcpp header names; in this repository it is compiled against the
// stand-ins under plugin/mock/ because the build image has no ROS 2.
#ifndef NEO_MPC_PLANNER2_B200_PLUGIN_NEOMPCPLANNER_H_
#define NEO_MPC_PLANNER2_B200_PLUGIN_NEOMPCPLANNER_H_

#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "geometry_msgs/msg/pose_stamped.hpp"
#include "geometry_msgs/msg/twist_stamped.hpp"
#include "nav2_core/controller.hpp"
#include "nav2_core/controller_exceptions.hpp"
#include "nav2_costmap_2d/costmap_2d_ros.hpp"
#include "nav_msgs/msg/path.hpp"
#include "rclcpp/rclcpp.hpp"
#include "rclcpp_lifecycle/lifecycle_node.hpp"
#include "tf2_ros/buffer.h"

#include "neompc.h"

namespace neo_mpc_planner {

class NeoMpcPlanner : public nav2_core::Controller {
public:
  NeoMpcPlanner() = default;
  ~NeoMpcPlanner() override;

  void configure(const rclcpp_lifecycle::LifecycleNode::WeakPtr & parent, std::string name,
                 std::shared_ptr<tf2_ros::Buffer> tf,
                 std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros) override;
  void cleanup() override;
  void activate() override;
  void deactivate() override;
  geometry_msgs::msg::TwistStamped computeVelocityCommands(const geometry_msgs::msg::PoseStamped & pose,
                                                           const geometry_msgs::msg::Twist & speed,
                                                           nav2_core::GoalChecker * goal_checker) override;
  void setPlan(const nav_msgs::msg::Path & path) override;
  void setSpeedLimit(const double & speed_limit, const bool & percentage) override;

  // diagnostics of the last solve (what the reference published as local_plan, srv.py:365)
  const std::vector<float> & lastPlan() const { return last_plan_; }
  const neompc_response & lastResponse() const { return last_response_; }
  const std::vector<neompc_plan_pose> & lastLocalPlan();       // poses of the 'local_plan' topic (srv.py:107), on demand
  const neompc_request & lastRequest() const { return last_request_; }      // the Optimizer request of the last tick (cpp:240-246)
  const neompc_carrot_info & lastCarrotInfo() const { return last_info_; }  // closest pose, carrot index, flags of the last tick
  bool lastTickUploadedCostmap() const { return last_upload_; }

private:
  bool uploadCostmapIfChanged();

  rclcpp_lifecycle::LifecycleNode::WeakPtr node_;
  std::shared_ptr<tf2_ros::Buffer> tf_;
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros_;
  rclcpp::Logger logger_{rclcpp::get_logger("MPC")};
  rclcpp::Clock::SharedPtr clock_;
  std::string plugin_name_;
  nav_msgs::msg::Path global_plan_;
  geometry_msgs::msg::Pose goal_pose_;
  bool closer_to_goal_ = false;
  bool slow_down_ = true;                // reference h:162
  uint32_t plan_start_ = 0;              // first plan pose not yet pruned (the reference erases the ones before, cpp:127)
  double lookahead_dist_min_ = 0.5, lookahead_dist_max_ = 0.5, lookahead_dist_close_to_goal_ = 0.5;
  double control_frequency_ = 20.0;
  double last_call_time_ = 0.0;          // srv.py:138
  neompc_handle * mpc_ = nullptr;        // replaces rclcpp::Client<neo_srvs2::srv::Optimizer> (reference h:150)
  neompc_params params_{};
  std::vector<float> last_plan_;
  std::vector<neompc_plan_pose> last_local_plan_;
  neompc_response last_response_{};
  neompc_request last_request_{};
  neompc_carrot_info last_info_{};
  bool local_plan_valid_ = false, last_upload_ = false;
  // what the device holds of the costmap (uploadCostmapIfChanged)
  bool costmap_loaded_ = false;
  uint64_t costmap_sum_ = 0;
  unsigned costmap_w_ = 0, costmap_h_ = 0;
  double costmap_res_ = 0.0, costmap_ox_ = 0.0, costmap_oy_ = 0.0;
  std::mutex mutex_;
};

}  // namespace neo_mpc_planner
#endif

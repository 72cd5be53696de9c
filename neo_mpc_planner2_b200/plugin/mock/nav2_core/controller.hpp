// Stand-in for nav2_core::Controller (Iron/Jazzy signature set, as overridden at reference include/NeoMpcPlanner.h:72-122).
#pragma once
#include <memory>
#include <string>
#include "geometry_msgs/msg/types.hpp"
#include "nav2_costmap_2d/costmap_2d_ros.hpp"
#include "rclcpp_lifecycle/lifecycle_node.hpp"
#include "tf2_ros/buffer.h"
namespace nav2_core {
class GoalChecker { public: virtual ~GoalChecker() = default; };
class Controller {
public:
  using Ptr = std::shared_ptr<Controller>;
  virtual ~Controller() = default;
  virtual void configure(const rclcpp_lifecycle::LifecycleNode::WeakPtr &, std::string name,
                         std::shared_ptr<tf2_ros::Buffer>, std::shared_ptr<nav2_costmap_2d::Costmap2DROS>) = 0;
  virtual void cleanup() = 0;
  virtual void activate() = 0;
  virtual void deactivate() = 0;
  virtual void setPlan(const nav_msgs::msg::Path & path) = 0;
  virtual geometry_msgs::msg::TwistStamped computeVelocityCommands(const geometry_msgs::msg::PoseStamped & pose,
                                                                    const geometry_msgs::msg::Twist & velocity,
                                                                    GoalChecker * goal_checker) = 0;
  virtual void setSpeedLimit(const double & speed_limit, const bool & percentage) = 0;
};
}

#pragma once
#include "rclcpp/rclcpp.hpp"
namespace rclcpp_lifecycle {
// parameters are doubles/ints keyed by name; get_parameter_or mirrors rclcpp's
class LifecycleNode {
public:
  using SharedPtr = std::shared_ptr<LifecycleNode>;
  using WeakPtr = std::weak_ptr<LifecycleNode>;
  std::map<std::string, double> params;
  template <typename T> bool get_parameter_or(const std::string & name, T & out, const T & dflt) const {
    auto it = params.find(name);
    if (it == params.end()) { out = dflt; return false; }
    out = static_cast<T>(it->second);
    return true;
  }
  rclcpp::Logger get_logger() const { return rclcpp::get_logger("controller_server"); }
  rclcpp::Clock::SharedPtr get_clock() const { return std::make_shared<rclcpp::Clock>(); }
};
}

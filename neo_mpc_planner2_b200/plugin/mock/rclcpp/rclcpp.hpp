// Stand-in for the slice of rclcpp the plugin uses: a logger, a clock and typed parameters on a node.
#pragma once
#include <chrono>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
namespace rclcpp {
struct Logger { std::string name; };
inline Logger get_logger(const std::string & n) { return Logger{n}; }
struct Time { double s = 0; double seconds() const { return s; } };
struct Clock {
  using SharedPtr = std::shared_ptr<Clock>;
  Time now() const {
    using namespace std::chrono;
    return Time{duration<double>(system_clock::now().time_since_epoch()).count()};
  }
};
}
#define RCLCPP_INFO(logger, ...) do { std::fprintf(stderr, "[INFO] [%s] ", (logger).name.c_str()); std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define RCLCPP_WARN RCLCPP_INFO
#define RCLCPP_ERROR RCLCPP_INFO

#pragma once
#include <stdexcept>
#include <string>
namespace nav2_core {
class ControllerException : public std::runtime_error {
public:
  explicit ControllerException(const std::string & d) : std::runtime_error(d) {}
};
}

// Stand-ins for the geometry_msgs / nav_msgs / std_msgs message structs the plugin touches (this image has no ROS 2).
// Field names and nesting follow the real messages so the plugin source compiles unchanged against a real Nav2.
#pragma once
#include <string>
#include <vector>
namespace builtin_interfaces { namespace msg { struct Time { int sec = 0; unsigned nanosec = 0; }; } }
namespace std_msgs { namespace msg { struct Header { builtin_interfaces::msg::Time stamp; std::string frame_id; }; } }
namespace geometry_msgs { namespace msg {
struct Point { double x = 0, y = 0, z = 0; };
struct Point32 { float x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
inline bool operator==(const Pose & a, const Pose & b) {
  return a.position.x == b.position.x && a.position.y == b.position.y && a.position.z == b.position.z &&
         a.orientation.x == b.orientation.x && a.orientation.y == b.orientation.y &&
         a.orientation.z == b.orientation.z && a.orientation.w == b.orientation.w;
}
inline bool operator!=(const Pose & a, const Pose & b) { return !(a == b); }
struct PoseStamped { std_msgs::msg::Header header; Pose pose; };
struct Twist { Vector3 linear, angular; };
struct TwistStamped { std_msgs::msg::Header header; Twist twist; };
}}
namespace nav_msgs { namespace msg { struct Path { std_msgs::msg::Header header; std::vector<geometry_msgs::msg::PoseStamped> poses; }; } }

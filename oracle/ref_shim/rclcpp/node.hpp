// ORACLE shim: just enough of rclcpp::Node for node.hpp to declare VINA_SLAM
#pragma once
#include <memory>
#include <string>
namespace rclcpp
{
struct Logger
{
};
class Node
{
public:
  typedef std::shared_ptr<Node> SharedPtr;
  Logger get_logger() const { return Logger(); }
  struct Now
  {
    double seconds() const { return 0.0; }
  };
  Now now() const { return Now(); }  // (initialization.cpp only stamps a wall-clock timer with it)
  template <typename T>
  bool get_parameter(const std::string&, T&) const
  {
    return false;
  }
};
}  // namespace rclcpp

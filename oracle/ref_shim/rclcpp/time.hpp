// ORACLE shim: rclcpp::Time / builtin_interfaces stamp as used by imu_ekf.cpp
#pragma once
#include <cstdint>
namespace builtin_interfaces
{
namespace msg
{
struct Time
{
  int32_t sec = 0;
  uint32_t nanosec = 0;
};
}  // namespace msg
}  // namespace builtin_interfaces
namespace rclcpp
{
class Time
{
  int64_t ns_;

public:
  Time() : ns_(0) {}
  explicit Time(int64_t ns) : ns_(ns) {}
  Time(const builtin_interfaces::msg::Time& s) : ns_((int64_t)s.sec * 1000000000LL + (int64_t)s.nanosec) {}
  double seconds() const { return (double)ns_ * 1e-9; }
  int64_t nanoseconds() const { return ns_; }
  operator builtin_interfaces::msg::Time() const
  {
    builtin_interfaces::msg::Time t;
    t.sec = (int32_t)(ns_ / 1000000000LL);
    t.nanosec = (uint32_t)(ns_ % 1000000000LL);
    return t;
  }
};
}  // namespace rclcpp

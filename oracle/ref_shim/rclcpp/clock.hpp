// ORACLE ref_shim - stand-in for <rclcpp/clock.hpp>: optimizers.cpp only stamps wall-clock timers with it
#pragma once
#include <chrono>
#include "rclcpp/time.hpp"
namespace rclcpp
{
struct ClockNow
{
  double s;
  double seconds() const { return s; }
};
class Clock
{
public:
  ClockNow now() const
  {
    return ClockNow{ std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() };
  }
};
}  // namespace rclcpp

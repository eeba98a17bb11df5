// ORACLE shim: sensor_msgs::msg::Imu fields used by imu_ekf.cpp
#pragma once
#include <memory>
#include <string>
#include "rclcpp/time.hpp"
namespace sensor_msgs
{
namespace msg
{
struct Imu
{
  typedef std::shared_ptr<Imu> SharedPtr;
  struct
  {
    builtin_interfaces::msg::Time stamp;
    std::string frame_id;
  } header;
  struct
  {
    double x = 0, y = 0, z = 0;
  } angular_velocity, linear_acceleration;
};
}  // namespace msg
}  // namespace sensor_msgs

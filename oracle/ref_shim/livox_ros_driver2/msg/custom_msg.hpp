// ORACLE shim: livox_ros_driver2::msg::CustomMsg as lidar_pointcloud_decoder.hpp / livox_handler use it
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include "rclcpp/time.hpp"
namespace livox_ros_driver2
{
namespace msg
{
struct CustomPoint
{
  uint32_t offset_time = 0;
  float x = 0, y = 0, z = 0;
  uint8_t reflectivity = 0, tag = 0, line = 0;
};
struct CustomMsg
{
  typedef std::shared_ptr<CustomMsg> SharedPtr;
  struct
  {
    builtin_interfaces::msg::Time stamp;
    std::string frame_id;
  } header;
  uint64_t timebase = 0;
  uint32_t point_num = 0;
  uint8_t lidar_id = 0;
  std::vector<CustomPoint> points;
};
}  // namespace msg
}  // namespace livox_ros_driver2

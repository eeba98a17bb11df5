// ORACLE shim: sensor_msgs::msg::PointCloud2 (declaration level)
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include "rclcpp/time.hpp"
namespace sensor_msgs
{
namespace msg
{
struct PointField
{
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 0;
  uint32_t count = 1;
};
struct PointCloud2
{
  typedef std::shared_ptr<PointCloud2> SharedPtr;
  struct
  {
    builtin_interfaces::msg::Time stamp;
    std::string frame_id;
  } header;
  uint32_t height = 1, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = true;
};
}  // namespace msg
}  // namespace sensor_msgs

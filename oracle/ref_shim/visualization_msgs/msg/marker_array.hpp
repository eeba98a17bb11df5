// ORACLE shim: visualization_msgs::msg::Marker(Array) fields written by OctoTree::collect_*_markers
// (visualisation only, out of scope; present so that octree.cpp compiles unmodified)
#pragma once
#include <cstdint>
#include <string>
#include <vector>
namespace geometry_msgs
{
namespace msg
{
struct Point
{
  double x = 0, y = 0, z = 0;
};
}  // namespace msg
}  // namespace geometry_msgs
namespace visualization_msgs
{
namespace msg
{
struct Marker
{
  enum
  {
    ARROW = 0,
    CUBE = 1,
    SPHERE = 2,
    CYLINDER = 3,
    ADD = 0,
    MODIFY = 0,
    DELETE = 2,
    DELETEALL = 3
  };
  struct
  {
    std::string frame_id;
  } header;
  std::string ns;
  int32_t id = 0, type = 0, action = 0;
  struct
  {
    geometry_msgs::msg::Point position;
    struct
    {
      double x = 0, y = 0, z = 0, w = 1;
    } orientation;
  } pose;
  struct
  {
    double x = 0, y = 0, z = 0;
  } scale;
  struct
  {
    float r = 0, g = 0, b = 0, a = 0;
  } color;
  std::vector<geometry_msgs::msg::Point> points;
};
struct MarkerArray
{
  std::vector<Marker> markers;
};
}  // namespace msg
}  // namespace visualization_msgs

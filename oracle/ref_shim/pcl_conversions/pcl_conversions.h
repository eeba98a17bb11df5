// ORACLE shim: pcl::fromROSMsg - every registered field of the point struct is looked up BY NAME in the message's
// field list and copied byte for byte (PCL's field map for fields of matching datatype; a field the message lacks
// stays value-initialised). Little-endian messages only, like the hosts the reference runs on.
#pragma once
#include <cstring>
#include "pcl/point_cloud.h"
#include "pcl/register_point_struct.h"
#include "sensor_msgs/msg/point_cloud2.hpp"
namespace pcl
{
template <typename P>
void fromROSMsg(const sensor_msgs::msg::PointCloud2& msg, PointCloud<P>& cloud)
{
  const std::vector<RosField> want = RosFieldMap<P>::fields();
  struct Map
  {
    size_t src, dst, size;
  };
  std::vector<Map> maps;
  for (const RosField& w : want)
    for (const auto& f : msg.fields)
      if (f.name == w.name) maps.push_back({ (size_t)f.offset, w.offset, w.size });
  const size_t n = (size_t)msg.width * msg.height;
  cloud.clear();
  cloud.resize(n);
  for (size_t i = 0; i < n; i++)
  {
    P p{};
    const uint8_t* src = msg.data.data() + i * msg.point_step;
    for (const Map& m : maps) std::memcpy(reinterpret_cast<char*>(&p) + m.dst, src + m.src, m.size);
    cloud[i] = p;
  }
}
}  // namespace pcl

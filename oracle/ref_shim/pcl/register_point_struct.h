// ORACLE shim: PCL's point-struct registration. POINT_CLOUD_REGISTER_POINT_STRUCT(name, (type, member, tag)...) records
// (tag, offset of the member, size of the type) per field - what pcl::fromROSMsg (pcl_conversions shim) needs to map the
// fields of a sensor_msgs/PointCloud2 by NAME into the struct, as PCL's field map does for fields of equal datatype.
#pragma once
#include <cstddef>
#include <string>
#include <vector>
#include "point_types.h"
#define PCL_ADD_POINT4D \
  float x;              \
  float y;              \
  float z;              \
  float data_pad_
#ifndef EIGEN_ALIGN16
#define EIGEN_ALIGN16 alignas(16)
#endif
namespace pcl
{
struct RosField
{
  std::string name;
  size_t offset;
  size_t size;
  RosField(const char* n, size_t o, size_t s) : name(n), offset(o), size(s) {}
};
struct RosFieldList
{
  std::vector<RosField> v;
  RosFieldList& operator<<(const RosField& f)
  {
    v.push_back(f);
    return *this;
  }
};
template <typename P>
struct RosFieldMap;
template <>
struct RosFieldMap<PointXYZ>
{
  static std::vector<RosField> fields()
  {
    return (RosFieldList() << RosField("x", offsetof(PointXYZ, x), 4) << RosField("y", offsetof(PointXYZ, y), 4)
                           << RosField("z", offsetof(PointXYZ, z), 4)).v;
  }
};
}  // namespace pcl
// sequence iteration over (a, b, c)(d, e, f)... without Boost.Preprocessor
#define VSHIM_SEQ_A(t, n, tag) VSHIM_ITEM(t, n, tag) VSHIM_SEQ_B
#define VSHIM_SEQ_B(t, n, tag) VSHIM_ITEM(t, n, tag) VSHIM_SEQ_A
#define VSHIM_SEQ_A_END
#define VSHIM_SEQ_B_END
#define VSHIM_CAT(a, b) VSHIM_CAT_(a, b)
#define VSHIM_CAT_(a, b) a##b
#define VSHIM_ITEM(t, n, tag) << pcl::RosField(#tag, offsetof(P_, n), sizeof(t))
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fseq)                 \
  namespace pcl                                                       \
  {                                                                   \
  template <>                                                         \
  struct RosFieldMap<name>                                            \
  {                                                                   \
    typedef name P_;                                                  \
    static std::vector<RosField> fields()                             \
    {                                                                 \
      return (RosFieldList() VSHIM_CAT(VSHIM_SEQ_A fseq, _END)).v;    \
    }                                                                 \
  };                                                                  \
  }

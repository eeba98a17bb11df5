// ORACLE shim: the fields of pcl::PointXYZINormal that the reference reads (48-byte layout kept)
#pragma once
namespace pcl
{
struct alignas(16) PointXYZINormal
{
  union
  {
    float data[4];
    struct
    {
      float x, y, z;
    };
  };
  union
  {
    float data_n[4];
    struct
    {
      float normal_x, normal_y, normal_z;
    };
  };
  float intensity = 0;
  float curvature = 0;
  float pad_[2];
  PointXYZINormal() : data{ 0, 0, 0, 1 }, data_n{ 0, 0, 0, 0 } {}
};
struct alignas(16) PointXYZ
{
  float x = 0, y = 0, z = 0;
  float pad_ = 1;
};
}  // namespace pcl

// ORACLE shim: the point-struct registration macros of PCL reduced to what the declarations in
// lidar_pointcloud_decoder.hpp need to parse (no field reflection: pcl::fromROSMsg is not available here)
#pragma once
#define PCL_ADD_POINT4D \
  float x;              \
  float y;              \
  float z;              \
  float data_pad_
#ifndef EIGEN_ALIGN16
#define EIGEN_ALIGN16 alignas(16)
#endif
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fseq) static_assert(sizeof(name) > 0, "");

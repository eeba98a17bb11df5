// ORACLE shim: pcl::KdTreeFLANN is used only by the init-phase kd-tree IEKF (odometry.cpp:267-439), which is
// out of scope. Brute-force stand-in so that odometry.cpp compiles unmodified.
#pragma once
#include <algorithm>
#include <vector>
#include "../point_cloud.h"
namespace pcl
{
template <typename T>
class KdTreeFLANN
{
  typename PointCloud<T>::Ptr cloud_;

public:
  void setInputCloud(const typename PointCloud<T>::Ptr& c) { cloud_ = c; }
  int nearestKSearch(const T& q, int k, std::vector<int>& idx, std::vector<float>& d2) const
  {
    std::vector<std::pair<float, int>> all;
    for (size_t i = 0; i < cloud_->size(); i++)
    {
      const T& p = (*cloud_)[i];
      float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
      all.push_back({ dx * dx + dy * dy + dz * dz, (int)i });
    }
    k = std::min<int>(k, (int)all.size());
    std::partial_sort(all.begin(), all.begin() + k, all.end());
    idx.resize(k);
    d2.resize(k);
    for (int i = 0; i < k; i++)
    {
      idx[i] = all[i].second;
      d2[i] = all[i].first;
    }
    return k;
  }
};
}  // namespace pcl

// ORACLE shim: the std::vector-like subset of pcl::PointCloud used by the reference's hot path
#pragma once
#include <memory>
#include <vector>
#include "point_types.h"
namespace pcl
{
template <typename T>
class PointCloud
{
public:
  typedef std::shared_ptr<PointCloud<T>> Ptr;
  typedef typename std::vector<T>::iterator iterator;
  typedef typename std::vector<T>::const_iterator const_iterator;
  std::vector<T> points;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); }
  void reserve(size_t n) { points.reserve(n); }
  void resize(size_t n) { points.resize(n); }
  void push_back(const T& p) { points.push_back(p); }
  T& operator[](size_t i) { return points[i]; }
  const T& operator[](size_t i) const { return points[i]; }
  T& back() { return points.back(); }
  T& front() { return points.front(); }
  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); }
  const_iterator end() const { return points.end(); }
  void swap(PointCloud& o) { points.swap(o.points); }
};
}  // namespace pcl

// Unpacking of the LiDAR driver messages: LidarPointCloudDecoder::process and its six handlers
// (src/sensor/lidar_pointcloud_decoder.cpp:21-240; point layouts include/vina_slam/lidar_pointcloud_decoder.hpp:44-109)
// as host functions behind the C ABI - a single pass over the message bytes that writes n x (x, y, z, curvature)
// float32 in arrival order with each handler's own time rule and keep rule applied. The result goes to
// vina_scan_prepare (point_filter_num = 1, blind2 < 0: sort + 0.11 s cut only). No ROS / PCL types: a PointCloud2 is
// its data pointer plus the offsets of the fields pcl::fromROSMsg would map by name.
#include <cmath>
#include <cstring>

#include "../../include/vina_b200.h"

namespace
{
template <typename T>
inline T field(const uint8_t* p, bool big)
{
  T v;
  if (big)
  {
    uint8_t r[sizeof(T)];
    for (size_t i = 0; i < sizeof(T); i++) r[i] = p[sizeof(T) - 1 - i];
    memcpy(&v, r, sizeof(T));
  }
  else
    memcpy(&v, p, sizeof(T));
  return v;
}

// the time field as the handler's point struct holds it
inline double time_field(const uint8_t* p, int datatype, bool big)
{
  switch (datatype)
  {
    case 6: return (double)field<uint32_t>(p, big);  // sensor_msgs::PointField::UINT32
    case 7: return (double)field<float>(p, big);     // FLOAT32
    default: return field<double>(p, big);           // FLOAT64
  }
}
}  // namespace

extern "C" {

int64_t vina_decode_pointcloud2(int lidar_type, const uint8_t* data, int64_t n_points, const vina_pc2_layout* L,
                                double header_stamp, double omega_l, double blind2, int point_filter_num, float* xyzt,
                                int64_t cap)
{
  if (!L || (!data && n_points > 0) || n_points < 0 || !xyzt || point_filter_num < 1 || L->point_step <= 0)
    return VINA_E_ARG;
  if (lidar_type < VINA_LIDAR_VELODYNE || lidar_type > VINA_LIDAR_TARTANAIR) return VINA_E_ARG;  // "Unsupported lidar type"
  if (lidar_type != VINA_LIDAR_TARTANAIR && (L->off_t < 0 || L->t_datatype < 6 || L->t_datatype > 8)) return VINA_E_ARG;
  const bool big = L->is_bigendian != 0;
  const size_t N = (size_t)n_points;
  int64_t out = 0;
  auto xyz = [&](size_t i, float& x, float& y, float& z) {
    const uint8_t* p = data + i * (size_t)L->point_step;
    x = field<float>(p + L->off_x, big);
    y = field<float>(p + L->off_y, big);
    z = field<float>(p + L->off_z, big);
  };
  auto tfield = [&](size_t i) { return time_field(data + i * (size_t)L->point_step + L->off_t, L->t_datatype, big); };
  auto emit = [&](float x, float y, float z, float c) -> bool {
    if (out >= cap) return false;
    float* o = xyzt + 4 * out++;
    o[0] = x, o[1] = y, o[2] = z, o[3] = c;
    return true;
  };
  if (N == 0) return 0;

  if (lidar_type == VINA_LIDAR_VELODYNE)
  {
    const float last_time = (float)tfield(N - 1);  // velodyne_ros::Point::time is a float
    if (last_time > 0.01 && last_time < 0.12)
    {
      for (size_t i = 0; i < N; ++i)
      {
        float x, y, z;
        xyz(i, x, y, z);
        const float c = (float)tfield(i);
        if ((i % point_filter_num) == 0 && (x * x + y * y + z * z) > blind2)
          if (!emit(x, y, z, c)) return VINA_E_CAPACITY;
      }
    }
    else
    {
      // no usable per-point time: stamps from the azimuth at omega_l deg/s. PROVENANCE: this loop RESTATES
      // velodyne_handler's azimuth branch (src/sensor/lidar_pointcloud_decoder.cpp:103-139) statement for statement
      // and with its names - a sequential unwrap state machine (yaw0, yaw_last, bias, cool) whose every output stamp
      // must equal the reference's; tests/test_decode_cpu.py replays it against the reference's own file
      bool first = true;
      double yaw0 = 0, yaw_last = 0, bias = 0;
      int cool = 0;
      for (size_t i = 0; i < N; ++i)
      {
        float x, y, z;
        xyz(i, x, y, z);
        if (std::fabs(x) < 0.1) continue;
        double yaw = std::atan2(y, x) * 57.2957795 - bias;
        if (first)
        {
          yaw0 = yaw_last = yaw;
          first = false;
        }
        if (x * x + y * y + z * z < blind2) continue;
        if ((yaw - yaw_last) > 180 && cool-- <= 0)
        {
          bias += 360;
          yaw -= 360;
          cool = 1000;
        }
        if (std::fabs(yaw - yaw_last) > 180) yaw += 360;
        const float c = (float)((yaw0 - yaw) / omega_l);
        yaw_last = yaw;
        if (c >= 0 && c < 0.1 && (i % point_filter_num) == 0)
          if (!emit(x, y, z, c)) return VINA_E_CAPACITY;
      }
    }
    return out;
  }

  double t0 = 0;
  if (lidar_type == VINA_LIDAR_HESAI) t0 = tfield(0);  // pl_orig.points.front().timestamp (:168)
  for (size_t i = 0; i < N; ++i)
  {
    float x, y, z, c = 0.f;
    xyz(i, x, y, z);
    bool keep;
    switch (lidar_type)
    {
      case VINA_LIDAR_OUSTER:  // uint32 nanoseconds (:157)
        c = (float)((double)field<uint32_t>(data + i * (size_t)L->point_step + L->off_t, big) / 1e9);
        keep = (i % point_filter_num) == 0 && (x * x + y * y + z * z) > blind2;
        break;
      case VINA_LIDAR_HESAI:  // absolute double stamps, relative to the first point (:184)
        c = (float)(tfield(i) - t0);
        keep = (i % point_filter_num) == 0 && (x * x + y * y + z * z) > blind2;
        break;
      case VINA_LIDAR_ROBOSENSE:  // absolute double stamps, relative to the header; PLANAR blind test (:214-217)
        c = (float)(tfield(i) - header_stamp);
        keep = ((i % point_filter_num) == 0) && ((x * x + y * y) > blind2);
        break;
      default:  // TartanAir: no time, no filter (:228-239)
        c = 0.f;
        keep = true;
        break;
    }
    if (keep && !emit(x, y, z, c)) return VINA_E_CAPACITY;
  }
  return out;
}

int64_t vina_decode_livox(const vina_livox_point* pts, int64_t n_points, double blind2, int point_filter_num,
                          float* xyzt, int64_t cap)
{
  if ((!pts && n_points > 0) || n_points < 0 || !xyzt || point_filter_num < 1) return VINA_E_ARG;
  int64_t out = 0;
  for (size_t i = 0; i < (size_t)n_points; ++i)
  {
    const float x = pts[i].x, y = pts[i].y, z = pts[i].z;
    const float c = (float)(pts[i].offset_time * (1e-9));  // (:66)
    if ((i % point_filter_num) == 0 && (x * x + y * y + z * z) > blind2)
    {
      if (out >= cap) return VINA_E_CAPACITY;
      float* o = xyzt + 4 * out++;
      o[0] = x, o[1] = y, o[2] = z, o[3] = c;
    }
  }
  return out;
}

// back().curvature of the scan pcl_handler would queue for these points (src/sensor/lidar_decoder.cpp:16-34): the
// largest time offset that is not beyond 0.11 s; 0.09 for an empty cloud (its two-point stand-in). What
// vina_sync_push_scan needs when the scan is only prepared (sorted, cut) later, on the device. Returns 0, or
// VINA_E_ARG where the reference would be left with an empty cloud (every stamp beyond 0.11 s).
int vina_scan_last_stamp(const float* xyzt, int64_t n, float* t_last)
{
  if ((!xyzt && n > 0) || n < 0 || !t_last) return VINA_E_ARG;
  if (n == 0)
  {
    *t_last = 0.09f;
    return VINA_OK;
  }
  bool any = false;
  float best = 0.f;
  for (int64_t i = 0; i < n; i++)
  {
    const float t = xyzt[4 * i + 3];
    if ((double)t > 0.11) continue;
    if (!any || t > best) best = t;
    any = true;
  }
  if (!any) return VINA_E_ARG;
  *t_last = best;
  return VINA_OK;
}
}

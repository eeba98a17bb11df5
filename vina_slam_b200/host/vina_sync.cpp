// Pairing of scans with their IMU samples in front of the per-scan step: the buffers of src/sensor/sync.cpp:5-16
// (filled by imu_handler, src/platform/ros2/subscribers.cpp:11-20, and the tail of pcl_handler,
// src/sensor/lidar_decoder.cpp:36-43) and sync_packages (src/sensor/sync.cpp:18-96) as a small host object
// behind the C ABI. Pure C++: no device, no context - the scan itself stays with the caller (or on the device,
// vina_scan_prepare); the queue carries its start time, its last time offset and an opaque tag.
// The reference's globals (imu_buf, pcl_buf + time_buf, last_pcl_time) and the function-local `static bool pl_ready`
// become the members imu_q, scan_q, prev_stamp and holding; the mutex stays (handlers
// and the odometry thread are different threads there), and exit(0) on a drained IMU buffer becomes VINA_E_STATE.
#include <deque>
#include <mutex>

#include "../../include/vina_b200.h"

struct vina_sync
{
  std::mutex mtx;
  std::deque<vina_imu> imu_q;
  struct Scan
  {
    double t_start;
    double t_last;
    int64_t tag;
  };
  std::deque<Scan> scan_q;  // scan_q + time_buf
  double imu_last_time = -1;
  int point_notime = 0;
  double prev_stamp = -1;
  bool holding = false;
  Scan cur = { 0, 0, 0 };            // pl_ptr of the caller, held across calls while the IMU lags behind
  double pcl_beg_time = 0, pcl_end_time = 0;  // p_imu.pcl_beg_time / pcl_end_time
};

extern "C" {

int vina_sync_create(int point_notime, vina_sync** out)
{
  if (!out) return VINA_E_ARG;
  vina_sync* s = new vina_sync();
  s->point_notime = point_notime;
  *out = s;
  return VINA_OK;
}

void vina_sync_destroy(vina_sync* s) { delete s; }

int vina_sync_push_imu(vina_sync* s, const vina_imu* imu)
{
  if (!s || !imu) return VINA_E_ARG;
  std::lock_guard<std::mutex> lk(s->mtx);
  s->imu_last_time = imu->t;
  s->imu_q.push_back(*imu);
  return VINA_OK;
}

int vina_sync_push_scan(vina_sync* s, double t_start, double t_last, int64_t tag)
{
  if (!s) return VINA_E_ARG;
  std::lock_guard<std::mutex> lk(s->mtx);
  s->scan_q.push_back({ t_start, t_last, tag });
  return VINA_OK;
}

int vina_sync_pending(vina_sync* s, int32_t* scans, int32_t* imus)
{
  if (!s) return VINA_E_ARG;
  std::lock_guard<std::mutex> lk(s->mtx);
  if (scans) *scans = (int32_t)s->scan_q.size() + (s->holding ? 1 : 0);
  if (imus) *imus = (int32_t)s->imu_q.size();
  return VINA_OK;
}

int vina_sync_next(vina_sync* s, int64_t* tag, double* pcl_beg_time, double* pcl_end_time, vina_imu* imus, int cap,
                   int32_t* m)
{
  if (!s || !tag || !pcl_beg_time || !pcl_end_time || !imus || !m || cap < 0) return VINA_E_ARG;
  *m = 0;
  if (!s->holding)
  {
    std::unique_lock<std::mutex> lk(s->mtx);
    if (s->scan_q.empty()) return 0;
    s->cur = s->scan_q.front();
    s->scan_q.pop_front();
    lk.unlock();
    s->pcl_beg_time = s->cur.t_start;
    s->pcl_end_time = s->pcl_beg_time + s->cur.t_last;  // + pl_ptr->back().curvature
    if (s->point_notime)
    {
      if (s->prev_stamp < 0)
      {
        s->prev_stamp = s->pcl_beg_time;
        *tag = s->cur.tag;
        return 2;  // the first scan only seeds the frame interval
      }
      s->pcl_end_time = s->pcl_beg_time;
      s->pcl_beg_time = s->prev_stamp;
      s->prev_stamp = s->pcl_end_time;
    }
    s->holding = true;
  }
  std::unique_lock<std::mutex> lk(s->mtx);
  if (s->imu_last_time <= s->pcl_end_time) return 0;
  int n = 0;
  bool overflow = false;
  double stamp = s->imu_q.front().t;  // (not empty: the sample stamped imu_last_time is never consumed below)
  while (!s->imu_q.empty() && stamp < s->pcl_end_time)
  {
    stamp = s->imu_q.front().t;
    if (stamp > s->pcl_end_time) break;
    if (n < cap)
      imus[n] = s->imu_q.front();
    else
      overflow = true;
    n++;
    s->imu_q.pop_front();
  }
  const bool drained = s->imu_q.empty();
  lk.unlock();
  s->holding = false;
  *tag = s->cur.tag;
  *pcl_beg_time = s->pcl_beg_time;
  *pcl_end_time = s->pcl_end_time;
  if (drained) return VINA_E_STATE;  // "the data flow is broken": the reference exit(0)s here
  if (overflow) return VINA_E_CAPACITY;
  *m = n;
  return n > 4 ? 1 : 2;
}
}

// Pairing of scans with their IMU samples in front of the per-scan step.
//
// PROVENANCE: this file RESTATES src/sensor/sync.cpp:5-96 of the reference - its buffers (imu_buf, pcl_buf + time_buf,
// imu_last_time, last_pcl_time, filled by imu_handler, src/platform/ros2/subscribers.cpp:11-20, and by the tail of
// pcl_handler, src/sensor/lidar_decoder.cpp:36-43) and the control flow of sync_packages (sync.cpp:18-96), statement
// for statement, because the entry points must return exactly what the reference returns for every interleaving of
// messages (tests/test_sync_cpu.py replays random streams against the reference's own sync.cpp). It keeps the
// reference's names. What differs: the globals and the function-local `static bool pl_ready` are members of a small
// object behind the C ABI, the scan itself stays with the caller (or on the device, vina_scan_prepare) and the queue
// carries its start time, last time offset and an opaque tag, and exit(0) on a drained IMU buffer is VINA_E_STATE.
// Pure C++: no device, no context.
#include <deque>
#include <mutex>

#include "../../include/vina_b200.h"

struct vina_sync
{
  std::mutex mBuf;
  std::deque<vina_imu> imu_buf;
  struct Scan
  {
    double t_start;
    double t_last;
    int64_t tag;
  };
  std::deque<Scan> pcl_buf;  // pcl_buf + time_buf
  double imu_last_time = -1;
  int point_notime = 0;
  double last_pcl_time = -1;
  bool pl_ready = false;
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
  std::lock_guard<std::mutex> lk(s->mBuf);
  s->imu_last_time = imu->t;
  s->imu_buf.push_back(*imu);
  return VINA_OK;
}

int vina_sync_push_scan(vina_sync* s, double t_start, double t_last, int64_t tag)
{
  if (!s) return VINA_E_ARG;
  std::lock_guard<std::mutex> lk(s->mBuf);
  s->pcl_buf.push_back({ t_start, t_last, tag });
  return VINA_OK;
}

int vina_sync_pending(vina_sync* s, int32_t* scans, int32_t* imus)
{
  if (!s) return VINA_E_ARG;
  std::lock_guard<std::mutex> lk(s->mBuf);
  if (scans) *scans = (int32_t)s->pcl_buf.size() + (s->pl_ready ? 1 : 0);
  if (imus) *imus = (int32_t)s->imu_buf.size();
  return VINA_OK;
}

int vina_sync_next(vina_sync* s, int64_t* tag, double* pcl_beg_time, double* pcl_end_time, vina_imu* imus, int cap,
                   int32_t* m)
{
  if (!s || !tag || !pcl_beg_time || !pcl_end_time || !imus || !m || cap < 0) return VINA_E_ARG;
  *m = 0;
  if (!s->pl_ready)
  {
    std::unique_lock<std::mutex> lk(s->mBuf);
    if (s->pcl_buf.empty()) return 0;
    s->cur = s->pcl_buf.front();
    s->pcl_buf.pop_front();
    lk.unlock();
    s->pcl_beg_time = s->cur.t_start;
    s->pcl_end_time = s->pcl_beg_time + s->cur.t_last;  // + pl_ptr->back().curvature
    if (s->point_notime)
    {
      if (s->last_pcl_time < 0)
      {
        s->last_pcl_time = s->pcl_beg_time;
        *tag = s->cur.tag;
        return 2;  // the first scan only seeds the frame interval
      }
      s->pcl_end_time = s->pcl_beg_time;
      s->pcl_beg_time = s->last_pcl_time;
      s->last_pcl_time = s->pcl_end_time;
    }
    s->pl_ready = true;
  }
  std::unique_lock<std::mutex> lk(s->mBuf);
  if (s->imu_last_time <= s->pcl_end_time) return 0;
  // (the sample stamped imu_last_time is the last one pushed and lies beyond the scan: it is never consumed, so the
  // buffer is not empty here - checked all the same, the reference would read front() of an empty deque)
  if (s->imu_buf.empty()) return VINA_E_STATE;
  // the samples this call hands over: counted first, so that a buffer that is too small loses nothing - the call
  // returns VINA_E_CAPACITY with the scan still held and can be repeated with a larger one
  int n = 0;
  double imu_time = s->imu_buf.front().t;
  for (const vina_imu& im : s->imu_buf)
  {
    if (!(imu_time < s->pcl_end_time)) break;  // while (!imu_buf.empty() && imu_time < pcl_end_time)
    imu_time = im.t;
    if (imu_time > s->pcl_end_time) break;
    n++;
  }
  if (n > cap)
  {
    *tag = s->cur.tag;  // (which scan is waiting, and for how many samples)
    *pcl_beg_time = s->pcl_beg_time;
    *pcl_end_time = s->pcl_end_time;
    *m = n;
    return VINA_E_CAPACITY;
  }
  for (int k = 0; k < n; k++) imus[k] = s->imu_buf[k];
  s->imu_buf.erase(s->imu_buf.begin(), s->imu_buf.begin() + n);
  const bool drained = s->imu_buf.empty();
  lk.unlock();
  s->pl_ready = false;
  *tag = s->cur.tag;
  *pcl_beg_time = s->pcl_beg_time;
  *pcl_end_time = s->pcl_end_time;
  if (drained) return VINA_E_STATE;  // "the data flow is broken": the reference exit(0)s here
  *m = n;
  return n > 4 ? 1 : 2;
}
}

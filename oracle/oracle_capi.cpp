// ORACLE — TEST INFRASTRUCTURE ONLY (pinned against oracle/_ref, see vina_oracle.hpp).
#include <algorithm>
#include <deque>
#include <vector>
#include "oracle_capi.h"
#include "vina_oracle.hpp"
#include <cstring>

using namespace vo;

static Mat3 m3(const double* a)
{
  Mat3 m;
  memcpy(m.d, a, sizeof(m.d));
  return m;
}
static Vec3 v3(const double* a) { return V3(a[0], a[1], a[2]); }

static void to_imust(const vo_state* s, IMUST& x)
{
  x.t = s->t;
  x.R = m3(s->R);
  x.p = v3(s->p);
  x.v = v3(s->v);
  x.bg = v3(s->bg);
  x.ba = v3(s->ba);
  x.g = v3(s->g);
  memcpy(x.cov.d, s->cov, sizeof(x.cov.d));
}
static void from_imust(const IMUST& x, vo_state* s)
{
  s->t = x.t;
  memcpy(s->R, x.R.d, sizeof(s->R));
  memcpy(s->p, x.p.d, 24);
  memcpy(s->v, x.v.d, 24);
  memcpy(s->bg, x.bg.d, 24);
  memcpy(s->ba, x.ba.d, 24);
  memcpy(s->g, x.g.d, 24);
  memcpy(s->cov, x.cov.d, sizeof(s->cov));
}
static Cloud to_cloud(const float* xyz4, int n)
{
  Cloud c(n);
  for (int i = 0; i < n; i++)
  {
    c[i].x = xyz4[4 * i + 0];
    c[i].y = xyz4[4 * i + 1];
    c[i].z = xyz4[4 * i + 2];
    c[i].curvature = xyz4[4 * i + 3];
  }
  return c;
}
static void from_cloud(const Cloud& c, float* xyz4)
{
  for (size_t i = 0; i < c.size(); i++)
  {
    xyz4[4 * i + 0] = c[i].x;
    xyz4[4 * i + 1] = c[i].y;
    xyz4[4 * i + 2] = c[i].z;
    xyz4[4 * i + 3] = c[i].curvature;
  }
}
static std::deque<ImuSample> to_imus(const double* imu7, int m)
{
  std::deque<ImuSample> d;
  for (int i = 0; i < m; i++)
  {
    ImuSample s;
    s.t = imu7[7 * i];
    for (int k = 0; k < 3; k++)
    {
      s.gyr[k] = imu7[7 * i + 1 + k];
      s.acc[k] = imu7[7 * i + 4 + k];
    }
    d.push_back(s);
  }
  return d;
}

extern "C" {

void vo_eig3(const double A[9], double vals[3], double vecs[9])
{
  SelfAdjointEigen3 s(m3(A));
  memcpy(vals, s.values.d, 24);
  memcpy(vecs, s.vectors.d, 72);
}
void vo_inverse15(const double A[225], double out[225])
{
  Mat15 m;
  memcpy(m.d, A, sizeof(m.d));
  Mat15 r = inverse(m);
  memcpy(out, r.d, sizeof(r.d));
}
void vo_exp(const double w[3], double R[9])
{
  Mat3 r = Exp(v3(w));
  memcpy(R, r.d, 72);
}
void vo_exp_dt(const double w[3], double dt, double R[9])
{
  Mat3 r = Exp(v3(w), dt);
  memcpy(R, r.d, 72);
}
void vo_log(const double R[9], double w[3])
{
  Vec3 r = Log(m3(R));
  memcpy(w, r.d, 24);
}
void vo_var_init(int n, const float* xyz4, const double ext_R[9], const double ext_t[3], double dept_err,
                 double beam_err, double* pnt, double* var)
{
  IMUST ext;
  ext.R = m3(ext_R);
  ext.p = v3(ext_t);
  Cloud c = to_cloud(xyz4, n);
  PVecPtr pptr(new PVec);
  var_init(ext, c, pptr, dept_err, beam_err);
  for (int i = 0; i < n; i++)
  {
    memcpy(pnt + 3 * (size_t)i, (*pptr)[i].pnt.d, 24);
    memcpy(var + 9 * (size_t)i, (*pptr)[i].var.d, 72);
  }
}
void vo_pvec_update(int n, const double* pnt, double* var, const double R[9], const double p[3],
                    const double cov[225], double* pwld)
{
  IMUST x;
  x.R = m3(R);
  x.p = v3(p);
  memcpy(x.cov.d, cov, sizeof(x.cov.d));
  PVecPtr pptr(new PVec(n));
  for (int i = 0; i < n; i++)
  {
    (*pptr)[i].pnt = v3(pnt + 3 * (size_t)i);
    (*pptr)[i].var = m3(var + 9 * (size_t)i);
  }
  std::vector<Vec3> pw;
  pvec_update(pptr, x, pw);
  for (int i = 0; i < n; i++)
  {
    memcpy(var + 9 * (size_t)i, (*pptr)[i].var.d, 72);
    memcpy(pwld + 3 * (size_t)i, pw[i].d, 24);
  }
}
void vo_voxel_keys(int n, const double* pw, double voxel_size, int64_t* keys)
{
  for (int i = 0; i < n; i++)
  {
    VOXEL_LOC k = voxel_key(v3(pw + 3 * (size_t)i), voxel_size);
    keys[3 * (size_t)i + 0] = k.x;
    keys[3 * (size_t)i + 1] = k.y;
    keys[3 * (size_t)i + 2] = k.z;
  }
}
int vo_down_sampling_voxel(int n, const float* xyz4_in, double voxel_size, float* xyz4_out)
{
  Cloud c = to_cloud(xyz4_in, n);
  down_sampling_voxel(c, voxel_size);
  from_cloud(c, xyz4_out);
  return (int)c.size();
}

// The scan front end between the sensor driver and IMUEKF::process: the keep rule every decoder handler applies
// (src/sensor/lidar_pointcloud_decoder.cpp:70, 96, 161, 190: point i stays iff i % point_filter_num == 0 and
// x*x + y*y + z*z > blind, float products compared with the double `blind`, which node.cpp:210 has already
// squared) followed by pcl_handler (src/sensor/lidar_decoder.cpp:16-34): two dummy points for an empty cloud,
// sort by time offset, drop the tail beyond 0.11 s. The reference's std::sort leaves the order of equal stamps
// unspecified; this restatement (and the device) fix it as the arrival order (stable sort).
// Returns the number of points written, or -1 where the reference would call back() on an empty cloud.
int vo_scan_prepare(int n, const float* xyz4_in, int point_filter_num, double blind2, float* xyz4_out)
{
  struct P
  {
    float x, y, z, t;
  };
  std::vector<P> pl;
  for (int i = 0; i < n; i++)
  {
    P pt = { xyz4_in[4 * i], xyz4_in[4 * i + 1], xyz4_in[4 * i + 2], xyz4_in[4 * i + 3] };
    if ((i % point_filter_num) == 0 && (pt.x * pt.x + pt.y * pt.y + pt.z * pt.z) > blind2) pl.push_back(pt);
  }
  if (pl.empty())
  {
    pl.push_back({ 0, 0, 0, 0 });
    pl.push_back({ 0, 0, 0, 0.09f });
  }
  std::stable_sort(pl.begin(), pl.end(), [](const P& a, const P& b) { return a.t < b.t; });
  while (!pl.empty() && pl.back().t > 0.11) pl.pop_back();
  if (pl.empty()) return -1;
  memcpy(xyz4_out, pl.data(), pl.size() * sizeof(P));
  return (int)pl.size();
}

// src/sensor/sync.cpp:5-96 restated with its globals (imu_buf, pcl_buf, time_buf, imu_last_time, point_notime,
// last_pcl_time) and the function-local static pl_ready as members of one object; a scan is represented by
// back().curvature and a tag. imu_handler = src/platform/ros2/subscribers.cpp:11-20, the scan push = the tail of
// pcl_handler (src/sensor/lidar_decoder.cpp:36-43). Return of vo_sync_next: 1 = sync_packages returned true,
// 0 = false with nothing consumed, 2 = false and the scan is gone, -6 = the reference exit(0)s (:79-82).
namespace
{
struct SyncRef
{
  std::deque<ImuSample> imu_buf;
  std::deque<std::pair<double, int64_t>> pcl_buf;  // (back().curvature, tag)
  std::deque<double> time_buf;
  double imu_last_time = -1;
  int point_notime = 0;
  double last_pcl_time = -1;
  bool pl_ready = false;
  std::pair<double, int64_t> pl;
  double pcl_beg_time = 0, pcl_end_time = 0;
};
}  // namespace
void* vo_sync_create(int point_notime)
{
  SyncRef* s = new SyncRef();
  s->point_notime = point_notime;
  return s;
}
void vo_sync_destroy(void* h) { delete (SyncRef*)h; }
void vo_sync_push_imu(void* h, const double imu7[7])
{
  SyncRef* s = (SyncRef*)h;
  ImuSample m;
  m.t = imu7[0];
  for (int k = 0; k < 3; k++) m.gyr[k] = imu7[1 + k], m.acc[k] = imu7[4 + k];
  s->imu_last_time = m.t;
  s->imu_buf.push_back(m);
}
void vo_sync_push_scan(void* h, double t_start, double t_last, int64_t tag)
{
  SyncRef* s = (SyncRef*)h;
  s->time_buf.push_back(t_start);
  s->pcl_buf.push_back({ t_last, tag });
}
int vo_sync_next(void* h, int64_t* tag, double* beg, double* end, double* imu7, int cap, int* m)
{
  SyncRef* s = (SyncRef*)h;
  *m = 0;
  if (!s->pl_ready)
  {
    if (s->pcl_buf.empty()) return 0;
    s->pl = s->pcl_buf.front();
    s->pcl_buf.pop_front();
    s->pcl_beg_time = s->time_buf.front();
    s->time_buf.pop_front();
    s->pcl_end_time = s->pcl_beg_time + s->pl.first;
    if (s->point_notime)
    {
      if (s->last_pcl_time < 0)
      {
        s->last_pcl_time = s->pcl_beg_time;
        *tag = s->pl.second;
        return 2;
      }
      s->pcl_end_time = s->pcl_beg_time;
      s->pcl_beg_time = s->last_pcl_time;
      s->last_pcl_time = s->pcl_end_time;
    }
    s->pl_ready = true;
  }
  if (!s->pl_ready || s->imu_last_time <= s->pcl_end_time) return 0;
  std::vector<ImuSample> imus;
  double imu_time = s->imu_buf.front().t;
  while ((!s->imu_buf.empty()) && (imu_time < s->pcl_end_time))
  {
    imu_time = s->imu_buf.front().t;
    if (imu_time > s->pcl_end_time) break;
    imus.push_back(s->imu_buf.front());
    s->imu_buf.pop_front();
  }
  *tag = s->pl.second;
  *beg = s->pcl_beg_time;
  *end = s->pcl_end_time;
  s->pl_ready = false;
  if (s->imu_buf.empty()) return -6;
  if ((int)imus.size() > cap) return -3;
  for (size_t i = 0; i < imus.size(); i++)
  {
    imu7[7 * i] = imus[i].t;
    for (int k = 0; k < 3; k++) imu7[7 * i + 1 + k] = imus[i].gyr[k], imu7[7 * i + 4 + k] = imus[i].acc[k];
  }
  *m = (int)imus.size();
  return imus.size() > 4 ? 1 : 2;
}

void* vo_odom_create(const vo_config* cfg)
{
  Globals g;
  g.voxel_size = cfg->voxel_size;
  g.min_eigen_value = cfg->min_eigen_value;
  for (int i = 0; i < 4; i++)
  {
    g.plane_eigen_value_thre[i] = 1.0 / cfg->plane_eigen_value_thre[i];  // node.cpp:256-259
    g.min_point[i] = cfg->min_point[i];
  }
  g.max_layer = cfg->max_layer;
  g.max_points = cfg->max_points;
  g.win_size = cfg->win_size;
  g.thread_num = cfg->thread_num;
  g.dept_err = cfg->dept_err;
  g.beam_err = cfg->beam_err;
  g.down_size = cfg->down_size;
  Odom* o = new Odom(g);
  o->extrin_para.R = m3(cfg->ext_R);
  o->extrin_para.p = v3(cfg->ext_t);
  o->odom_ekf.Lid_rot_to_IMU = m3(cfg->ext_R);
  o->odom_ekf.Lid_offset_to_IMU = v3(cfg->ext_t);
  o->odom_ekf.cov_gyr = V3(cfg->cov_gyr, cfg->cov_gyr, cfg->cov_gyr);
  o->odom_ekf.cov_acc = V3(cfg->cov_acc, cfg->cov_acc, cfg->cov_acc);
  o->odom_ekf.cov_bias_gyr = V3(cfg->rdw_gyr, cfg->rdw_gyr, cfg->rdw_gyr);
  o->odom_ekf.cov_bias_acc = V3(cfg->rdw_acc, cfg->rdw_acc, cfg->rdw_acc);
  for (int k = 0; k < 3; k++)  // node.cpp:262-265
  {
    o->ba_noise.noiseMeas(k, k) = cfg->cov_gyr;
    o->ba_noise.noiseMeas(3 + k, 3 + k) = cfg->cov_acc;
    o->ba_noise.noiseWalk(k, k) = cfg->rdw_gyr;
    o->ba_noise.noiseWalk(3 + k, 3 + k) = cfg->rdw_acc;
  }
  return o;
}
void vo_odom_destroy(void* h) { delete (Odom*)h; }
void vo_odom_set_state(void* h, const vo_state* s) { to_imust(s, ((Odom*)h)->x_curr); }
void vo_odom_get_state(void* h, vo_state* s) { from_imust(((Odom*)h)->x_curr, s); }
void vo_odom_set_imu_anchor(void* h, double last_pcl_end_time, const double last_imu7[7], double scale_gravity)
{
  Odom* o = (Odom*)h;
  o->odom_ekf.last_pcl_end_time = last_pcl_end_time;
  o->odom_ekf.last_imu.t = last_imu7[0];
  for (int k = 0; k < 3; k++)
  {
    o->odom_ekf.last_imu.gyr[k] = last_imu7[1 + k];
    o->odom_ekf.last_imu.acc[k] = last_imu7[4 + k];
  }
  o->odom_ekf.scale_gravity = scale_gravity;
  o->ba_noise.scale_gravity = scale_gravity;  // node.cpp:309
}
void vo_odom_bootstrap(void* h, const float* xyz4, int n, const vo_state* x_known)
{
  Odom* o = (Odom*)h;
  Cloud c = to_cloud(xyz4, n);
  IMUST x;
  to_imust(x_known, x);
  o->bootstrap(c, x);
}
int vo_odom_step(void* h, float* xyz4, int n, double pcl_beg_time, const double* imu7, int m, int iekf_on_full,
                 int max_iter)
{
  Odom* o = (Odom*)h;
  Cloud c = to_cloud(xyz4, n);
  std::deque<ImuSample> imus = to_imus(imu7, m);
  int r = o->step(c, pcl_beg_time, imus, iekf_on_full != 0, max_iter);
  from_cloud(c, xyz4);
  return r;
}
// the start-up phase (VINA_SLAM::initialization, node.cpp:293-366): switch a freshly created odometry to a cold start,
// then feed scans until 1 comes back
void vo_odom_cold_start(void* h)
{
  Odom* o = (Odom*)h;
  o->odom_ekf.init_flag = false;
  o->odom_ekf.init_num = 0;
  o->odom_ekf.mean_acc.setZero();
  o->odom_ekf.mean_gyr.setZero();
  o->x_curr.setZero();
  o->pl_tree.clear();
}
int vo_odom_init_scan(void* h, const float* xyz4, int n, double beg, const double* imu7, int m)
{
  Odom* o = (Odom*)h;
  Cloud c = to_cloud(xyz4, n);
  std::deque<ImuSample> imus = to_imus(imu7, m);
  return o->init_scan(c, beg, imus);
}
void vo_odom_stage_times(void* h, double t[4])
{
  Odom* o = (Odom*)h;
  t[0] = o->t_odom;
  t[1] = o->t_insert;
  t[2] = o->t_recut;
  t[3] = o->t_margi;
}
int vo_odom_last_iters(void* h) { return ((Odom*)h)->last_iters; }
int vo_odom_last_down(void* h, float* xyz4, int cap)
{
  Odom* o = (Odom*)h;
  int n = (int)o->last_down.size();
  if (xyz4 && cap >= n) from_cloud(o->last_down, xyz4);
  return n;
}

int vo_odom_propagate(void* h, double pcl_beg_time, double pcl_end_time, const double* imu7, int m)
{
  Odom* o = (Odom*)h;
  o->odom_ekf.pcl_beg_time = pcl_beg_time;
  o->odom_ekf.pcl_end_time = pcl_end_time;
  std::deque<ImuSample> imus = to_imus(imu7, m);
  return o->odom_ekf.propagate(o->x_curr, imus);
}
int vo_odom_imu_poses(void* h, double* poses22, int cap)
{
  Odom* o = (Odom*)h;
  int n = (int)o->odom_ekf.imu_poses.size();
  if (!poses22 || cap < n) return n;
  for (int i = 0; i < n; i++)
  {
    IMUST& s = o->odom_ekf.imu_poses[i];
    double* q = poses22 + 22 * (size_t)i;
    q[0] = s.t;
    memcpy(q + 1, s.R.d, 72);
    memcpy(q + 10, s.p.d, 24);
    memcpy(q + 13, s.v.d, 24);
    memcpy(q + 16, s.bg.d, 24);  // angvel_avr
    memcpy(q + 19, s.ba.d, 24);  // acc_imu
  }
  return n;
}
void vo_odom_deskew(void* h, float* xyz4, int n)
{
  Odom* o = (Odom*)h;
  Cloud c = to_cloud(xyz4, n);
  o->odom_ekf.deskew(o->x_curr, c);
  from_cloud(c, xyz4);
}
int vo_odom_motion_blur(void* h, float* xyz4, int n, double beg, double end, const double* imu7, int m)
{
  Odom* o = (Odom*)h;
  Cloud c = to_cloud(xyz4, n);
  std::deque<ImuSample> imus = to_imus(imu7, m);
  o->odom_ekf.pcl_beg_time = beg;
  o->odom_ekf.pcl_end_time = end;
  int r = o->odom_ekf.motion_blur(o->x_curr, c, imus);
  from_cloud(c, xyz4);
  return r;
}
int vo_odom_match(void* h, int n, const double* wld, const double* var, uint8_t* flags, double* sigma, double* centers)
{
  Odom* o = (Odom*)h;
  int cnt = 0;
  for (int i = 0; i < n; i++)
  {
    Vec3 w = v3(wld + 3 * (size_t)i);
    Mat3 v = m3(var + 9 * (size_t)i);
    Plane* pla = nullptr;
    double sd = 0;
    OctoTree* oc = nullptr;
    int f = match(&o->G, o->surf_map, w, pla, v, sd, oc);
    flags[i] = f ? 1 : 0;
    sigma[i] = f ? sd : 0.0;
    for (int k = 0; k < 3; k++) centers[3 * (size_t)i + k] = f ? pla->center[k] : 0.0;
    cnt += f ? 1 : 0;
  }
  return cnt;
}
void vo_odom_set_dump(void* h, int on) { ((Odom*)h)->dump_iters = on != 0; }
int vo_odom_iekf(void* h, int n, const double* pnt, const double* var, int max_iter)
{
  Odom* o = (Odom*)h;
  PVecPtr pptr(new PVec(n));
  for (int i = 0; i < n; i++)
  {
    (*pptr)[i].pnt = v3(pnt + 3 * (size_t)i);
    (*pptr)[i].var = m3(var + 9 * (size_t)i);
  }
  return o->LioStateEstimation(pptr, max_iter) ? 1 : 0;
}
int vo_odom_iter_dump(void* h, int it, double HTH[36], double HTz[6], double nnt[9], int32_t* match_num,
                      int64_t* keys, int32_t* codes, uint8_t* flags, double* sigma, double R[9], double p[3])
{
  Odom* o = (Odom*)h;
  if (it < 0 || it >= (int)o->iter_dumps.size()) return -1;
  IekfIterDump& d = o->iter_dumps[it];
  memcpy(HTH, d.HTH, sizeof(d.HTH));
  memcpy(HTz, d.HTz, sizeof(d.HTz));
  memcpy(nnt, d.nnt, sizeof(d.nnt));
  *match_num = d.match_num;
  if (keys) memcpy(keys, d.keys.data(), d.keys.size() * sizeof(int64_t));
  if (codes) memcpy(codes, d.codes.data(), d.codes.size() * sizeof(int32_t));
  if (flags) memcpy(flags, d.flags.data(), d.flags.size());
  if (sigma) memcpy(sigma, d.sigma.data(), d.sigma.size() * sizeof(double));
  memcpy(R, d.R, sizeof(d.R));
  memcpy(p, d.p, sizeof(d.p));
  return (int)d.flags.size();
}
void vo_odom_map_update(void* h, int n, const double* pnt, const double* var)
{
  Odom* o = (Odom*)h;
  PVecPtr pptr(new PVec(n));
  for (int i = 0; i < n; i++)
  {
    (*pptr)[i].pnt = v3(pnt + 3 * (size_t)i);
    (*pptr)[i].var = m3(var + 9 * (size_t)i);
  }
  o->pwld.clear();
  pvec_update(pptr, o->x_curr, o->pwld);
  o->map_update(pptr);
}

static void count_nodes(OctoTree* n, int64_t& c)
{
  c++;
  for (int i = 0; i < 8; i++)
    if (n->leaves[i]) count_nodes(n->leaves[i], c);
}
int64_t vo_odom_map_count(void* h, int64_t* n_roots, int64_t* n_slide)
{
  Odom* o = (Odom*)h;
  int64_t c = 0;
  for (auto& kv : o->surf_map) count_nodes(kv.second, c);
  if (n_roots) *n_roots = (int64_t)o->surf_map.size();
  if (n_slide) *n_slide = (int64_t)o->surf_map_slide.size();
  return c;
}
static void export_node(Odom* o, OctoTree* n, vo_node_record* out, int64_t cap, int64_t& c)
{
  if (c < cap)
  {
    vo_node_record& r = out[c];
    memset(&r, 0, sizeof(r));
    r.key[0] = n->root_key.x;
    r.key[1] = n->root_key.y;
    r.key[2] = n->root_key.z;
    r.code = n->code();
    r.layer = n->layer;
    r.octo_state = n->octo_state;
    r.isexist = n->isexist;
    r.has_sw = n->sw != nullptr;
    r.is_plane = n->plane.is_plane;
    r.last_num = n->last_num;
    r.opt_state = n->opt_state >= 0 ? 1 : 0;
    r.N_add = n->pcr_add.N;
    r.N_fix = n->pcr_fix.N;
    r.n_point_fix = (int)n->point_fix.size();
    if (n->sw)
    {
      for (int i = 0; i < o->G.win_size && i < 16; i++)
      {
        r.N_local[i] = n->sw->pcrs_local[o->G.mp[i]].N;
        r.n_win_points += (int)n->sw->points[o->G.mp[i]].size();
      }
    }
    memcpy(r.P_add, n->pcr_add.P.d, 72);
    memcpy(r.v_add, n->pcr_add.v.d, 24);
    memcpy(r.P_fix, n->pcr_fix.P.d, 72);
    memcpy(r.v_fix, n->pcr_fix.v.d, 24);
    memcpy(r.eig_value, n->eig_value.d, 24);
    memcpy(r.eig_vector, n->eig_vector.d, 72);
    memcpy(r.center, n->plane.center.d, 24);
    memcpy(r.normal, n->plane.normal.d, 24);
    memcpy(r.plane_var, n->plane.plane_var.d, 288);
    r.radius = n->plane.radius;
    memcpy(r.cov_add, n->cov_add.d, 648);
    memcpy(r.voxel_center, n->voxel_center, 24);
    r.quater_length = n->quater_length;
  }
  c++;
  for (int i = 0; i < 8; i++)
    if (n->leaves[i]) export_node(o, n->leaves[i], out, cap, c);
}
int64_t vo_odom_map_export(void* h, vo_node_record* out, int64_t cap)
{
  Odom* o = (Odom*)h;
  int64_t c = 0;
  for (auto& kv : o->surf_map) export_node(o, kv.second, out, cap, c);
  return c;
}
// ---- BA probe (oracle_capi.h): the restated LidarFactor on the factors captured by the last map update
void vo_odom_ba_probe(void* h, int on) { ((Odom*)h)->ba_probe = on != 0; }
int vo_odom_ba_count(void* h) { return (int)((Odom*)h)->ba_factors.plvec_voxels.size(); }
int vo_odom_ba_poses(void* h, double* poses12, int cap)
{
  Odom* o = (Odom*)h;
  const int n = (int)o->ba_xs.size();
  for (int i = 0; i < n && i < cap; i++)
  {
    memcpy(poses12 + 12 * i, o->ba_xs[i].R.d, 72);
    memcpy(poses12 + 12 * i + 9, o->ba_xs[i].p.d, 24);
  }
  return n;
}
static std::vector<IMUST> ba_pose_vec(const double* poses12, int win)
{
  std::vector<IMUST> xs(win);
  for (int i = 0; i < win; i++)
  {
    memcpy(xs[i].R.d, poses12 + 12 * i, 72);
    memcpy(xs[i].p.d, poses12 + 12 * i + 9, 24);
  }
  return xs;
}
int vo_odom_ba_hess(void* h, const double* poses12, int win, double* Hess, double* JacT, double* residual)
{
  Odom* o = (Odom*)h;
  if (win != o->ba_factors.win_size) return -1;
  std::vector<IMUST> xs = ba_pose_vec(poses12, win);
  std::vector<double> H, J;
  double r = 0;
  o->ba_factors.acc_evaluate2(xs, 0, (int)o->ba_factors.plvec_voxels.size(), H, J, r);
  memcpy(Hess, H.data(), H.size() * sizeof(double));
  memcpy(JacT, J.data(), J.size() * sizeof(double));
  *residual = r;
  return 0;
}
int vo_odom_ba_residual(void* h, const double* poses12, int win, double* residual, double* lam0, int cap)
{
  Odom* o = (Odom*)h;
  if (win != o->ba_factors.win_size) return -1;
  std::vector<IMUST> xs = ba_pose_vec(poses12, win);
  double r = 0;
  const int n = (int)o->ba_factors.plvec_voxels.size();
  o->ba_factors.evaluate_only_residual(xs, 0, n, r);
  *residual = r;
  if (lam0)
    for (int a = 0; a < n && a < cap; a++) lam0[a] = o->ba_factors.eig_values[a][0];
  return 0;
}
// one IMU_PRE built from m samples and evaluated between two states (imu_preintegration.cpp:32-163)
double vo_ba_imu_evaluate(const vo_config* cfg, const double bg[3], const double ba[3], const double* imu7, int m,
                          double scale_gravity, const vo_state* s1, const vo_state* s2, double* jtj, double* gg)
{
  BaNoise nz;
  for (int k = 0; k < 3; k++)
  {
    nz.noiseMeas(k, k) = cfg->cov_gyr;
    nz.noiseMeas(3 + k, 3 + k) = cfg->cov_acc;
    nz.noiseWalk(k, k) = cfg->rdw_gyr;
    nz.noiseWalk(3 + k, 3 + k) = cfg->rdw_acc;
  }
  nz.scale_gravity = scale_gravity;
  std::deque<ImuSample> buf;
  for (int i = 0; i < m; i++)
  {
    ImuSample s;
    s.t = imu7[7 * i];
    for (int k = 0; k < 3; k++)
    {
      s.gyr[k] = imu7[7 * i + 1 + k];
      s.acc[k] = imu7[7 * i + 4 + k];
    }
    buf.push_back(s);
  }
  IMU_PRE f(v3(bg), v3(ba));
  f.push_imu(buf, nz);
  IMUST a, b;
  to_imust(s1, a);
  to_imust(s2, b);
  Mat<30, 30> J;
  Mat<30, 1> g;
  J.setZero();
  g.setZero();
  double r = f.give_evaluate(a, b, J, g, jtj != nullptr && gg != nullptr);
  if (jtj && gg)
  {
    memcpy(jtj, J.d, sizeof(J.d));
    memcpy(gg, g.d, sizeof(g.d));
  }
  return r;
}
void vo_odom_set_ba(void* h, int on, double imu_coef)
{
  Odom* o = (Odom*)h;
  o->if_BA = on != 0;
  if (imu_coef > 0) o->imu_coef = imu_coef;
}
void vo_odom_ba_stats(void* h, int* runs, int* last_iters)
{
  Odom* o = (Odom*)h;
  *runs = o->ba_runs;
  *last_iters = o->ba_last_iters;
}
int vo_odom_window(void* h, int* win_count, int* mp, int cap)
{
  Odom* o = (Odom*)h;
  *win_count = o->win_count;
  for (int i = 0; i < o->G.win_size && i < cap; i++) mp[i] = o->G.mp[i];
  return o->G.win_size;
}
void vo_odom_journey(void* h, double* jour, int* release_flag)
{
  Odom* o = (Odom*)h;
  if (jour) *jour = o->jour;
  if (release_flag) *release_flag = o->release_flag ? 1 : 0;
}
int vo_odom_idle(void* h, int horizon, int* nodes_freed) { return ((Odom*)h)->idle_release(horizon, nodes_freed); }
}

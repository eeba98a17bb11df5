// ORACLE — TEST INFRASTRUCTURE ONLY.
// Harness around the REFERENCE'S OWN SOURCES. The Makefile target `_ref` compiles, unmodified and where
// they lie under /root/reference,
//     src/core/point_utils.cpp   src/mapping/octree.cpp   src/mapping/voxel_map.cpp
//     src/estimation/imu_ekf.cpp src/pipeline/odometry.cpp   src/mapping/factors.cpp
// against the header shims in oracle/ref_shim/ (a minimal eager Eigen, the PCL point/cloud types and the
// ROS message fields those files touch - none of Eigen / PCL / ROS 2 is installed here), and links them
// with this file into oracle/_ref/libvina_ref.so. This file supplies only what lives in reference files that
// cannot be compiled here (they need the whole ROS 2 node):
//   * the two globals of node.cpp:38 and the VINA_SLAM constructor (node.cpp:52) reduced to nothing,
//   * the fan-out drivers multi_recut / multi_margi (local_mapping.cpp:17-84, 144-201) and the scan body
//     of thd_odometry_localmapping (local_mapping.cpp:389-546) - loops that call the reference's real
//     OctoTree::recut / tras_opt / margi, cut_voxel_multi, var_init, pvec_update, IMUEKF::process and
//     VINA_SLAM::VNC_lio.
// It exports the same plain-C API as oracle_capi.cpp, so the Python tests drive the restatement and the
// reference with identical code and compare them (tests/test_oracle_vs_ref.py).
// Because the reference keeps its configuration in mutable globals (octree.cpp:67-75) only ONE instance
// may exist per process.
#include "vina_slam/platform/ros2/node.hpp"
#include "vina_slam/core/point_utils.hpp"
#include "vina_slam/mapping/voxel_map.hpp"
#include "vina_slam/mapping/optimizers.hpp"
#include "vina_slam/preintegration.hpp"
#include "vina_slam/sensor/sync.hpp"
#include "vina_slam/sensor/lidar_decoder.hpp"
#include "vina_slam/pipeline/initialization.hpp"

#include <chrono>
#include <cstring>
#include <thread>

#include "oracle_capi.h"

double dept_err, beam_err;  // node.cpp:38

VINA_SLAM::VINA_SLAM(const rclcpp::Node::SharedPtr& node_in) : node(node_in) {}

namespace
{
double now_s()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct RefOdom
{
  VINA_SLAM vs;
  LidarFactor voxhess;
  NormalFactor normalFactor;
  PLV(3) pwld;
  int degrade_cnt = 0;
  double t_odom = 0, t_insert = 0, t_recut = 0, t_margi = 0;
  pcl::PointCloud<PointType> last_down;
  // BA probe: a copy of the LiDAR factors and the window poses as they are between multi_recut and multi_margi
  // (where LI_BA_Optimizer::damping_iter consumes them, local_mapping.cpp:492-497)
  bool ba_probe = false;
  LidarFactor ba_factors;
  vector<IMUST> ba_xs;
  // sliding-window BA (local_mapping.cpp:437-441, 492-497, 541-546); frames inserted by bootstrap carry no
  // pre-integration factor (nullptr): BA runs once every pair of consecutive window frames has one
  bool if_BA = false;
  deque<IMU_PRE*> imu_pre_buf;
  int ba_runs = 0, ba_last_iters = 0;
  // locals of thd_odometry_localmapping (local_mapping.cpp:262-263, 272)
  Eigen::Vector3d last_pos = Eigen::Vector3d(0, 0, 0);
  double jour = 0.0;
  bool release_flag = false;
  RefOdom(int win) : vs(nullptr), voxhess(win), normalFactor(win), ba_factors(win) {}

  // local_mapping.cpp:144-201 (the overload the per-scan loop calls, :451)
  void multi_recut(unordered_map<VOXEL_LOC, OctoTree*>& feat_map, int win_count, vector<IMUST>& xs)
  {
    auto& sws = vs.sws;
    int thd_num = vs.thread_num;
    vector<vector<OctoTree*>> octss(thd_num);
    int g_size = feat_map.size();
    if (g_size < thd_num) return;
    vector<thread*> mthreads(thd_num);
    double part = 1.0 * g_size / thd_num;
    int cnt = 0;
    for (auto iter = feat_map.begin(); iter != feat_map.end(); iter++)
    {
      octss[cnt].push_back(iter->second);
      if (octss[cnt].size() >= part && cnt < thd_num - 1) cnt++;
    }
    auto recut_func = [](int win_count, vector<OctoTree*>& oct, vector<IMUST> xxs, vector<SlideWindow*>& sw) {
      for (OctoTree* oc : oct) oc->recut(win_count, xxs, sw);
    };
    for (int i = 1; i < thd_num; i++)
      mthreads[i] = new thread(recut_func, win_count, ref(octss[i]), xs, ref(sws[i]));
    for (int i = 0; i < thd_num; i++)
    {
      if (i == 0)
        recut_func(win_count, octss[i], xs, sws[i]);
      else
      {
        mthreads[i]->join();
        delete mthreads[i];
      }
    }
    for (size_t i = 1; i < sws.size(); i++)
    {
      sws[0].insert(sws[0].end(), sws[i].begin(), sws[i].end());
      sws[i].clear();
    }
    for (auto iter = feat_map.begin(); iter != feat_map.end(); iter++)
    {
      iter->second->tras_opt(voxhess);
      iter->second->tras_opt(normalFactor);
    }
  }

  // local_mapping.cpp:17-84
  void multi_margi(unordered_map<VOXEL_LOC, OctoTree*>& feat_map, double jour, int win_count, vector<IMUST>& xs,
                   vector<SlideWindow*>& sw)
  {
    int thd_num = vs.thread_num;
    vector<vector<OctoTree*>*> octs;
    for (int i = 0; i < thd_num; i++) octs.push_back(new vector<OctoTree*>());
    int g_size = feat_map.size();
    if (g_size < thd_num) return;
    vector<thread*> mthreads(thd_num);
    double part = 1.0 * g_size / thd_num;
    int cnt = 0;
    for (auto iter = feat_map.begin(); iter != feat_map.end(); iter++)
    {
      iter->second->jour = jour;
      octs[cnt]->push_back(iter->second);
      if (octs[cnt]->size() >= part && cnt < thd_num - 1) cnt++;
    }
    auto margi_func = [](int win_cnt, vector<OctoTree*>* oct, vector<IMUST> xxs, LidarFactor& vh) {
      for (OctoTree* oc : *oct) oc->margi(win_cnt, 1, xxs, vh);
    };
    for (int i = 1; i < thd_num; i++) mthreads[i] = new thread(margi_func, win_count, octs[i], xs, ref(voxhess));
    for (int i = 0; i < thd_num; i++)
    {
      if (i == 0)
        margi_func(win_count, octs[i], xs, voxhess);
      else
      {
        mthreads[i]->join();
        delete mthreads[i];
      }
    }
    for (auto iter = feat_map.begin(); iter != feat_map.end();)
    {
      if (iter->second->isexist)
        iter++;
      else
      {
        iter->second->clear_slwd(sw);
        feat_map.erase(iter++);
      }
    }
    for (int i = 0; i < thd_num; i++) delete octs[i];
  }

  // local_mapping.cpp:434-451 and 489-546 with if_BA == 0
  void map_update(PVecPtr pptr, deque<std::shared_ptr<sensor_msgs::msg::Imu>>* imus = nullptr)
  {
    const int mgsize = 1;
    vs.win_count++;
    vs.x_buf.push_back(vs.x_curr);
    vs.pvec_buf.push_back(pptr);
    if (vs.win_count > 1)
    {
      IMU_PRE* f = nullptr;
      if (imus)
      {
        f = new IMU_PRE(vs.x_buf[vs.win_count - 2].bg, vs.x_buf[vs.win_count - 2].ba);
        f->push_imu(*imus);
      }
      imu_pre_buf.push_back(f);
    }
    voxhess.clear();
    voxhess.win_size = vs.win_size;
    normalFactor.clear();
    normalFactor.win_size = vs.win_size;
    double t1 = now_s();
    cut_voxel_multi(vs.surf_map, vs.pvec_buf[vs.win_count - 1], vs.win_count - 1, vs.surf_map_slide, vs.win_size, pwld,
                    vs.sws);
    double t2 = now_s();
    multi_recut(vs.surf_map_slide, vs.win_count, vs.x_buf);
    double t3 = now_s();
    t_insert = t2 - t1;
    t_recut = t3 - t2;
    t_margi = 0;
    if (ba_probe && vs.win_count >= vs.win_size)
    {
      ba_factors = voxhess;
      ba_xs = vs.x_buf;
    }
    window_tail();
  }

  // local_mapping.cpp:489-546: BA, margi, journey, window shift - once the window is full
  void window_tail()
  {
    const int mgsize = 1;
    if (vs.win_count >= vs.win_size)
    {
      bool all_imu = (int)imu_pre_buf.size() == vs.win_count - 1;
      for (IMU_PRE* f : imu_pre_buf) all_imu = all_imu && f != nullptr;
      if (if_BA && all_imu)
      {
        LI_BA_Optimizer opt_lsv;  // local_mapping.cpp:492-497
        Eigen::MatrixXd hess;
        opt_lsv.damping_iter(vs.x_buf, voxhess, imu_pre_buf, &hess);
        ba_runs++;
      }
      vs.x_curr.R = vs.x_buf[vs.win_count - 1].R;
      vs.x_curr.p = vs.x_buf[vs.win_count - 1].p;
      double t5 = now_s();
      multi_margi(vs.surf_map_slide, jour, vs.win_count, vs.x_buf, vs.sws[0]);
      t_margi = now_s() - t5;
      if ((vs.win_base + vs.win_count) % 10 == 0)  // local_mapping.cpp:509-519
      {
        double spat = (vs.x_curr.p - last_pos).norm();
        if (spat > 0.5)
        {
          jour += spat;
          last_pos = vs.x_curr.p;
          release_flag = true;
        }
      }
      for (int i = 0; i < vs.win_size; i++)
      {
        mp[i] += mgsize;
        if (mp[i] >= vs.win_size) mp[i] -= vs.win_size;
      }
      for (int i = mgsize; i < vs.win_count; i++)
      {
        vs.x_buf[i - mgsize] = vs.x_buf[i];
        PVecPtr pvec_tem = vs.pvec_buf[i - mgsize];
        vs.pvec_buf[i - mgsize] = vs.pvec_buf[i];
        vs.pvec_buf[i] = pvec_tem;
      }
      for (int i = vs.win_count - mgsize; i < vs.win_count; i++)
      {
        vs.x_buf.pop_back();
        vs.pvec_buf.pop_back();
        delete imu_pre_buf.front();
        imu_pre_buf.pop_front();
      }
      vs.win_base += mgsize;
      vs.win_count -= mgsize;
    }
  }

  // ---- VINA_SLAM::initialization (node.cpp:293-366), system_reset (node.cpp:368-408) and the loop's handling of
  // the result (local_mapping.cpp:362-388, then :489-546 on success) - restated here because node.cpp needs the
  // whole ROS node; everything they call is the reference's own code: IMUEKF::process / IMU_init, down_sampling_voxel,
  // down_sampling_close, var_init, VINA_SLAM::lio_state_estimation_kdtree (odometry.cpp:267-439), pvec_update,
  // IMU_PRE and Initialization::motion_init (initialization.cpp:158-367: re-deskew, cut_voxel, recut, tras_opt,
  // LI_BA_OptimizerGravity::damping_iter, align_gravity).
  vector<pcl::PointCloud<PointType>::Ptr> pl_origs;
  vector<double> beg_times;
  vector<deque<std::shared_ptr<sensor_msgs::msg::Imu>>> vec_imus;
  Eigen::MatrixXd init_hess;
  int motion_init_rounds = 0;

  int initialization(deque<std::shared_ptr<sensor_msgs::msg::Imu>>& imus, pcl::PointCloud<PointType>::Ptr pcl_curr)
  {
    pcl::PointCloud<PointType>::Ptr orig(new pcl::PointCloud<PointType>(*pcl_curr));
    if (vs.odom_ekf.process(vs.x_curr, *pcl_curr, imus) == 0) return 0;
    if (vs.win_count == 0) imupre_scale_gravity = vs.odom_ekf.scale_gravity;
    PVecPtr pptr(new PVec);
    double downkd = vs.down_size >= 0.5 ? vs.down_size : 0.5;
    down_sampling_voxel(*pcl_curr, downkd);
    var_init(vs.extrin_para, *pcl_curr, pptr, dept_err, beam_err);
    vs.lio_state_estimation_kdtree(pptr);
    pwld.clear();
    pvec_update(pptr, vs.x_curr, pwld);
    vs.win_count++;
    vs.x_buf.push_back(vs.x_curr);
    vs.pvec_buf.push_back(pptr);
    if (vs.win_count > 1)
    {
      imu_pre_buf.push_back(new IMU_PRE(vs.x_buf[vs.win_count - 2].bg, vs.x_buf[vs.win_count - 2].ba));
      imu_pre_buf[vs.win_count - 2]->push_imu(imus);
    }
    pcl::PointCloud<PointType> pl_mid = *orig;
    down_sampling_close(*orig, vs.down_size);
    if (orig->size() < 1000)
    {
      *orig = pl_mid;
      down_sampling_close(*orig, vs.down_size / 2);
    }
    sort(orig->begin(), orig->end(), [](PointType& x, PointType& y) { return x.curvature < y.curvature; });
    pl_origs.push_back(orig);
    beg_times.push_back(vs.odom_ekf.pcl_beg_time);
    vec_imus.push_back(imus);
    if (vs.win_count >= vs.win_size)
    {
      int ok = Initialization::instance(init_node).motion_init(pl_origs, vec_imus, beg_times, &init_hess, voxhess, vs.x_buf,
                                                               vs.surf_map, vs.surf_map_slide, vs.pvec_buf, vs.win_size,
                                                               vs.sws, vs.x_curr, imu_pre_buf, vs.extrin_para);
      return ok == 0 ? -1 : 1;
    }
    return 0;
  }
  rclcpp::Node::SharedPtr init_node = std::make_shared<rclcpp::Node>();

  void system_reset(deque<std::shared_ptr<sensor_msgs::msg::Imu>>& imus)
  {
    for (auto iter = vs.surf_map.begin(); iter != vs.surf_map.end(); iter++)
    {
      iter->second->tras_ptr(vs.octos_release);
      iter->second->clear_slwd(vs.sws[0]);
      delete iter->second;
    }
    for (OctoTree* ot : vs.octos_release) delete ot;  // (the idle path of the loop deletes these, local_mapping.cpp:552-571)
    vs.octos_release.clear();
    vs.surf_map.clear();
    vs.surf_map_slide.clear();
    vs.x_curr.setZero();
    vs.x_curr.p = Eigen::Vector3d(0, 0, 30);
    vs.odom_ekf.mean_acc.setZero();
    vs.odom_ekf.init_num = 0;
    vs.odom_ekf.IMU_init(imus);
    vs.x_curr.g = -vs.odom_ekf.mean_acc * imupre_scale_gravity;
    for (size_t i = 0; i < imu_pre_buf.size(); i++) delete imu_pre_buf[i];
    vs.x_buf.clear();
    vs.pvec_buf.clear();
    imu_pre_buf.clear();
    vs.pl_tree->clear();
    for (int i = 0; i < vs.win_size; i++) mp[i] = i;
    vs.win_base = 0;
    vs.win_count = 0;
  }

  // one scan of the start-up phase: 0 = still collecting, 1 = initialised (the window is full and has been
  // marginalised once: the next scan goes to step()), -1 = motion_init failed and the system was reset
  int init_scan(pcl::PointCloud<PointType>::Ptr pcl_curr, double beg, deque<std::shared_ptr<sensor_msgs::msg::Imu>>& imus)
  {
    vs.odom_ekf.pcl_beg_time = beg;
    vs.odom_ekf.pcl_end_time = beg + pcl_curr->back().curvature;  // sync.cpp:40
    int init = initialization(imus, pcl_curr);
    if (init == 1)
    {
      window_tail();
      return 1;
    }
    if (init == -1) system_reset(imus);
    return init;
  }

  // local_mapping.cpp:317-341, the `else if (release_flag)` branch of the idle path, with the 700 as a
  // parameter; OctoTree::tras_ptr and the deletes are the reference's own. Roots still in surf_map_slide are
  // kept (they would dangle there; unreachable with the reference's 700 m, see vina_oracle.cpp).
  int idle_release(int horizon, int* nodes_freed)
  {
    if (nodes_freed) *nodes_freed = 0;
    if (!release_flag) return 0;
    release_flag = false;
    auto& surf_map = vs.surf_map;
    vector<OctoTree*> octos;
    int roots = 0;
    for (auto iter = surf_map.begin(); iter != surf_map.end();)
    {
      int dis = jour - iter->second->jour;
      if (dis < horizon || vs.surf_map_slide.count(iter->first))
      {
        iter++;
      }
      else
      {
        octos.push_back(iter->second);
        iter->second->tras_ptr(octos);
        surf_map.erase(iter++);
        roots++;
      }
    }
    int ocsize = octos.size();
    if (nodes_freed) *nodes_freed = ocsize;
    for (int i = 0; i < ocsize; i++)
    {
      delete octos[i];
    }
    return roots;
  }

  void downsample(pcl::PointCloud<PointType>& pcl_curr, pcl::PointCloud<PointType>& pl_down)
  {
    pl_down = pcl_curr;  // local_mapping.cpp:396-403
    down_sampling_voxel(pl_down, vs.down_size);
    if (pl_down.size() < 2000)
    {
      pl_down = pcl_curr;
      down_sampling_voxel(pl_down, vs.down_size / 2);
    }
    last_down = pl_down;
  }

  // local_mapping.cpp:389-546
  int step(pcl::PointCloud<PointType>& pcl_curr, double beg, deque<std::shared_ptr<sensor_msgs::msg::Imu>>& imus,
           bool iekf_on_full)
  {
    double t0 = now_s();
    vs.odom_ekf.pcl_beg_time = beg;
    vs.odom_ekf.pcl_end_time = beg + pcl_curr.back().curvature;  // sync.cpp:40
    if (vs.odom_ekf.last_pcl_end_time - vs.odom_ekf.pcl_beg_time > 0.01) return -1;  // the reference exit(0)s
    if (vs.odom_ekf.process(vs.x_curr, pcl_curr, imus) == 0) return -2;
    pcl::PointCloud<PointType> pl_down;
    downsample(pcl_curr, pl_down);
    PVecPtr pptr(new PVec);
    var_init(vs.extrin_para, pl_down, pptr, dept_err, beam_err);
    auto pcl_curr_temp = pcl_curr;
    PVecPtr no_ds_pptr(new PVec);
    var_init(vs.extrin_para, pcl_curr_temp, no_ds_pptr, dept_err, beam_err);
    bool ok = iekf_on_full ? vs.VNC_lio(no_ds_pptr) : vs.VNC_lio(pptr);  // :413 (production) / 4-iteration budget
    if (ok)
    {
      if (degrade_cnt > 0) degrade_cnt--;
    }
    else
      degrade_cnt++;
    pwld.clear();
    pvec_update(pptr, vs.x_curr, pwld);
    t_odom = now_s() - t0;
    map_update(pptr, &imus);
    return 0;
  }

  void bootstrap(pcl::PointCloud<PointType>& pcl_deskewed, const IMUST& x_known)
  {
    vs.x_curr = x_known;
    pcl::PointCloud<PointType> pl_down;
    downsample(pcl_deskewed, pl_down);
    PVecPtr pptr(new PVec);
    var_init(vs.extrin_para, pl_down, pptr, dept_err, beam_err);
    pwld.clear();
    pvec_update(pptr, vs.x_curr, pwld);
    map_update(pptr);
  }
};

RefOdom* g_inst = nullptr;

void to_imust(const vo_state* s, IMUST& x)
{
  x.t = s->t;
  memcpy(x.R.data(), s->R, 72);
  memcpy(x.p.data(), s->p, 24);
  memcpy(x.v.data(), s->v, 24);
  memcpy(x.bg.data(), s->bg, 24);
  memcpy(x.ba.data(), s->ba, 24);
  memcpy(x.g.data(), s->g, 24);
  memcpy(x.cov.data(), s->cov, sizeof(s->cov));
}
void from_imust(const IMUST& x, vo_state* s)
{
  s->t = x.t;
  memcpy(s->R, x.R.data(), 72);
  memcpy(s->p, x.p.data(), 24);
  memcpy(s->v, x.v.data(), 24);
  memcpy(s->bg, x.bg.data(), 24);
  memcpy(s->ba, x.ba.data(), 24);
  memcpy(s->g, x.g.data(), 24);
  memcpy(s->cov, x.cov.data(), sizeof(s->cov));
}
pcl::PointCloud<PointType> to_cloud(const float* xyz4, int n)
{
  pcl::PointCloud<PointType> c;
  c.resize(n);
  for (int i = 0; i < n; i++)
  {
    c[i].x = xyz4[4 * i + 0];
    c[i].y = xyz4[4 * i + 1];
    c[i].z = xyz4[4 * i + 2];
    c[i].curvature = xyz4[4 * i + 3];
  }
  return c;
}
void from_cloud(const pcl::PointCloud<PointType>& c, float* xyz4)
{
  for (size_t i = 0; i < c.size(); i++)
  {
    xyz4[4 * i + 0] = c[i].x;
    xyz4[4 * i + 1] = c[i].y;
    xyz4[4 * i + 2] = c[i].z;
    xyz4[4 * i + 3] = c[i].curvature;
  }
}
std::shared_ptr<sensor_msgs::msg::Imu> to_imu(const double* q)
{
  auto m = std::make_shared<sensor_msgs::msg::Imu>();
  m->header.stamp = rclcpp::Time((int64_t)llround(q[0] * 1e9));
  m->angular_velocity.x = q[1];
  m->angular_velocity.y = q[2];
  m->angular_velocity.z = q[3];
  m->linear_acceleration.x = q[4];
  m->linear_acceleration.y = q[5];
  m->linear_acceleration.z = q[6];
  return m;
}
}  // namespace

extern "C" {

void vo_var_init(int n, const float* xyz4, const double ext_R[9], const double ext_t[3], double dept, double beam,
                 double* pnt, double* var)
{
  IMUST ext;
  memcpy(ext.R.data(), ext_R, 72);
  memcpy(ext.p.data(), ext_t, 24);
  auto c = to_cloud(xyz4, n);
  PVecPtr pptr(new PVec);
  var_init(ext, c, pptr, dept, beam);
  for (int i = 0; i < n; i++)
  {
    memcpy(pnt + 3 * (size_t)i, (*pptr)[i].pnt.data(), 24);
    memcpy(var + 9 * (size_t)i, (*pptr)[i].var.data(), 72);
  }
}
void vo_pvec_update(int n, const double* pnt, double* var, const double R[9], const double p[3], const double cov[225],
                    double* pw)
{
  IMUST x;
  memcpy(x.R.data(), R, 72);
  memcpy(x.p.data(), p, 24);
  memcpy(x.cov.data(), cov, 225 * 8);
  PVecPtr pptr(new PVec(n));
  for (int i = 0; i < n; i++)
  {
    memcpy((*pptr)[i].pnt.data(), pnt + 3 * (size_t)i, 24);
    memcpy((*pptr)[i].var.data(), var + 9 * (size_t)i, 72);
  }
  PLV(3) out;
  pvec_update(pptr, x, out);
  for (int i = 0; i < n; i++)
  {
    memcpy(var + 9 * (size_t)i, (*pptr)[i].var.data(), 72);
    memcpy(pw + 3 * (size_t)i, out[i].data(), 24);
  }
}
int vo_down_sampling_voxel(int n, const float* in, double vs, float* out)
{
  auto c = to_cloud(in, n);
  down_sampling_voxel(c, vs);
  from_cloud(c, out);
  return (int)c.size();
}
void vo_exp(const double w[3], double R[9])
{
  Eigen::Matrix3d r = Exp(Eigen::Vector3d(w[0], w[1], w[2]));
  memcpy(R, r.data(), 72);
}
void vo_exp_dt(const double w[3], double dt, double R[9])
{
  Eigen::Matrix3d r = Exp(Eigen::Vector3d(w[0], w[1], w[2]), dt);
  memcpy(R, r.data(), 72);
}
void vo_log(const double R[9], double w[3])
{
  Eigen::Matrix3d m;
  memcpy(m.data(), R, 72);
  Eigen::Vector3d r = Log(m);
  memcpy(w, r.data(), 24);
}

void* vo_odom_create(const vo_config* cfg)
{
  if (g_inst) return nullptr;  // the reference's globals allow one instance per process
  // the reference reports on std::cout (IMU init, motion_init); this library's copy of the stream has no business
  // writing into the test output: a stream without a buffer ignores everything it is sent
  std::cout.rdbuf(nullptr);
  RefOdom* o = new RefOdom(cfg->win_size);
  VINA_SLAM& vs = o->vs;
  voxel_size = cfg->voxel_size;
  min_eigen_value = cfg->min_eigen_value;
  plane_eigen_value_thre.assign(cfg->plane_eigen_value_thre, cfg->plane_eigen_value_thre + 4);
  for (double& it : plane_eigen_value_thre) it = 1.0 / it;  // node.cpp:256-259
  min_point << cfg->min_point[0], cfg->min_point[1], cfg->min_point[2], cfg->min_point[3];
  max_layer = cfg->max_layer;
  max_points = cfg->max_points;
  dept_err = cfg->dept_err;
  beam_err = cfg->beam_err;
  vs.down_size = cfg->down_size;
  vs.win_size = cfg->win_size;
  vs.thread_num = cfg->thread_num;
  vs.if_BA = 0;
  vs.sws.resize(cfg->thread_num);  // node.cpp:289
  mp.resize(cfg->win_size);
  for (int i = 0; i < cfg->win_size; i++) mp[i] = i;  // node.cpp:431-435
  memcpy(vs.extrin_para.R.data(), cfg->ext_R, 72);
  memcpy(vs.extrin_para.p.data(), cfg->ext_t, 24);
  IMUEKF& e = vs.odom_ekf;
  e.Lid_rot_to_IMU = vs.extrin_para.R;
  e.Lid_offset_to_IMU = vs.extrin_para.p;
  e.cov_gyr << cfg->cov_gyr, cfg->cov_gyr, cfg->cov_gyr;  // node.cpp:211-214
  e.cov_acc << cfg->cov_acc, cfg->cov_acc, cfg->cov_acc;
  e.cov_bias_gyr << cfg->rdw_gyr, cfg->rdw_gyr, cfg->rdw_gyr;
  e.cov_bias_acc << cfg->rdw_acc, cfg->rdw_acc, cfg->rdw_acc;
  noiseMeas.setZero();  // node.cpp:262-265
  noiseWalk.setZero();
  noiseMeas.diagonal() << cfg->cov_gyr, cfg->cov_gyr, cfg->cov_gyr, cfg->cov_acc, cfg->cov_acc, cfg->cov_acc;
  noiseWalk.diagonal() << cfg->rdw_gyr, cfg->rdw_gyr, cfg->rdw_gyr, cfg->rdw_acc, cfg->rdw_acc, cfg->rdw_acc;
  imupre_scale_gravity = 1.0;
  e.init_flag = true;  // the harness bootstraps instead of IMU_init
  e.pcl_beg_time = e.pcl_end_time = e.last_pcl_end_time = 0;
  vs.x_curr.g = Eigen::Vector3d(0, 0, -9.8);
  g_inst = o;
  return o;
}
void vo_odom_destroy(void* h)
{
  RefOdom* o = (RefOdom*)h;
  for (auto& kv : o->vs.surf_map)
  {
    kv.second->delete_ptr();
    delete kv.second;
  }
  for (auto& v : o->vs.sws)
    for (SlideWindow* s : v) delete s;
  delete o;
  g_inst = nullptr;
}
void vo_odom_set_state(void* h, const vo_state* s) { to_imust(s, ((RefOdom*)h)->vs.x_curr); }
void vo_odom_get_state(void* h, vo_state* s) { from_imust(((RefOdom*)h)->vs.x_curr, s); }
void vo_odom_set_imu_anchor(void* h, double last_end, const double last_imu7[7], double scale_gravity)
{
  IMUEKF& e = ((RefOdom*)h)->vs.odom_ekf;
  e.last_pcl_end_time = last_end;
  e.last_imu = to_imu(last_imu7);
  e.scale_gravity = scale_gravity;
  imupre_scale_gravity = scale_gravity;  // node.cpp:309
}
void vo_odom_bootstrap(void* h, const float* xyz4, int n, const vo_state* x_known)
{
  auto c = to_cloud(xyz4, n);
  IMUST x;
  to_imust(x_known, x);
  ((RefOdom*)h)->bootstrap(c, x);
}
int vo_odom_step(void* h, float* xyz4, int n, double beg, const double* imu7, int m, int iekf_on_full, int max_iter)
{
  (void)max_iter;  // the reference's VNC_lio budget is fixed at 4 (odometry.cpp:68)
  auto c = to_cloud(xyz4, n);
  deque<std::shared_ptr<sensor_msgs::msg::Imu>> imus;
  for (int i = 0; i < m; i++) imus.push_back(to_imu(imu7 + 7 * (size_t)i));
  int r = ((RefOdom*)h)->step(c, beg, imus, iekf_on_full != 0);
  from_cloud(c, xyz4);
  return r;
}
// one scan of the start-up phase (see RefOdom::init_scan); the context must have been switched to a cold start with
// vo_odom_cold_start (the default set-up bootstraps at known states instead)
void vo_odom_cold_start(void* h)
{
  RefOdom* o = (RefOdom*)h;
  IMUEKF& e = o->vs.odom_ekf;
  e.init_flag = false;
  e.init_num = 0;
  e.mean_acc.setZero();
  e.mean_gyr.setZero();
  o->vs.x_curr.setZero();  // (the state a freshly constructed node starts from)
  if (!o->vs.pl_tree) o->vs.pl_tree.reset(new pcl::PointCloud<PointType>());
  o->vs.pl_tree->clear();
}
int vo_odom_init_scan(void* h, const float* xyz4, int n, double beg, const double* imu7, int m)
{
  RefOdom* o = (RefOdom*)h;
  pcl::PointCloud<PointType>::Ptr c(new pcl::PointCloud<PointType>(to_cloud(xyz4, n)));
  deque<std::shared_ptr<sensor_msgs::msg::Imu>> imus;
  for (int i = 0; i < m; i++) imus.push_back(to_imu(imu7 + 7 * (size_t)i));
  return o->init_scan(c, beg, imus);
}
void vo_odom_stage_times(void* h, double t[4])
{
  RefOdom* o = (RefOdom*)h;
  t[0] = o->t_odom;
  t[1] = o->t_insert;
  t[2] = o->t_recut;
  t[3] = o->t_margi;
}
int vo_odom_last_down(void* h, float* xyz4, int cap)
{
  RefOdom* o = (RefOdom*)h;
  int n = (int)o->last_down.size();
  if (xyz4 && cap >= n) from_cloud(o->last_down, xyz4);
  return n;
}
// motion_blur in one piece (the reference does not split propagation and deskew)
int vo_odom_motion_blur(void* h, float* xyz4, int n, double beg, double end, const double* imu7, int m)
{
  RefOdom* o = (RefOdom*)h;
  auto c = to_cloud(xyz4, n);
  deque<std::shared_ptr<sensor_msgs::msg::Imu>> imus;
  for (int i = 0; i < m; i++) imus.push_back(to_imu(imu7 + 7 * (size_t)i));
  o->vs.odom_ekf.pcl_beg_time = beg;
  o->vs.odom_ekf.pcl_end_time = end;
  if (o->vs.odom_ekf.last_pcl_end_time - beg > 0.01) return -1;
  o->vs.odom_ekf.motion_blur(o->vs.x_curr, c, imus);
  from_cloud(c, xyz4);
  return 0;
}
int vo_odom_imu_poses(void* h, double* poses22, int cap)
{
  RefOdom* o = (RefOdom*)h;
  auto& ps = o->vs.odom_ekf.imu_poses;
  int n = (int)ps.size();
  if (!poses22 || cap < n) return n;
  for (int i = 0; i < n; i++)
  {
    double* q = poses22 + 22 * (size_t)i;
    q[0] = ps[i].t;
    memcpy(q + 1, ps[i].R.data(), 72);
    memcpy(q + 10, ps[i].p.data(), 24);
    memcpy(q + 13, ps[i].v.data(), 24);
    memcpy(q + 16, ps[i].bg.data(), 24);
    memcpy(q + 19, ps[i].ba.data(), 24);
  }
  return n;
}
// VINA_SLAM::VNC_lio on caller-provided pointVar arrays
int vo_odom_iekf(void* h, int n, const double* pnt, const double* var, int max_iter)
{
  (void)max_iter;
  RefOdom* o = (RefOdom*)h;
  PVecPtr pptr(new PVec(n));
  for (int i = 0; i < n; i++)
  {
    memcpy((*pptr)[i].pnt.data(), pnt + 3 * (size_t)i, 24);
    memcpy((*pptr)[i].var.data(), var + 9 * (size_t)i, 72);
  }
  return o->vs.VNC_lio(pptr) ? 1 : 0;
}
// match() (voxel_map.cpp:241-266) for world points with given world covariances
int vo_odom_match(void* h, int n, const double* wld, const double* var, uint8_t* flags, double* sigma, double* centers)
{
  RefOdom* o = (RefOdom*)h;
  int cnt = 0;
  for (int i = 0; i < n; i++)
  {
    Eigen::Vector3d w(wld[3 * (size_t)i], wld[3 * (size_t)i + 1], wld[3 * (size_t)i + 2]);
    Eigen::Matrix3d v;
    memcpy(v.data(), var + 9 * (size_t)i, 72);
    Plane* pla = nullptr;
    double sd = 0;
    OctoTree* oc = nullptr;
    int f = match(o->vs.surf_map, w, pla, v, sd, oc);
    flags[i] = f ? 1 : 0;
    sigma[i] = f ? sd : 0.0;
    for (int k = 0; k < 3; k++) centers[3 * (size_t)i + k] = f ? pla->center[k] : 0.0;
    cnt += f ? 1 : 0;
  }
  return cnt;
}
void vo_odom_map_update(void* h, int n, const double* pnt, const double* var)
{
  RefOdom* o = (RefOdom*)h;
  PVecPtr pptr(new PVec(n));
  for (int i = 0; i < n; i++)
  {
    memcpy((*pptr)[i].pnt.data(), pnt + 3 * (size_t)i, 24);
    memcpy((*pptr)[i].var.data(), var + 9 * (size_t)i, 72);
  }
  o->pwld.clear();
  pvec_update(pptr, o->vs.x_curr, o->pwld);
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
  RefOdom* o = (RefOdom*)h;
  int64_t c = 0;
  for (auto& kv : o->vs.surf_map) count_nodes(kv.second, c);
  if (n_roots) *n_roots = (int64_t)o->vs.surf_map.size();
  if (n_slide) *n_slide = (int64_t)o->vs.surf_map_slide.size();
  return c;
}
static void export_node(RefOdom* o, OctoTree* n, const VOXEL_LOC& key, int path, vo_node_record* out, int64_t cap,
                        int64_t& c)
{
  if (c < cap)
  {
    vo_node_record& r = out[c];
    memset(&r, 0, sizeof(r));
    r.key[0] = key.x;
    r.key[1] = key.y;
    r.key[2] = key.z;
    r.code = n->layer | (path << 2);
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
      for (int i = 0; i < o->vs.win_size && i < 16; i++)
      {
        r.N_local[i] = n->sw->pcrs_local[mp[i]].N;
        r.n_win_points += (int)n->sw->points[mp[i]].size();
      }
    memcpy(r.P_add, n->pcr_add.P.data(), 72);
    memcpy(r.v_add, n->pcr_add.v.data(), 24);
    memcpy(r.P_fix, n->pcr_fix.P.data(), 72);
    memcpy(r.v_fix, n->pcr_fix.v.data(), 24);
    memcpy(r.eig_value, n->eig_value.data(), 24);
    memcpy(r.eig_vector, n->eig_vector.data(), 72);
    memcpy(r.center, n->plane.center.data(), 24);
    memcpy(r.normal, n->plane.normal.data(), 24);
    memcpy(r.plane_var, n->plane.plane_var.data(), 288);
    r.radius = n->plane.radius;
    memcpy(r.cov_add, n->cov_add.data(), 648);
    memcpy(r.voxel_center, n->voxel_center, 24);
    r.quater_length = n->quater_length;
  }
  c++;
  for (int i = 0; i < 8; i++)
    if (n->leaves[i]) export_node(o, n->leaves[i], key, path | (i << (3 * n->layer)), out, cap, c);
}
int64_t vo_odom_map_export(void* h, vo_node_record* out, int64_t cap)
{
  RefOdom* o = (RefOdom*)h;
  int64_t c = 0;
  for (auto& kv : o->vs.surf_map) export_node(o, kv.second, kv.first, 0, out, cap, c);
  return c;
}
void vo_odom_set_ba(void* h, int on, double coef)
{
  ((RefOdom*)h)->if_BA = on != 0;
  if (coef > 0) imu_coef = coef;  // optimizers.cpp:8
}
void vo_odom_ba_stats(void* h, int* runs, int* last_iters)
{
  *runs = ((RefOdom*)h)->ba_runs;
  *last_iters = ((RefOdom*)h)->ba_last_iters;
}
// ---- BA probe: the reference's LidarFactor::acc_evaluate2 / evaluate_only_residual (factors.cpp:22-158) on
// the factors captured by the last map update
void vo_odom_ba_probe(void* h, int on) { ((RefOdom*)h)->ba_probe = on != 0; }
int vo_odom_ba_count(void* h) { return (int)((RefOdom*)h)->ba_factors.plvec_voxels.size(); }
int vo_odom_ba_poses(void* h, double* poses12, int cap)
{
  RefOdom* o = (RefOdom*)h;
  const int n = (int)o->ba_xs.size();
  for (int i = 0; i < n && i < cap; i++)
  {
    memcpy(poses12 + 12 * i, o->ba_xs[i].R.data(), 72);
    memcpy(poses12 + 12 * i + 9, o->ba_xs[i].p.data(), 24);
  }
  return n;
}
static vector<IMUST> ba_pose_vec(const double* poses12, int win)
{
  vector<IMUST> xs(win);
  for (int i = 0; i < win; i++)
  {
    memcpy(xs[i].R.data(), poses12 + 12 * i, 72);
    memcpy(xs[i].p.data(), poses12 + 12 * i + 9, 24);
  }
  return xs;
}
int vo_odom_ba_hess(void* h, const double* poses12, int win, double* Hess, double* JacT, double* residual)
{
  RefOdom* o = (RefOdom*)h;
  if (win != o->ba_factors.win_size) return -1;
  vector<IMUST> xs = ba_pose_vec(poses12, win);
  Eigen::MatrixXd H(6 * win, 6 * win);
  Eigen::VectorXd J(6 * win);
  double r = 0;
  o->ba_factors.acc_evaluate2(xs, 0, (int)o->ba_factors.plvec_voxels.size(), H, J, r);
  for (int j = 0; j < 6 * win; j++)
    for (int i = 0; i < 6 * win; i++) Hess[i + 6 * win * j] = H(i, j);
  for (int i = 0; i < 6 * win; i++) JacT[i] = J(i);
  *residual = r;
  return 0;
}
// evaluate_only_residual overwrites the factors' eig_values / eig_vectors / pcr_adds (factors.cpp:150-153): the
// captured copy is updated like the reference's container; lam0 (optional) receives every factor's lambda_0
int vo_odom_ba_residual(void* h, const double* poses12, int win, double* residual, double* lam0, int cap)
{
  RefOdom* o = (RefOdom*)h;
  if (win != o->ba_factors.win_size) return -1;
  vector<IMUST> xs = ba_pose_vec(poses12, win);
  double r = 0;
  const int n = (int)o->ba_factors.plvec_voxels.size();
  o->ba_factors.evaluate_only_residual(xs, 0, n, r);
  *residual = r;
  if (lam0)
    for (int a = 0; a < n && a < cap; a++) lam0[a] = o->ba_factors.eig_values[a][0];
  return 0;
}

int vo_odom_window(void* h, int* win_count, int* mpo, int cap)
{
  RefOdom* o = (RefOdom*)h;
  *win_count = o->vs.win_count;
  for (int i = 0; i < o->vs.win_size && i < cap; i++) mpo[i] = mp[i];
  return o->vs.win_size;
}
// ---- the reference's own sync_packages (src/sensor/sync.cpp, compiled unmodified) behind the vo_sync_* API. Its
// state is the file's globals plus a function-local static, so only ONE instance may ever exist per process. A scan is
// a one-point cloud: curvature = the time offset of the scan's last point, normal_x = the tag (small integers).
namespace
{
bool g_sync_made = false, g_sync_held = false;
IMUEKF g_sync_ekf;
pcl::PointCloud<PointType>::Ptr g_sync_pl;
}  // namespace
void* vo_sync_create(int notime)
{
  if (g_sync_made) return nullptr;
  g_sync_made = true;
  point_notime = notime;
  return &g_sync_ekf;
}
void vo_sync_destroy(void*) {}
void vo_sync_push_imu(void*, const double imu7[7])
{
  // imu_handler (src/platform/ros2/subscribers.cpp:11-20); the stamp is integer nanoseconds like a ROS stamp
  auto msg = std::make_shared<sensor_msgs::msg::Imu>();
  const int64_t ns = (int64_t)llround(imu7[0] * 1e9);
  msg->header.stamp.sec = (int32_t)(ns / 1000000000LL);
  msg->header.stamp.nanosec = (uint32_t)(ns % 1000000000LL);
  msg->angular_velocity.x = imu7[1], msg->angular_velocity.y = imu7[2], msg->angular_velocity.z = imu7[3];
  msg->linear_acceleration.x = imu7[4], msg->linear_acceleration.y = imu7[5], msg->linear_acceleration.z = imu7[6];
  mBuf.lock();
  imu_last_time = rclcpp::Time(msg->header.stamp).seconds();
  imu_buf.push_back(msg);
  mBuf.unlock();
}
void vo_sync_push_scan(void*, double t_start, double t_last, int64_t tag)
{
  // the tail of pcl_handler (src/sensor/lidar_decoder.cpp:36-43)
  pcl::PointCloud<PointType>::Ptr pl_ptr(new pcl::PointCloud<PointType>());
  PointType ap;
  ap.x = ap.y = ap.z = 0;
  ap.curvature = (float)t_last;
  ap.normal_x = (float)tag;
  pl_ptr->push_back(ap);
  mBuf.lock();
  time_buf.push_back(t_start);
  pcl_buf.push_back(pl_ptr);
  mBuf.unlock();
}
int vo_sync_next(void*, int64_t* tag, double* beg, double* end, double* imu7, int cap, int* m)
{
  *m = 0;
  const size_t nbuf = pcl_buf.size();
  const bool seeding = point_notime && !g_sync_held && last_pcl_time < 0;
  deque<std::shared_ptr<sensor_msgs::msg::Imu>> imus;  // a fresh deque per loop iteration (local_mapping.cpp:300)
  const bool ok = sync_packages(g_sync_pl, imus, g_sync_ekf);
  const bool popped = pcl_buf.size() < nbuf;
  if (!ok && !popped && !g_sync_held) return 0;
  *tag = (int64_t)g_sync_pl->back().normal_x;
  if (!ok && popped && seeding)
  {
    g_sync_held = false;
    return 2;
  }
  if (!ok && !(imu_last_time > g_sync_ekf.pcl_end_time))
  {
    g_sync_held = true;  // the IMU stream has not passed the scan's end: sync_packages keeps the scan
    return 0;
  }
  g_sync_held = false;
  *beg = g_sync_ekf.pcl_beg_time;
  *end = g_sync_ekf.pcl_end_time;
  if ((int)imus.size() > cap) return -3;
  for (size_t i = 0; i < imus.size(); i++)
  {
    imu7[7 * i] = rclcpp::Time(imus[i]->header.stamp).seconds();
    imu7[7 * i + 1] = imus[i]->angular_velocity.x, imu7[7 * i + 2] = imus[i]->angular_velocity.y;
    imu7[7 * i + 3] = imus[i]->angular_velocity.z;
    imu7[7 * i + 4] = imus[i]->linear_acceleration.x, imu7[7 * i + 5] = imus[i]->linear_acceleration.y;
    imu7[7 * i + 6] = imus[i]->linear_acceleration.z;
  }
  *m = (int)imus.size();
  return ok ? 1 : 2;
}

// ---- the reference's own message handlers (src/sensor/lidar_pointcloud_decoder.cpp) and pcl_handler
// (src/sensor/lidar_decoder.cpp), compiled unmodified; pcl::fromROSMsg comes from the shim (fields mapped by name).
namespace
{
builtin_interfaces::msg::Time stamp_of(double t)
{
  builtin_interfaces::msg::Time s;
  const int64_t ns = (int64_t)llround(t * 1e9);
  s.sec = (int32_t)(ns / 1000000000LL);
  s.nanosec = (uint32_t)(ns % 1000000000LL);
  return s;
}
int64_t cloud_out(pcl::PointCloud<PointType>& pl, float* xyz4_out, int64_t cap)
{
  if ((int64_t)pl.size() > cap) return -3;
  for (size_t i = 0; i < pl.size(); i++)
  {
    xyz4_out[4 * i] = pl[i].x, xyz4_out[4 * i + 1] = pl[i].y, xyz4_out[4 * i + 2] = pl[i].z;
    xyz4_out[4 * i + 3] = pl[i].curvature;
  }
  return (int64_t)pl.size();
}
}  // namespace
int64_t vo_decode_handler(int lidar_type, const uint8_t* data, int64_t n, int point_step, int off_x, int off_y, int off_z,
                          int off_t, int t_datatype, double header_stamp, double omega_l, double blind2,
                          int point_filter_num, float* xyz4_out, int64_t cap)
{
  auto msg = std::make_shared<sensor_msgs::msg::PointCloud2>();
  msg->header.stamp = stamp_of(header_stamp);
  msg->width = (uint32_t)n;
  msg->height = 1;
  msg->point_step = (uint32_t)point_step;
  msg->row_step = (uint32_t)(point_step * n);
  msg->data.assign(data, data + (size_t)n * point_step);
  const char* tname = lidar_type == VELODYNE ? "time" : lidar_type == OUSTER ? "t" : "timestamp";
  msg->fields = { { "x", (uint32_t)off_x, 7, 1 }, { "y", (uint32_t)off_y, 7, 1 }, { "z", (uint32_t)off_z, 7, 1 } };
  if (off_t >= 0) msg->fields.push_back({ tname, (uint32_t)off_t, (uint8_t)t_datatype, 1 });
  feat.lidar_type = lidar_type;
  feat.point_filter_num = point_filter_num;
  feat.blind = blind2;
  feat.omega_l = omega_l;
  pcl::PointCloud<PointType> pl_full;
  const sensor_msgs::msg::PointCloud2::SharedPtr cmsg = msg;
  feat.process(cmsg, pl_full);
  return cloud_out(pl_full, xyz4_out, cap);
}
int64_t vo_decode_livox(const uint32_t* offset_time, const float* xyz, int64_t n, double blind2, int point_filter_num,
                        float* xyz4_out, int64_t cap)
{
  auto msg = std::make_shared<livox_ros_driver2::msg::CustomMsg>();
  msg->point_num = (uint32_t)n;
  msg->points.resize(n);
  for (int64_t i = 0; i < n; i++)
  {
    msg->points[i].offset_time = offset_time[i];
    msg->points[i].x = xyz[3 * i], msg->points[i].y = xyz[3 * i + 1], msg->points[i].z = xyz[3 * i + 2];
  }
  feat.lidar_type = LIVOX;
  feat.point_filter_num = point_filter_num;
  feat.blind = blind2;
  pcl::PointCloud<PointType> pl_full;
  const livox_ros_driver2::msg::CustomMsg::SharedPtr cmsg = msg;
  feat.process(cmsg, pl_full);
  return cloud_out(pl_full, xyz4_out, cap);
}
// pcl_handler on a scan given as n x (x, y, z, time offset): packed as a Velodyne message with a float `time` field
// (the handler that passes the stamps through unchanged; the LAST input point's stamp must lie in (0.01, 0.12),
// lidar_pointcloud_decoder.cpp:86), result taken back out of pcl_buf / time_buf.
int vo_scan_prepare(int n, const float* xyz4_in, int point_filter_num, double blind2, float* xyz4_out)
{
  auto msg = std::make_shared<sensor_msgs::msg::PointCloud2>();
  msg->width = (uint32_t)n;
  msg->height = 1;
  msg->point_step = 16;
  msg->data.resize((size_t)n * 16);
  if (n > 0) memcpy(msg->data.data(), xyz4_in, (size_t)n * 16);
  msg->fields = { { "x", 0, 7, 1 }, { "y", 4, 7, 1 }, { "z", 8, 7, 1 }, { "time", 12, 7, 1 } };
  feat.lidar_type = VELODYNE;
  feat.point_filter_num = point_filter_num;
  feat.blind = blind2;
  const sensor_msgs::msg::PointCloud2::SharedPtr cmsg = msg;
  pcl_handler(cmsg);
  mBuf.lock();
  pcl::PointCloud<PointType>::Ptr pl = pcl_buf.back();
  pcl_buf.pop_back();
  time_buf.pop_back();
  mBuf.unlock();
  return (int)cloud_out(*pl, xyz4_out, (int64_t)(n > 2 ? n : 2));
}

void vo_odom_journey(void* h, double* jour, int* release_flag)
{
  RefOdom* o = (RefOdom*)h;
  if (jour) *jour = o->jour;
  if (release_flag) *release_flag = o->release_flag ? 1 : 0;
}
int vo_odom_idle(void* h, int horizon, int* nodes_freed) { return ((RefOdom*)h)->idle_release(horizon, nodes_freed); }
}

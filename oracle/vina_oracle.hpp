// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the VINA-SLAM per-scan hot path (SURVEY.md §8a rows
// a1-a13). The reference (/root/reference, C++17 on ROS 2 + PCL + Eigen3) has
// no tests, golden vectors or fixtures and its own build cannot run here (no
// Eigen/PCL/ROS). PINNING: the reference's own source files
// (point_utils.cpp, octree.cpp, voxel_map.cpp, imu_ekf.cpp, odometry.cpp) are
// compiled unmodified against the header shims of oracle/ref_shim into
// oracle/_ref/libvina_ref.so (oracle/Makefile, target `ref`); this restatement
// reproduces that build BIT FOR BIT on whole synthetic sequences
// (tests/test_oracle_vs_ref.py, golden vectors tests/golden/ref_small.npz).
// What stays restated on both sides is Eigen itself (absent, version unpinned
// by the reference): SelfAdjointEigenSolver and inverse() follow Eigen 3.4.0 and
// are checked against numpy (tests/test_oracle.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use anything in this directory. The product
// (vina_slam_b200/) never links, imports or calls it.
//
// Every function cites the reference file:line it follows. Data structures
// keep the reference's shape on purpose (std::unordered_map<VOXEL_LOC,
// OctoTree*>, pointer octree, per-node SlideWindow, fork-join std::thread over
// root voxels) so that the timed "-O3 -ffast-math" build of this file is a
// fair CPU baseline of the reference's own design.
#pragma once
#include "omat.hpp"
#include <cstdint>
#include <deque>
#include <memory>
#include <unordered_map>
#include <vector>

namespace vo
{
#define VO_HASH_P 1000033
#define VO_MAX_N 100000000000
#define VO_DIM 15

// include/vina_slam/core/types.hpp:13-41
struct VOXEL_LOC
{
  int64_t x, y, z;
  VOXEL_LOC(int64_t vx = 0, int64_t vy = 0, int64_t vz = 0) : x(vx), y(vy), z(vz) {}
  bool operator==(const VOXEL_LOC& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct VoxelHash
{
  size_t operator()(const VOXEL_LOC& s) const
  {
    using std::hash;
    return (((hash<int64_t>()(s.z) * VO_HASH_P) % VO_MAX_N + hash<int64_t>()(s.y)) * VO_HASH_P) % VO_MAX_N +
           hash<int64_t>()(s.x);
  }
};

// include/vina_slam/core/types.hpp:43-113
struct IMUST
{
  double t;
  Mat3 R;
  Vec3 p, v, bg, ba, g;
  Mat15 cov;
  IMUST() { setZero(); }
  void setZero()
  {
    t = 0;
    R.setIdentity();
    p.setZero();
    v.setZero();
    bg.setZero();
    ba.setZero();
    g = V3(0, 0, -9.8);
    cov.setIdentity();
    cov = cov * 0.0001;
    for (int i = 9; i < 15; i++) cov(i, i) = 0.00001;
  }
  IMUST& operator+=(const Vec15& ist)  // types.hpp:67-75
  {
    R = R * Exp(ist.block<3, 1>(0, 0));
    p += ist.block<3, 1>(3, 0);
    v += ist.block<3, 1>(6, 0);
    bg += ist.block<3, 1>(9, 0);
    ba += ist.block<3, 1>(12, 0);
    return *this;
  }
  Vec15 operator-(const IMUST& b) const  // types.hpp:77-86
  {
    Vec15 a;
    a.setBlock<3, 1>(0, 0, Log(b.R.transpose() * R));
    a.setBlock<3, 1>(3, 0, p - b.p);
    a.setBlock<3, 1>(6, 0, v - b.v);
    a.setBlock<3, 1>(9, 0, bg - b.bg);
    a.setBlock<3, 1>(12, 0, ba - b.ba);
    return a;
  }
};

// include/vina_slam/core/types.hpp:115-175
struct PointCluster
{
  Mat3 P;
  Vec3 v;
  int N;
  PointCluster() { clear(); }
  void clear()
  {
    P.setZero();
    v.setZero();
    N = 0;
  }
  void push(const Vec3& vec)
  {
    N++;
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 3; i++) P(i, j) = P(i, j) + vec[i] * vec[j];
    v += vec;
  }
  Mat3 cov() const
  {
    Vec3 center = v / (double)N;
    Mat3 r;
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 3; i++) r(i, j) = P(i, j) / (double)N - center[i] * center[j];
    return r;
  }
  PointCluster& operator+=(const PointCluster& s)
  {
    P += s.P;
    v += s.v;
    N += s.N;
    return *this;
  }
  PointCluster& operator-=(const PointCluster& s)
  {
    P -= s.P;
    v -= s.v;
    N -= s.N;
    return *this;
  }
  void transform(const PointCluster& sigv, const IMUST& stat)  // types.hpp:168-174
  {
    N = sigv.N;
    v = stat.R * sigv.v + (double)N * stat.p;
    Vec3 Rv = stat.R * sigv.v;
    Mat3 rp;
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 3; i++) rp(i, j) = Rv[i] * stat.p[j];
    Mat3 RPRt = stat.R * sigv.P * stat.R.transpose();
    Vec3 Np = (double)N * stat.p;
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 3; i++) P(i, j) = ((RPRt(i, j) + rp(i, j)) + rp(j, i)) + Np[i] * stat.p[j];
  }
};

// include/vina_slam/core/types.hpp:177-182
struct pointVar
{
  Vec3 pnt;
  Mat3 var;
  float intensity = 0;
};
typedef std::vector<pointVar> PVec;
typedef std::shared_ptr<PVec> PVecPtr;

// the four PointType (pcl::PointXYZINormal) fields the path reads
struct PointXYZT
{
  float x, y, z, curvature;
};
typedef std::vector<PointXYZT> Cloud;

// one sensor_msgs::Imu as used by imu_ekf.cpp: stamp (s), gyro, accel
struct ImuSample
{
  double t;
  double gyr[3];
  double acc[3];
};

// include/vina_slam/mapping/plane.hpp:6-24 (hot-path fields only)
struct Plane
{
  Vec3 center = Vec3::Zero();
  Vec3 normal = Vec3::Zero();
  Mat6 plane_var = Mat6::Zero();
  float radius = 0;
  bool is_plane = false;
};

// mutable globals of the reference (octree.cpp:67-75, node.cpp:38,219,256-259)
struct Globals
{
  double min_point[4] = { 20, 20, 15, 10 };
  double min_eigen_value = 0.0025;
  int max_layer = 2;
  int max_points = 100;
  double voxel_size = 1.0;
  double plane_eigen_value_thre[4] = { 1, 1, 1, 1 };  // already inverted (node.cpp:256-259)
  std::vector<int> mp;                                // ring map, octree.cpp:75
  int thread_num = 5;
  int win_size = 10;
  double dept_err = 0.02, beam_err = 0.05;
  double down_size = 0.1;
};

// include/vina_slam/mapping/slide_window.hpp:7-20, octree.cpp:115-140
struct SlideWindow
{
  std::vector<PVec> points;
  std::vector<PointCluster> pcrs_local;
  explicit SlideWindow(int wdsize)
  {
    pcrs_local.resize(wdsize);
    points.resize(wdsize);
    for (int i = 0; i < wdsize; i++) points[i].reserve(20);
  }
  void resize(int wdsize)
  {
    if ((int)points.size() != wdsize)
    {
      points.resize(wdsize);
      pcrs_local.resize(wdsize);
    }
  }
  void clear()
  {
    for (size_t i = 0; i < points.size(); i++)
    {
      points[i].clear();
      pcrs_local[i].clear();
    }
  }
};

void Bf_var(const pointVar& pv, Mat9& bcov, const Vec3& vec);  // octree.cpp:83-92

// include/vina_slam/mapping/octree.hpp:21-97 (hot-path members)
class OctoTree
{
public:
  Globals* G;
  SlideWindow* sw = nullptr;
  PointCluster pcr_add;
  Mat9 cov_add;
  PointCluster pcr_fix;
  PVec point_fix;
  int layer, octo_state, wdsize;
  OctoTree* leaves[8];
  double voxel_center[3];
  double jour = 0;
  float quater_length;
  Plane plane;
  bool isexist = false;
  Vec3 eig_value;
  Mat3 eig_vector;
  int last_num = 0, opt_state = -1;
  // bookkeeping that is NOT in the reference: identity of the node for parity
  // dumps (root key + child path), never read by the algorithm
  VOXEL_LOC root_key;
  int path = 0;

  OctoTree(Globals* g, int _l, int _w);
  void push(int ord, const pointVar& pv, const Vec3& pw, std::vector<SlideWindow*>& sws);
  void push_fix(pointVar& pv);
  bool plane_judge(Vec3& eig_values);
  void allocate(int ord, const pointVar& pv, const Vec3& pw, std::vector<SlideWindow*>& sws);
  OctoTree* make_child(int leafnum, const int xyz[3]);
  void fix_divide(std::vector<SlideWindow*>& sws);
  void subdivide(int si, IMUST& xx, std::vector<SlideWindow*>& sws);
  void plane_update();
  void recut(int win_count, std::vector<IMUST>& x_buf, std::vector<SlideWindow*>& sws);
  void margi(int win_count, int mgsize, std::vector<IMUST>& x_buf);
  int match(Vec3& wld, Plane*& pla, double& max_prob, Mat3& var_wld, double& sigma_d, OctoTree*& oc);
  bool inside(Vec3& wld);
  void clear_slwd(std::vector<SlideWindow*>& sws);
  void delete_ptr();
  int code() const { return layer | (path << 2); }
};

typedef std::unordered_map<VOXEL_LOC, OctoTree*, VoxelHash> VoxelMap;

// src/core/point_utils.cpp:3-65, include/vina_slam/core/point_utils.hpp:7-44
void calcBodyVar(Vec3& pb, const float range_inc, const float degree_inc, Mat3& var);
void var_init(IMUST& ext, Cloud& pl_cur, PVecPtr pptr, double dept_err, double beam_err);
void pvec_update(PVecPtr pptr, IMUST& x_curr, std::vector<Vec3>& pwld);
void down_sampling_voxel(Cloud& pl_feat, double voxel_size);
void down_sampling_close(Cloud& pl_feat, double voxel_size);  // point_utils.hpp:47-113 (vina_oracle_init.cpp)

// voxel key, src/mapping/voxel_map.cpp:56-65 / 246-253
VOXEL_LOC voxel_key(const Vec3& pw, double voxel_size);

// src/mapping/voxel_map.cpp:47-135, 241-266
void cut_voxel_multi(Globals* G, VoxelMap& feat_map, PVecPtr pvec, int win_count, VoxelMap& feat_tem_map, int wdsize,
                     std::vector<Vec3>& pwld, std::vector<std::vector<SlideWindow*>>& sws);
int match(Globals* G, VoxelMap& feat_map, Vec3& wld, Plane*& pla, Mat3& var_wld, double& sigma_d, OctoTree*& oc);

// src/estimation/imu_ekf.cpp:13-145 (+ ekf_imu.hpp:12-42)
class IMUEKF
{
public:
  double pcl_beg_time = 0, pcl_end_time = 0, last_pcl_end_time = 0;
  ImuSample last_imu;
  Vec3 cov_acc, cov_gyr, cov_bias_gyr, cov_bias_acc;
  Mat3 Lid_rot_to_IMU;
  Vec3 Lid_offset_to_IMU;
  double scale_gravity = 1.0;
  std::vector<IMUST> imu_poses;
  int point_notime = 0;
  // start-up (imu_ekf.cpp:147-201, ekf_imu.hpp:16-20); the harness bootstrap sets init_flag and skips it
  bool init_flag = true;
  int init_num = 0, min_init_num = 30;
  Vec3 mean_acc = Vec3::Zero(), mean_gyr = Vec3::Zero();
  void IMU_init(std::deque<ImuSample>& imus);
  // 0 = still initialising, 1 = scan deskewed, -1 = "LiDAR time regress" (the reference exit(0)s)
  int process(IMUST& x_curr, Cloud& pcl_in, std::deque<ImuSample>& imus);
  IMUEKF();
  // returns 0 on success, -1 for "LiDAR time regress" (the reference exit(0)s)
  int motion_blur(IMUST& xc, Cloud& pcl_in, std::deque<ImuSample>& imus);
  // the two halves of motion_blur, exposed for stage-wise parity:
  int propagate(IMUST& xc, std::deque<ImuSample>& imus);  // imu_ekf.cpp:17-104
  void deskew(const IMUST& xc, Cloud& pcl_in);            // imu_ekf.cpp:106-144
};

// per-iteration debug record of LioStateEstimation (not in the reference)
struct IekfIterDump
{
  double HTH[36];  // column-major 6x6
  double HTz[6];
  double nnt[9];
  int match_num;
  std::vector<int64_t> keys;   // 3 per point
  std::vector<int32_t> codes;  // OctoTree::code() of the associated leaf, -1 if none
  std::vector<uint8_t> flags;
  std::vector<double> sigma;   // sigma_d where flag, else 0
  double R[9], p[3];           // state the iteration was evaluated at
};

// include/vina_slam/mapping/factors.hpp:10-40, src/mapping/factors.cpp:7-168: the LiDAR BA factor container
// (one entry per plane leaf that tras_opt accepted) and its two evaluations
struct LidarFactor
{
  std::vector<PointCluster> sig_vecs;                   // pcr_fix (world frame)
  std::vector<std::vector<PointCluster>> plvec_voxels;  // pcrs_local[mp[i]] (body frame of window frame i)
  std::vector<double> coeffs;
  std::vector<Vec3> eig_values;
  std::vector<Mat3> eig_vectors;
  std::vector<PointCluster> pcr_adds;
  int win_size = 0;
  void push_voxel(std::vector<PointCluster>& vec_orig, PointCluster& fix, double coe, Vec3& eig_value, Mat3& eig_vector,
                  PointCluster& pcr_add);                                                        // factors.cpp:11-20
  void clear();
  // Hess: (6 win)^2 column-major, JacT: 6 win
  void acc_evaluate2(const std::vector<IMUST>& xs, int head, int end, std::vector<double>& Hess,
                     std::vector<double>& JacT, double& residual);                               // factors.cpp:22-126
  void evaluate_only_residual(const std::vector<IMUST>& xs, int head, int end, double& residual);  // factors.cpp:128-158
};

// src/estimation/imu_preintegration.cpp, include/vina_slam/preintegration.hpp: one IMU pre-integration factor
// between two consecutive window frames (vina_oracle_ba.cpp)
struct BaNoise  // the globals of imu_preintegration.cpp:3-5 (node.cpp:262-265, 309)
{
  Mat6 noiseMeas = Mat6::Zero(), noiseWalk = Mat6::Zero();
  double scale_gravity = 1.0;
};
struct IMU_PRE
{
  Mat3 R_delta;
  Vec3 p_delta, v_delta, bg, ba;
  Mat3 R_bg, p_bg, p_ba, v_bg, v_ba;
  double dtime;
  Mat15 cov;
  Vec3 dbg, dba, dbg_buf, dba_buf;
  IMU_PRE(const Vec3& bg1, const Vec3& ba1);
  void push_imu(std::deque<ImuSample>& imu_buffer, const BaNoise& nz);
  void add_imu(Vec3& cur_gyr, Vec3& cur_acc, double dt, const BaNoise& nz);
  double give_evaluate(IMUST& st1, IMUST& st2, Mat<30, 30>& jtj, Mat<30, 1>& gg, bool jac_enable);
  // imu_preintegration.cpp:165-237: with the gravity Jacobian (3 more columns)
  double give_evaluate_g(IMUST& st1, IMUST& st2, Mat<33, 33>& jtj, Mat<33, 1>& gg, bool jac_enable);
  void update_state(const Vec15& dxi);
};
// LI_BA_Optimizer::damping_iter (src/mapping/optimizers.cpp:430-517); returns the number of LM iterations
int ba_damping_iter(std::vector<IMUST>& x_stats, LidarFactor& voxhess, std::deque<IMU_PRE*>& imus_factor, double imu_coef,
                    std::vector<double>* hess_out);

// LI_BA_OptimizerGravity::damping_iter (src/mapping/optimizers.cpp:746-826), vina_oracle_init.cpp
void ba_damping_iter_gravity(std::vector<IMUST>& x_stats, LidarFactor& voxhess, std::deque<IMU_PRE*>& imus_factor,
                             std::vector<double>& resis, int max_iter, double imu_coef);

// The owner of everything VINA_SLAM keeps for the per-scan loop
// (include/vina_slam/platform/ros2/node.hpp:30-96; src/pipeline/local_mapping.cpp:258-550)
class Odom
{
public:
  Globals G;
  IMUST x_curr, extrin_para;
  IMUEKF odom_ekf;
  VoxelMap surf_map, surf_map_slide;
  std::vector<std::vector<SlideWindow*>> sws;
  std::vector<IMUST> x_buf;
  std::vector<PVecPtr> pvec_buf;
  int win_count = 0, win_base = 0;
  std::vector<Vec3> pwld;
  int degrade_cnt = 0;

  // debugging / parity hooks
  bool dump_iters = false;
  std::vector<IekfIterDump> iter_dumps;
  int last_iters = 0;
  // stage timers (seconds, steady_clock) at the reference's own stamps
  // local_mapping.cpp:359-360, 432, 449, 452, 503, 508
  double t_odom = 0, t_insert = 0, t_recut = 0, t_margi = 0;
  PVecPtr last_pptr, last_full_pptr;
  Cloud last_down;
  // BA probe: copy of the LiDAR factors (tras_opt, octree.cpp:498-521) and of the window poses between
  // multi_recut and multi_margi, where LI_BA_Optimizer::damping_iter consumes them (local_mapping.cpp:492-497)
  bool ba_probe = false;
  LidarFactor ba_factors;
  std::vector<IMUST> ba_xs;
  // sliding-window BA (local_mapping.cpp:437-441, 492-497, 541-546). The harness bootstraps the window without
  // IMU data: those frames carry no pre-integration factor (nullptr) and BA only runs once every pair of
  // consecutive window frames has one.
  bool if_BA = false;
  double imu_coef = 1e-4;
  BaNoise ba_noise;
  std::deque<IMU_PRE*> imu_pre_buf;
  int ba_runs = 0, ba_last_iters = 0;
  // distance travelled and map pruning (local_mapping.cpp:262-263, 272, 317-341, 509-519)
  double jour = 0;
  Vec3 last_pos = Vec3::Zero();
  bool release_flag = false;

  // start-up phase (node.cpp:293-408, initialization.cpp, odometry.cpp:267-439; vina_oracle_init.cpp)
  Cloud pl_tree, kd_cloud;
  std::vector<std::shared_ptr<Cloud>> pl_origs;
  std::vector<double> beg_times;
  std::vector<std::deque<ImuSample>> vec_imus;
  LidarFactor init_voxhess;
  std::vector<OctoTree*> init_nodes;
  int init_rounds = 0;
  Vec3 init_eig = Vec3::Zero();
  void lio_state_estimation_kdtree(PVecPtr pptr);
  int motion_init();
  void clear_map();
  int initialization(std::deque<ImuSample>& imus, Cloud& pcl_curr);
  void system_reset(std::deque<ImuSample>& imus);
  // one scan of the start-up phase: 0 = collecting, 1 = initialised, -1 = failed (system reset)
  int init_scan(Cloud& pcl_curr, double beg, std::deque<ImuSample>& imus);
  // local_mapping.cpp:489-546 (BA, margi, journey, window shift); vh / nodes: the factor container margi takes its
  // values from (motion_init's), nullptr = the one multi_recut of this scan marked
  void window_tail(LidarFactor* vh = nullptr, std::vector<OctoTree*>* nodes = nullptr);

  explicit Odom(const Globals& g);
  ~Odom();

  // src/pipeline/odometry.cpp:64-255 with use_vnc=false; max_iter<=0 keeps the
  // reference's 20, a positive value overrides (4 = the VNC_lio budget).
  bool LioStateEstimation(PVecPtr pptr, int max_iter_override);
  void multi_recut(VoxelMap& feat_map, int win_count, std::vector<IMUST>& xs,
                   std::vector<std::vector<SlideWindow*>>& sws);                     // local_mapping.cpp:203-254
  void multi_margi(VoxelMap& feat_map, int win_count, std::vector<IMUST>& xs,
                   std::vector<SlideWindow*>& sw);                                   // local_mapping.cpp:17-84
  // local_mapping.cpp:434-451 + 489-546: push frame, insert, recut, margi, window shift
  void map_update(PVecPtr pptr, std::deque<ImuSample>* imus = nullptr);
  // local_mapping.cpp:389-546, one scan. iekf_on_full: feed the un-downsampled
  // scan to the IEKF (production, :413) or the down-sampled one (:412).
  int step(Cloud& pcl_curr, double pcl_beg_time, std::deque<ImuSample>& imus, bool iekf_on_full, int max_iter);
  // harness bootstrap (replaces initialization()): deskewed scan
  // at a known state -> downsample, var_init, pvec_update, map_update.
  void bootstrap(Cloud& pcl_deskewed, const IMUST& x_known);
  // the `else if (release_flag)` branch of the idle path, local_mapping.cpp:317-341: erase every root voxel (and
  // its subtree) whose last marginalisation is `horizon` metres of travel or more behind (700 in the reference).
  // Returns the number of roots erased, nodes_freed counts the subtrees' nodes too.
  int idle_release(int horizon, int* nodes_freed);
  void tras_ptr(OctoTree* ot, std::vector<OctoTree*>& octos_release);  // octree.cpp:597-608
};
}  // namespace vo

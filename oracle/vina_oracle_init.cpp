// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's start-up phase:
//   IMUEKF::IMU_init / process                      src/estimation/imu_ekf.cpp:147-201
//   down_sampling_close                             include/vina_slam/core/point_utils.hpp:47-113
//   VINA_SLAM::lio_state_estimation_kdtree          src/pipeline/odometry.cpp:267-439
//   VINA_SLAM::initialization / system_reset        src/platform/ros2/node.cpp:293-408
//   Initialization::align_gravity / motion_blur /
//   motion_init                                     src/pipeline/initialization.cpp:28-367
//   cut_voxel (serial)                              src/mapping/voxel_map.cpp:4-45
//   IMU_PRE::give_evaluate_g                        src/estimation/imu_preintegration.cpp:165-237
//   LI_BA_OptimizerGravity                          src/mapping/optimizers.cpp:624-826
// PINNING: the reference's own files (initialization.cpp, odometry.cpp, optimizers.cpp, imu_preintegration.cpp,
// imu_ekf.cpp, ...) compile unmodified into oracle/_ref (ref_harness.cpp restates only node.cpp:293-408, which needs
// the ROS node); tests/test_oracle_vs_ref.py drives both through a cold start and compares states and maps.
// Third-party pieces restated on both sides (absent here, see oracle/README.md): pcl::KdTreeFLANN::nearestKSearch
// (exact k nearest neighbours by squared float distance, ties by index), Eigen's colPivHouseholderQr().solve of the
// 5 x 3 plane fit (normal equations on both sides), AngleAxisd(angle, axis).toRotationMatrix().
#include "vina_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

namespace vo
{
std::vector<double> ldlt_solve(std::vector<double> L, int n, const std::vector<double>& b);  // vina_oracle_ba.cpp
void tras_opt_collect(OctoTree* n, LidarFactor& vox_opt, std::vector<OctoTree*>* nodes);  // vina_oracle.cpp
Mat3 jr(Vec3 vec);
Mat3 jr_inv(const Mat3& rotR);

// ---- imu_ekf.cpp:147-201 --------------------------------------------------------------------------------------
void IMUEKF::IMU_init(std::deque<ImuSample>& imus)
{
  Vec3 cur_acc, cur_gyr;
  for (const ImuSample& imu : imus)
  {
    cur_acc = V3(imu.acc[0], imu.acc[1], imu.acc[2]);
    cur_gyr = V3(imu.gyr[0], imu.gyr[1], imu.gyr[2]);
    if (init_num != 0)
    {
      mean_acc += (cur_acc - mean_acc) / (double)init_num;
      mean_gyr += (cur_gyr - mean_gyr) / (double)init_num;
    }
    else
    {
      mean_acc = cur_acc;
      mean_gyr = cur_gyr;
      init_num = 1;
    }
    init_num++;
  }
  last_imu = imus.back();
}

int IMUEKF::process(IMUST& x_curr, Cloud& pcl_in, std::deque<ImuSample>& imus)
{
  if (!init_flag)
  {
    IMU_init(imus);
    if (norm(mean_acc) < 2) scale_gravity = 9.8;  // G_m_s2
    x_curr.g = (-1.0 * mean_acc) * scale_gravity;
    if (init_num > min_init_num) init_flag = true;
    last_pcl_end_time = pcl_end_time;
    return 0;
  }
  return motion_blur(x_curr, pcl_in, imus) == 0 ? 1 : -1;
}

// ---- point_utils.hpp:47-113 -----------------------------------------------------------------------------------
void down_sampling_close(Cloud& pl_feat, double voxel_size)
{
  if (voxel_size < 0.001) return;
  std::unordered_map<VOXEL_LOC, std::shared_ptr<Cloud>, VoxelHash> feat_map;
  float loc_xyz[3];
  for (PointXYZT& p_c : pl_feat)
  {
    const float data[3] = { p_c.x, p_c.y, p_c.z };
    for (int j = 0; j < 3; j++)
    {
      loc_xyz[j] = data[j] / voxel_size;
      if (loc_xyz[j] < 0) loc_xyz[j] -= 1.0;
    }
    VOXEL_LOC position((int64_t)loc_xyz[0], (int64_t)loc_xyz[1], (int64_t)loc_xyz[2]);
    auto iter = feat_map.find(position);
    if (iter == feat_map.end())
    {
      std::shared_ptr<Cloud> pl_ptr(new Cloud);
      pl_ptr->push_back(p_c);
      feat_map[position] = pl_ptr;
    }
    else
      iter->second->push_back(p_c);
  }
  pl_feat.clear();
  for (auto iter = feat_map.begin(); iter != feat_map.end(); ++iter)
  {
    std::shared_ptr<Cloud> pl_ptr = iter->second;
    PointXYZT pb = (*pl_ptr)[0];
    int plsize = (int)pl_ptr->size();
    for (int i = 1; i < plsize; i++)
    {
      PointXYZT& pp = (*pl_ptr)[i];
      pb.x += pp.x;
      pb.y += pp.y;
      pb.z += pp.z;
    }
    pb.x /= plsize;
    pb.y /= plsize;
    pb.z /= plsize;
    double ndis = 100;
    int mnum = 0;
    for (int i = 0; i < plsize; i++)
    {
      PointXYZT& pp = (*pl_ptr)[i];
      double xx = pb.x - pp.x;
      double yy = pb.y - pp.y;
      double zz = pb.z - pp.z;
      double dis = xx * xx + yy * yy + zz * zz;
      if (dis < ndis)
      {
        mnum = i;
        ndis = dis;
      }
    }
    pl_feat.push_back((*pl_ptr)[mnum]);
  }
}

// ---- odometry.cpp:267-439 -------------------------------------------------------------------------------------
namespace
{
const int NMATCH = 5;  // include/vina_slam/core/constants.hpp

// pcl::KdTreeFLANN::nearestKSearch as ref_shim/pcl/kdtree/kdtree_flann.h restates it: exact, float distances
void nearest_k(const Cloud& cloud, const PointXYZT& q, int k, std::vector<int>& idx, std::vector<float>& d2)
{
  std::vector<std::pair<float, int>> all;
  all.reserve(cloud.size());
  for (size_t i = 0; i < cloud.size(); i++)
  {
    const PointXYZT& p = cloud[i];
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
}
}  // namespace

void Odom::lio_state_estimation_kdtree(PVecPtr pptr)
{
  if (!pptr || pptr->empty()) return;
  if (pl_tree.size() < 100)
  {
    for (pointVar pv : *pptr)
    {
      pv.pnt = x_curr.R * pv.pnt + x_curr.p;
      PointXYZT pp;
      pp.x = pv.pnt[0];
      pp.y = pv.pnt[1];
      pp.z = pv.pnt[2];
      pp.curvature = 0;
      pl_tree.push_back(pp);
    }
    kd_cloud = pl_tree;  // kd_map.setInputCloud(pl_tree): the tree shares the cloud
    return;
  }
  // (the kd-tree indexes the shared cloud object: points pushed later are not searched until the next setInputCloud,
  // but the cloud the indices refer to IS pl_tree - with an exact brute-force search both are the same array here)
  const int num_max_iter = 4;
  IMUST x_prop = x_curr;
  int psize = (int)pptr->size();
  bool EKF_stop_flg = false;
  bool flg_EKF_converged = false;
  Mat15 G, H_T_H, I_STATE;
  G.setZero();
  H_T_H.setZero();
  I_STATE.setIdentity();
  std::vector<float> sqdis(NMATCH);
  std::vector<int> nearInd(NMATCH);
  int rematch_num = 0;
  Mat15 cov_inv = inverse(x_curr.cov);
  Mat<NMATCH, 1> b;
  for (int i = 0; i < NMATCH; i++) b[i] = -1.0;
  std::vector<double> ds(psize, -1);
  std::vector<Vec3> directs(psize);
  bool refind = true;
  for (int iterCount = 0; iterCount < num_max_iter; iterCount++)
  {
    Mat6 HTH;
    HTH.setZero();
    Vec6 HTz;
    HTz.setZero();
    int valid = 0;
    for (int i = 0; i < psize; i++)
    {
      pointVar& pv = pptr->at(i);
      Mat3 phat = hat(pv.pnt);
      Vec3 wld = x_curr.R * pv.pnt + x_curr.p;
      if (refind)
      {
        PointXYZT apx;
        apx.x = wld[0];
        apx.y = wld[1];
        apx.z = wld[2];
        apx.curvature = 0;
        nearest_k(pl_tree, apx, NMATCH, nearInd, sqdis);
        Mat<NMATCH, 3> A;
        for (int k = 0; k < NMATCH; k++)
        {
          const PointXYZT& pp = pl_tree[nearInd[k]];
          A(k, 0) = pp.x;
          A(k, 1) = pp.y;
          A(k, 2) = pp.z;
        }
        // A.colPivHouseholderQr().solve(b) as ref_shim/mini_eigen.hpp evaluates it: (A^T A)^-1 (A^T b)
        Mat<3, NMATCH> At = A.transpose();
        Mat3 AtA = At * A;
        Vec3 Atb = At * b;
        Vec3 direct = inverse(AtA) * Atb;
        bool check_flag = false;
        for (int k = 0; k < NMATCH; k++)
        {
          double d = (direct[0] * A(k, 0) + direct[1] * A(k, 1)) + direct[2] * A(k, 2);
          if (std::fabs(d + 1.0) > 0.1) check_flag = true;
        }
        if (check_flag)
        {
          ds[i] = -1;
          continue;
        }
        double d = 1.0 / norm(direct);
        ds[i] = d;
        directs[i] = direct * d;
      }
      if (ds[i] >= 0)
      {
        double pd2 = dot(directs[i], wld) + ds[i];
        Vec6 jac_s;
        Vec3 h3 = phat * x_curr.R.transpose() * directs[i];
        for (int k = 0; k < 3; k++)
        {
          jac_s[k] = h3[k];
          jac_s[3 + k] = directs[i][k];
        }
        for (int c = 0; c < 6; c++)
          for (int r = 0; r < 6; r++) HTH(r, c) = HTH(r, c) + jac_s[r] * jac_s[c];
        for (int r = 0; r < 6; r++) HTz[r] = HTz[r] + jac_s[r] * (-pd2);
        valid++;
      }
    }
    H_T_H.setBlock<6, 6>(0, 0, HTH);
    Mat15 K_1 = inverse(Mat15(H_T_H + cov_inv / 1000.0));
    G.setBlock<15, 6>(0, 0, K_1.block<15, 6>(0, 0) * HTH);
    Vec15 vec = x_prop - x_curr;
    Vec15 solution = K_1.block<15, 6>(0, 0) * HTz + vec - G.block<15, 6>(0, 0) * vec.block<6, 1>(0, 0);
    x_curr += solution;
    Vec3 rot_add = solution.block<3, 1>(0, 0);
    Vec3 tra_add = solution.block<3, 1>(3, 0);
    refind = false;
    if ((norm(rot_add) * 57.3 < 0.01) && (norm(tra_add) * 100 < 0.015))
    {
      refind = true;
      flg_EKF_converged = true;
      rematch_num++;
    }
    if (iterCount == num_max_iter - 2 && !flg_EKF_converged) refind = true;
    if (rematch_num >= 2 || (iterCount == num_max_iter - 1))
    {
      x_curr.cov = (I_STATE - G) * x_curr.cov;
      EKF_stop_flg = true;
    }
    if (EKF_stop_flg) break;
  }
  for (pointVar pv : *pptr)
  {
    pv.pnt = x_curr.R * pv.pnt + x_curr.p;
    PointXYZT ap;
    ap.x = pv.pnt[0];
    ap.y = pv.pnt[1];
    ap.z = pv.pnt[2];
    ap.curvature = 0;
    pl_tree.push_back(ap);
  }
  down_sampling_voxel(pl_tree, 0.5);
}

// ---- voxel_map.cpp:4-45 ---------------------------------------------------------------------------------------
static void cut_voxel(Globals* G, VoxelMap& feat_map, PVecPtr pvec, int win_count, VoxelMap& feat_tem_map, int wdsize,
                      std::vector<Vec3>& pwld, std::vector<SlideWindow*>& sws)
{
  int plsize = (int)pvec->size();
  for (int i = 0; i < plsize; i++)
  {
    pointVar& pv = (*pvec)[i];
    Vec3& pw = pwld[i];
    VOXEL_LOC position = voxel_key(pw, G->voxel_size);
    auto iter_feat_map = feat_map.find(position);
    if (iter_feat_map != feat_map.end())
    {
      iter_feat_map->second->allocate(win_count, pv, pw, sws);
      iter_feat_map->second->isexist = true;
      if (feat_tem_map.find(position) == feat_tem_map.end()) feat_tem_map[position] = iter_feat_map->second;
    }
    else
    {
      OctoTree* ot = new OctoTree(G, 0, wdsize);
      ot->root_key = position;
      ot->allocate(win_count, pv, pw, sws);
      ot->voxel_center[0] = (0.5 + position.x) * G->voxel_size;
      ot->voxel_center[1] = (0.5 + position.y) * G->voxel_size;
      ot->voxel_center[2] = (0.5 + position.z) * G->voxel_size;
      ot->quater_length = G->voxel_size / 4.0;
      feat_map[position] = ot;
      feat_tem_map[position] = ot;
    }
  }
}

// ---- imu_preintegration.cpp:165-237 ---------------------------------------------------------------------------
double IMU_PRE::give_evaluate_g(IMUST& st1, IMUST& st2, Mat<33, 33>& jtj, Mat<33, 1>& gg, bool jac_enable)
{
  Mat15 joca, jocb;
  Vec15 rr;
  joca.setZero();
  jocb.setZero();
  rr.setZero();
  Mat<15, 3> jocg;
  jocg.setZero();
  const Mat3 I33 = Mat3::Identity();

  Mat3 R_correct = R_delta * Exp(R_bg * dbg);
  Vec3 t_correct = p_delta + p_bg * dbg + p_ba * dba;
  Vec3 v_correct = v_delta + v_bg * dbg + v_ba * dba;

  Mat3 res_r = R_correct.transpose() * st1.R.transpose() * st2.R;
  Vec3 exp_v = st1.R.transpose() * (st2.v - st1.v - dtime * st1.g);
  Vec3 res_v = exp_v - v_correct;
  Vec3 exp_t = st1.R.transpose() * (st2.p - st1.p - st1.v * dtime - 0.5 * dtime * dtime * st1.g);
  Vec3 res_t = exp_t - t_correct;
  Vec3 res_bg = st2.bg - st1.bg;
  Vec3 res_ba = st2.ba - st1.ba;
  double b_wei = 1;

  rr.setBlock<3, 1>(0, 0, Log(res_r));
  rr.setBlock<3, 1>(3, 0, res_t);
  rr.setBlock<3, 1>(6, 0, res_v);
  rr.setBlock<3, 1>(9, 0, res_bg * b_wei);
  rr.setBlock<3, 1>(12, 0, res_ba * b_wei);

  Mat15 cov_inv = inverse(cov);

  if (jac_enable)
  {
    Mat3 JR_inv = jr_inv(res_r);
    joca.setBlock<3, 3>(0, 0, -JR_inv * st2.R.transpose() * st1.R);
    jocb.setBlock<3, 3>(0, 0, JR_inv);
    joca.setBlock<3, 3>(0, 9, -JR_inv * res_r.transpose() * jr(R_bg * dbg) * R_bg);

    joca.setBlock<3, 3>(3, 0, hat(exp_t));
    joca.setBlock<3, 3>(3, 3, -st1.R.transpose());
    joca.setBlock<3, 3>(3, 6, -st1.R.transpose() * dtime);
    joca.setBlock<3, 3>(3, 9, -p_bg);
    joca.setBlock<3, 3>(3, 12, -p_ba);
    jocb.setBlock<3, 3>(3, 3, st1.R.transpose());

    joca.setBlock<3, 3>(6, 0, hat(exp_v));
    joca.setBlock<3, 3>(6, 6, -st1.R.transpose());
    joca.setBlock<3, 3>(6, 9, -v_bg);
    joca.setBlock<3, 3>(6, 12, -v_ba);
    jocb.setBlock<3, 3>(6, 6, st1.R.transpose());

    joca.setBlock<3, 3>(9, 9, -I33 * b_wei);
    joca.setBlock<3, 3>(12, 12, -I33 * b_wei);
    jocb.setBlock<3, 3>(9, 9, I33 * b_wei);
    jocb.setBlock<3, 3>(12, 12, I33 * b_wei);

    jocg.setBlock<3, 3>(3, 0, st1.R.transpose() * (-0.5 * dtime * dtime));
    jocg.setBlock<3, 3>(6, 0, st1.R.transpose() * (-dtime));

    Mat<15, 33> joc;
    joc.setBlock<15, 15>(0, 0, joca);
    joc.setBlock<15, 15>(0, 15, jocb);
    joc.setBlock<15, 3>(0, 30, jocg);
    jtj = joc.transpose() * cov_inv * joc;
    gg = joc.transpose() * cov_inv * rr;
  }
  return dot(rr, Vec15(cov_inv * rr));
}

// ---- optimizers.cpp:624-826: LI_BA_OptimizerGravity -----------------------------------------------------------
namespace
{
const int DIM = 15, DVEL = 6;

double divide_thread_g(int win_size, std::vector<IMUST>& x_stats, LidarFactor& voxhess, std::deque<IMU_PRE*>& imus_factor,
                       std::vector<double>& Hess, std::vector<double>& JacT, double imu_coef)
{
  const int imu_leng = win_size * DIM + 3;
  int thd_num = 5;
  double residual = 0;
  Hess.assign((size_t)imu_leng * imu_leng, 0.0);
  JacT.assign(imu_leng, 0.0);
  std::vector<std::vector<double>> hessians(thd_num), jacobins(thd_num);
  std::vector<double> resis(thd_num, 0);
  int tthd_num = thd_num;
  int g_size = (int)voxhess.plvec_voxels.size();
  if (g_size < tthd_num) tthd_num = 1;
  double part = 1.0 * g_size / tthd_num;
  std::vector<std::thread*> mthreads(tthd_num, nullptr);
  for (int i = 1; i < tthd_num; i++)
    mthreads[i] = new std::thread(&LidarFactor::acc_evaluate2, &voxhess, x_stats, (int)(part * i), (int)(part * (i + 1)),
                                  std::ref(hessians[i]), std::ref(jacobins[i]), std::ref(resis[i]));
  auto H = [&](int r, int c) -> double& { return Hess[r + (size_t)imu_leng * c]; };
  Mat<33, 33> jtj;
  Mat<33, 1> gg;
  const int g0 = imu_leng - 3;
  for (int i = 0; i < win_size - 1; i++)
  {
    jtj.setZero();
    gg.setZero();
    residual += imus_factor[i]->give_evaluate_g(x_stats[i], x_stats[i + 1], jtj, gg, true);
    for (int c = 0; c < 2 * DIM; c++)
      for (int r = 0; r < 2 * DIM; r++) H(i * DIM + r, i * DIM + c) = H(i * DIM + r, i * DIM + c) + jtj(r, c);
    for (int c = 0; c < 3; c++)
      for (int r = 0; r < 2 * DIM; r++) H(i * DIM + r, g0 + c) = H(i * DIM + r, g0 + c) + jtj(r, 2 * DIM + c);
    for (int c = 0; c < 2 * DIM; c++)
      for (int r = 0; r < 3; r++) H(g0 + r, i * DIM + c) = H(g0 + r, i * DIM + c) + jtj(2 * DIM + r, c);
    for (int c = 0; c < 3; c++)
      for (int r = 0; r < 3; r++) H(g0 + r, g0 + c) = H(g0 + r, g0 + c) + jtj(2 * DIM + r, 2 * DIM + c);
    for (int r = 0; r < 2 * DIM; r++) JacT[i * DIM + r] = JacT[i * DIM + r] + gg[r];
    for (int r = 0; r < 3; r++) JacT[g0 + r] = JacT[g0 + r] + gg[2 * DIM + r];
  }
  for (double& h : Hess) h = h * imu_coef;
  for (double& j : JacT) j = j * imu_coef;
  residual *= (imu_coef * 0.5);

  const int jac_leng = win_size * DVEL;
  for (int i = 0; i < tthd_num; i++)
  {
    if (i != 0)
    {
      mthreads[i]->join();
      delete mthreads[i];
    }
    else
      voxhess.acc_evaluate2(x_stats, 0, (int)part, hessians[0], jacobins[0], resis[0]);
    const std::vector<double>& hs = hessians[i];
    const std::vector<double>& js = jacobins[i];
    for (int a = 0; a < win_size; a++)
    {
      for (int r = 0; r < DVEL; r++) JacT[a * DIM + r] = JacT[a * DIM + r] + js[a * DVEL + r];
      for (int b = 0; b < win_size; b++)
        for (int c = 0; c < DVEL; c++)
          for (int r = 0; r < DVEL; r++)
            H(a * DIM + r, b * DIM + c) = H(a * DIM + r, b * DIM + c) + hs[(a * DVEL + r) + (size_t)jac_leng * (b * DVEL + c)];
    }
    residual += resis[i];
  }
  return residual;
}

double only_residual_g(int win_size, std::vector<IMUST>& x_stats, LidarFactor& voxhess, std::deque<IMU_PRE*>& imus_factor,
                       double imu_coef)
{
  double residual1 = 0, residual2 = 0;
  Mat<33, 33> jtj;
  Mat<33, 1> gg;
  int thd_num = 5;
  std::vector<double> residuals(thd_num, 0);
  int g_size = (int)voxhess.plvec_voxels.size();
  if (g_size < thd_num) thd_num = 1;
  std::vector<std::thread*> mthreads(thd_num, nullptr);
  double part = 1.0 * g_size / thd_num;
  for (int i = 1; i < thd_num; i++)
    mthreads[i] = new std::thread(&LidarFactor::evaluate_only_residual, &voxhess, x_stats, (int)(part * i),
                                  (int)(part * (i + 1)), std::ref(residuals[i]));
  for (int i = 0; i < win_size - 1; i++)
    residual1 += imus_factor[i]->give_evaluate_g(x_stats[i], x_stats[i + 1], jtj, gg, false);
  residual1 *= (imu_coef * 0.5);
  for (int i = 0; i < thd_num; i++)
  {
    if (i != 0)
    {
      mthreads[i]->join();
      delete mthreads[i];
    }
    else
      voxhess.evaluate_only_residual(x_stats, (int)(part * i), (int)(part * (i + 1)), residuals[i]);
    residual2 += residuals[i];
  }
  return (residual1 + residual2);
}
}  // namespace

// optimizers.cpp:746-826
void ba_damping_iter_gravity(std::vector<IMUST>& x_stats, LidarFactor& voxhess, std::deque<IMU_PRE*>& imus_factor,
                             std::vector<double>& resis, int max_iter, double imu_coef)
{
  const int win_size = voxhess.win_size;
  const int imu_leng = win_size * DIM + 3;
  double u = 0.01, v = 2;
  std::vector<double> D((size_t)imu_leng * imu_leng, 0.0), Hess, JacT, dxi(imu_leng);
  for (int i = 0; i < imu_leng; i++) D[i + (size_t)imu_leng * i] = 1.0;
  double residual1 = 0, residual2 = 0, q;
  bool is_calc_hess = true;
  std::vector<IMUST> x_stats_temp = x_stats;
  auto H = [&](int r, int c) -> double& { return Hess[r + (size_t)imu_leng * c]; };
  for (int i = 0; i < max_iter; i++)
  {
    if (is_calc_hess) residual1 = divide_thread_g(win_size, x_stats, voxhess, imus_factor, Hess, JacT, imu_coef);
    if (i == 0) resis.push_back(residual1);
    for (int c = 0; c < imu_leng; c++)
      for (int r = 0; r < 6; r++) H(r, c) = 0.0;  // topRows(6).setZero()
    for (int c = 0; c < 6; c++)
      for (int r = 0; r < imu_leng; r++) H(r, c) = 0.0;  // leftCols(6).setZero()
    for (int c = 0; c < 6; c++)
      for (int r = 0; r < 6; r++) H(r, c) = r == c ? 1.0 : 0.0;
    for (int r = 0; r < 6; r++) JacT[r] = 0.0;
    for (int k = 0; k < imu_leng; k++) D[k + (size_t)imu_leng * k] = H(k, k);
    std::vector<double> A((size_t)imu_leng * imu_leng), nb(imu_leng);
    for (size_t k = 0; k < A.size(); k++) A[k] = Hess[k] + u * D[k];
    for (int k = 0; k < imu_leng; k++) nb[k] = -JacT[k];
    dxi = ldlt_solve(A, imu_leng, nb);

    x_stats_temp[0].g += V3(dxi[imu_leng - 3], dxi[imu_leng - 2], dxi[imu_leng - 1]);
    for (int j = 0; j < win_size; j++)
    {
      Vec3 d0 = V3(dxi[DIM * j], dxi[DIM * j + 1], dxi[DIM * j + 2]);
      x_stats_temp[j].R = x_stats[j].R * Exp(d0);
      x_stats_temp[j].p = x_stats[j].p + V3(dxi[DIM * j + 3], dxi[DIM * j + 4], dxi[DIM * j + 5]);
      x_stats_temp[j].v = x_stats[j].v + V3(dxi[DIM * j + 6], dxi[DIM * j + 7], dxi[DIM * j + 8]);
      x_stats_temp[j].bg = x_stats[j].bg + V3(dxi[DIM * j + 9], dxi[DIM * j + 10], dxi[DIM * j + 11]);
      x_stats_temp[j].ba = x_stats[j].ba + V3(dxi[DIM * j + 12], dxi[DIM * j + 13], dxi[DIM * j + 14]);
      x_stats_temp[j].g = x_stats_temp[0].g;
    }
    for (int j = 0; j < win_size - 1; j++)
    {
      Vec15 dj;
      for (int k = 0; k < DIM; k++) dj[k] = dxi[DIM * j + k];
      imus_factor[j]->update_state(dj);
    }
    double q1;
    {
      std::vector<double> w(imu_leng);
      for (int r = 0; r < imu_leng; r++)
      {
        double s = (u * D[r]) * dxi[0];
        for (int k = 1; k < imu_leng; k++) s = s + (u * D[r + (size_t)imu_leng * k]) * dxi[k];
        w[r] = s - JacT[r];
      }
      double s = dxi[0] * w[0];
      for (int k = 1; k < imu_leng; k++) s = s + dxi[k] * w[k];
      q1 = 0.5 * s;
    }
    residual2 = only_residual_g(win_size, x_stats_temp, voxhess, imus_factor, imu_coef);
    q = (residual1 - residual2);
    if (q > 0)
    {
      x_stats = x_stats_temp;
      double one_three = 1.0 / 3;
      q = q / q1;
      v = 2;
      q = 1 - std::pow(2 * q - 1, 3);
      u *= (q < one_three ? one_three : q);
      is_calc_hess = true;
    }
    else
    {
      u = u * v;
      v = 2 * v;
      is_calc_hess = false;
      for (int j = 0; j < win_size - 1; j++)
      {
        imus_factor[j]->dbg = imus_factor[j]->dbg_buf;
        imus_factor[j]->dba = imus_factor[j]->dba_buf;
      }
    }
    if (std::fabs((residual1 - residual2) / residual1) < 1e-6) break;
  }
  resis.push_back(residual2);
}

// ---- initialization.cpp:28-156 --------------------------------------------------------------------------------
// Eigen::AngleAxisd(angle, axis).toRotationMatrix() (Eigen 3.4.0 Geometry/AngleAxis.h; ref_shim/mini_eigen.hpp)
static Mat3 angle_axis_to_matrix(double ang, const Vec3& ax)
{
  Mat3 res;
  const double s = std::sin(ang), c = std::cos(ang);
  const Vec3 sin_axis = V3(s * ax[0], s * ax[1], s * ax[2]);
  const Vec3 cos1_axis = V3((1.0 - c) * ax[0], (1.0 - c) * ax[1], (1.0 - c) * ax[2]);
  double tmp;
  tmp = cos1_axis[0] * ax[1];
  res(0, 1) = tmp - sin_axis[2];
  res(1, 0) = tmp + sin_axis[2];
  tmp = cos1_axis[0] * ax[2];
  res(0, 2) = tmp + sin_axis[1];
  res(2, 0) = tmp - sin_axis[1];
  tmp = cos1_axis[1] * ax[2];
  res(1, 2) = tmp - sin_axis[0];
  res(2, 1) = tmp + sin_axis[0];
  res(0, 0) = cos1_axis[0] * ax[0] + c;
  res(1, 1) = cos1_axis[1] * ax[1] + c;
  res(2, 2) = cos1_axis[2] * ax[2] + c;
  return res;
}

static void align_gravity(std::vector<IMUST>& xs)
{
  Vec3 g0 = xs[0].g;
  Vec3 n0 = normalized(g0);
  Vec3 n1 = V3(0, 0, 1);
  if (n0[2] < 0) n1[2] = -1;
  Vec3 rotvec = cross(n0, n1);
  double rnorm = norm(rotvec);
  rotvec = rotvec / rnorm;
  Mat3 rot = angle_axis_to_matrix(std::asin(rnorm), rotvec);
  g0 = rot * g0;
  Vec3 p0 = xs[0].p;
  for (size_t i = 0; i < xs.size(); i++)
  {
    xs[i].p = rot * (xs[i].p - p0) + p0;
    xs[i].R = rot * xs[i].R;
    xs[i].v = rot * xs[i].v;
    xs[i].g = g0;
  }
}

// Initialization::motion_blur (initialization.cpp:64-156): backward integration from the frame's end state
static void init_motion_blur(Cloud& pl, PVec& pvec, IMUST xc, IMUST xl, std::deque<ImuSample>& imus, double pcl_beg_time,
                             IMUST& extrin_para, double scale_gravity, int point_notime)
{
  xc.bg = xl.bg;
  xc.ba = xl.ba;
  Vec3 acc_imu, angvel_avr, acc_avr, vel_imu(xc.v), pos_imu(xc.p);
  Mat3 R_imu(xc.R);
  std::vector<IMUST> imu_poses;
  for (size_t it = imus.size() - 1; it != 0; it--)
  {
    ImuSample& head = imus[it - 1];
    ImuSample& tail = imus[it];
    angvel_avr = V3(0.5 * (head.gyr[0] + tail.gyr[0]), 0.5 * (head.gyr[1] + tail.gyr[1]),
                    0.5 * (head.gyr[2] + tail.gyr[2]));
    acc_avr = V3(0.5 * (head.acc[0] + tail.acc[0]), 0.5 * (head.acc[1] + tail.acc[1]),
                 0.5 * (head.acc[2] + tail.acc[2]));
    angvel_avr -= xc.bg;
    acc_avr = acc_avr * scale_gravity - xc.ba;
    double dt = head.t - tail.t;
    Mat3 Exp_f = Exp(angvel_avr, dt);
    acc_imu = R_imu * acc_avr + xc.g;
    pos_imu = pos_imu + vel_imu * dt + 0.5 * acc_imu * dt * dt;
    vel_imu = vel_imu + acc_imu * dt;
    R_imu = R_imu * Exp_f;
    double offt = head.t - pcl_beg_time;
    IMUST pose;
    pose.t = offt;
    pose.R = R_imu;
    pose.p = pos_imu;
    pose.v = vel_imu;
    pose.bg = angvel_avr;
    pose.ba = acc_imu;
    imu_poses.push_back(pose);
  }
  pointVar pv;
  pv.var.setIdentity();
  if (point_notime)
  {
    for (PointXYZT& ap : pl)
    {
      pv.pnt = V3(ap.x, ap.y, ap.z);
      pv.pnt = extrin_para.R * pv.pnt + extrin_para.p;
      pvec.push_back(pv);
    }
    return;
  }
  long it_pcl = (long)pl.size() - 1;
  for (size_t k = 0; k < imu_poses.size(); k++)
  {
    IMUST& head = imu_poses[k];
    R_imu = head.R;
    acc_imu = head.ba;
    vel_imu = head.v;
    pos_imu = head.p;
    angvel_avr = head.bg;
    for (; pl[it_pcl].curvature > head.t; it_pcl--)
    {
      double dt = pl[it_pcl].curvature - head.t;
      Mat3 R_i = R_imu * Exp(angvel_avr, dt);
      Vec3 T_ei = pos_imu + vel_imu * dt + 0.5 * acc_imu * dt * dt - xc.p;
      Vec3 P_i = V3(pl[it_pcl].x, pl[it_pcl].y, pl[it_pcl].z);
      Vec3 P_compensate = xc.R.transpose() * (R_i * (extrin_para.R * P_i + extrin_para.p) + T_ei);
      pv.pnt = P_compensate;
      pvec.push_back(pv);
      if (it_pcl == 0) break;
    }
  }
}

// ---- initialization.cpp:158-367 -------------------------------------------------------------------------------
int Odom::motion_init()
{
  std::vector<Vec3> pw;
  int converge_flag = 0;
  double min_eigen_value_orig = G.min_eigen_value;
  double thre_orig[4];
  for (int k = 0; k < 4; k++) thre_orig[k] = G.plane_eigen_value_thre[k];
  G.min_eigen_value = 0.02;
  for (int k = 0; k < 4; k++) G.plane_eigen_value_thre[k] = 1.0 / 4;
  double converge_thre = 0.05;
  bool is_degrade = true;
  Vec3 eigvalue = Vec3::Zero();
  const int win_size = G.win_size;
  init_rounds = 0;
  for (int iterCnt = 0; iterCnt < 10; iterCnt++)
  {
    init_rounds++;
    if (converge_flag == 1)
    {
      G.min_eigen_value = min_eigen_value_orig;
      for (int k = 0; k < 4; k++) G.plane_eigen_value_thre[k] = thre_orig[k];
    }
    clear_map();
    for (int i = 0; i < win_size; i++)
    {
      pw.clear();
      pvec_buf[i]->clear();
      int l = i == 0 ? i : i - 1;
      init_motion_blur(*pl_origs[i], *pvec_buf[i], x_buf[i], x_buf[l], vec_imus[i], beg_times[i], extrin_para,
                       ba_noise.scale_gravity, odom_ekf.point_notime);
      if (converge_flag == 1)
      {
        for (pointVar& pv : *pvec_buf[i]) calcBodyVar(pv.pnt, G.dept_err, G.beam_err, pv.var);
        pvec_update(pvec_buf[i], x_buf[i], pw);
      }
      else
      {
        for (pointVar& pv : *pvec_buf[i]) pw.push_back(x_buf[i].R * pv.pnt + x_buf[i].p);
      }
      cut_voxel(&G, surf_map, pvec_buf[i], i, surf_map_slide, win_size, pw, sws[0]);
    }
    init_voxhess.clear();
    init_voxhess.win_size = win_size;
    init_nodes.clear();
    for (auto iter = surf_map.begin(); iter != surf_map.end(); ++iter)
    {
      iter->second->recut(win_size, x_buf, sws[0]);
      tras_opt_collect(iter->second, init_voxhess, &init_nodes);
    }
    for (size_t a = 0; a < init_nodes.size(); a++) init_nodes[a]->opt_state = (int)a;  // tras_opt: opt_state = index
    if (init_voxhess.plvec_voxels.size() < 10) break;
    std::vector<double> resis;
    ba_damping_iter_gravity(x_buf, init_voxhess, imu_pre_buf, resis, 3, imu_coef);
    Mat3 nnt = Mat3::Zero();
    for (int i = 0; i < win_size - 1; i++) delete imu_pre_buf[i];
    imu_pre_buf.clear();
    for (int i = 1; i < win_size; i++)
    {
      imu_pre_buf.push_back(new IMU_PRE(x_buf[i - 1].bg, x_buf[i - 1].ba));
      imu_pre_buf.back()->push_imu(vec_imus[i], ba_noise);
    }
    if (std::fabs(resis[0] - resis[1]) / resis[0] < converge_thre && iterCnt >= 2)
    {
      for (Mat3& m : init_voxhess.eig_vectors)
      {
        Vec3 v3 = m.block<3, 1>(0, 0);
        for (int c = 0; c < 3; c++)
          for (int r = 0; r < 3; r++) nnt(r, c) = nnt(r, c) + v3[r] * v3[c];
      }
      SelfAdjointEigen3 saes(nnt);
      eigvalue = saes.values;
      is_degrade = eigvalue[0] < 15 ? true : false;
      converge_thre = 0.01;
      if (converge_flag == 0)
      {
        align_gravity(x_buf);
        converge_flag = 1;
        continue;
      }
      else
        break;
    }
  }
  x_curr = x_buf[win_size - 1];
  double gnm = norm(x_curr.g);
  init_eig = eigvalue;
  if (is_degrade) converge_flag = 0;
  if (gnm < 9.6 || gnm > 10.0) converge_flag = 0;
  if (converge_flag == 0) clear_map();
  pl_origs.clear();
  vec_imus.clear();
  beg_times.clear();
  return converge_flag;
}

// the map teardown motion_init repeats (initialization.cpp:202-217, 319-334)
void Odom::clear_map()
{
  std::vector<OctoTree*> octos;
  for (auto iter = surf_map.begin(); iter != surf_map.end(); ++iter)
  {
    tras_ptr(iter->second, octos);
    iter->second->clear_slwd(sws[0]);
    delete iter->second;
  }
  for (size_t i = 0; i < octos.size(); i++) delete octos[i];
  surf_map.clear();
  surf_map_slide.clear();
}

// ---- node.cpp:293-408 + local_mapping.cpp:362-388 -------------------------------------------------------------
int Odom::initialization(std::deque<ImuSample>& imus, Cloud& pcl_curr)
{
  std::shared_ptr<Cloud> orig(new Cloud(pcl_curr));
  if (odom_ekf.process(x_curr, pcl_curr, imus) == 0) return 0;
  if (win_count == 0) ba_noise.scale_gravity = odom_ekf.scale_gravity;
  PVecPtr pptr(new PVec);
  double downkd = G.down_size >= 0.5 ? G.down_size : 0.5;
  down_sampling_voxel(pcl_curr, downkd);
  var_init(extrin_para, pcl_curr, pptr, G.dept_err, G.beam_err);
  lio_state_estimation_kdtree(pptr);
  pwld.clear();
  pvec_update(pptr, x_curr, pwld);
  win_count++;
  x_buf.push_back(x_curr);
  pvec_buf.push_back(pptr);
  if (win_count > 1)
  {
    imu_pre_buf.push_back(new IMU_PRE(x_buf[win_count - 2].bg, x_buf[win_count - 2].ba));
    imu_pre_buf[win_count - 2]->push_imu(imus, ba_noise);
  }
  Cloud pl_mid = *orig;
  down_sampling_close(*orig, G.down_size);
  if (orig->size() < 1000)
  {
    *orig = pl_mid;
    down_sampling_close(*orig, G.down_size / 2);
  }
  std::sort(orig->begin(), orig->end(), [](const PointXYZT& x, const PointXYZT& y) { return x.curvature < y.curvature; });
  pl_origs.push_back(orig);
  beg_times.push_back(odom_ekf.pcl_beg_time);
  vec_imus.push_back(imus);
  if (win_count >= G.win_size) return motion_init() == 0 ? -1 : 1;
  return 0;
}

void Odom::system_reset(std::deque<ImuSample>& imus)
{
  std::vector<OctoTree*> octos;
  for (auto iter = surf_map.begin(); iter != surf_map.end(); iter++)
  {
    tras_ptr(iter->second, octos);
    iter->second->clear_slwd(sws[0]);
    delete iter->second;
  }
  for (OctoTree* ot : octos) delete ot;
  surf_map.clear();
  surf_map_slide.clear();
  x_curr.setZero();
  x_curr.p = V3(0, 0, 30);
  odom_ekf.mean_acc.setZero();
  odom_ekf.init_num = 0;
  odom_ekf.IMU_init(imus);
  x_curr.g = (-1.0 * odom_ekf.mean_acc) * ba_noise.scale_gravity;
  for (size_t i = 0; i < imu_pre_buf.size(); i++) delete imu_pre_buf[i];
  x_buf.clear();
  pvec_buf.clear();
  imu_pre_buf.clear();
  pl_tree.clear();
  for (int i = 0; i < G.win_size; i++) G.mp[i] = i;
  win_base = 0;
  win_count = 0;
}

int Odom::init_scan(Cloud& pcl_curr, double beg, std::deque<ImuSample>& imus)
{
  odom_ekf.pcl_beg_time = beg;
  odom_ekf.pcl_end_time = beg + pcl_curr.back().curvature;  // sync.cpp:40
  int init = initialization(imus, pcl_curr);
  if (init == 1)
  {
    window_tail(&init_voxhess, &init_nodes);
    return 1;
  }
  if (init == -1) system_reset(imus);
  return init;
}
}  // namespace vo

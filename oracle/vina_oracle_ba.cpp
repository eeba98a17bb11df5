// ORACLE — TEST INFRASTRUCTURE ONLY (pinned against the reference build oracle/_ref: see oracle/README.md).
// The sliding-window LiDAR-inertial BA of the reference, restated:
//   * IMU_PRE            src/estimation/imu_preintegration.cpp:7-163 (+ update_state :235-242)
//   * LI_BA_Optimizer    src/mapping/optimizers.cpp:171-245, 340-376, 430-517 (LiDAR + IMU factors, the overload
//                        thd_odometry_localmapping calls, local_mapping.cpp:492-497)
// Expressions keep the reference's grouping and omat.hpp's eager left-to-right products, which is what the
// reference build evaluates (ref_shim/mini_eigen.hpp) - bit-identical in the strict builds.
// Third-party arithmetic restated on both sides (Eigen is absent) - parity UNPINNED at
// these calls: Matrix<15,15>::inverse() (PartialPivLU, omat.hpp), LDLT::solve (pivoted LDL^T, ldlt_solve below),
// AngleAxisd(Matrix3d) (trace / antisymmetric-part formula, angle_axis below).
#include <cmath>
#include <thread>

#include "vina_oracle.hpp"

namespace vo
{
// (jr, jr_inv, angle_axis and ldlt_solve are shared with vina_oracle_init.cpp)

// include/vina_slam/core/math.hpp:57-71
Mat3 jr(Vec3 vec)
{
  double ang = norm(vec);
  if (ang < 1e-9) return Mat3::Identity();
  vec = vec / ang;
  double ra = std::sin(ang) / ang;
  return ra * Mat3::Identity() + (1 - ra) * vec * vec.transpose() - (1 - std::cos(ang)) / ang * hat(vec);
}
// Eigen::AngleAxisd(const Matrix3d&) as restated in ref_shim/mini_eigen.hpp
void angle_axis(const Mat3& R, Vec3& ax, double& ang)
{
  double c = 0.5 * (trace(R) - 1.0);
  c = std::max(-1.0, std::min(1.0, c));
  ang = std::acos(c);
  Vec3 k = V3(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1));
  double n = norm(k);
  ax = n > 0 ? Vec3(k / n) : V3(1, 0, 0);
}
// math.hpp:73-88
Mat3 jr_inv(const Mat3& rotR)
{
  Vec3 axi;
  double ang;
  angle_axis(rotR, axi, ang);
  if (ang < 1e-9) return Mat3::Identity();
  double ctt = ang / 2 / std::tan(ang / 2);
  return ctt * Mat3::Identity() + (1 - ctt) * axi * axi.transpose() + ang / 2 * hat(axi);
}

// (A).ldlt().solve(b) - the routine of ref_shim/mini_eigen.hpp (DynLDLT), same operation order
std::vector<double> ldlt_solve(std::vector<double> L, int n, const std::vector<double>& b)
{
  auto a = [&](int i, int j) -> double& { return L[i + (size_t)j * n]; };
  std::vector<int> perm(n);
  std::vector<double> temp(n);
  for (int k = 0; k < n; k++)
  {
    int p = k;
    double best = std::fabs(a(k, k));
    for (int i = k + 1; i < n; i++)
      if (std::fabs(a(i, i)) > best)
      {
        best = std::fabs(a(i, i));
        p = i;
      }
    perm[k] = p;
    if (p != k)
    {
      for (int j = 0; j < k; j++) std::swap(a(k, j), a(p, j));
      for (int i = p + 1; i < n; i++) std::swap(a(i, k), a(i, p));
      std::swap(a(k, k), a(p, p));
      for (int i = k + 1; i < p; i++) std::swap(a(i, k), a(p, i));
    }
    if (k > 0)
    {
      for (int j = 0; j < k; j++) temp[j] = a(j, j) * a(k, j);
      double s = a(k, k);
      for (int j = 0; j < k; j++) s = s - a(k, j) * temp[j];
      a(k, k) = s;
      for (int i = k + 1; i < n; i++)
      {
        double t = a(i, k);
        for (int j = 0; j < k; j++) t = t - a(i, j) * temp[j];
        a(i, k) = t;
      }
    }
    const double d = a(k, k);
    if (std::fabs(d) > 0.0)
      for (int i = k + 1; i < n; i++) a(i, k) = a(i, k) / d;
  }
  std::vector<double> x = b;
  for (int k = 0; k < n; k++)
    if (perm[k] != k) std::swap(x[k], x[perm[k]]);
  for (int i = 0; i < n; i++)
  {
    double s = x[i];
    for (int j = 0; j < i; j++) s = s - a(i, j) * x[j];
    x[i] = s;
  }
  for (int i = 0; i < n; i++)
  {
    const double d = a(i, i);
    x[i] = std::fabs(d) > 2.2250738585072014e-308 ? x[i] / d : 0.0;
  }
  for (int i = n - 1; i >= 0; i--)
  {
    double s = x[i];
    for (int j = i + 1; j < n; j++) s = s - a(j, i) * x[j];
    x[i] = s;
  }
  for (int k = n - 1; k >= 0; k--)
    if (perm[k] != k) std::swap(x[k], x[perm[k]]);
  return x;
}


// ---- imu_preintegration.cpp ---------------------------------------------------------------------------------
IMU_PRE::IMU_PRE(const Vec3& bg1, const Vec3& ba1)
{
  bg = bg1;
  ba = ba1;
  R_delta.setIdentity();
  p_delta.setZero();
  v_delta.setZero();
  R_bg.setZero();
  p_bg.setZero();
  p_ba.setZero();
  v_bg.setZero();
  v_ba.setZero();
  dtime = 0;
  dbg.setZero();
  dba.setZero();
  dbg_buf.setZero();
  dba_buf.setZero();
  cov.setZero();
}

// imu_preintegration.cpp:32-57; stamps are integer nanoseconds in the reference (rclcpp::Time)
void IMU_PRE::push_imu(std::deque<ImuSample>& imu_buffer, const BaNoise& nz)
{
  Vec3 cur_gyr, cur_acc;
  for (size_t k = 1; k < imu_buffer.size(); k++)
  {
    const ImuSample& imu_prev = imu_buffer[k - 1];
    const ImuSample& imu_curr = imu_buffer[k];
    double dt = imu_curr.t - imu_prev.t;
    cur_gyr = V3(0.5 * (imu_prev.gyr[0] + imu_curr.gyr[0]), 0.5 * (imu_prev.gyr[1] + imu_curr.gyr[1]),
                 0.5 * (imu_prev.gyr[2] + imu_curr.gyr[2]));
    cur_acc = V3(0.5 * (imu_prev.acc[0] + imu_curr.acc[0]), 0.5 * (imu_prev.acc[1] + imu_curr.acc[1]),
                 0.5 * (imu_prev.acc[2] + imu_curr.acc[2]));
    cur_gyr = cur_gyr - bg;
    cur_acc = cur_acc * nz.scale_gravity - ba;
    add_imu(cur_gyr, cur_acc, dt, nz);
  }
}

// imu_preintegration.cpp:59-100
void IMU_PRE::add_imu(Vec3& cur_gyr, Vec3& cur_acc, double dt, const BaNoise& nz)
{
  dtime += dt;
  Mat3 rotation_increment = Exp(cur_gyr, dt);
  Mat3 right_jacobian = jr(cur_gyr * dt);
  Mat3 rotation_dt = dt * R_delta;
  Mat3 rotation_dt2_half = 0.5 * dt * dt * R_delta;
  Mat3 acc_skew = hat(cur_acc);

  p_ba = p_ba + v_ba * dt - rotation_dt2_half;
  p_bg = p_bg + v_bg * dt - rotation_dt2_half * acc_skew * R_bg;
  v_ba = v_ba - rotation_dt;
  v_bg = v_bg - rotation_dt * acc_skew * R_bg;
  R_bg = rotation_increment.transpose() * R_bg - right_jacobian * dt;

  Mat9 jacobian_a = Mat9::Identity();
  Mat<9, 6> jacobian_b = Mat<9, 6>::Zero();
  jacobian_a.setBlock<3, 3>(0, 0, rotation_increment.transpose());
  jacobian_a.setBlock<3, 3>(3, 0, -rotation_dt2_half * acc_skew);
  jacobian_a.setBlock<3, 3>(3, 6, Mat3::Identity() * dt);
  jacobian_a.setBlock<3, 3>(6, 0, -rotation_dt * acc_skew);
  jacobian_b.setBlock<3, 3>(0, 0, right_jacobian * dt);
  jacobian_b.setBlock<3, 3>(3, 3, rotation_dt2_half);
  jacobian_b.setBlock<3, 3>(6, 3, rotation_dt);

  Mat9 c99 = cov.block<9, 9>(0, 0);
  cov.setBlock<9, 9>(0, 0, jacobian_a * c99 * jacobian_a.transpose() + jacobian_b * nz.noiseMeas * jacobian_b.transpose());
  Mat6 c66 = cov.block<6, 6>(9, 9);
  cov.setBlock<6, 6>(9, 9, c66 + nz.noiseWalk * dt);

  p_delta += v_delta * dt + rotation_dt2_half * cur_acc;
  v_delta += rotation_dt * cur_acc;
  R_delta = R_delta * rotation_increment;
}

// imu_preintegration.cpp:102-163
double IMU_PRE::give_evaluate(IMUST& st1, IMUST& st2, Mat<30, 30>& jtj, Mat<30, 1>& gg, bool jac_enable)
{
  Mat15 joca, jocb;
  Vec15 rr;
  joca.setZero();
  jocb.setZero();
  rr.setZero();
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

    Mat<15, 30> joc;
    joc.setBlock<15, 15>(0, 0, joca);
    joc.setBlock<15, 15>(0, 15, jocb);
    jtj = joc.transpose() * cov_inv * joc;
    gg = joc.transpose() * cov_inv * rr;
  }
  return dot(rr, Vec15(cov_inv * rr));
}

// imu_preintegration.cpp:235-242
void IMU_PRE::update_state(const Vec15& dxi)
{
  dbg_buf = dbg;
  dba_buf = dba;
  dbg += dxi.block<3, 1>(9, 0);
  dba += dxi.block<3, 1>(12, 0);
}

// ---- optimizers.cpp: LI_BA_Optimizer -------------------------------------------------------------------------
namespace
{
const int DIM = 15, DVEL = 6;

// optimizers.cpp:181-245
double divide_thread(int win_size, std::vector<IMUST>& x_stats, LidarFactor& voxhess, std::deque<IMU_PRE*>& imus_factor,
                     std::vector<double>& Hess, std::vector<double>& JacT, double imu_coef)
{
  const int imu_leng = win_size * DIM;
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
  Mat<30, 30> jtj;
  Mat<30, 1> gg;
  for (int i = 0; i < win_size - 1; i++)
  {
    jtj.setZero();
    gg.setZero();
    residual += imus_factor[i]->give_evaluate(x_stats[i], x_stats[i + 1], jtj, gg, true);
    for (int c = 0; c < 2 * DIM; c++)
      for (int r = 0; r < 2 * DIM; r++) H(i * DIM + r, i * DIM + c) = H(i * DIM + r, i * DIM + c) + jtj(r, c);
    for (int r = 0; r < 2 * DIM; r++) JacT[i * DIM + r] = JacT[i * DIM + r] + gg[r];
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
    // hess_plus (optimizers.cpp:171-179)
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

// optimizers.cpp:340-376
double only_residual(int win_size, std::vector<IMUST>& x_stats, LidarFactor& voxhess, std::deque<IMU_PRE*>& imus_factor,
                     double imu_coef)
{
  double residual1 = 0, residual2 = 0;
  Mat<30, 30> jtj;
  Mat<30, 1> gg;
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
    residual1 += imus_factor[i]->give_evaluate(x_stats[i], x_stats[i + 1], jtj, gg, false);
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

// optimizers.cpp:430-517
int ba_damping_iter(std::vector<IMUST>& x_stats, LidarFactor& voxhess, std::deque<IMU_PRE*>& imus_factor, double imu_coef,
                    std::vector<double>* hess_out)
{
  const int win_size = voxhess.win_size;
  const int imu_leng = win_size * DIM;
  double u = 0.01, v = 2;
  std::vector<double> D((size_t)imu_leng * imu_leng, 0.0), Hess, JacT, dxi(imu_leng);
  for (int i = 0; i < imu_leng; i++) D[i + (size_t)imu_leng * i] = 1.0;
  double residual1 = 0, residual2 = 0, q;
  bool is_calc_hess = true;
  std::vector<IMUST> x_stats_temp = x_stats;
  auto H = [&](int r, int c) -> double& { return Hess[r + (size_t)imu_leng * c]; };
  int iters = 0;
  int max_iter = 10;
  for (int i = 0; i < max_iter; i++)
  {
    iters++;
    if (is_calc_hess)
    {
      residual1 = divide_thread(win_size, x_stats, voxhess, imus_factor, Hess, JacT, imu_coef);
      if (hess_out) *hess_out = Hess;
    }
    for (int c = 0; c < imu_leng; c++)
      for (int r = 0; r < DIM; r++) H(r, c) = 0.0;  // topRows(DIM).setZero()
    for (int c = 0; c < DIM; c++)
      for (int r = 0; r < imu_leng; r++) H(r, c) = 0.0;  // leftCols(DIM).setZero()
    for (int c = 0; c < DIM; c++)
      for (int r = 0; r < DIM; r++) H(r, c) = r == c ? 1.0 : 0.0;
    for (int r = 0; r < DIM; r++) JacT[r] = 0.0;
    for (int k = 0; k < imu_leng; k++) D[k + (size_t)imu_leng * k] = H(k, k);
    // dxi = (Hess + u * D).ldlt().solve(-JacT)
    std::vector<double> A((size_t)imu_leng * imu_leng), nb(imu_leng);
    for (size_t k = 0; k < A.size(); k++) A[k] = Hess[k] + u * D[k];
    for (int k = 0; k < imu_leng; k++) nb[k] = -JacT[k];
    dxi = ldlt_solve(A, imu_leng, nb);

    for (int j = 0; j < win_size; j++)
    {
      Vec3 d0 = V3(dxi[DIM * j], dxi[DIM * j + 1], dxi[DIM * j + 2]);
      x_stats_temp[j].R = x_stats[j].R * Exp(d0);
      x_stats_temp[j].p = x_stats[j].p + V3(dxi[DIM * j + 3], dxi[DIM * j + 4], dxi[DIM * j + 5]);
      x_stats_temp[j].v = x_stats[j].v + V3(dxi[DIM * j + 6], dxi[DIM * j + 7], dxi[DIM * j + 8]);
      x_stats_temp[j].bg = x_stats[j].bg + V3(dxi[DIM * j + 9], dxi[DIM * j + 10], dxi[DIM * j + 11]);
      x_stats_temp[j].ba = x_stats[j].ba + V3(dxi[DIM * j + 12], dxi[DIM * j + 13], dxi[DIM * j + 14]);
    }
    for (int j = 0; j < win_size - 1; j++)
    {
      Vec15 dj;
      for (int k = 0; k < DIM; k++) dj[k] = dxi[DIM * j + k];
      imus_factor[j]->update_state(dj);
    }
    // q1 = 0.5 * dxi.dot(u * D * dxi - JacT): (u * D) is a full matrix product in the reference
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
    residual2 = only_residual(win_size, x_stats_temp, voxhess, imus_factor, imu_coef);
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
  return iters;
}
}  // namespace vo

// Sliding-window LiDAR-inertial bundle adjustment, host side. Mirrors
//   IMU_PRE                         src/estimation/imu_preintegration.cpp:7-163, 235-242  (pre-integration factor
//                                   between consecutive window frames; 9 of them, 15x15 algebra: stays on the host)
//   LI_BA_Optimizer::damping_iter   src/mapping/optimizers.cpp:171-245, 340-376, 430-517   (Levenberg-Marquardt over
//                                   10 frames x 15 states; the 150x150 solve stays on the host)
// and takes the LiDAR factor - the part that is data-parallel over thousands of plane voxels
// (LidarFactor::acc_evaluate2 / evaluate_only_residual, factors.cpp:22-158) - from the device kernels of
// csrc/ba_kernels.cu through vina_ba_lidar_hessian / vina_ba_lidar_residual.
// Third-party arithmetic (Eigen is not a dependency here): Matrix<15,15>::inverse() = LU with partial pivoting,
// LDLT::solve = LDL^T with diagonal pivoting, AngleAxisd(Matrix3d) = trace / antisymmetric-part formula.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <vector>

#include "../csrc/vn_ctx.h"
#include "vina_ba.h"

namespace
{
// ---- a small fixed-size column-major matrix (host only) ------------------------------------------------------
template <int R, int C>
struct HM
{
  double d[R * C];
  double& operator()(int i, int j) { return d[i + j * R]; }
  double operator()(int i, int j) const { return d[i + j * R]; }
  double& operator[](int i) { return d[i]; }
  double operator[](int i) const { return d[i]; }
  static HM zero()
  {
    HM m;
    for (int i = 0; i < R * C; i++) m.d[i] = 0.0;
    return m;
  }
  static HM eye()
  {
    HM m = zero();
    for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = 1.0;
    return m;
  }
  HM<C, R> T() const
  {
    HM<C, R> t;
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C; j++) t(j, i) = (*this)(i, j);
    return t;
  }
  template <int BR, int BC>
  HM<BR, BC> blk(int r0, int c0) const
  {
    HM<BR, BC> b;
    for (int j = 0; j < BC; j++)
      for (int i = 0; i < BR; i++) b(i, j) = (*this)(r0 + i, c0 + j);
    return b;
  }
  template <int BR, int BC>
  void set(int r0, int c0, const HM<BR, BC>& b)
  {
    for (int j = 0; j < BC; j++)
      for (int i = 0; i < BR; i++) (*this)(r0 + i, c0 + j) = b(i, j);
  }
};
template <int R, int C>
HM<R, C> operator+(const HM<R, C>& a, const HM<R, C>& b)
{
  HM<R, C> r;
  for (int i = 0; i < R * C; i++) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int R, int C>
HM<R, C> operator-(const HM<R, C>& a, const HM<R, C>& b)
{
  HM<R, C> r;
  for (int i = 0; i < R * C; i++) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int R, int C>
HM<R, C> operator-(const HM<R, C>& a)
{
  HM<R, C> r;
  for (int i = 0; i < R * C; i++) r.d[i] = -a.d[i];
  return r;
}
template <int R, int C>
HM<R, C> operator*(double s, const HM<R, C>& a)
{
  HM<R, C> r;
  for (int i = 0; i < R * C; i++) r.d[i] = s * a.d[i];
  return r;
}
template <int R, int C>
HM<R, C> operator*(const HM<R, C>& a, double s)
{
  HM<R, C> r;
  for (int i = 0; i < R * C; i++) r.d[i] = a.d[i] * s;
  return r;
}
template <int R, int K, int C>
HM<R, C> operator*(const HM<R, K>& a, const HM<K, C>& b)
{
  HM<R, C> r;
  for (int j = 0; j < C; j++)
    for (int i = 0; i < R; i++)
    {
      double s = a(i, 0) * b(0, j);
      for (int k = 1; k < K; k++) s = s + a(i, k) * b(k, j);
      r(i, j) = s;
    }
  return r;
}
typedef HM<3, 1> V3;
typedef HM<3, 3> M3;
typedef HM<15, 15> M15;
typedef HM<15, 1> V15;

V3 v3(const double* p)
{
  V3 v;
  v[0] = p[0], v[1] = p[1], v[2] = p[2];
  return v;
}
M3 m3(const double* p)
{
  M3 m;
  memcpy(m.d, p, 72);
  return m;
}
double dot(const V3& a, const V3& b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
double nrm(const V3& a) { return std::sqrt(dot(a, a)); }
M3 hat(const V3& v)
{
  M3 O = M3::zero();
  O(0, 1) = -v[2], O(0, 2) = v[1];
  O(1, 0) = v[2], O(1, 2) = -v[0];
  O(2, 0) = -v[1], O(2, 1) = v[0];
  return O;
}
// include/vina_slam/core/math.hpp:12-24, 26-41, 43-48, 57-88
M3 Exp(const V3& ang)
{
  double n = nrm(ang);
  if (n >= 1e-9)
  {
    V3 ax;
    for (int k = 0; k < 3; k++) ax[k] = ang[k] / n;
    M3 K = hat(ax);
    return M3::eye() + std::sin(n) * K + (1.0 - std::cos(n)) * K * K;
  }
  return M3::eye();
}
M3 Exp(const V3& w, double dt)
{
  double n = nrm(w);
  if (n > 1e-7)
  {
    V3 ax;
    for (int k = 0; k < 3; k++) ax[k] = w[k] / n;
    M3 K = hat(ax);
    double r = n * dt;
    return M3::eye() + std::sin(r) * K + (1.0 - std::cos(r)) * K * K;
  }
  return M3::eye();
}
V3 Log(const M3& R)
{
  double tr = (R(0, 0) + R(1, 1)) + R(2, 2);
  double theta = (tr > 3.0 - 1e-6) ? 0.0 : std::acos(0.5 * (tr - 1));
  V3 K;
  K[0] = R(2, 1) - R(1, 2), K[1] = R(0, 2) - R(2, 0), K[2] = R(1, 0) - R(0, 1);
  return (std::fabs(theta) < 0.001) ? (0.5 * K) : (0.5 * theta / std::sin(theta) * K);
}
M3 jr(V3 vec)
{
  double ang = nrm(vec);
  if (ang < 1e-9) return M3::eye();
  for (int k = 0; k < 3; k++) vec[k] = vec[k] / ang;
  double ra = std::sin(ang) / ang;
  return ra * M3::eye() + (1 - ra) * vec * vec.T() - (1 - std::cos(ang)) / ang * hat(vec);
}
M3 jr_inv(const M3& R)
{
  double c = 0.5 * (((R(0, 0) + R(1, 1)) + R(2, 2)) - 1.0);
  c = std::max(-1.0, std::min(1.0, c));
  const double ang = std::acos(c);
  V3 k;
  k[0] = R(2, 1) - R(1, 2), k[1] = R(0, 2) - R(2, 0), k[2] = R(1, 0) - R(0, 1);
  const double n = nrm(k);
  V3 axi;
  axi[0] = 1, axi[1] = 0, axi[2] = 0;
  if (n > 0)
    for (int q = 0; q < 3; q++) axi[q] = k[q] / n;
  if (ang < 1e-9) return M3::eye();
  double ctt = ang / 2 / std::tan(ang / 2);
  return ctt * M3::eye() + (1 - ctt) * axi * axi.T() + ang / 2 * hat(axi);
}
// LU with partial pivoting, inverse by substitution of the identity (Eigen: PartialPivLU for sizes > 4)
M15 inverse15(const M15& A)
{
  const int N = 15;
  M15 lu = A;
  int perm[N];
  for (int i = 0; i < N; i++) perm[i] = i;
  for (int k = 0; k < N; k++)
  {
    int piv = k;
    double best = std::fabs(lu(k, k));
    for (int i = k + 1; i < N; i++)
      if (std::fabs(lu(i, k)) > best) best = std::fabs(lu(i, k)), piv = i;
    if (piv != k)
    {
      for (int j = 0; j < N; j++) std::swap(lu(k, j), lu(piv, j));
      std::swap(perm[k], perm[piv]);
    }
    for (int i = k + 1; i < N; i++)
    {
      lu(i, k) = lu(i, k) / lu(k, k);
      for (int j = k + 1; j < N; j++) lu(i, j) = lu(i, j) - lu(i, k) * lu(k, j);
    }
  }
  M15 inv;
  for (int c = 0; c < N; c++)
  {
    double y[N];
    for (int i = 0; i < N; i++)
    {
      double s = (perm[i] == c) ? 1.0 : 0.0;
      for (int j = 0; j < i; j++) s = s - lu(i, j) * y[j];
      y[i] = s;
    }
    for (int i = N - 1; i >= 0; i--)
    {
      double s = y[i];
      for (int j = i + 1; j < N; j++) s = s - lu(i, j) * inv(j, c);
      inv(i, c) = s / lu(i, i);
    }
  }
  return inv;
}
// x = A^-1 b for the symmetric (damped) normal matrix: LDL^T with diagonal pivoting on the lower triangle.
// Right-looking with DELAYED updates: the pivots of a panel of LDLT_NB columns are eliminated against an eagerly
// maintained diagonal, and the trailing matrix receives the panel's rank-NB update in one pass - the matrix
// (146 KB for a 10-frame window) is streamed from L2 once per panel instead of once per column.
// (compiled twice, AVX2 + FMA and baseline x86-64; the loader picks one; contraction is allowed in this routine
// only: the solve is toleranced, nothing decision-bearing)
#define LDLT_NB 8
__attribute__((target_clones("avx2,fma", "default"), optimize("fp-contract=fast"))) std::vector<double> ldlt_solve(
    std::vector<double> L, int n, const std::vector<double>& b)
{
  auto a = [&](int i, int j) -> double& { return L[i + (size_t)j * n]; };
  std::vector<int> perm(n);
  std::vector<double> diag(n), W((size_t)n * LDLT_NB);  // W(i, s) = d_s * l_(i, k0 + s) of the current panel
  for (int i = 0; i < n; i++) diag[i] = a(i, i);
  for (int k0 = 0; k0 < n; k0 += LDLT_NB)
  {
    const int nb = std::min(LDLT_NB, n - k0);
    for (int t = 0; t < nb; t++)
    {
      const int k = k0 + t;
      int p = k;
      double best = std::fabs(diag[k]);
      for (int i = k + 1; i < n; i++)
        if (std::fabs(diag[i]) > best) best = std::fabs(diag[i]), p = i;
      perm[k] = p;
      if (p != k)
      {
        // symmetric swap of rows / columns k and p: finished columns, the panel's pending factors, the not yet
        // updated trailing entries and the diagonal
        for (int j = 0; j < k; j++) std::swap(a(k, j), a(p, j));
        for (int s = 0; s < t; s++) std::swap(W[k + (size_t)n * s], W[p + (size_t)n * s]);
        for (int i = p + 1; i < n; i++) std::swap(a(i, k), a(i, p));
        std::swap(diag[k], diag[p]);
        for (int i = k + 1; i < p; i++) std::swap(a(i, k), a(p, i));
      }
      // column k with the panel's earlier columns applied
      double* __restrict__ ck = &L[(size_t)k * n];
      for (int s = 0; s < t; s++)
      {
        const double f = W[k + (size_t)n * s];
        const double* __restrict__ ls = &L[(size_t)(k0 + s) * n];
        for (int i = k + 1; i < n; i++) ck[i] -= ls[i] * f;
      }
      const double d = diag[k];
      ck[k] = d;
      double* __restrict__ wt = &W[(size_t)n * t];
      if (std::fabs(d) > 0.0)
      {
        for (int i = k + 1; i < n; i++) wt[i] = ck[i];       // d * l_i
        for (int i = k + 1; i < n; i++) ck[i] = ck[i] / d;   // l_i
        for (int i = k + 1; i < n; i++) diag[i] -= ck[i] * wt[i];
      }
      else
        for (int i = k + 1; i < n; i++) wt[i] = 0.0;
    }
    // the panel's rank-nb update of the trailing columns (strictly below the diagonal: the diagonal is in diag[])
    const int j0 = k0 + nb;
    if (nb == LDLT_NB)
    {
      const double* ls[LDLT_NB];
      for (int s = 0; s < LDLT_NB; s++) ls[s] = &L[(size_t)(k0 + s) * n];
      for (int j = j0; j < n; j++)
      {
        double f[LDLT_NB];
        for (int s = 0; s < LDLT_NB; s++) f[s] = W[j + (size_t)n * s];
        double* __restrict__ cj = &L[(size_t)j * n];
        for (int i = j + 1; i < n; i++)
        {
          double acc = 0.0;
#pragma unroll
          for (int s = 0; s < LDLT_NB; s++) acc += ls[s][i] * f[s];
          cj[i] -= acc;
        }
      }
    }
    else
      for (int s = 0; s < nb; s++)
      {
        const double* __restrict__ ls = &L[(size_t)(k0 + s) * n];
        for (int j = j0; j < n; j++)
        {
          const double f = W[j + (size_t)n * s];
          double* __restrict__ cj = &L[(size_t)j * n];
          for (int i = j + 1; i < n; i++) cj[i] -= ls[i] * f;
        }
      }
  }
  std::vector<double> x = b;
  for (int k = 0; k < n; k++)
    if (perm[k] != k) std::swap(x[k], x[perm[k]]);
  for (int j = 0; j < n; j++)  // L y = b, column oriented
  {
    const double xj = x[j];
    const double* __restrict__ cj = &L[(size_t)j * n];
    for (int i = j + 1; i < n; i++) x[i] -= cj[i] * xj;
  }
  for (int i = 0; i < n; i++)
  {
    const double d = a(i, i);
    x[i] = std::fabs(d) > 2.2250738585072014e-308 ? x[i] / d : 0.0;
  }
  for (int i = n - 1; i >= 0; i--)  // L^T x = y
  {
    double s = x[i];
    const double* __restrict__ ci = &L[(size_t)i * n];
    for (int j = i + 1; j < n; j++) s -= ci[j] * x[j];
    x[i] = s;
  }
  for (int k = n - 1; k >= 0; k--)
    if (perm[k] != k) std::swap(x[k], x[perm[k]]);
  return x;
}
}  // namespace

// ---- the IMU factor between two consecutive window frames ------------------------------------------------------
// Role of the reference's IMU_PRE (src/estimation/imu_preintegration.cpp:7-163, 235-242); written from the
// pre-integration model (Forster et al., on-manifold pre-integration) with the reference's conventions, so that
// residual / J^T W J / J^T W r agree with its code to rounding (tests/test_ba_host_cpu.py, 1e-9):
//   error state of the increment  e = (dtheta, dp, dv), biases (bg, ba) as random walks;
//   per IMU sample (w, a, dt) with E = Exp(w dt), Jr = Jr(w dt), G = DR [a]x:
//     e' = A e + B n,   A = [ E^T 0 0 ; -dt^2/2 G  I  dt I ; -dt G 0 I ],   B = [ Jr dt 0 ; 0 dt^2/2 DR ; 0 dt DR ]
//     bias Jacobians    dR/dbg' = E^T dR/dbg - Jr dt,
//                       dp/dbg' = dp/dbg + dt dv/dbg - dt^2/2 G dR/dbg,   dp/dba' = dp/dba + dt dv/dba - dt^2/2 DR,
//                       dv/dbg' = dv/dbg - dt G dR/dbg,                   dv/dba' = dv/dba - dt DR
//     increments        Dp += Dv dt + dt^2/2 DR a,  Dv += dt DR a,  DR = DR E.
// Everything is kept as 3x3 blocks: A and B are block-sparse, so the 9x9 covariance recursion costs 21 block
// products instead of two dense 9x9x9 ones, and the 15x30 Jacobian of the factor has 16 non-zero blocks of 50.
struct ImuPre
{
  M3 DR, dR_dbg, dp_dbg, dp_dba, dv_dbg, dv_dba;
  V3 Dp, Dv, bg, ba, dbg, dba, dbg_buf, dba_buf;
  double dtime = 0;
  M3 S[3][3];  // covariance of (dtheta, dp, dv), 3x3 blocks
  double walk_g = 0, walk_a = 0;  // accumulated bias random walk (isotropic)
  M15 cov, cov_inv;               // assembled once the batch is integrated
  ImuPre(const double* bg1, const double* ba1)
  {
    bg = v3(bg1), ba = v3(ba1);
    DR = M3::eye();
    dR_dbg = dp_dbg = dp_dba = dv_dbg = dv_dba = M3::zero();
    Dp = Dv = dbg = dba = dbg_buf = dba_buf = V3::zero();
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) S[i][j] = M3::zero();
    cov = M15::zero();
  }
  void integrate(const V3& w, const V3& acc, double dt, const vina_config& cfg)
  {
    dtime += dt;
    const double h = 0.5 * dt * dt;
    const M3 E = Exp(w, dt), Et = E.T();
    const M3 Jr = jr(w * dt);
    const M3 G = DR * hat(acc);
    const M3 GR = G * dR_dbg;
    // bias Jacobians (position before velocity: it uses the old velocity Jacobians)
    dp_dba = dp_dba + dv_dba * dt - h * DR;
    dp_dbg = dp_dbg + dv_dbg * dt - h * GR;
    dv_dba = dv_dba - dt * DR;
    dv_dbg = dv_dbg - dt * GR;
    dR_dbg = Et * dR_dbg - Jr * dt;
    // covariance: T = A S, then S' = T A^T + B Q B^T, block by block
    const M3 F2 = -h * G, F3 = -dt * G;
    M3 T[3][3];
    for (int j = 0; j < 3; j++)
    {
      T[0][j] = Et * S[0][j];
      T[1][j] = F2 * S[0][j] + S[1][j] + dt * S[2][j];
      T[2][j] = F3 * S[0][j] + S[2][j];
    }
    const M3 F2t = F2.T(), F3t = F3.T();
    for (int i = 0; i < 3; i++)
    {
      S[i][0] = T[i][0] * E;
      S[i][1] = T[i][0] * F2t + T[i][1] + dt * T[i][2];
      S[i][2] = T[i][0] * F3t + T[i][2];
    }
    // B Q B^T, Q = diag(cov_gyr I, cov_acc I) (node.cpp:262-265): gyro noise enters dtheta through Jr dt, accelerometer
    // noise enters (dp, dv) through (dt^2/2, dt) DR
    const M3 RRt = DR * DR.T();
    S[0][0] = S[0][0] + (cfg.cov_gyr * dt * dt) * (Jr * Jr.T());
    S[1][1] = S[1][1] + (cfg.cov_acc * h * h) * RRt;
    S[1][2] = S[1][2] + (cfg.cov_acc * h * dt) * RRt;
    S[2][1] = S[2][1] + (cfg.cov_acc * h * dt) * RRt;
    S[2][2] = S[2][2] + (cfg.cov_acc * dt * dt) * RRt;
    walk_g += cfg.rdw_gyr * dt;
    walk_a += cfg.rdw_acc * dt;
    // increments
    const V3 Ra = DR * acc;
    Dp = Dp + (Dv * dt + h * Ra);
    Dv = Dv + dt * Ra;
    DR = DR * E;
  }
  // the scan's IMU batch (ends re-stamped to the scan boundaries): mid-point samples, bias-corrected
  // (imu_preintegration.cpp:32-57)
  void push_imu(const std::deque<vina_imu>& buf, double scale_gravity, const vina_config& cfg)
  {
    for (size_t k = 1; k < buf.size(); k++)
    {
      const vina_imu &s0 = buf[k - 1], &s1 = buf[k];
      V3 w, acc;
      for (int q = 0; q < 3; q++)
      {
        w[q] = 0.5 * (s0.gyr[q] + s1.gyr[q]) - bg[q];
        acc[q] = 0.5 * (s0.acc[q] + s1.acc[q]) * scale_gravity - ba[q];
      }
      integrate(w, acc, s1.t - s0.t, cfg);
    }
    cov = M15::zero();
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) cov.set<3, 3>(3 * i, 3 * j, S[i][j]);
    for (int k = 0; k < 3; k++) cov(9 + k, 9 + k) = walk_g, cov(12 + k, 12 + k) = walk_a;
    cov_inv = inverse15(cov);
  }
  // residual r = (Log(DR~^T R1^T R2), R1^T(p2 - p1 - v1 T - g T^2/2) - Dp~, R1^T(v2 - v1 - g T) - Dv~, bg2 - bg1,
  // ba2 - ba1) with the increments corrected to first order for the bias change since integration; returns r^T W r
  // and, on request, J^T W J (30x30) and J^T W r (30) for the two frames' 15-dim states (theta, p, v, bg, ba)
  // (imu_preintegration.cpp:102-163)
  double evaluate(const vina_state& s1, const vina_state& s2, HM<30, 30>* jtj, HM<30, 1>* gg) const
  {
    if (!jtj || !gg) return evaluate_n(s1, s2, nullptr, nullptr, 10);
    return evaluate_n(s1, s2, jtj->d, gg->d, 10);
  }
  // ncols = 10: the two frames' states; ncols = 11: plus the gravity vector of the window (IMU_PRE::give_evaluate_g,
  // imu_preintegration.cpp:165-237: g enters the position and velocity residuals through -R1^T g T^2/2 and -R1^T g T).
  // jtj: (3 ncols)^2 column-major, gg: 3 ncols
  double evaluate_n(const vina_state& s1, const vina_state& s2, double* jtj, double* gg, int ncols) const
  {
    const M3 R1 = m3(s1.R), R2 = m3(s2.R), R1t = R1.T();
    const V3 p1 = v3(s1.p), p2 = v3(s2.p), v1 = v3(s1.v), v2 = v3(s2.v), g = v3(s1.g);
    const V3 th = dR_dbg * dbg;
    const M3 DRc = DR * Exp(th);
    const V3 Dpc = Dp + dp_dbg * dbg + dp_dba * dba;
    const V3 Dvc = Dv + dv_dbg * dbg + dv_dba * dba;
    const M3 Rerr = DRc.T() * R1t * R2;
    const V3 pv = R1t * (v2 - v1 - dtime * g);
    const V3 pt = R1t * (p2 - p1 - v1 * dtime - 0.5 * dtime * dtime * g);
    V15 r = V15::zero();
    r.set<3, 1>(0, 0, Log(Rerr));
    r.set<3, 1>(3, 0, pt - Dpc);
    r.set<3, 1>(6, 0, pv - Dvc);
    r.set<3, 1>(9, 0, v3(s2.bg) - v3(s1.bg));
    r.set<3, 1>(12, 0, v3(s2.ba) - v3(s1.ba));
    const V15 Wr = cov_inv * r;
    double cost = r[0] * Wr[0];
    for (int k = 1; k < 15; k++) cost = cost + r[k] * Wr[k];
    if (!jtj || !gg) return cost;

    // the Jacobian as 5 x 10 (11) blocks of 3 x 3 (block row = residual part, block column = state part of frame 1,
    // 2, then gravity)
    M3 J[5][11];
    bool nz[5][11] = { { false } };
    auto put = [&](int i, int c, const M3& m) {
      J[i][c] = m;
      nz[i][c] = true;
    };
    const M3 Jri = jr_inv(Rerr), I3 = M3::eye();
    put(0, 0, -Jri * R2.T() * R1);
    put(0, 3, -Jri * Rerr.T() * jr(th) * dR_dbg);
    put(0, 5, Jri);
    put(1, 0, hat(pt));
    put(1, 1, -R1t);
    put(1, 2, -R1t * dtime);
    put(1, 3, -dp_dbg);
    put(1, 4, -dp_dba);
    put(1, 6, R1t);
    put(2, 0, hat(pv));
    put(2, 2, -R1t);
    put(2, 3, -dv_dbg);
    put(2, 4, -dv_dba);
    put(2, 7, R1t);
    put(3, 3, -I3);
    put(3, 8, I3);
    put(4, 4, -I3);
    put(4, 9, I3);
    if (ncols == 11)
    {
      put(1, 10, R1t * (-0.5 * dtime * dtime));
      put(2, 10, R1t * (-dtime));
    }
    // WJ = W J: only the non-zero blocks of each block column contribute
    M3 WJ[5][11];
    for (int c = 0; c < ncols; c++)
      for (int i = 0; i < 5; i++)
      {
        M3 acc = M3::zero();
        for (int k = 0; k < 5; k++)
          if (nz[k][c]) acc = acc + cov_inv.blk<3, 3>(3 * i, 3 * k) * J[k][c];
        WJ[i][c] = acc;
      }
    const int ld = 3 * ncols;
    for (int a2 = 0; a2 < ncols; a2++)
    {
      for (int c = 0; c < ncols; c++)
      {
        M3 acc = M3::zero();
        for (int i = 0; i < 5; i++)
          if (nz[i][a2]) acc = acc + J[i][a2].T() * WJ[i][c];
        for (int cc = 0; cc < 3; cc++)
          for (int rr = 0; rr < 3; rr++) jtj[(3 * a2 + rr) + (size_t)ld * (3 * c + cc)] = acc(rr, cc);
      }
      V3 ga = V3::zero();
      for (int i = 0; i < 5; i++)
        if (nz[i][a2]) ga = ga + J[i][a2].T() * Wr.blk<3, 1>(3 * i, 0);
      for (int rr = 0; rr < 3; rr++) gg[3 * a2 + rr] = ga[rr];
    }
    return cost;
  }
  void update_state(const double* dxi15)  // the LM step's bias increments (imu_preintegration.cpp:235-242)
  {
    dbg_buf = dbg;
    dba_buf = dba;
    for (int k = 0; k < 3; k++)
    {
      dbg[k] += dxi15[9 + k];
      dba[k] += dxi15[12 + k];
    }
  }
};

ImuPre* ba_imu_factor_new(const double* bg, const double* ba, const std::deque<vina_imu>& imus, double scale_gravity,
                          const vina_config& cfg)
{
  ImuPre* f = new ImuPre(bg, ba);
  f->push_imu(imus, scale_gravity, cfg);
  return f;
}
void ba_imu_factor_delete(ImuPre* f) { delete f; }

// LI_BA_Optimizer::damping_iter (optimizers.cpp:430-517). xs: the window's states (R, p, v, bg, ba, g); the LiDAR
// factor store of ctx must hold this scan's factors (vina_ba_collect). Returns VINA_OK and the iteration count.
int ba_damping_iter(vina_ctx* ctx, std::vector<vina_state>& xs, std::deque<ImuPre*>& imus_factor, double imu_coef,
                    int* iters_out)
{
  return ba_damping_iter_ex(ctx, xs, imus_factor, imu_coef, iters_out, false, 10, nullptr);
}

// gravity = the start-up variant (LI_BA_OptimizerGravity::damping_iter, optimizers.cpp:746-826): the window's gravity
// vector is a 3-dim unknown behind the states, only the POSE of the first frame is fixed (its v, bg, ba are free),
// max_iter iterations; resis (if given) receives the cost before the first and after the last iteration.
int ba_damping_iter_ex(vina_ctx* ctx, std::vector<vina_state>& xs, std::deque<ImuPre*>& imus_factor, double imu_coef,
                       int* iters_out, bool gravity, int max_iter, double* resis)
{
  // Levenberg-Marquardt with Nielsen's damping update on the window's 15-dim states, first frame fixed. What must
  // equal the reference (LI_BA_Optimizer::damping_iter, optimizers.cpp:430-517) is the SEQUENCE OF DECISIONS - which
  // steps are accepted, when the Hessian is re-evaluated, when the loop stops - because the oracle comparison counts
  // BA runs and LM iterations per scan; the constants below (lambda0 = 0.01, nu = 2, gain -> 1 - (2 rho - 1)^3
  // clamped at 1/3, relative-decrease stop at 1e-6, at most 10 iterations) are that contract. The algebra is
  // organised for this solver: the fixed frame never enters the system (the reference zeroes its rows and columns of
  // a 150 x 150 matrix), only the free (win - 1) * 15 block is assembled and factorised, D = diag(H) is a vector,
  // and the device evaluates the LiDAR factor while the host evaluates the IMU factors.
  const int SD = 15, PD = 6;  // state / pose dimension per frame
  const int win = (int)xs.size();
  const int FX = gravity ? PD : SD;             // fixed leading unknowns (frame 0: its pose, or its whole state)
  const int NG = gravity ? 3 : 0;               // gravity unknowns behind the states
  const int ntot = win * SD + NG;
  const int m = ntot - FX, nl = win * PD;
  std::vector<double> Hf((size_t)m * m), gf(m), dvec(m), step((size_t)ntot, 0.0);
  std::vector<double> hl((size_t)nl * nl), jl(nl);
  std::vector<vina_pose> poses(win);
  auto set_poses = [&](const std::vector<vina_state>& st) {
    for (int i = 0; i < win; i++)
    {
      memcpy(poses[i].R, st[i].R, 72);
      memcpy(poses[i].p, st[i].p, 24);
    }
  };
  // entry (r, c) of the full system lands in the free block when neither index belongs to frame 0
  auto addH = [&](int r, int c, double v) {
    if (r >= FX && c >= FX) Hf[(r - FX) + (size_t)m * (c - FX)] += v;
  };
  auto addg = [&](int r, double v) {
    if (r >= FX) gf[r - FX] += v;
  };
  // index of local unknown k of the IMU factor between frames i, i + 1 (30 states, then gravity) in the system
  auto gidx = [&](int i, int k) { return k < 2 * SD ? i * SD + k : win * SD + (k - 2 * SD); };
  const int NC = gravity ? 11 : 10, NL = 3 * NC;
  std::vector<double> jtj((size_t)NL * NL), gg(NL);
  double lambda = 0.01, nu = 2;
  double cost_cur = 0, cost_try = 0;
  bool relinearise = true;
  std::vector<vina_state> xt = xs;
  double g_try[3] = { xs[0].g[0], xs[0].g[1], xs[0].g[2] };
  int iters = 0;
  // VINA_TRACE: where a BA run spends its time (host IMU factors / device LiDAR factor / solve)
  static double tr_us[5] = { 0, 0, 0, 0, 0 };
  static int tr_calls = 0;
  auto now_us = []() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
  };
  double t0 = 0;
  for (int it = 0; it < max_iter; it++)
  {
    iters++;
    if (relinearise)
    {
      // normal equations at xs: LiDAR factor on the device, IMU factors on the host meanwhile (the reference
      // overlaps the same two with its worker threads, divide_thread, optimizers.cpp:181-245)
      std::fill(Hf.begin(), Hf.end(), 0.0);
      std::fill(gf.begin(), gf.end(), 0.0);
      set_poses(xs);
      t0 = now_us();
      int r = vn_ba_hess_enqueue(ctx, poses.data(), win);
      if (r) return r;
      double cost_imu = 0;
      for (int i = 0; i < win - 1; i++)
      {
        cost_imu += imus_factor[i]->evaluate_n(xs[i], xs[i + 1], jtj.data(), gg.data(), NC);
        for (int c = 0; c < NL; c++)
          for (int r2 = 0; r2 < NL; r2++) addH(gidx(i, r2), gidx(i, c), imu_coef * jtj[r2 + (size_t)NL * c]);
        for (int r2 = 0; r2 < NL; r2++) addg(gidx(i, r2), imu_coef * gg[r2]);
      }
      cost_imu *= (imu_coef * 0.5);
      tr_us[0] += now_us() - t0;
      double cost_lidar = 0;
      t0 = now_us();
      r = vn_ba_hess_finish(ctx, win, hl.data(), jl.data(), &cost_lidar);
      if (r) return r;
      tr_us[1] += now_us() - t0;
      for (int a = 0; a < win; a++)  // the pose blocks of the LiDAR factor (hess_plus, optimizers.cpp:171-179)
      {
        for (int k = 0; k < PD; k++) addg(a * SD + k, jl[a * PD + k]);
        for (int b2 = 0; b2 < win; b2++)
          for (int c = 0; c < PD; c++)
            for (int k = 0; k < PD; k++) addH(a * SD + k, b2 * SD + c, hl[(a * PD + k) + (size_t)nl * (b2 * PD + c)]);
      }
      cost_cur = cost_imu + cost_lidar;
      for (int k = 0; k < m; k++) dvec[k] = Hf[k + (size_t)m * k];
    }
    if (it == 0 && resis) resis[0] = cost_cur;
    t0 = now_us();
    {
      // step = -(H + lambda D)^-1 g on the free block
      std::vector<double> A = Hf, nb(m);
      for (int k = 0; k < m; k++)
      {
        A[k + (size_t)m * k] += lambda * dvec[k];
        nb[k] = -gf[k];
      }
      std::vector<double> sol = ldlt_solve(A, m, nb);
      for (int k = 0; k < m; k++) step[FX + k] = sol[k];
    }
    tr_us[2] += now_us() - t0;
    for (int j = 0; j < win; j++)  // xt = xs (+) step
    {
      const double* d = &step[SD * j];
      V3 d0 = v3(d);
      M3 Rn = m3(xs[j].R) * Exp(d0);
      memcpy(xt[j].R, Rn.d, 72);
      for (int k = 0; k < 3; k++)
      {
        xt[j].p[k] = xs[j].p[k] + d[3 + k];
        xt[j].v[k] = xs[j].v[k] + d[6 + k];
        xt[j].bg[k] = xs[j].bg[k] + d[9 + k];
        xt[j].ba[k] = xs[j].ba[k] + d[12 + k];
      }
    }
    if (gravity)
    {
      // one gravity vector for the whole window. The reference adds the increment to the CANDIDATE's gravity
      // (x_stats_temp[0].g += dxi.tail(3), optimizers.cpp:781), which keeps the increments of rejected steps: the
      // sequence of LM decisions has to be the reference's, so the same bookkeeping here
      for (int k = 0; k < 3; k++) g_try[k] += step[win * SD + k];
      for (int j = 0; j < win; j++)
        for (int k = 0; k < 3; k++) xt[j].g[k] = g_try[k];
    }
    for (int j = 0; j < win - 1; j++) imus_factor[j]->update_state(&step[SD * j]);
    // predicted decrease of the quadratic model: 1/2 step^T (lambda D step - g)
    double predicted = 0;
    for (int k = 0; k < m; k++) predicted += step[FX + k] * ((lambda * dvec[k]) * step[FX + k] - gf[k]);
    predicted *= 0.5;
    {
      // cost at the candidate (only_residual, optimizers.cpp:340-376)
      double c_imu = 0;
      t0 = now_us();
      for (int i = 0; i < win - 1; i++) c_imu += imus_factor[i]->evaluate(xt[i], xt[i + 1], nullptr, nullptr);
      c_imu *= (imu_coef * 0.5);
      tr_us[3] += now_us() - t0;
      set_poses(xt);
      double c_lidar = 0;
      t0 = now_us();
      int r = vina_ba_lidar_residual(ctx, poses.data(), win, &c_lidar, nullptr, 0);
      if (r) return r;
      tr_us[4] += now_us() - t0;
      cost_try = c_imu + c_lidar;
    }
    const double decrease = cost_cur - cost_try;
    if (decrease > 0)
    {
      xs = xt;
      const double rho = decrease / predicted;
      const double shrink = 1 - std::pow(2 * rho - 1, 3);
      lambda *= (shrink < 1.0 / 3 ? 1.0 / 3 : shrink);
      nu = 2;
      relinearise = true;
    }
    else
    {
      lambda *= nu;
      nu *= 2;
      relinearise = false;
      for (int j = 0; j < win - 1; j++)  // the factors go back to the biases of xs
      {
        imus_factor[j]->dbg = imus_factor[j]->dbg_buf;
        imus_factor[j]->dba = imus_factor[j]->dba_buf;
      }
    }
    if (std::fabs(decrease / cost_cur) < 1e-6) break;
  }
  if (resis) resis[1] = cost_try;
  if (iters_out) *iters_out = iters;
  if (ctx->trace && (++tr_calls % 10) == 0)
    fprintf(stderr, "[vina trace] BA, us per run over %d runs: imu jac (device Hessian in flight) %.1f, wait for the device Hessian %.1f, solve %.1f, imu res %.1f, "
                    "lidar res (device) %.1f\n", tr_calls, tr_us[0] / tr_calls, tr_us[1] / tr_calls, tr_us[2] / tr_calls,
            tr_us[3] / tr_calls, tr_us[4] / tr_calls);
  return VINA_OK;
}

// ---- stateless host entry points (no CUDA context needed): the CPU test-suite checks the host side of the BA
// against the oracle with them
extern "C" int vina_ba_imu_evaluate(const vina_config* cfg, const double bg[3], const double ba[3], const vina_imu* imus, int m,
                                    double scale_gravity, const vina_state* s1, const vina_state* s2, double* residual,
                                    double* jtj, double* gg)
{
  if (!cfg || !bg || !ba || !imus || m < 2 || !s1 || !s2 || !residual) return VINA_E_ARG;
  std::deque<vina_imu> buf(imus, imus + m);
  ImuPre f(bg, ba);
  f.push_imu(buf, scale_gravity, *cfg);
  HM<30, 30> J;
  HM<30, 1> g;
  if (jtj && gg)
  {
    *residual = f.evaluate(*s1, *s2, &J, &g);
    memcpy(jtj, J.d, sizeof(J.d));
    memcpy(gg, g.d, sizeof(g.d));
  }
  else
    *residual = f.evaluate(*s1, *s2, nullptr, nullptr);
  return VINA_OK;
}

extern "C" int vina_ba_solve(const double* A, int n, const double* b, double* x)
{
  if (!A || !b || !x || n < 1) return VINA_E_ARG;
  std::vector<double> sol = ldlt_solve(std::vector<double>(A, A + (size_t)n * n), n, std::vector<double>(b, b + n));
  memcpy(x, sol.data(), (size_t)n * sizeof(double));
  return VINA_OK;
}

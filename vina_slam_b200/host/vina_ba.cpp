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

// ---- IMU_PRE ------------------------------------------------------------------------------------------------
struct ImuPre
{
  M3 R_delta, R_bg, p_bg, p_ba, v_bg, v_ba;
  V3 p_delta, v_delta, bg, ba, dbg, dba, dbg_buf, dba_buf;
  double dtime = 0;
  M15 cov, cov_inv;  // cov_inv: the covariance does not change once the batch is integrated
  ImuPre(const double* bg1, const double* ba1)
  {
    bg = v3(bg1), ba = v3(ba1);
    R_delta = M3::eye();
    R_bg = p_bg = p_ba = v_bg = v_ba = M3::zero();
    p_delta = v_delta = dbg = dba = dbg_buf = dba_buf = V3::zero();
    cov = M15::zero();
  }
  // imu_preintegration.cpp:59-100
  void add_imu(const V3& cur_gyr, const V3& cur_acc, double dt, const vina_config& cfg)
  {
    dtime += dt;
    M3 rotation_increment = Exp(cur_gyr, dt);
    M3 right_jacobian = jr(cur_gyr * dt);
    M3 rotation_dt = dt * R_delta;
    M3 rotation_dt2_half = 0.5 * dt * dt * R_delta;
    M3 acc_skew = hat(cur_acc);
    p_ba = p_ba + v_ba * dt - rotation_dt2_half;
    p_bg = p_bg + v_bg * dt - rotation_dt2_half * acc_skew * R_bg;
    v_ba = v_ba - rotation_dt;
    v_bg = v_bg - rotation_dt * acc_skew * R_bg;
    R_bg = rotation_increment.T() * R_bg - right_jacobian * dt;
    HM<9, 9> ja = HM<9, 9>::eye();
    HM<9, 6> jb = HM<9, 6>::zero();
    ja.set<3, 3>(0, 0, rotation_increment.T());
    ja.set<3, 3>(3, 0, -rotation_dt2_half * acc_skew);
    ja.set<3, 3>(3, 6, M3::eye() * dt);
    ja.set<3, 3>(6, 0, -rotation_dt * acc_skew);
    jb.set<3, 3>(0, 0, right_jacobian * dt);
    jb.set<3, 3>(3, 3, rotation_dt2_half);
    jb.set<3, 3>(6, 3, rotation_dt);
    HM<6, 6> nm = HM<6, 6>::zero(), nw = HM<6, 6>::zero();  // node.cpp:262-265
    for (int k = 0; k < 3; k++)
    {
      nm(k, k) = cfg.cov_gyr, nm(3 + k, 3 + k) = cfg.cov_acc;
      nw(k, k) = cfg.rdw_gyr, nw(3 + k, 3 + k) = cfg.rdw_acc;
    }
    HM<9, 9> c99 = cov.blk<9, 9>(0, 0);
    cov.set<9, 9>(0, 0, ja * c99 * ja.T() + jb * nm * jb.T());
    HM<6, 6> c66 = cov.blk<6, 6>(9, 9);
    cov.set<6, 6>(9, 9, c66 + nw * dt);
    p_delta = p_delta + (v_delta * dt + rotation_dt2_half * cur_acc);
    v_delta = v_delta + rotation_dt * cur_acc;
    R_delta = R_delta * rotation_increment;
  }
  // imu_preintegration.cpp:32-57 (the deque is the scan's IMU batch with its ends re-stamped to the scan boundaries)
  void push_imu(const std::deque<vina_imu>& buf, double scale_gravity, const vina_config& cfg)
  {
    for (size_t k = 1; k < buf.size(); k++)
    {
      const vina_imu &a = buf[k - 1], &b = buf[k];
      const double dt = b.t - a.t;
      V3 g, c;
      for (int q = 0; q < 3; q++)
      {
        g[q] = 0.5 * (a.gyr[q] + b.gyr[q]);
        c[q] = 0.5 * (a.acc[q] + b.acc[q]);
      }
      g = g - bg;
      c = c * scale_gravity - ba;
      add_imu(g, c, dt, cfg);
    }
    cov_inv = inverse15(cov);
  }
  // imu_preintegration.cpp:102-163; jtj 30x30 column-major, gg 30
  double evaluate(const vina_state& s1, const vina_state& s2, HM<30, 30>* jtj, HM<30, 1>* gg) const
  {
    const M3 R1 = m3(s1.R), R2 = m3(s2.R), I33 = M3::eye();
    const V3 p1 = v3(s1.p), p2 = v3(s2.p), v1 = v3(s1.v), v2 = v3(s2.v), g1 = v3(s1.g);
    M3 R_correct = R_delta * Exp(R_bg * dbg);
    V3 t_correct = p_delta + p_bg * dbg + p_ba * dba;
    V3 v_correct = v_delta + v_bg * dbg + v_ba * dba;
    M3 res_r = R_correct.T() * R1.T() * R2;
    V3 exp_v = R1.T() * (v2 - v1 - dtime * g1);
    V3 res_v = exp_v - v_correct;
    V3 exp_t = R1.T() * (p2 - p1 - v1 * dtime - 0.5 * dtime * dtime * g1);
    V3 res_t = exp_t - t_correct;
    V15 rr = V15::zero();
    rr.set<3, 1>(0, 0, Log(res_r));
    rr.set<3, 1>(3, 0, res_t);
    rr.set<3, 1>(6, 0, res_v);
    rr.set<3, 1>(9, 0, v3(s2.bg) - v3(s1.bg));
    rr.set<3, 1>(12, 0, v3(s2.ba) - v3(s1.ba));
    if (jtj && gg)
    {
      M15 joca = M15::zero(), jocb = M15::zero();
      M3 JR_inv = jr_inv(res_r);
      joca.set<3, 3>(0, 0, -JR_inv * R2.T() * R1);
      jocb.set<3, 3>(0, 0, JR_inv);
      joca.set<3, 3>(0, 9, -JR_inv * res_r.T() * jr(R_bg * dbg) * R_bg);
      joca.set<3, 3>(3, 0, hat(exp_t));
      joca.set<3, 3>(3, 3, -R1.T());
      joca.set<3, 3>(3, 6, -R1.T() * dtime);
      joca.set<3, 3>(3, 9, -p_bg);
      joca.set<3, 3>(3, 12, -p_ba);
      jocb.set<3, 3>(3, 3, R1.T());
      joca.set<3, 3>(6, 0, hat(exp_v));
      joca.set<3, 3>(6, 6, -R1.T());
      joca.set<3, 3>(6, 9, -v_bg);
      joca.set<3, 3>(6, 12, -v_ba);
      jocb.set<3, 3>(6, 6, R1.T());
      joca.set<3, 3>(9, 9, -I33);
      joca.set<3, 3>(12, 12, -I33);
      jocb.set<3, 3>(9, 9, I33);
      jocb.set<3, 3>(12, 12, I33);
      HM<15, 30> joc;
      joc.set<15, 15>(0, 0, joca);
      joc.set<15, 15>(0, 15, jocb);
      *jtj = joc.T() * cov_inv * joc;
      *gg = joc.T() * cov_inv * rr;
    }
    V15 w = cov_inv * rr;
    double s = rr[0] * w[0];
    for (int k = 1; k < 15; k++) s = s + rr[k] * w[k];
    return s;
  }
  void update_state(const double* dxi15)  // imu_preintegration.cpp:235-242
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
  const int DIM = 15, DVEL = 6;
  const int win = (int)xs.size();
  const int n = win * DIM, nl = win * DVEL;
  double u = 0.01, v = 2;
  std::vector<double> D((size_t)n * n, 0.0), Hess((size_t)n * n, 0.0), JacT(n, 0.0), dxi(n, 0.0);
  for (int i = 0; i < n; i++) D[i + (size_t)n * i] = 1.0;
  auto H = [&](int r, int c) -> double& { return Hess[r + (size_t)n * c]; };
  std::vector<double> hl((size_t)nl * nl), jl(nl);
  std::vector<vina_pose> poses(win);
  auto set_poses = [&](const std::vector<vina_state>& s) {
    for (int i = 0; i < win; i++)
    {
      memcpy(poses[i].R, s[i].R, 72);
      memcpy(poses[i].p, s[i].p, 24);
    }
  };
  double residual1 = 0, residual2 = 0, q;
  bool is_calc_hess = true;
  std::vector<vina_state> xt = xs;
  int iters = 0;
  // VINA_TRACE: where a BA run spends its time (host IMU factors / device LiDAR factor / solve)
  static double tr_us[5] = { 0, 0, 0, 0, 0 };
  static int tr_calls = 0;
  auto now_us = []() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
  };
  double t0 = 0;
  for (int it = 0; it < 10; it++)
  {
    iters++;
    if (is_calc_hess)
    {
      // divide_thread (optimizers.cpp:181-245): the LiDAR factor on the device, the IMU factors on the host
      // meanwhile (the reference overlaps the same two with its worker threads)
      std::fill(Hess.begin(), Hess.end(), 0.0);
      std::fill(JacT.begin(), JacT.end(), 0.0);
      set_poses(xs);
      t0 = now_us();
      int r = vn_ba_hess_enqueue(ctx, poses.data(), win);
      if (r) return r;
      double residual = 0;
      HM<30, 30> jtj;
      HM<30, 1> gg;
      for (int i = 0; i < win - 1; i++)
      {
        residual += imus_factor[i]->evaluate(xs[i], xs[i + 1], &jtj, &gg);
        for (int c = 0; c < 2 * DIM; c++)
          for (int r2 = 0; r2 < 2 * DIM; r2++) H(i * DIM + r2, i * DIM + c) += jtj(r2, c);
        for (int r2 = 0; r2 < 2 * DIM; r2++) JacT[i * DIM + r2] += gg[r2];
      }
      for (double& h : Hess) h *= imu_coef;
      for (double& j : JacT) j *= imu_coef;
      residual *= (imu_coef * 0.5);
      tr_us[0] += now_us() - t0;
      double rl = 0;
      t0 = now_us();
      r = vn_ba_hess_finish(ctx, win, hl.data(), jl.data(), &rl);
      if (r) return r;
      tr_us[1] += now_us() - t0;
      for (int a = 0; a < win; a++)  // hess_plus (optimizers.cpp:171-179)
      {
        for (int k = 0; k < DVEL; k++) JacT[a * DIM + k] += jl[a * DVEL + k];
        for (int b = 0; b < win; b++)
          for (int c = 0; c < DVEL; c++)
            for (int k = 0; k < DVEL; k++) H(a * DIM + k, b * DIM + c) += hl[(a * DVEL + k) + (size_t)nl * (b * DVEL + c)];
      }
      residual1 = residual + rl;
    }
    for (int c = 0; c < n; c++)
      for (int r = 0; r < DIM; r++) H(r, c) = 0.0;
    for (int c = 0; c < DIM; c++)
      for (int r = 0; r < n; r++) H(r, c) = 0.0;
    for (int c = 0; c < DIM; c++) H(c, c) = 1.0;
    for (int r = 0; r < DIM; r++) JacT[r] = 0.0;
    for (int k = 0; k < n; k++) D[k + (size_t)n * k] = H(k, k);
    t0 = now_us();
    {
      // dxi = (Hess + u D).ldlt().solve(-JacT). The first frame is fixed: its rows / columns are the identity
      // and its gradient is zero, so dxi(0:15) = 0 and the remaining (n - 15)^2 block is solved on its own.
      const int m = n - DIM;
      std::vector<double> A((size_t)m * m), nb(m);
      for (int c = 0; c < m; c++)
        for (int r = 0; r < m; r++)
          A[r + (size_t)m * c] = Hess[(DIM + r) + (size_t)n * (DIM + c)] + u * D[(DIM + r) + (size_t)n * (DIM + c)];
      for (int k = 0; k < m; k++) nb[k] = -JacT[DIM + k];
      std::vector<double> sol = ldlt_solve(A, m, nb);
      for (int k = 0; k < DIM; k++) dxi[k] = 0.0;
      for (int k = 0; k < m; k++) dxi[DIM + k] = sol[k];
    }
    tr_us[2] += now_us() - t0;
    for (int j = 0; j < win; j++)
    {
      V3 d0;
      for (int k = 0; k < 3; k++) d0[k] = dxi[DIM * j + k];
      M3 Rn = m3(xs[j].R) * Exp(d0);
      memcpy(xt[j].R, Rn.d, 72);
      for (int k = 0; k < 3; k++)
      {
        xt[j].p[k] = xs[j].p[k] + dxi[DIM * j + 3 + k];
        xt[j].v[k] = xs[j].v[k] + dxi[DIM * j + 6 + k];
        xt[j].bg[k] = xs[j].bg[k] + dxi[DIM * j + 9 + k];
        xt[j].ba[k] = xs[j].ba[k] + dxi[DIM * j + 12 + k];
      }
    }
    for (int j = 0; j < win - 1; j++) imus_factor[j]->update_state(&dxi[DIM * j]);
    double q1 = 0;
    {
      // D is diagonal: u D dxi - JacT needs no 150x150 product
      double s = 0;
      for (int k = 0; k < n; k++) s += dxi[k] * ((u * D[k + (size_t)n * k]) * dxi[k] - JacT[k]);
      q1 = 0.5 * s;
    }
    // only_residual (optimizers.cpp:340-376)
    {
      double r1 = 0;
      t0 = now_us();
      for (int i = 0; i < win - 1; i++) r1 += imus_factor[i]->evaluate(xt[i], xt[i + 1], nullptr, nullptr);
      r1 *= (imu_coef * 0.5);
      tr_us[3] += now_us() - t0;
      set_poses(xt);
      double rl = 0;
      t0 = now_us();
      int r = vina_ba_lidar_residual(ctx, poses.data(), win, &rl, nullptr, 0);
      if (r) return r;
      tr_us[4] += now_us() - t0;
      residual2 = r1 + rl;
    }
    q = residual1 - residual2;
    if (q > 0)
    {
      xs = xt;
      const double one_three = 1.0 / 3;
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
      for (int j = 0; j < win - 1; j++)
      {
        imus_factor[j]->dbg = imus_factor[j]->dbg_buf;
        imus_factor[j]->dba = imus_factor[j]->dba_buf;
      }
    }
    if (std::fabs((residual1 - residual2) / residual1) < 1e-6) break;
  }
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

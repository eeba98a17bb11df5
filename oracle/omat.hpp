// ORACLE — TEST INFRASTRUCTURE ONLY (pinned against the reference build oracle/_ref: see oracle/README.md).
// Minimal fixed-size dense linear algebra used by the CPU restatement of the
// VINA-SLAM hot path. It stands in for the Eigen3 expressions the reference
// uses (Eigen is absent from this image). Storage is
// column-major like Eigen's default; every product coefficient is evaluated
// left-to-right, sum_k a(i,k)*b(k,j), with no FMA contraction (the oracle is
// built with -ffp-contract=off), so decision-bearing expressions have one
// fixed rounding sequence.
#pragma once
#include <cmath>
#include <cstring>

namespace vo
{
template <int R, int C>
struct Mat
{
  double d[R * C];

  double& operator()(int i, int j) { return d[i + j * R]; }
  const double& operator()(int i, int j) const { return d[i + j * R]; }
  double& operator[](int i) { return d[i]; }
  const double& operator[](int i) const { return d[i]; }

  static Mat Zero()
  {
    Mat m;
    for (int i = 0; i < R * C; i++) m.d[i] = 0.0;
    return m;
  }
  static Mat Identity()
  {
    Mat m = Zero();
    for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = 1.0;
    return m;
  }
  void setZero() { *this = Zero(); }
  void setIdentity() { *this = Identity(); }

  Mat<C, R> transpose() const
  {
    Mat<C, R> t;
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C; j++) t(j, i) = (*this)(i, j);
    return t;
  }

  template <int BR, int BC>
  Mat<BR, BC> block(int r0, int c0) const
  {
    Mat<BR, BC> b;
    for (int i = 0; i < BR; i++)
      for (int j = 0; j < BC; j++) b(i, j) = (*this)(r0 + i, c0 + j);
    return b;
  }
  template <int BR, int BC>
  void setBlock(int r0, int c0, const Mat<BR, BC>& b)
  {
    for (int i = 0; i < BR; i++)
      for (int j = 0; j < BC; j++) (*this)(r0 + i, c0 + j) = b(i, j);
  }
  Mat<R, 1> col(int j) const { return block<R, 1>(0, j); }

  Mat& operator+=(const Mat& o)
  {
    for (int i = 0; i < R * C; i++) d[i] = d[i] + o.d[i];
    return *this;
  }
  Mat& operator-=(const Mat& o)
  {
    for (int i = 0; i < R * C; i++) d[i] = d[i] - o.d[i];
    return *this;
  }
};

template <int R, int C>
inline Mat<R, C> operator+(const Mat<R, C>& a, const Mat<R, C>& b)
{
  Mat<R, C> r;
  for (int i = 0; i < R * C; i++) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int R, int C>
inline Mat<R, C> operator-(const Mat<R, C>& a, const Mat<R, C>& b)
{
  Mat<R, C> r;
  for (int i = 0; i < R * C; i++) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int R, int C>
inline Mat<R, C> operator-(const Mat<R, C>& a)
{
  Mat<R, C> r;
  for (int i = 0; i < R * C; i++) r.d[i] = -a.d[i];
  return r;
}
template <int R, int C>
inline Mat<R, C> operator*(double s, const Mat<R, C>& a)
{
  Mat<R, C> r;
  for (int i = 0; i < R * C; i++) r.d[i] = s * a.d[i];
  return r;
}
template <int R, int C>
inline Mat<R, C> operator*(const Mat<R, C>& a, double s)
{
  Mat<R, C> r;
  for (int i = 0; i < R * C; i++) r.d[i] = a.d[i] * s;
  return r;
}
template <int R, int C>
inline Mat<R, C> operator/(const Mat<R, C>& a, double s)
{
  Mat<R, C> r;
  for (int i = 0; i < R * C; i++) r.d[i] = a.d[i] / s;
  return r;
}
template <int R, int K, int C>
inline Mat<R, C> operator*(const Mat<R, K>& a, const Mat<K, C>& b)
{
  Mat<R, C> r;
  for (int j = 0; j < C; j++)
    for (int i = 0; i < R; i++)
    {
      double s = a(i, 0) * b(0, j);
      for (int k = 1; k < K; k++) s = s + a(i, k) * b(k, j);
      r(i, j) = s;
    }
  return r;
}

typedef Mat<3, 1> Vec3;
typedef Mat<3, 3> Mat3;
typedef Mat<6, 1> Vec6;
typedef Mat<6, 6> Mat6;
typedef Mat<9, 9> Mat9;
typedef Mat<15, 1> Vec15;
typedef Mat<15, 15> Mat15;

inline Vec3 V3(double x, double y, double z)
{
  Vec3 v;
  v[0] = x;
  v[1] = y;
  v[2] = z;
  return v;
}
template <int N>
inline double dot(const Mat<N, 1>& a, const Mat<N, 1>& b)
{
  double s = a[0] * b[0];
  for (int i = 1; i < N; i++) s = s + a[i] * b[i];
  return s;
}
template <int N>
inline double squaredNorm(const Mat<N, 1>& a)
{
  return dot(a, a);
}
template <int N>
inline double norm(const Mat<N, 1>& a)
{
  return std::sqrt(dot(a, a));
}
inline Vec3 cross(const Vec3& a, const Vec3& b)
{
  return V3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
// Eigen's normalize(): divide by the norm when it is > 0.
inline void normalize(Vec3& a)
{
  double n2 = squaredNorm(a);
  if (n2 > 0) a = a / std::sqrt(n2);
}
inline Vec3 normalized(const Vec3& a)
{
  Vec3 r = a;
  normalize(r);
  return r;
}
inline double trace(const Mat3& m) { return (m(0, 0) + m(1, 1)) + m(2, 2); }

// vina_slam/core/math.hpp:50-55 (hat) — row-major comma initialiser
inline Mat3 hat(const Vec3& v)
{
  Mat3 O;
  O(0, 0) = 0;
  O(0, 1) = -v[2];
  O(0, 2) = v[1];
  O(1, 0) = v[2];
  O(1, 1) = 0;
  O(1, 2) = -v[0];
  O(2, 0) = -v[1];
  O(2, 1) = v[0];
  O(2, 2) = 0;
  return O;
}

// vina_slam/core/math.hpp:12-24 : Exp(ang), threshold ||ang|| >= 1e-9
inline Mat3 Exp(const Vec3& ang)
{
  double ang_norm = norm(ang);
  if (ang_norm >= 1e-9)
  {
    Vec3 r_axis = ang / ang_norm;
    Mat3 K = hat(r_axis);
    return Mat3::Identity() + std::sin(ang_norm) * K + (1.0 - std::cos(ang_norm)) * K * K;
  }
  return Mat3::Identity();
}

// vina_slam/core/math.hpp:26-41 : Exp(ang_vel, dt), threshold ||w|| > 1e-7
inline Mat3 Exp(const Vec3& ang_vel, double dt)
{
  double ang_vel_norm = norm(ang_vel);
  if (ang_vel_norm > 1e-7)
  {
    Vec3 r_axis = ang_vel / ang_vel_norm;
    Mat3 K = hat(r_axis);
    double r_ang = ang_vel_norm * dt;
    return Mat3::Identity() + std::sin(r_ang) * K + (1.0 - std::cos(r_ang)) * K * K;
  }
  return Mat3::Identity();
}

// vina_slam/core/math.hpp:43-48
inline Vec3 Log(const Mat3& R)
{
  double tr = trace(R);
  double theta = (tr > 3.0 - 1e-6) ? 0.0 : std::acos(0.5 * (tr - 1));
  Vec3 K = V3(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1));
  return (std::fabs(theta) < 0.001) ? (0.5 * K) : (0.5 * theta / std::sin(theta) * K);
}

// ---------------------------------------------------------------------------
// Restatement of Eigen 3.4.0 SelfAdjointEigenSolver<Matrix3d>::compute()
// (Eigen/src/Eigenvalues/SelfAdjointEigenSolver.h: scaling by max|a_ij| of the
// lower triangle, tridiagonalization_inplace_selector<.,3,false>,
// computeFromTridiagonal_impl with tridiagonal_qr_step / Wilkinson shift,
// ascending sort). Eigen is a third-party dependency of the reference whose
// version is NOT pinned (CMakeLists.txt:35 find_package(Eigen3 REQUIRED)); the
// published 3.4.0 algorithm is restated. Call sites in the reference:
// octree.cpp:362, 435, 651; odometry.cpp:244. Only the lower triangle is read.
// ---------------------------------------------------------------------------
struct SelfAdjointEigen3
{
  Vec3 values;   // ascending
  Mat3 vectors;  // columns

  static void makeGivens(double p, double q, double& c, double& s)
  {
    if (q == 0.0)
    {
      c = p < 0 ? -1.0 : 1.0;
      s = 0.0;
    }
    else if (p == 0.0)
    {
      c = 0.0;
      s = q < 0 ? 1.0 : -1.0;
    }
    else if (std::fabs(p) > std::fabs(q))
    {
      double t = q / p;
      double u = std::sqrt(1.0 + t * t);
      if (p < 0) u = -u;
      c = 1.0 / u;
      s = -t * c;
    }
    else
    {
      double t = p / q;
      double u = std::sqrt(1.0 + t * t);
      if (q < 0) u = -u;
      s = -1.0 / u;
      c = -t * s;
    }
  }
  static double hypot_pos(double x, double y)
  {
    double ax = std::fabs(x), ay = std::fabs(y);
    double p = ax > ay ? ax : ay;
    if (p == 0.0) return 0.0;
    double qp = (ax > ay ? ay : ax) / p;
    return p * std::sqrt(1.0 + qp * qp);
  }
  static void qr_step(double* diag, double* subdiag, int start, int end, Mat3& Q)
  {
    double td = (diag[end - 1] - diag[end]) * 0.5;
    double e = subdiag[end - 1];
    double mu = diag[end];
    if (td == 0.0)
    {
      mu -= std::fabs(e);
    }
    else if (e != 0.0)
    {
      const double e2 = e * e;
      const double h = hypot_pos(td, e);
      if (e2 == 0.0)
        mu -= e / ((td + (td > 0.0 ? h : -h)) / e);
      else
        mu -= e2 / (td + (td > 0.0 ? h : -h));
    }
    double x = diag[start] - mu;
    double z = subdiag[start];
    for (int k = start; k < end && z != 0.0; ++k)
    {
      double c, s;
      makeGivens(x, z, c, s);
      double sdk = s * diag[k] + c * subdiag[k];
      double dkp1 = s * subdiag[k] + c * diag[k + 1];
      diag[k] = c * (c * diag[k] - s * subdiag[k]) - s * (c * subdiag[k] - s * diag[k + 1]);
      diag[k + 1] = s * sdk + c * dkp1;
      subdiag[k] = c * sdk - s * dkp1;
      if (k > start) subdiag[k - 1] = c * subdiag[k - 1] - s * z;
      x = subdiag[k];
      if (k < end - 1)
      {
        z = -s * subdiag[k + 1];
        subdiag[k + 1] = c * subdiag[k + 1];
      }
      // Q = Q * G : applyOnTheRight(k, k+1, rot)
      for (int i = 0; i < 3; i++)
      {
        double xi = Q(i, k), yi = Q(i, k + 1);
        Q(i, k) = c * xi - s * yi;
        Q(i, k + 1) = s * xi + c * yi;
      }
    }
  }

  explicit SelfAdjointEigen3(const Mat3& A)
  {
    Mat3 mat = Mat3::Zero();
    for (int j = 0; j < 3; j++)
      for (int i = j; i < 3; i++) mat(i, j) = A(i, j);
    double scale = 0.0;
    for (int j = 0; j < 3; j++)
      for (int i = j; i < 3; i++)
        if (std::fabs(mat(i, j)) > scale) scale = std::fabs(mat(i, j));
    if (scale == 0.0) scale = 1.0;
    for (int j = 0; j < 3; j++)
      for (int i = j; i < 3; i++) mat(i, j) = mat(i, j) / scale;

    double diag[3], subdiag[2];
    const double tol = 2.2250738585072014e-308;  // numeric_limits<double>::min()
    diag[0] = mat(0, 0);
    double v1norm2 = mat(2, 0) * mat(2, 0);
    if (v1norm2 <= tol)
    {
      diag[1] = mat(1, 1);
      diag[2] = mat(2, 2);
      subdiag[0] = mat(1, 0);
      subdiag[1] = mat(2, 1);
      mat.setIdentity();
    }
    else
    {
      double beta = std::sqrt(mat(1, 0) * mat(1, 0) + v1norm2);
      double invBeta = 1.0 / beta;
      double m01 = mat(1, 0) * invBeta;
      double m02 = mat(2, 0) * invBeta;
      double q = 2.0 * m01 * mat(2, 1) + m02 * (mat(2, 2) - mat(1, 1));
      diag[1] = mat(1, 1) + m02 * q;
      diag[2] = mat(2, 2) - m02 * q;
      subdiag[0] = beta;
      subdiag[1] = mat(2, 1) - m01 * q;
      mat.setZero();
      mat(0, 0) = 1;
      mat(1, 1) = m01;
      mat(1, 2) = m02;
      mat(2, 1) = m02;
      mat(2, 2) = -m01;
    }

    const int n = 3, maxIterations = 30;
    int end = n - 1, start = 0, iter = 0;
    const double considerAsZero = 2.2250738585072014e-308;
    const double precision_inv = 1.0 / 2.220446049250313e-16;
    while (end > 0)
    {
      for (int i = start; i < end; ++i)
      {
        if (std::fabs(subdiag[i]) < considerAsZero)
          subdiag[i] = 0.0;
        else
        {
          const double scaled_subdiag = precision_inv * subdiag[i];
          if (scaled_subdiag * scaled_subdiag <= (std::fabs(diag[i]) + std::fabs(diag[i + 1]))) subdiag[i] = 0.0;
        }
      }
      while (end > 0 && subdiag[end - 1] == 0.0) end--;
      if (end <= 0) break;
      iter++;
      if (iter > maxIterations * n) break;
      start = end - 1;
      while (start > 0 && subdiag[start - 1] != 0.0) start--;
      qr_step(diag, subdiag, start, end, mat);
    }
    if (iter <= maxIterations * n)
    {
      for (int i = 0; i < n - 1; ++i)
      {
        int k = 0;
        double mn = diag[i];
        for (int j = 1; j < n - i; j++)
          if (diag[i + j] < mn)
          {
            mn = diag[i + j];
            k = j;
          }
        if (k > 0)
        {
          double t = diag[i];
          diag[i] = diag[k + i];
          diag[k + i] = t;
          for (int r = 0; r < 3; r++)
          {
            double tt = mat(r, i);
            mat(r, i) = mat(r, k + i);
            mat(r, k + i) = tt;
          }
        }
      }
    }
    for (int i = 0; i < 3; i++) values[i] = diag[i] * scale;
    vectors = mat;
  }
};

// Matrix<double,15,15>::inverse() in Eigen is PartialPivLU-based for sizes > 4
// (Eigen/src/LU/InverseImpl.h). Restated as LU with partial pivoting followed
// by forward/back substitution of the identity. Call sites:
// odometry.cpp:82, 194.
template <int N>
inline Mat<N, N> inverse(const Mat<N, N>& A)
{
  Mat<N, N> lu = A;
  int perm[N];
  for (int i = 0; i < N; i++) perm[i] = i;
  for (int k = 0; k < N; k++)
  {
    int piv = k;
    double best = std::fabs(lu(k, k));
    for (int i = k + 1; i < N; i++)
      if (std::fabs(lu(i, k)) > best)
      {
        best = std::fabs(lu(i, k));
        piv = i;
      }
    if (piv != k)
    {
      for (int j = 0; j < N; j++)
      {
        double t = lu(k, j);
        lu(k, j) = lu(piv, j);
        lu(piv, j) = t;
      }
      int t = perm[k];
      perm[k] = perm[piv];
      perm[piv] = t;
    }
    for (int i = k + 1; i < N; i++)
    {
      lu(i, k) = lu(i, k) / lu(k, k);
      for (int j = k + 1; j < N; j++) lu(i, j) = lu(i, j) - lu(i, k) * lu(k, j);
    }
  }
  Mat<N, N> inv;
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
}  // namespace vo

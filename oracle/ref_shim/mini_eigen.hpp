// ORACLE — TEST INFRASTRUCTURE ONLY.
// A minimal, eager stand-in for the subset of the Eigen 3 API that the reference's hot-path sources use
// (types.hpp, math.hpp, point_utils.cpp, octree.cpp, voxel_map.cpp, imu_ekf.cpp). Eigen itself is not
// installed in this image. With this header the reference's OWN source files compile unmodified into
// oracle/_ref/libvina_ref.so, which pins the hand-written restatement (vina_oracle.cpp) against the
// reference's real control flow. Every operator evaluates eagerly into a dense column-major Matrix, product
// coefficients are summed left to right (Eigen's order for small fixed sizes), and
// SelfAdjointEigenSolver / inverse() restate Eigen 3.4.0 (see oracle/omat.hpp).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen
{
constexpr int Dynamic = -1;
template <typename T>
using aligned_allocator = std::allocator<T>;

template <typename S, int R, int C>
class Matrix;

// CRTP base: anything with rows(), cols(), coeff(i,j)
template <typename D, int R, int C>
struct Base
{
  const D& d() const { return static_cast<const D&>(*this); }
  D& d() { return static_cast<D&>(*this); }
  double coeff(int i, int j) const { return d().coeff(i, j); }
  Matrix<double, R, C> eval() const;
  Matrix<double, C, R> transpose() const;
  double norm() const { return std::sqrt(squaredNorm()); }
  double squaredNorm() const
  {
    static_assert(R == 1 || C == 1, "vector");
    const int n = R * C;
    double s = lin(0) * lin(0);
    for (int i = 1; i < n; i++) s = s + lin(i) * lin(i);
    return s;
  }
  double lin(int i) const { return C == 1 ? coeff(i, 0) : coeff(0, i); }
  template <typename O, int R2, int C2>
  double dot(const Base<O, R2, C2>& o) const
  {
    static_assert(R2 * C2 == R * C, "same length");
    const int n = R * C;
    double s = lin(0) * o.lin(0);
    for (int i = 1; i < n; i++) s = s + lin(i) * o.lin(i);
    return s;
  }
  template <typename O>
  Matrix<double, 3, 1> cross(const Base<O, 3, 1>& o) const;
  Matrix<double, R, C> normalized() const;
  double trace() const
  {
    double s = coeff(0, 0);
    for (int i = 1; i < R; i++) s = s + coeff(i, i);
    return s;
  }
  double sum() const
  {
    double s = 0;
    for (int j = 0; j < C; j++)
      for (int i = 0; i < R; i++) s = s + coeff(i, j);
    return s;
  }
  Matrix<double, R, C> inverse() const;
  Matrix<double, (R < C ? R : C), 1> diagonal() const;
  Matrix<double, R, R> asDiagonal() const;  // for vectors
  template <int BR, int BC>
  Matrix<double, BR, BC> block(int r0, int c0) const;
  Matrix<double, R, 1> col(int j) const;
  Matrix<double, 1, C> row(int i) const;
  double operator()(int i, int j) const { return coeff(i, j); }
  double operator()(int i) const { return lin(i); }
  double operator[](int i) const { return lin(i); }
  double x() const { return lin(0); }
  double y() const { return lin(1); }
  double z() const { return lin(2); }
  operator double() const
  {
    static_assert(R == 1 && C == 1, "only 1x1 converts to a scalar");
    return coeff(0, 0);
  }
};

template <typename P, int BR, int BC>
struct Block;
template <typename P>
struct DiagRef;
template <typename P>
struct SegRef;

template <typename P>
struct CommaInit
{
  P& p;
  int k;
  CommaInit(P& p_, double v) : p(p_), k(0) { put(v); }
  void put(double v)
  {
    const int cols = p.cols();
    p.ref(k / cols, k % cols) = v;  // row-major fill order
    k++;
  }
  template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  CommaInit& operator,(T v)
  {
    put((double)v);
    return *this;
  }
};

template <int R, int C>
class Matrix<double, R, C> : public Base<Matrix<double, R, C>, R, C>
{
public:
  double m[R * C];
  Matrix() {}
  Matrix(const Matrix&) = default;
  Matrix& operator=(const Matrix&) = default;
  template <typename O>
  Matrix(const Base<O, R, C>& o)
  {
    for (int j = 0; j < C; j++)
      for (int i = 0; i < R; i++) m[i + j * R] = o.coeff(i, j);
  }
  // row vector <-> column vector (Eigen transposes vectors implicitly on assignment)
  template <typename O, int R2 = R, int C2 = C, typename = typename std::enable_if<(R2 == 1 || C2 == 1) && R2 != C2>::type>
  Matrix(const Base<O, C, R>& o)
  {
    for (int i = 0; i < R * C; i++) m[i] = o.lin(i);
  }
  Matrix(double x, double y, double z)
  {
    static_assert(R * C == 3, "3-vector");
    m[0] = x;
    m[1] = y;
    m[2] = z;
  }
  Matrix(double x, double y)
  {
    static_assert(R * C == 2, "2-vector");
    m[0] = x;
    m[1] = y;
  }
  Matrix(double x, double y, double z, double w)
  {
    static_assert(R * C == 4, "4-vector");
    m[0] = x;
    m[1] = y;
    m[2] = z;
    m[3] = w;
  }
  int rows() const { return R; }
  int cols() const { return C; }
  double coeff(int i, int j) const { return m[i + j * R]; }
  double& ref(int i, int j) { return m[i + j * R]; }
  double& operator()(int i, int j) { return m[i + j * R]; }
  double operator()(int i, int j) const { return m[i + j * R]; }
  double& operator()(int i) { return m[i]; }
  double operator()(int i) const { return m[i]; }
  double& operator[](int i) { return m[i]; }
  double operator[](int i) const { return m[i]; }
  double* data() { return m; }
  const double* data() const { return m; }
  void setZero()
  {
    for (int i = 0; i < R * C; i++) m[i] = 0.0;
  }
  void setIdentity()
  {
    setZero();
    for (int i = 0; i < (R < C ? R : C); i++) m[i + i * R] = 1.0;
  }
  static Matrix Zero()
  {
    Matrix z;
    z.setZero();
    return z;
  }
  static Matrix Identity()
  {
    Matrix z;
    z.setIdentity();
    return z;
  }
  void setOnes()
  {
    for (int i = 0; i < R * C; i++) m[i] = 1.0;
  }
  // least squares through the normal equations: only the out-of-scope kd-tree IEKF calls this
  struct LsqSolver
  {
    const Matrix& A;
    template <typename O>
    Matrix<double, C, 1> solve(const Base<O, R, 1>& b) const
    {
      Matrix<double, C, R> At = A.transpose();
      Matrix<double, C, C> AtA = At * A;
      Matrix<double, C, 1> Atb = At * b;
      return AtA.inverse() * Atb;
    }
  };
  LsqSolver colPivHouseholderQr() const { return LsqSolver{ *this }; }
  static Matrix UnitZ()
  {
    Matrix z;
    z.setZero();
    z.m[2] = 1.0;
    return z;
  }
  template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  CommaInit<Matrix> operator<<(T v)
  {
    return CommaInit<Matrix>(*this, (double)v);
  }
  template <typename O>
  Matrix& operator+=(const Base<O, R, C>& o)
  {
    for (int j = 0; j < C; j++)
      for (int i = 0; i < R; i++) m[i + j * R] = m[i + j * R] + o.coeff(i, j);
    return *this;
  }
  template <typename O>
  Matrix& operator-=(const Base<O, R, C>& o)
  {
    for (int j = 0; j < C; j++)
      for (int i = 0; i < R; i++) m[i + j * R] = m[i + j * R] - o.coeff(i, j);
    return *this;
  }
  Matrix& operator*=(double s)
  {
    for (int i = 0; i < R * C; i++) m[i] = m[i] * s;
    return *this;
  }
  Matrix& operator/=(double s)
  {
    for (int i = 0; i < R * C; i++) m[i] = m[i] / s;
    return *this;
  }
  void normalize()
  {
    double z = this->squaredNorm();
    if (z > 0) *this /= std::sqrt(z);
  }
  // writable views
  using Base<Matrix, R, C>::block;
  using Base<Matrix, R, C>::col;
  using Base<Matrix, R, C>::row;
  using Base<Matrix, R, C>::diagonal;
  template <int BR, int BC>
  Block<Matrix, BR, BC> block(int r0, int c0)
  {
    return Block<Matrix, BR, BC>(*this, r0, c0);
  }
  Block<Matrix, R, 1> col(int j) { return Block<Matrix, R, 1>(*this, 0, j); }
  Block<Matrix, 1, C> row(int i) { return Block<Matrix, 1, C>(*this, i, 0); }  // writable: A.row(i) << x, y, z
  DiagRef<Matrix> diagonal() { return DiagRef<Matrix>(*this); }
  SegRef<Matrix> head(int n) { return SegRef<Matrix>(*this, 0, n); }
  SegRef<Matrix> tail(int n) { return SegRef<Matrix>(*this, R * C - n, n); }
  Matrix<double, Dynamic, 1> head(int n) const;
};

// dynamic matrix / vector (eager, runtime dimensions): what factors.cpp, imu_preintegration.cpp and optimizers.cpp
// use - fixed-size block views, topRows / leftCols / head / tail views, diagonal(), +, -, scalar and matrix
// products, dot, ldlt().solve()
template <typename P>
struct DynView  // rows [r0, r0+nr) x cols [c0, c0+nc) of a dynamic matrix
{
  P& p;
  int r0, c0, nr, nc;
  DynView(P& p_, int r, int c, int rr, int cc) : p(p_), r0(r), c0(c), nr(rr), nc(cc) {}
  void setZero()
  {
    for (int j = 0; j < nc; j++)
      for (int i = 0; i < nr; i++) p.ref(r0 + i, c0 + j) = 0.0;
  }
  template <typename Q>
  DynView& operator+=(const DynView<Q>& o)
  {
    for (int j = 0; j < nc; j++)
      for (int i = 0; i < nr; i++) p.ref(r0 + i, c0 + j) = p.ref(r0 + i, c0 + j) + o.p.coeff(o.r0 + i, o.c0 + j);
    return *this;
  }
  double coeff(int i, int j = 0) const { return p.coeff(r0 + i, c0 + j); }
};
template <typename P>
struct DynDiag
{
  P& p;
  explicit DynDiag(P& p_) : p(p_) {}
  DynDiag& operator=(const DynDiag& o)
  {
    for (int i = 0; i < p.rows(); i++) p.ref(i, i) = o.p.coeff(i, i);
    return *this;
  }
};
template <int C>
class Matrix<double, Dynamic, C>;
struct DynLDLT;

template <int C>
class Matrix<double, Dynamic, C>
{
public:
  std::vector<double> v;
  int r = 0, c = 0;
  Matrix() {}
  Matrix(int rr, int cc = 1) : v((size_t)rr * cc, 0.0), r(rr), c(cc) {}
  template <typename O, int R2, int C2>
  Matrix(const Base<O, R2, C2>& o) : v((size_t)R2 * C2), r(R2), c(C2)
  {
    for (int j = 0; j < C2; j++)
      for (int i = 0; i < R2; i++) v[i + (size_t)j * r] = o.coeff(i, j);
  }
  template <typename O, int R2, int C2>
  Matrix& operator=(const Base<O, R2, C2>& o)
  {
    Matrix<double, R2, C2> t(o);
    resize(R2, C2);
    for (int j = 0; j < C2; j++)
      for (int i = 0; i < R2; i++) v[i + (size_t)j * r] = t.coeff(i, j);
    return *this;
  }
  void resize(int rr, int cc = 1)
  {
    if (rr != r || cc != c)
    {
      v.assign((size_t)rr * cc, 0.0);
      r = rr;
      c = cc;
    }
  }
  void setZero() { std::fill(v.begin(), v.end(), 0.0); }
  void setIdentity()
  {
    setZero();
    for (int i = 0; i < (r < c ? r : c); i++) v[i + (size_t)i * r] = 1.0;
  }
  int rows() const { return r; }
  int cols() const { return c; }
  int size() const { return r * c; }
  double& operator()(int i, int j = 0) { return v[i + (size_t)j * r]; }
  double operator()(int i, int j = 0) const { return v[i + (size_t)j * r]; }
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double coeff(int i, int j = 0) const { return v[i + (size_t)j * r]; }
  double& ref(int i, int j = 0) { return v[i + (size_t)j * r]; }
  const double* data() const { return v.data(); }
  double* data() { return v.data(); }
  // fixed-size views (factors.cpp: Hess.block<6, 6>(6 * i, 6 * j) += ...)
  template <int BR, int BC>
  Block<Matrix, BR, BC> block(int r0, int c0)
  {
    return Block<Matrix, BR, BC>(*this, r0, c0);
  }
  template <int BR, int BC>
  Matrix<double, BR, BC> block(int r0, int c0) const
  {
    Matrix<double, BR, BC> t;
    for (int j = 0; j < BC; j++)
      for (int i = 0; i < BR; i++) t.ref(i, j) = coeff(r0 + i, c0 + j);
    return t;
  }
  DynView<Matrix> topRows(int n) { return DynView<Matrix>(*this, 0, 0, n, c); }
  DynView<Matrix> leftCols(int n) { return DynView<Matrix>(*this, 0, 0, r, n); }
  DynView<Matrix> head(int n) { return DynView<Matrix>(*this, 0, 0, n, 1); }
  DynView<Matrix> tail(int n) { return DynView<Matrix>(*this, r - n, 0, n, 1); }
  DynDiag<Matrix> diagonal() { return DynDiag<Matrix>(*this); }
  Matrix& operator+=(const Matrix& o)
  {
    for (size_t i = 0; i < v.size(); i++) v[i] = v[i] + o.v[i];
    return *this;
  }
  Matrix& operator*=(double s)
  {
    for (size_t i = 0; i < v.size(); i++) v[i] = v[i] * s;
    return *this;
  }
  double dot(const Matrix& o) const
  {
    double s = v[0] * o.v[0];
    for (size_t i = 1; i < v.size(); i++) s = s + v[i] * o.v[i];
    return s;
  }
  inline DynLDLT ldlt() const;
};
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;

// a fixed vector += a runtime tail / head of a dynamic one (optimizers.cpp: x.g += dxi.tail(3))
template <int C, int R2>
Matrix<double, R2, 1>& operator+=(Matrix<double, R2, 1>& a, const DynView<Matrix<double, Dynamic, C>>& o)
{
  for (int i = 0; i < R2; i++) a[i] = a[i] + o.coeff(i, 0);
  return a;
}
template <int C>
Matrix<double, Dynamic, C> operator+(const Matrix<double, Dynamic, C>& a, const Matrix<double, Dynamic, C>& b)
{
  Matrix<double, Dynamic, C> t(a.r, a.c);
  for (size_t i = 0; i < t.v.size(); i++) t.v[i] = a.v[i] + b.v[i];
  return t;
}
template <int C>
Matrix<double, Dynamic, C> operator-(const Matrix<double, Dynamic, C>& a, const Matrix<double, Dynamic, C>& b)
{
  Matrix<double, Dynamic, C> t(a.r, a.c);
  for (size_t i = 0; i < t.v.size(); i++) t.v[i] = a.v[i] - b.v[i];
  return t;
}
template <int C>
Matrix<double, Dynamic, C> operator-(const Matrix<double, Dynamic, C>& a)
{
  Matrix<double, Dynamic, C> t(a.r, a.c);
  for (size_t i = 0; i < t.v.size(); i++) t.v[i] = -a.v[i];
  return t;
}
template <int C>
Matrix<double, Dynamic, C> operator*(double s, const Matrix<double, Dynamic, C>& a)
{
  Matrix<double, Dynamic, C> t(a.r, a.c);
  for (size_t i = 0; i < t.v.size(); i++) t.v[i] = s * a.v[i];
  return t;
}
template <int C>
Matrix<double, Dynamic, C> operator*(const Matrix<double, Dynamic, C>& a, double s)
{
  Matrix<double, Dynamic, C> t(a.r, a.c);
  for (size_t i = 0; i < t.v.size(); i++) t.v[i] = a.v[i] * s;
  return t;
}
template <int C1, int C2>
Matrix<double, Dynamic, C2> operator*(const Matrix<double, Dynamic, C1>& a, const Matrix<double, Dynamic, C2>& b)
{
  Matrix<double, Dynamic, C2> t(a.r, b.c);
  for (int j = 0; j < b.c; j++)
    for (int i = 0; i < a.r; i++)
    {
      double s = a.coeff(i, 0) * b.coeff(0, j);
      for (int k = 1; k < a.c; k++) s = s + a.coeff(i, k) * b.coeff(k, j);
      t.ref(i, j) = s;
    }
  return t;
}

// A.ldlt().solve(b): Eigen's LDLT (robust Cholesky with diagonal pivoting) restated as the textbook unblocked
// algorithm on the lower triangle: pivot = largest |diagonal| of the trailing block, symmetric swap,
// A21 <- (A21 - A20 (D .* A10^T)) / d_k. Eigen itself is absent: parity UNPINNED at this call;
// oracle/omat.hpp holds the same routine so restatement and reference build agree with each other.
struct DynLDLT
{
  int n = 0;
  std::vector<double> L;  // column-major n x n, unit lower triangle below the diagonal, D on the diagonal
  std::vector<int> perm;  // transpositions
  explicit DynLDLT(const MatrixXd& A) : n(A.rows()), L(A.v), perm(A.rows())
  {
    auto a = [&](int i, int j) -> double& { return L[i + (size_t)j * n]; };
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
        // symmetric swap of rows / columns k and p within the lower triangle
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
  }
  VectorXd solve(const VectorXd& b) const
  {
    auto a = [&](int i, int j) -> double { return L[i + (size_t)j * n]; };
    VectorXd x = b;
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
};
template <int C>
inline DynLDLT Matrix<double, Dynamic, C>::ldlt() const
{
  return DynLDLT(*this);
}

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;

// ---- views -------------------------------------------------------------------------------------------
template <typename P, int BR, int BC>
struct Block : public Base<Block<P, BR, BC>, BR, BC>
{
  P& p;
  int r0, c0;
  Block(P& p_, int r, int c) : p(p_), r0(r), c0(c) {}
  int rows() const { return BR; }
  int cols() const { return BC; }
  double coeff(int i, int j) const { return p.coeff(r0 + i, c0 + j); }
  double& ref(int i, int j) { return p.ref(r0 + i, c0 + j); }
  template <typename O>
  Block& operator=(const Base<O, BR, BC>& o)
  {
    Matrix<double, BR, BC> t(o);  // evaluate first (aliasing-safe)
    for (int j = 0; j < BC; j++)
      for (int i = 0; i < BR; i++) ref(i, j) = t.coeff(i, j);
    return *this;
  }
  Block& operator=(const Block& o) { return operator=<Block>(o); }
  // vector assigned to a vector block of the other orientation
  template <typename O, int R2 = BR, int C2 = BC, typename = typename std::enable_if<(R2 == 1 || C2 == 1) && R2 != C2>::type>
  Block& operator=(const Base<O, BC, BR>& o)
  {
    Matrix<double, BC, BR> t(o);
    for (int i = 0; i < BR * BC; i++) (BC == 1 ? ref(i, 0) : ref(0, i)) = t.lin(i);
    return *this;
  }
  template <typename O>
  Block& operator+=(const Base<O, BR, BC>& o)
  {
    Matrix<double, BR, BC> t(o);
    for (int j = 0; j < BC; j++)
      for (int i = 0; i < BR; i++) ref(i, j) = ref(i, j) + t.coeff(i, j);
    return *this;
  }
  template <typename O>
  void swap(Base<O, BR, BC>&& o)
  {
    O& od = o.d();
    for (int j = 0; j < BC; j++)
      for (int i = 0; i < BR; i++) std::swap(ref(i, j), od.ref(i, j));
  }
  DiagRef<Block> diagonal() { return DiagRef<Block>(*this); }
  void setZero()
  {
    for (int j = 0; j < BC; j++)
      for (int i = 0; i < BR; i++) ref(i, j) = 0.0;
  }
  void setIdentity()
  {
    for (int j = 0; j < BC; j++)
      for (int i = 0; i < BR; i++) ref(i, j) = i == j ? 1.0 : 0.0;
  }
  template <int C2>
  Block& operator+=(const Matrix<double, Dynamic, C2>& o)  // a dynamic matrix of the block's size
  {
    for (int j = 0; j < BC; j++)
      for (int i = 0; i < BR; i++) ref(i, j) = ref(i, j) + o.coeff(i, j);
    return *this;
  }
  template <typename Q>
  Block& operator+=(const DynView<Q>& o)
  {
    for (int j = 0; j < BC; j++)
      for (int i = 0; i < BR; i++) ref(i, j) = ref(i, j) + o.coeff(i, j);
    return *this;
  }
  template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  CommaInit<Block> operator<<(T v)
  {
    return CommaInit<Block>(*this, (double)v);
  }
};

template <typename P>
struct DiagRef
{
  P& p;
  explicit DiagRef(P& p_) : p(p_) {}
  template <typename O, int N>
  DiagRef& operator=(const Base<O, N, 1>& o)
  {
    for (int i = 0; i < N; i++) p.ref(i, i) = o.coeff(i, 0);
    return *this;
  }
  template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  CommaInit<DiagRef> operator<<(T v)
  {
    return CommaInit<DiagRef>(*this, (double)v);
  }
  int cols() const { return 1; }
  double& ref(int i, int) { return p.ref(i, i); }
  template <int N>
  operator Matrix<double, N, 1>() const
  {
    Matrix<double, N, 1> t;
    for (int i = 0; i < N; i++) t[i] = p.coeff(i, i);
    return t;
  }
};

// runtime-length segment of a fixed vector: v.head(n) = ..., v.tail(n) = ..., v.head(n) << ...
template <typename P>
struct SegRef
{
  P& p;
  int off, n;
  SegRef(P& p_, int o, int n_) : p(p_), off(o), n(n_) {}
  int cols() const { return p.cols() == 1 ? 1 : n; }
  double& ref(int i, int j) { return p[off + (p.cols() == 1 ? i : j)]; }
  template <typename O, int R2, int C2>
  SegRef& operator=(const Base<O, R2, C2>& o)
  {
    Matrix<double, R2, C2> t(o);
    for (int i = 0; i < n; i++) p[off + i] = t.lin(i);
    return *this;
  }
  template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  CommaInit<SegRef> operator<<(T v)
  {
    return CommaInit<SegRef>(*this, (double)v);
  }
};

// ---- Base members ------------------------------------------------------------------------------------
template <typename D, int R, int C>
Matrix<double, R, C> Base<D, R, C>::eval() const
{
  return Matrix<double, R, C>(*this);
}
template <typename D, int R, int C>
Matrix<double, C, R> Base<D, R, C>::transpose() const
{
  Matrix<double, C, R> t;
  for (int j = 0; j < C; j++)
    for (int i = 0; i < R; i++) t.ref(j, i) = coeff(i, j);
  return t;
}
template <typename D, int R, int C>
template <typename O>
Matrix<double, 3, 1> Base<D, R, C>::cross(const Base<O, 3, 1>& o) const
{
  const double a0 = lin(0), a1 = lin(1), a2 = lin(2), b0 = o.lin(0), b1 = o.lin(1), b2 = o.lin(2);
  return Matrix<double, 3, 1>(a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0);
}
template <typename D, int R, int C>
Matrix<double, R, C> Base<D, R, C>::normalized() const
{
  Matrix<double, R, C> t(*this);
  t.normalize();
  return t;
}
template <typename D, int R, int C>
Matrix<double, (R < C ? R : C), 1> Base<D, R, C>::diagonal() const
{
  Matrix<double, (R < C ? R : C), 1> t;
  for (int i = 0; i < (R < C ? R : C); i++) t[i] = coeff(i, i);
  return t;
}
template <typename D, int R, int C>
Matrix<double, R, R> Base<D, R, C>::asDiagonal() const
{
  static_assert(C == 1, "asDiagonal of a column vector");
  Matrix<double, R, R> t;
  t.setZero();
  for (int i = 0; i < R; i++) t.ref(i, i) = coeff(i, 0);
  return t;
}
template <typename D, int R, int C>
template <int BR, int BC>
Matrix<double, BR, BC> Base<D, R, C>::block(int r0, int c0) const
{
  Matrix<double, BR, BC> t;
  for (int j = 0; j < BC; j++)
    for (int i = 0; i < BR; i++) t.ref(i, j) = coeff(r0 + i, c0 + j);
  return t;
}
template <typename D, int R, int C>
Matrix<double, R, 1> Base<D, R, C>::col(int j) const
{
  return block<R, 1>(0, j);
}
template <typename D, int R, int C>
Matrix<double, 1, C> Base<D, R, C>::row(int i) const
{
  return block<1, C>(i, 0);
}
// inverse(): PartialPivLU for N > 4 as in Eigen; the hot path only inverts 15x15 (odometry.cpp:82, 194)
template <typename D, int R, int C>
Matrix<double, R, C> Base<D, R, C>::inverse() const
{
  static_assert(R == C, "square");
  constexpr int N = R;
  Matrix<double, N, N> lu(*this), inv;
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
      for (int j = 0; j < N; j++) std::swap(lu(k, j), lu(piv, j));
      std::swap(perm[k], perm[piv]);
    }
    for (int i = k + 1; i < N; i++)
    {
      lu(i, k) = lu(i, k) / lu(k, k);
      for (int j = k + 1; j < N; j++) lu(i, j) = lu(i, j) - lu(i, k) * lu(k, j);
    }
  }
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

// ---- operators (eager) -------------------------------------------------------------------------------
template <typename A, typename B, int R, int C>
Matrix<double, R, C> operator+(const Base<A, R, C>& a, const Base<B, R, C>& b)
{
  Matrix<double, R, C> t;
  for (int j = 0; j < C; j++)
    for (int i = 0; i < R; i++) t.ref(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return t;
}
template <typename A, typename B, int R, int C>
Matrix<double, R, C> operator-(const Base<A, R, C>& a, const Base<B, R, C>& b)
{
  Matrix<double, R, C> t;
  for (int j = 0; j < C; j++)
    for (int i = 0; i < R; i++) t.ref(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return t;
}
template <typename A, int R, int C>
Matrix<double, R, C> operator-(const Base<A, R, C>& a)
{
  Matrix<double, R, C> t;
  for (int j = 0; j < C; j++)
    for (int i = 0; i < R; i++) t.ref(i, j) = -a.coeff(i, j);
  return t;
}
template <typename A, int R, int C, typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
Matrix<double, R, C> operator*(const Base<A, R, C>& a, T s)
{
  Matrix<double, R, C> t;
  for (int j = 0; j < C; j++)
    for (int i = 0; i < R; i++) t.ref(i, j) = a.coeff(i, j) * (double)s;
  return t;
}
template <typename A, int R, int C, typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
Matrix<double, R, C> operator*(T s, const Base<A, R, C>& a)
{
  Matrix<double, R, C> t;
  for (int j = 0; j < C; j++)
    for (int i = 0; i < R; i++) t.ref(i, j) = (double)s * a.coeff(i, j);
  return t;
}
template <typename A, int R, int C, typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
Matrix<double, R, C> operator/(const Base<A, R, C>& a, T s)
{
  Matrix<double, R, C> t;
  for (int j = 0; j < C; j++)
    for (int i = 0; i < R; i++) t.ref(i, j) = a.coeff(i, j) / (double)s;
  return t;
}
template <typename A, typename B, int R, int K, int C>
Matrix<double, R, C> operator*(const Base<A, R, K>& a, const Base<B, K, C>& b)
{
  Matrix<double, R, C> t;
  for (int j = 0; j < C; j++)
    for (int i = 0; i < R; i++)
    {
      double s = a.coeff(i, 0) * b.coeff(0, j);
      for (int k = 1; k < K; k++) s = s + a.coeff(i, k) * b.coeff(k, j);
      t.ref(i, j) = s;
    }
  return t;
}

// ---- SelfAdjointEigenSolver<Matrix3d>: Eigen 3.4.0 compute() restated --------------------------------
template <typename M>
class SelfAdjointEigenSolver;
template <>
class SelfAdjointEigenSolver<Matrix3d>
{
  Vector3d vals;
  Matrix3d vecs;
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

public:
  template <typename O>
  explicit SelfAdjointEigenSolver(const Base<O, 3, 3>& A)
  {
    Matrix3d mat = Matrix3d::Zero();
    for (int j = 0; j < 3; j++)
      for (int i = j; i < 3; i++) mat(i, j) = A.coeff(i, j);
    double scale = 0.0;
    for (int j = 0; j < 3; j++)
      for (int i = j; i < 3; i++) scale = std::max(scale, std::fabs(mat(i, j)));
    if (scale == 0.0) scale = 1.0;
    for (int j = 0; j < 3; j++)
      for (int i = j; i < 3; i++) mat(i, j) = mat(i, j) / scale;
    double diag[3], sub[2];
    const double tol = 2.2250738585072014e-308;
    diag[0] = mat(0, 0);
    double v1norm2 = mat(2, 0) * mat(2, 0);
    if (v1norm2 <= tol)
    {
      diag[1] = mat(1, 1);
      diag[2] = mat(2, 2);
      sub[0] = mat(1, 0);
      sub[1] = mat(2, 1);
      mat.setIdentity();
    }
    else
    {
      double beta = std::sqrt(mat(1, 0) * mat(1, 0) + v1norm2);
      double invBeta = 1.0 / beta;
      double m01 = mat(1, 0) * invBeta, m02 = mat(2, 0) * invBeta;
      double q = 2.0 * m01 * mat(2, 1) + m02 * (mat(2, 2) - mat(1, 1));
      diag[1] = mat(1, 1) + m02 * q;
      diag[2] = mat(2, 2) - m02 * q;
      sub[0] = beta;
      sub[1] = mat(2, 1) - m01 * q;
      mat.setZero();
      mat(0, 0) = 1;
      mat(1, 1) = m01;
      mat(1, 2) = m02;
      mat(2, 1) = m02;
      mat(2, 2) = -m01;
    }
    int end = 2, start = 0, iter = 0;
    const double pinv = 1.0 / 2.220446049250313e-16;
    while (end > 0)
    {
      for (int i = start; i < end; ++i)
      {
        if (std::fabs(sub[i]) < tol)
          sub[i] = 0.0;
        else
        {
          const double ss = pinv * sub[i];
          if (ss * ss <= (std::fabs(diag[i]) + std::fabs(diag[i + 1]))) sub[i] = 0.0;
        }
      }
      while (end > 0 && sub[end - 1] == 0.0) end--;
      if (end <= 0) break;
      iter++;
      if (iter > 90) break;
      start = end - 1;
      while (start > 0 && sub[start - 1] != 0.0) start--;
      double td = (diag[end - 1] - diag[end]) * 0.5;
      double e = sub[end - 1];
      double mu = diag[end];
      if (td == 0.0)
        mu -= std::fabs(e);
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
      double z = sub[start];
      for (int k = start; k < end && z != 0.0; ++k)
      {
        double c, s;
        makeGivens(x, z, c, s);
        double sdk = s * diag[k] + c * sub[k];
        double dkp1 = s * sub[k] + c * diag[k + 1];
        diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
        diag[k + 1] = s * sdk + c * dkp1;
        sub[k] = c * sdk - s * dkp1;
        if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
        x = sub[k];
        if (k < end - 1)
        {
          z = -s * sub[k + 1];
          sub[k + 1] = c * sub[k + 1];
        }
        for (int i = 0; i < 3; i++)
        {
          double xi = mat(i, k), yi = mat(i, k + 1);
          mat(i, k) = c * xi - s * yi;
          mat(i, k + 1) = s * xi + c * yi;
        }
      }
    }
    if (iter <= 90)
      for (int i = 0; i < 2; ++i)
      {
        int k = 0;
        double mn = diag[i];
        for (int j = 1; j < 3 - i; j++)
          if (diag[i + j] < mn)
          {
            mn = diag[i + j];
            k = j;
          }
        if (k > 0)
        {
          std::swap(diag[i], diag[k + i]);
          for (int r = 0; r < 3; r++) std::swap(mat(r, i), mat(r, k + i));
        }
      }
    for (int i = 0; i < 3; i++) vals[i] = diag[i] * scale;
    vecs = mat;
  }
  const Vector3d& eigenvalues() const { return vals; }
  const Matrix3d& eigenvectors() const { return vecs; }
};

// ---- Geometry bits the reference touches outside the hot path (compile-only fidelity) -----------------
class AngleAxisd
{
  Vector3d ax;
  double ang;

public:
  explicit AngleAxisd(const Matrix3d& R)
  {
    double c = 0.5 * (R.trace() - 1.0);
    c = std::max(-1.0, std::min(1.0, c));
    ang = std::acos(c);
    Vector3d k(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1));
    double n = k.norm();
    ax = n > 0 ? Vector3d(k / n) : Vector3d(1, 0, 0);
  }
  // (angle, axis) and toRotationMatrix(): Eigen 3.4.0 Geometry/AngleAxis.h (used by Initialization::align_gravity)
  AngleAxisd(double angle, const Vector3d& axis) : ax(axis), ang(angle) {}
  Matrix3d toRotationMatrix() const
  {
    Matrix3d res;
    const double s = std::sin(ang), c = std::cos(ang);
    const Vector3d sin_axis(s * ax[0], s * ax[1], s * ax[2]);
    const Vector3d cos1_axis((1.0 - c) * ax[0], (1.0 - c) * ax[1], (1.0 - c) * ax[2]);
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
  const Vector3d& axis() const { return ax; }
  double angle() const { return ang; }
};
class Quaterniond
{
  double q[4];

public:
  Quaterniond() : q{ 0, 0, 0, 1 } {}
  static Quaterniond FromTwoVectors(const Vector3d& a, const Vector3d& b)
  {
    Quaterniond r;
    Vector3d v0 = a.normalized(), v1 = b.normalized();
    double c = v0.dot(v1);
    Vector3d ax = v0.cross(v1);
    double s = std::sqrt(std::max(0.0, (1.0 + c) * 2.0));
    if (s > 1e-12)
    {
      r.q[0] = ax[0] / s;
      r.q[1] = ax[1] / s;
      r.q[2] = ax[2] / s;
      r.q[3] = s * 0.5;
    }
    return r;
  }
  double x() const { return q[0]; }
  double y() const { return q[1]; }
  double z() const { return q[2]; }
  double w() const { return q[3]; }
};
}  // namespace Eigen

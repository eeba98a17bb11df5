// Small fixed-size math shared by the kernels and the host-side solve.
// Decision-bearing expressions (world point, voxel key, child index, inside(),
// the fp32 gate, cluster sums and the 3x3 eigen-decomposition that feeds
// plane_judge) are written with dm/da/ds = single-rounding multiply/add/subtract
// so that neither nvcc (-fmad) nor g++ (-ffp-contract) can fuse them: the
// rounding sequence is the one fixed in SURVEY.md Appendix A.
#pragma once
#include <cmath>
#include <cstdint>
#include "vn_types.cuh"

#ifdef __CUDACC__
#define VN_HD __host__ __device__ __forceinline__
#else
#define VN_HD inline
#endif

#ifdef __CUDA_ARCH__
VN_HD double dm(double a, double b) { return __dmul_rn(a, b); }
VN_HD double da(double a, double b) { return __dadd_rn(a, b); }
VN_HD double ds(double a, double b) { return __dsub_rn(a, b); }
VN_HD float fm(float a, float b) { return __fmul_rn(a, b); }
VN_HD float fa(float a, float b) { return __fadd_rn(a, b); }
VN_HD float fs(float a, float b) { return __fsub_rn(a, b); }
VN_HD float fdv(float a, float b) { return __fdiv_rn(a, b); }
#else
// host objects are built with -ffp-contract=off
VN_HD double dm(double a, double b) { return a * b; }
VN_HD double da(double a, double b) { return a + b; }
VN_HD double ds(double a, double b) { return a - b; }
VN_HD float fm(float a, float b) { return a * b; }
VN_HD float fa(float a, float b) { return a + b; }
VN_HD float fs(float a, float b) { return a - b; }
VN_HD float fdv(float a, float b) { return a / b; }
#endif

// packed index of a symmetric 3x3: [0 1 2; 1 3 4; 2 4 5]
VN_HD int s3(int i, int j)
{
  int a = i < j ? i : j, b = i < j ? j : i;
  return a * 3 - (a * (a - 1)) / 2 + (b - a);
}
// packed index (upper triangle by rows) of a symmetric n x n
VN_HD int sN(int n, int i, int j)
{
  int a = i < j ? i : j, b = i < j ? j : i;
  return a * n - (a * (a - 1)) / 2 + (b - a);
}

// dot of 3-vectors, left to right (Eigen's fixed-size redux order)
VN_HD double dot3(const double* a, const double* b) { return da(da(dm(a[0], b[0]), dm(a[1], b[1])), dm(a[2], b[2])); }

// w = R*p + t with R column-major: ((R(i,0)p0 + R(i,1)p1) + R(i,2)p2) + t(i)   (Appendix A.1)
VN_HD void rot_trans(const double* R, const double* t, const double* p, double* w)
{
#pragma unroll
  for (int i = 0; i < 3; i++) w[i] = da(da(da(dm(R[i], p[0]), dm(R[i + 3], p[1])), dm(R[i + 6], p[2])), t[i]);
}
VN_HD void rot_vec(const double* R, const double* p, double* w)
{
#pragma unroll
  for (int i = 0; i < 3; i++) w[i] = da(da(dm(R[i], p[0]), dm(R[i + 3], p[1])), dm(R[i + 6], p[2]));
}
// w = R^T * p
VN_HD void rotT_vec(const double* R, const double* p, double* w)
{
#pragma unroll
  for (int i = 0; i < 3; i++) w[i] = da(da(dm(R[3 * i], p[0]), dm(R[3 * i + 1], p[1])), dm(R[3 * i + 2], p[2]));
}
// C = A*B, 3x3 column-major, coefficients left to right
VN_HD void mat3_mul(const double* A, const double* B, double* C)
{
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++)
      C[i + 3 * j] = da(da(dm(A[i], B[3 * j]), dm(A[i + 3], B[3 * j + 1])), dm(A[i + 6], B[3 * j + 2]));
}
// C = A*B^T
VN_HD void mat3_mulT(const double* A, const double* B, double* C)
{
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++)
      C[i + 3 * j] = da(da(dm(A[i], B[j]), dm(A[i + 3], B[j + 3])), dm(A[i + 6], B[j + 6]));
}
VN_HD void hat3(const double* v, double* O)
{
  O[0] = 0;
  O[3] = -v[2];
  O[6] = v[1];
  O[1] = v[2];
  O[4] = 0;
  O[7] = -v[0];
  O[2] = -v[1];
  O[5] = v[0];
  O[8] = 0;
}
VN_HD void sym6_to_full(const double* s, double* M)
{
  M[0] = s[0];
  M[3] = s[1];
  M[6] = s[2];
  M[1] = s[1];
  M[4] = s[3];
  M[7] = s[4];
  M[2] = s[2];
  M[5] = s[4];
  M[8] = s[5];
}

// voxel key (src/mapping/voxel_map.cpp:246-253, Appendix A.2):
// double divide -> round to float -> "-1 if negative" in float -> truncate
VN_HD long long voxel_coord(double w, double voxel_size)
{
  float f = (float)(w / voxel_size);
  if (f < 0) f = fs(f, 1.0f);
  return (long long)f;
}
// 3 x 21-bit packing with bias; bit 63 marks "occupied". false when out of range.
VN_HD bool pack_key(long long x, long long y, long long z, unsigned long long* out)
{
  long long bx = x + VN_KEY_BIAS, by = y + VN_KEY_BIAS, bz = z + VN_KEY_BIAS;
  if ((unsigned long long)bx > VN_KEY_MASK || (unsigned long long)by > VN_KEY_MASK ||
      (unsigned long long)bz > VN_KEY_MASK)
    return false;
  *out = (1ull << 63) | ((unsigned long long)bx << 42) | ((unsigned long long)by << 21) | (unsigned long long)bz;
  return true;
}
VN_HD void unpack_key(unsigned long long k, long long* xyz)
{
  xyz[0] = (long long)((k >> 42) & VN_KEY_MASK) - VN_KEY_BIAS;
  xyz[1] = (long long)((k >> 21) & VN_KEY_MASK) - VN_KEY_BIAS;
  xyz[2] = (long long)(k & VN_KEY_MASK) - VN_KEY_BIAS;
}
VN_HD unsigned int hash_key(unsigned long long k)
{
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (unsigned int)k;
}
// owner rank of a root voxel in a map sharded by hash range over `world` ranks
VN_HD int shard_owner(unsigned long long key, int world)
{
  return (int)(((unsigned long long)hash_key(key) * (unsigned long long)world) >> 32);
}
// child index (octree.cpp:211-215 == 584-588): strict >
VN_HD int child_index(const double* w, const double* c)
{
  return 4 * (w[0] > c[0] ? 1 : 0) + 2 * (w[1] > c[1] ? 1 : 0) + (w[2] > c[2] ? 1 : 0);
}
// OctoTree::inside (octree.cpp:732-737): hl = 2*quater_length (float -> double), closed box
VN_HD bool inside_box(const double* w, const double* c, float ql)
{
  double hl = (double)fm(ql, 2.0f);
  return (w[0] >= ds(c[0], hl) && w[0] <= da(c[0], hl) && w[1] >= ds(c[1], hl) && w[1] <= da(c[1], hl) &&
          w[2] >= ds(c[2], hl) && w[2] <= da(c[2], hl));
}

// ---------------------------------------------------------------------------
// Symmetric 3x3 eigen-decomposition: Eigen 3.4.0 SelfAdjointEigenSolver<Matrix3d>::compute
// (scaling, 3x3 tridiagonalisation, implicit symmetric QR with Wilkinson shift,
// ascending sort). The reference calls it at octree.cpp:362, 435 and
// odometry.cpp:244; plane_judge (octree.cpp:198-201) consumes the result, so
// this routine is decision-bearing. Input: lower triangle L = (0,0),(1,0),(2,0),(1,1),(2,1),(2,2).
// Output: values ascending, vectors column-major.
VN_HD void givens(double p, double q, double& c, double& s)
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
  else if (fabs(p) > fabs(q))
  {
    double t = q / p;
    double u = sqrt(da(1.0, dm(t, t)));
    if (p < 0) u = -u;
    c = 1.0 / u;
    s = dm(-t, c);
  }
  else
  {
    double t = p / q;
    double u = sqrt(da(1.0, dm(t, t)));
    if (q < 0) u = -u;
    s = -1.0 / u;
    c = dm(-t, s);
  }
}
VN_HD double hypot_pos(double x, double y)
{
  double ax = fabs(x), ay = fabs(y);
  double p = ax > ay ? ax : ay;
  if (p == 0.0) return 0.0;
  double qp = (ax > ay ? ay : ax) / p;
  return dm(p, sqrt(da(1.0, dm(qp, qp))));
}
// One implicit QR step on rows/columns START..END of the tridiagonal matrix (Eigen's tridiagonal_qr_step). START / END
// are compile-time constants - a 3x3 matrix only ever sees (0,2), (1,2) and (0,1) - so that every diag / sub / Q index
// is static and the whole solver lives in registers (dynamically indexed arrays would go to local memory: the
// solver is a single dependent chain, every local-memory round trip is on it). Same operations, same order.
template <int START, int END>
VN_HD void eig3_qr_step(double (&diag)[3], double (&sub)[2], double (&Q)[9])
{
  double td = dm(ds(diag[END - 1], diag[END]), 0.5);
  double e = sub[END - 1];
  double mu = diag[END];
  if (td == 0.0)
    mu = ds(mu, fabs(e));
  else if (e != 0.0)
  {
    const double e2 = dm(e, e);
    const double h = hypot_pos(td, e);
    if (e2 == 0.0)
      mu = ds(mu, e / (da(td, (td > 0.0 ? h : -h)) / e));
    else
      mu = ds(mu, e2 / da(td, (td > 0.0 ? h : -h)));
  }
  double x = ds(diag[START], mu);
  double z = sub[START];
#pragma unroll
  for (int k = START; k < END; ++k)
  {
    if (z == 0.0) break;
    double c, s;
    givens(x, z, c, s);
    double sdk = da(dm(s, diag[k]), dm(c, sub[k]));
    double dkp1 = da(dm(s, sub[k]), dm(c, diag[k + 1]));
    diag[k] = ds(dm(c, ds(dm(c, diag[k]), dm(s, sub[k]))), dm(s, ds(dm(c, sub[k]), dm(s, diag[k + 1]))));
    diag[k + 1] = da(dm(s, sdk), dm(c, dkp1));
    sub[k] = ds(dm(c, sdk), dm(s, dkp1));
    if (k > START) sub[k - 1] = ds(dm(c, sub[k - 1]), dm(s, z));
    x = sub[k];
    if (k < END - 1)
    {
      z = dm(-s, sub[k + 1]);
      sub[k + 1] = dm(c, sub[k + 1]);
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
      double xi = Q[i + 3 * k], yi = Q[i + 3 * (k + 1)];
      Q[i + 3 * k] = ds(dm(c, xi), dm(s, yi));
      Q[i + 3 * (k + 1)] = da(dm(s, xi), dm(c, yi));
    }
  }
}
// sub[I] is negligible against its diagonal neighbours (Eigen's deflation test)
template <int I>
VN_HD void eig3_deflate(const double (&diag)[3], double (&sub)[2])
{
  const double tol = 2.2250738585072014e-308;
  const double precision_inv = 1.0 / 2.220446049250313e-16;
  if (fabs(sub[I]) < tol)
    sub[I] = 0.0;
  else
  {
    const double ss = dm(precision_inv, sub[I]);
    if (dm(ss, ss) <= da(fabs(diag[I]), fabs(diag[I + 1]))) sub[I] = 0.0;
  }
}
VN_HD void eig3_swap_cols(double (&diag)[3], double (&Q)[9], int a, int b)
{
  double t = diag[a];
  diag[a] = diag[b];
  diag[b] = t;
#pragma unroll
  for (int r = 0; r < 3; r++)
  {
    double tt = Q[r + 3 * a];
    Q[r + 3 * a] = Q[r + 3 * b];
    Q[r + 3 * b] = tt;
  }
}
VN_HD void eig3_sym(const double* L, double* values, double* Qout)
{
  double m00 = L[0], m10 = L[1], m20 = L[2], m11 = L[3], m21 = L[4], m22 = L[5];
  double scale = fabs(m00);
  if (fabs(m10) > scale) scale = fabs(m10);
  if (fabs(m20) > scale) scale = fabs(m20);
  if (fabs(m11) > scale) scale = fabs(m11);
  if (fabs(m21) > scale) scale = fabs(m21);
  if (fabs(m22) > scale) scale = fabs(m22);
  if (scale == 0.0) scale = 1.0;
  m00 = m00 / scale;
  m10 = m10 / scale;
  m20 = m20 / scale;
  m11 = m11 / scale;
  m21 = m21 / scale;
  m22 = m22 / scale;

  double diag[3], sub[2], Q[9];
  const double tol = 2.2250738585072014e-308;
  diag[0] = m00;
  double v1norm2 = dm(m20, m20);
#pragma unroll
  for (int i = 0; i < 9; i++) Q[i] = 0.0;
  if (v1norm2 <= tol)
  {
    diag[1] = m11;
    diag[2] = m22;
    sub[0] = m10;
    sub[1] = m21;
    Q[0] = 1.0;
    Q[4] = 1.0;
    Q[8] = 1.0;
  }
  else
  {
    double beta = sqrt(da(dm(m10, m10), v1norm2));
    double invBeta = 1.0 / beta;
    double m01 = dm(m10, invBeta);
    double m02 = dm(m20, invBeta);
    double q = da(dm(dm(2.0, m01), m21), dm(m02, ds(m22, m11)));
    diag[1] = da(m11, dm(m02, q));
    diag[2] = ds(m22, dm(m02, q));
    sub[0] = beta;
    sub[1] = ds(m21, dm(m01, q));
    Q[0] = 1.0;
    Q[4] = m01;
    Q[7] = m02;  // (1,2)
    Q[5] = m02;  // (2,1)
    Q[8] = -m01;
  }

  int end = 2, start = 0, iter = 0;
  while (end > 0)
  {
    // for (i = start; i < end; ++i) deflation test of sub[i]
    if (start <= 0 && 0 < end) eig3_deflate<0>(diag, sub);
    if (start <= 1 && 1 < end) eig3_deflate<1>(diag, sub);
    // while (end > 0 && sub[end - 1] == 0) end--
    if (end == 2 && sub[1] == 0.0) end = 1;
    if (end == 1 && sub[0] == 0.0) end = 0;
    if (end <= 0) break;
    iter++;
    if (iter > 90) break;
    // start = end - 1; while (start > 0 && sub[start - 1] != 0) start--
    start = end - 1;
    if (start == 1 && sub[0] != 0.0) start = 0;
    if (end == 2)
    {
      if (start == 0)
        eig3_qr_step<0, 2>(diag, sub, Q);
      else
        eig3_qr_step<1, 2>(diag, sub, Q);
    }
    else
      eig3_qr_step<0, 1>(diag, sub, Q);
  }
  if (iter <= 90)
  {
    // selection sort, ascending, first minimum wins (Eigen's sort of the eigenvalues with their vectors)
    int k = 0;
    double mn = diag[0];
    if (diag[1] < mn)
    {
      mn = diag[1];
      k = 1;
    }
    if (diag[2] < mn) k = 2;
    if (k == 1)
      eig3_swap_cols(diag, Q, 0, 1);
    else if (k == 2)
      eig3_swap_cols(diag, Q, 0, 2);
    if (diag[2] < diag[1]) eig3_swap_cols(diag, Q, 1, 2);
  }
#pragma unroll
  for (int i = 0; i < 3; i++) values[i] = dm(diag[i], scale);
#pragma unroll
  for (int i = 0; i < 9; i++) Qout[i] = Q[i];
}

// PointCluster::cov() lower triangle (types.hpp:144-148)
VN_HD void cluster_cov(const Cluster& c, double* L)
{
  double n = (double)c.N;
  double ctr[3] = { c.v[0] / n, c.v[1] / n, c.v[2] / n };
  L[0] = ds(c.P[0] / n, dm(ctr[0], ctr[0]));
  L[1] = ds(c.P[1] / n, dm(ctr[1], ctr[0]));
  L[2] = ds(c.P[2] / n, dm(ctr[2], ctr[0]));
  L[3] = ds(c.P[3] / n, dm(ctr[1], ctr[1]));
  L[4] = ds(c.P[4] / n, dm(ctr[2], ctr[1]));
  L[5] = ds(c.P[5] / n, dm(ctr[2], ctr[2]));
}
// PointCluster::push (types.hpp:137-142)
VN_HD void cluster_push(Cluster& c, const double* p)
{
  c.N++;
  c.P[0] = da(c.P[0], dm(p[0], p[0]));
  c.P[1] = da(c.P[1], dm(p[1], p[0]));
  c.P[2] = da(c.P[2], dm(p[2], p[0]));
  c.P[3] = da(c.P[3], dm(p[1], p[1]));
  c.P[4] = da(c.P[4], dm(p[2], p[1]));
  c.P[5] = da(c.P[5], dm(p[2], p[2]));
  c.v[0] = da(c.v[0], p[0]);
  c.v[1] = da(c.v[1], p[1]);
  c.v[2] = da(c.v[2], p[2]);
}
VN_HD void cluster_clear(Cluster& c)
{
  for (int i = 0; i < 6; i++) c.P[i] = 0;
  for (int i = 0; i < 3; i++) c.v[i] = 0;
  c.N = 0;
}
VN_HD void cluster_add(Cluster& a, const Cluster& b)
{
  for (int i = 0; i < 6; i++) a.P[i] = da(a.P[i], b.P[i]);
  for (int i = 0; i < 3; i++) a.v[i] = da(a.v[i], b.v[i]);
  a.N += b.N;
}
VN_HD void cluster_sub(Cluster& a, const Cluster& b)
{
  for (int i = 0; i < 6; i++) a.P[i] = ds(a.P[i], b.P[i]);
  for (int i = 0; i < 3; i++) a.v[i] = ds(a.v[i], b.v[i]);
  a.N -= b.N;
}
// PointCluster::transform (types.hpp:168-174), lower triangle of the result
VN_HD void cluster_transform(Cluster& out, const Cluster& sig, const double* R, const double* p)
{
  double n = (double)sig.N;
  double Rv[3];
  rot_vec(R, sig.v, Rv);
  double Np[3] = { dm(n, p[0]), dm(n, p[1]), dm(n, p[2]) };
  out.N = sig.N;
  for (int i = 0; i < 3; i++) out.v[i] = da(Rv[i], Np[i]);
  double Pf[9], T[9];
  sym6_to_full(sig.P, Pf);
  mat3_mul(R, Pf, T);
  const int li[6] = { 0, 1, 2, 1, 2, 2 }, lj[6] = { 0, 0, 0, 1, 1, 2 };
  for (int k = 0; k < 6; k++)
  {
    int i = li[k], j = lj[k];
    double rprt = da(da(dm(T[i], R[j]), dm(T[i + 3], R[j + 3])), dm(T[i + 6], R[j + 6]));
    double rp_ij = dm(Rv[i], p[j]);
    double rp_ji = dm(Rv[j], p[i]);
    out.P[k] = da(da(da(rprt, rp_ij), rp_ji), dm(Np[i], p[j]));
  }
}

// pvec_update covariance (point_utils.cpp:61-62), symmetric storage in/out:
// var_w = R var R^T + hat(p) rot_var hat(p)^T + tsl_var, upper triangle
VN_HD void world_var(const double* R, const double* pnt, const double* var6, const double* rot_var,
                     const double* tsl_var, double* out6)
{
  double V[9], A[9], B[9], ph[9], Cm[9], D[9];
  sym6_to_full(var6, V);
  mat3_mul(R, V, A);
  mat3_mulT(A, R, B);
  hat3(pnt, ph);
  mat3_mul(ph, rot_var, Cm);
  mat3_mulT(Cm, ph, D);
  const int ui[6] = { 0, 0, 0, 1, 1, 2 }, uj[6] = { 0, 1, 2, 1, 2, 2 };
  for (int k = 0; k < 6; k++)
  {
    int i = ui[k] + 3 * uj[k];
    out6[k] = da(da(B[i], D[i]), tsl_var[i]);
  }
}

// cov_add += Bf_var(pv, vec) (octree.cpp:83-92); cov packed upper 9x9, var symmetric
VN_HD void bf_var_add(double* cov45, const double* var6, const double* vec)
{
  double V[9];
  sym6_to_full(var6, V);
  // Bi rows: d(xx,xy,xz,yy,yz,zz)/dp
  double Bi[6][3] = { { 2 * vec[0], 0, 0 }, { vec[1], vec[0], 0 }, { vec[2], 0, vec[0] },
                      { 0, 2 * vec[1], 0 }, { 0, vec[2], vec[1] }, { 0, 0, 2 * vec[2] } };
  double Biup[6][3];
  for (int r = 0; r < 6; r++)
    for (int c = 0; c < 3; c++) Biup[r][c] = Bi[r][0] * V[0 + 3 * c] + Bi[r][1] * V[1 + 3 * c] + Bi[r][2] * V[2 + 3 * c];
  for (int r = 0; r < 6; r++)
  {
    for (int c = r; c < 6; c++)
      cov45[sN(9, r, c)] += Biup[r][0] * Bi[c][0] + Biup[r][1] * Bi[c][1] + Biup[r][2] * Bi[c][2];
    for (int c = 0; c < 3; c++) cov45[sN(9, r, 6 + c)] += Biup[r][c];
  }
  for (int r = 0; r < 3; r++)
    for (int c = r; c < 3; c++) cov45[sN(9, 6 + r, 6 + c)] += V[r + 3 * c];
}

// 15x15 inverse: LU with partial pivoting (what Eigen's Matrix<15,15>::inverse() does);
// host-side a7 (odometry.cpp:82, 194)
template <int N>
inline void inverse_lu(const double* A, double* inv)  // column-major
{
  double lu[N * N];
  int perm[N];
  for (int i = 0; i < N * N; i++) lu[i] = A[i];
  for (int i = 0; i < N; i++) perm[i] = i;
  auto at = [&](int i, int j) -> double& { return lu[i + j * N]; };
  for (int k = 0; k < N; k++)
  {
    int piv = k;
    double best = fabs(at(k, k));
    for (int i = k + 1; i < N; i++)
      if (fabs(at(i, k)) > best)
      {
        best = fabs(at(i, k));
        piv = i;
      }
    if (piv != k)
    {
      for (int j = 0; j < N; j++)
      {
        double t = at(k, j);
        at(k, j) = at(piv, j);
        at(piv, j) = t;
      }
      int t = perm[k];
      perm[k] = perm[piv];
      perm[piv] = t;
    }
    for (int i = k + 1; i < N; i++)
    {
      at(i, k) = at(i, k) / at(k, k);
      for (int j = k + 1; j < N; j++) at(i, j) = at(i, j) - at(i, k) * at(k, j);
    }
  }
  for (int c = 0; c < N; c++)
  {
    double y[N];
    for (int i = 0; i < N; i++)
    {
      double s = (perm[i] == c) ? 1.0 : 0.0;
      for (int j = 0; j < i; j++) s = s - at(i, j) * y[j];
      y[i] = s;
    }
    for (int i = N - 1; i >= 0; i--)
    {
      double s = y[i];
      for (int j = i + 1; j < N; j++) s = s - at(i, j) * inv[j + c * N];
      inv[i + c * N] = s / at(i, i);
    }
  }
}

// Host orchestration of the per-scan loop — the C++ that stays on the CPU
// (north_star: "host code stays C++ in src/pipeline"). It mirrors
//   IMUEKF::motion_blur, IMU part        src/estimation/imu_ekf.cpp:13-104   (a1, sequential, ~20-40 steps)
//   VINA_SLAM::LioStateEstimation        src/pipeline/odometry.cpp:64-255   (iteration loop + a7 15x15 solve)
//   thd_odometry_localmapping, scan body src/pipeline/local_mapping.cpp:389-546
// and drives the CUDA kernels only through the low-level C ABI of
// include/vina_b200.h — the same calls a reference-side adapter would make
// (INTEGRATION.md). No point-level work happens here.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <unordered_map>
#include <vector>
#include <chrono>
#include "../csrc/vn_ctx.h"
#include "vina_ba.h"

int vn_finish_downsample(vina_ctx* ctx);
int vn_iekf_launch(vina_ctx* ctx, const double R[9], const double p[3], bool debug);
extern "C" int vina_odom_iekf_host(vina_ctx* ctx, int which, int max_iter, int* iters_out, int* not_degenerate);
void vn_iekf_unpack(const double* r, double HTH[36], double HTz[6], double nnt[9], int32_t* match_num);

namespace
{
// column-major helpers -------------------------------------------------------
inline void m3_mul(const double* A, const double* B, double* C) { mat3_mul(A, B, C); }
inline void m3_vec(const double* A, const double* v, double* w) { rot_vec(A, v, w); }
inline void m3_T(const double* A, double* T)
{
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T[j + 3 * i] = A[i + 3 * j];
}
inline double norm3(const double* a) { return std::sqrt(dot3(a, a)); }

// Exp(ang_vel, dt) — include/vina_slam/core/math.hpp:26-41
void Exp_dt(const double* w, double dt, double* E)
{
  for (int i = 0; i < 9; i++) E[i] = 0;
  E[0] = E[4] = E[8] = 1;
  double nrm = norm3(w);
  if (nrm > 1e-7)
  {
    double ax[3] = { w[0] / nrm, w[1] / nrm, w[2] / nrm };
    double K[9], sK[9], KK[9];
    hat3(ax, K);
    double r = nrm * dt;
    double s = std::sin(r), c1 = 1.0 - std::cos(r);
    for (int i = 0; i < 9; i++) sK[i] = c1 * K[i];
    m3_mul(sK, K, KK);
    for (int i = 0; i < 9; i++) E[i] = (E[i] + s * K[i]) + KK[i];
  }
}
// Exp(ang) — math.hpp:12-24
void Exp_v(const double* a, double* E)
{
  for (int i = 0; i < 9; i++) E[i] = 0;
  E[0] = E[4] = E[8] = 1;
  double nrm = norm3(a);
  if (nrm >= 1e-9)
  {
    double ax[3] = { a[0] / nrm, a[1] / nrm, a[2] / nrm };
    double K[9], sK[9], KK[9];
    hat3(ax, K);
    double s = std::sin(nrm), c1 = 1.0 - std::cos(nrm);
    for (int i = 0; i < 9; i++) sK[i] = c1 * K[i];
    m3_mul(sK, K, KK);
    for (int i = 0; i < 9; i++) E[i] = (E[i] + s * K[i]) + KK[i];
  }
}
// Log(R) — math.hpp:43-48
void Log_R(const double* R, double* w)
{
  double tr = (R[0] + R[4]) + R[8];
  double theta = (tr > 3.0 - 1e-6) ? 0.0 : std::acos(0.5 * (tr - 1));
  double K[3] = { R[5] - R[7], R[6] - R[2], R[1] - R[3] };
  double f = (std::fabs(theta) < 0.001) ? 0.5 : (0.5 * theta / std::sin(theta));
  for (int i = 0; i < 3; i++) w[i] = f * K[i];
}

// dense column-major n x n helpers for the 15-dim state
void mat_mul(int n, int k, int m, const double* A, const double* B, double* C)  // (n x k)(k x m)
{
  for (int j = 0; j < m; j++)
    for (int i = 0; i < n; i++)
    {
      double s = A[i] * B[k * j];
      for (int t = 1; t < k; t++) s = s + A[i + n * t] * B[t + k * j];
      C[i + n * j] = s;
    }
}
}  // namespace

struct OdomHost
{
  vina_state x_curr;
  // IMUEKF members (include/vina_slam/ekf_imu.hpp:12-42)
  double pcl_beg_time = 0, pcl_end_time = 0, last_pcl_end_time = 0;
  vina_imu last_imu;
  double scale_gravity = 1.0;
  std::vector<vina_imu_pose> imu_poses;
  // sliding window (node.hpp: x_buf, win_count, win_base)
  std::vector<vina_pose> x_buf;
  int win_count = 0, win_base = 0;
  int degrade_cnt = 0;
  int last_iters = 0;
  // host IEKF driven one iteration at a time (vina_odom_iekf_host_begin / _update)
  vina_state hi_prop;
  double hi_cov_inv[225];
  int hi_iter = 0, hi_max = 0, hi_rematch = 0;
  bool hi_active = false;
  // sliding-window BA (local_mapping.cpp:437-441, 492-497, 541-546); vina_odom_set_ba
  bool if_BA = false;
  double imu_coef = 1e-4;               // LocalBA.imu_coef (node.cpp:247)
  std::vector<vina_state> xs_buf;       // x_buf with the full states (v, bg, ba, g): what BA optimises
  std::deque<ImuPre*> imu_pre_buf;      // imu_pre_buf; nullptr for frames that came without IMU data (bootstrap)
  std::deque<vina_imu> ba_imus;         // the scan's IMU batch, ends re-stamped to the scan boundaries (imu_ekf.cpp:95-104)
  bool ba_imus_valid = false;
  int ba_runs = 0, ba_last_iters = 0;
  // distance travelled / map pruning (local_mapping.cpp:262-263, 272, 317-341, 509-519)
  double jour = 0.0;
  double last_pos[3] = { 0, 0, 0 };
  bool release_flag = false;
  // start-up phase (VINA_SLAM::initialization, node.cpp:293-366; vina_odom_cold_start / vina_odom_init_scan)
  bool in_init = false;
  bool imu_init_flag = true;  // IMUEKF::init_flag: a context that is bootstrapped at known states never needs IMU_init
  int imu_init_num = 0;
  double mean_acc[3] = { 0, 0, 0 }, mean_gyr[3] = { 0, 0, 0 };
  std::vector<vina_state> init_xs;                 // x_buf with the full states
  std::vector<std::vector<float>> pl_origs;        // per frame: retained raw points (x, y, z, t), time-sorted
  std::vector<double> beg_times;
  std::vector<std::deque<vina_imu>> vec_imus;      // per frame: its IMU batch, ends re-stamped
  int init_rounds = 0;
  ~OdomHost()
  {
    for (ImuPre* f : imu_pre_buf) ba_imu_factor_delete(f);
  }
  OdomHost()
  {
    memset(&x_curr, 0, sizeof(x_curr));
    x_curr.R[0] = x_curr.R[4] = x_curr.R[8] = 1;
    x_curr.g[2] = -9.8;
    for (int i = 0; i < 15; i++) x_curr.cov[i + 15 * i] = i < 9 ? 1e-4 : 1e-5;  // IMUST::setZero, types.hpp:101-112
    memset(&last_imu, 0, sizeof(last_imu));
  }
};

void odom_host_destroy(OdomHost* o) { delete o; }

static OdomHost* odom(vina_ctx* ctx)
{
  if (!ctx->odom) ctx->odom = new OdomHost();
  return ctx->odom;
}

// a1 — IMUEKF::motion_blur, IMU forward propagation (imu_ekf.cpp:17-104)
static int imu_propagate(vina_ctx* ctx, OdomHost* o, const vina_imu* imus_in, int m)
{
  std::deque<vina_imu> imus(imus_in, imus_in + m);
  imus.push_front(o->last_imu);
  if (o->last_pcl_end_time - o->pcl_beg_time > 0.01)
    return vn_fail(ctx, VINA_E_TIME, "LiDAR time regress: beg %.6f < last end %.6f", o->pcl_beg_time,
                   o->last_pcl_end_time);
  vina_state& xc = o->x_curr;
  const vina_config& cfg = ctx->cfg;
  o->imu_poses.clear();
  double acc_imu[3] = { 0, 0, 0 }, angvel_avr[3] = { 0, 0, 0 }, acc_avr[3];
  double vel_imu[3], pos_imu[3], R_imu[9];
  memcpy(vel_imu, xc.v, 24);
  memcpy(pos_imu, xc.p, 24);
  memcpy(R_imu, xc.R, 72);
  // F is the identity plus five 3x3 blocks (imu_ekf.cpp:60-73): cov = F cov F^T + W is evaluated over F's non-zero
  // entries only - the skipped terms are products with an exact 0.0, and the kept ones are added in the dense
  // product's order (k ascending), so every non-zero result is the dense one bit for bit at a fifth of the work
  double F[225], W[225], T1[225], T2[225];
  double dt = 0;
  for (size_t it = 0; it + 1 < imus.size(); it++)
  {
    const vina_imu& head = imus[it];
    const vina_imu& tail = imus[it + 1];
    if (head.t < o->last_pcl_end_time) continue;
    for (int k = 0; k < 3; k++)
    {
      angvel_avr[k] = 0.5 * (head.gyr[k] + tail.gyr[k]);
      acc_avr[k] = 0.5 * (head.acc[k] + tail.acc[k]);
    }
    for (int k = 0; k < 3; k++)
    {
      angvel_avr[k] = angvel_avr[k] - xc.bg[k];
      acc_avr[k] = acc_avr[k] * o->scale_gravity - xc.ba[k];
    }
    rot_trans(R_imu, xc.g, acc_avr, acc_imu);
    double cur_time = head.t;
    if (cur_time < o->last_pcl_end_time) cur_time = o->last_pcl_end_time;
    dt = tail.t - cur_time;
    double offt = cur_time - o->pcl_beg_time;

    vina_imu_pose ps;
    ps.t = offt;
    memcpy(ps.R, R_imu, 72);
    memcpy(ps.p, pos_imu, 24);
    memcpy(ps.v, vel_imu, 24);
    memcpy(ps.w, angvel_avr, 24);
    memcpy(ps.a, acc_imu, 24);
    o->imu_poses.push_back(ps);

    double acc_skew[9], Exp_f[9], Exp_b[9];
    hat3(acc_avr, acc_skew);
    Exp_dt(angvel_avr, dt, Exp_f);
    Exp_dt(angvel_avr, -dt, Exp_b);
    std::fill(F, F + 225, 0.0);
    std::fill(W, W + 225, 0.0);
    for (int i = 0; i < 15; i++) F[i + 15 * i] = 1.0;
    auto setblk = [&](double* M, int r0, int c0, const double* B) {
      for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) M[(r0 + i) + 15 * (c0 + j)] = B[i + 3 * j];
    };
    double B[9], nR[9], RA[9];
    setblk(F, 0, 0, Exp_b);
    for (int i = 0; i < 9; i++) B[i] = 0;
    B[0] = B[4] = B[8] = -1.0 * dt;
    setblk(F, 0, 9, B);
    B[0] = B[4] = B[8] = 1.0 * dt;
    setblk(F, 3, 6, B);
    for (int i = 0; i < 9; i++) nR[i] = -1.0 * R_imu[i];
    m3_mul(nR, acc_skew, RA);
    for (int i = 0; i < 9; i++) B[i] = RA[i] * dt;
    setblk(F, 6, 0, B);
    for (int i = 0; i < 9; i++) B[i] = nR[i] * dt;
    setblk(F, 6, 12, B);
    for (int k = 0; k < 3; k++) W[k + 15 * k] = cfg.cov_gyr * dt * dt;
    {
      double D[9] = { cfg.cov_acc, 0, 0, 0, cfg.cov_acc, 0, 0, 0, cfg.cov_acc }, RD[9], RDRt[9];
      m3_mul(R_imu, D, RD);
      mat3_mulT(RD, R_imu, RDRt);
      for (int i = 0; i < 9; i++) B[i] = RDRt[i] * dt * dt;
      setblk(W, 6, 6, B);
    }
    for (int k = 0; k < 3; k++) W[(9 + k) + 15 * (9 + k)] = cfg.rdw_gyr * dt * dt;
    for (int k = 0; k < 3; k++) W[(12 + k) + 15 * (12 + k)] = cfg.rdw_acc * dt * dt;
    // cov = F cov F^T + W
    {
      int nz[15][15], nnz[15];
      for (int i = 0; i < 15; i++)
      {
        nnz[i] = 0;
        for (int k = 0; k < 15; k++)
          if (F[i + 15 * k] != 0.0) nz[i][nnz[i]++] = k;
      }
      const double* P = xc.cov;
      for (int j = 0; j < 15; j++)
        for (int i = 0; i < 15; i++)
        {
          double s = F[i + 15 * nz[i][0]] * P[nz[i][0] + 15 * j];  // (every row holds its diagonal 1 or a rotation)
          for (int a = 1; a < nnz[i]; a++) s = s + F[i + 15 * nz[i][a]] * P[nz[i][a] + 15 * j];
          T1[i + 15 * j] = s;
        }
      for (int j = 0; j < 15; j++)  // (T1 F^T)(i, j) = sum_k T1(i, k) F(j, k)
        for (int i = 0; i < 15; i++)
        {
          double s = T1[i + 15 * nz[j][0]] * F[j + 15 * nz[j][0]];
          for (int a = 1; a < nnz[j]; a++) s = s + T1[i + 15 * nz[j][a]] * F[j + 15 * nz[j][a]];
          T2[i + 15 * j] = s;
        }
    }
    for (int i = 0; i < 225; i++) xc.cov[i] = T2[i] + W[i];

    for (int k = 0; k < 3; k++) pos_imu[k] = (pos_imu[k] + vel_imu[k] * dt) + ((0.5 * acc_imu[k]) * dt) * dt;
    for (int k = 0; k < 3; k++) vel_imu[k] = vel_imu[k] + acc_imu[k] * dt;
    double Rn[9];
    m3_mul(R_imu, Exp_f, Rn);
    memcpy(R_imu, Rn, 72);
  }
  double imu_end_time = imus.back().t;
  double note = o->pcl_end_time > imu_end_time ? 1.0 : -1.0;
  dt = note * (o->pcl_end_time - imu_end_time);
  double nw[3] = { note * angvel_avr[0], note * angvel_avr[1], note * angvel_avr[2] }, E[9], Rn[9];
  for (int k = 0; k < 3; k++) xc.v[k] = vel_imu[k] + (note * acc_imu[k]) * dt;
  Exp_dt(nw, dt, E);
  m3_mul(R_imu, E, Rn);
  memcpy(xc.R, Rn, 72);
  for (int k = 0; k < 3; k++)
    xc.p[k] = (pos_imu[k] + (note * vel_imu[k]) * dt) + (((note * 0.5) * acc_imu[k]) * dt) * dt;
  xc.t = o->pcl_end_time;
  if (o->if_BA || o->in_init)
  {
    // imu_ekf.cpp:95-104: the batch handed to the pre-integration, first / last sample re-stamped to the scan
    // boundaries (integer nanoseconds, rclcpp::Time)
    o->ba_imus.assign(imus.begin(), imus.end());
    o->ba_imus.front().t = (double)static_cast<int64_t>(o->last_pcl_end_time * 1e9) * 1e-9;
    o->ba_imus.back().t = (double)static_cast<int64_t>(o->pcl_end_time * 1e9) * 1e-9;
    o->ba_imus_valid = true;
  }
  o->last_imu = imus.back();
  o->last_pcl_end_time = o->pcl_end_time;
  return VINA_OK;
}

// x (+)= delta — IMUST::operator+= (types.hpp:67-75)
static void state_boxplus(vina_state& x, const double* d)
{
  double E[9], Rn[9];
  Exp_v(d, E);
  m3_mul(x.R, E, Rn);
  memcpy(x.R, Rn, 72);
  for (int k = 0; k < 3; k++)
  {
    x.p[k] += d[3 + k];
    x.v[k] += d[6 + k];
    x.bg[k] += d[9 + k];
    x.ba[k] += d[12 + k];
  }
}
// a (-) b — IMUST::operator- (types.hpp:77-86)
static void state_boxminus(const vina_state& a, const vina_state& b, double* out)
{
  double bt[9], M[9];
  m3_T(b.R, bt);
  m3_mul(bt, a.R, M);
  Log_R(M, out);
  for (int k = 0; k < 3; k++)
  {
    out[3 + k] = a.p[k] - b.p[k];
    out[6 + k] = a.v[k] - b.v[k];
    out[9 + k] = a.bg[k] - b.bg[k];
    out[12 + k] = a.ba[k] - b.ba[k];
  }
}

// VINA_SLAM::LioStateEstimation (odometry.cpp:64-255), use_vnc == false terms - host variant: the per-point
// loop of each iteration runs on the GPU (k_iekf publishing its 34 sums to mapped memory), the 15x15 update
// (a7) here. Kept for vina_odom_iekf(..., host_solve) / cross-checks; the per-scan path uses the device loop.
static int lio_state_estimation_host(vina_ctx* ctx, OdomHost* o, int which, int max_iter_override, int* not_degenerate)
{
  vina_state& x_curr = o->x_curr;
  const vina_state x_prop = x_curr;
  const int num_max_iter = max_iter_override > 0 ? max_iter_override : 20;
  double G[225], HTH15[225], cov_inv[225], K1[225], tmp[225];
  memset(G, 0, sizeof(G));
  memset(HTH15, 0, sizeof(HTH15));
  int rematch_num = 0;
  double nnt[9] = { 0 };
  inverse_lu<15>(x_curr.cov, cov_inv);
  double rot_var[9], tsl_var[9];
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++)
    {
      rot_var[i + 3 * j] = x_curr.cov[i + 15 * j];
      tsl_var[i + 3 * j] = x_curr.cov[(3 + i) + 15 * (3 + j)];
    }
  int r = vina_iekf_begin(ctx, which, rot_var, tsl_var);
  if (r) return r;
  o->last_iters = 0;

  for (int iterCount = 0; iterCount < num_max_iter; iterCount++)
  {
    double HTH[36], HTz[6];
    int32_t match_num = 0;
    r = vn_iekf_launch(ctx, x_curr.R, x_curr.p, false);
    if (r) return r;
    r = vn_iekf_wait(ctx);
    if (r) return r;
    vn_iekf_unpack(ctx->h_result, HTH, HTz, nnt, &match_num);
    o->last_iters = iterCount + 1;

    for (int j = 0; j < 6; j++)
      for (int i = 0; i < 6; i++) HTH15[i + 15 * j] = HTH[i + 6 * j];
    for (int i = 0; i < 225; i++) tmp[i] = HTH15[i] + cov_inv[i];
    inverse_lu<15>(tmp, K1);
    // G(:,0:6) = K1(:,0:6) * HTH
    double G6[90];
    mat_mul(15, 6, 6, K1, HTH, G6);  // first 6 columns of K1 are the first 90 entries
    memcpy(G, G6, sizeof(G6));
    double vec[15], sol[15], a[15], b[15];
    state_boxminus(x_prop, x_curr, vec);
    mat_mul(15, 6, 1, K1, HTz, a);
    mat_mul(15, 6, 1, G6, vec, b);
    for (int i = 0; i < 15; i++) sol[i] = (a[i] + vec[i]) - b[i];
    state_boxplus(x_curr, sol);

    bool conv = (norm3(sol) * 57.3 < 0.01) && (norm3(sol + 3) * 100 < 0.015);
    if (conv || ((rematch_num == 0) && (iterCount == num_max_iter - 2))) rematch_num++;
    if (rematch_num >= 2 || (iterCount == num_max_iter - 1))
    {
      // cov = (I - G) cov
      double IG[225], nc[225];
      for (int i = 0; i < 225; i++) IG[i] = -G[i];
      for (int i = 0; i < 15; i++) IG[i + 15 * i] = 1.0 - G[i + 15 * i];
      mat_mul(15, 15, 15, IG, x_curr.cov, nc);
      memcpy(x_curr.cov, nc, sizeof(nc));
      break;
    }
  }
  ctx->tm.iekf_kernel_ms = 0;
  ctx->tm.iekf_iters = o->last_iters;
  // degeneracy test on the last iteration's nnt (odometry.cpp:244-254)
  double L[6] = { nnt[0], nnt[1], nnt[2], nnt[4], nnt[5], nnt[8] }, ev[3], Q[9];
  eig3_sym(L, ev, Q);
  *not_degenerate = !(ev[0] < 14);
  return VINA_OK;
}

// stage x_curr as the device iterate (x_prop = x_curr, prior covariance and its blocks, loop counters)
static void stage_iterate(const vina_state& x, int max_iter, IekfDev* h)
{
  memcpy(h->R, x.R, 72);
  memcpy(h->p, x.p, 24);
  memcpy(h->v, x.v, 24);
  memcpy(h->bg, x.bg, 24);
  memcpy(h->ba, x.ba, 24);
  memcpy(h->Rp, x.R, 72);
  memcpy(h->pp, x.p, 24);
  memcpy(h->vp, x.v, 24);
  memcpy(h->bgp, x.bg, 24);
  memcpy(h->bap, x.ba, 24);
  memcpy(h->cov, x.cov, sizeof(x.cov));
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++)
    {
      h->rot_var[i + 3 * j] = x.cov[i + 15 * j];
      h->tsl_var[i + 3 * j] = x.cov[(3 + i) + 15 * (3 + j)];
    }
  h->iter = 0;
  h->rematch = 0;
  h->done = 0;
  h->max_iter = max_iter;
}

// take the converged iterate back (after the stream has been synchronised)
static void unstage_iterate(const IekfDev* h, vina_state& x, int* iters, int* not_degenerate)
{
  memcpy(x.R, h->R, 72);
  memcpy(x.p, h->p, 24);
  memcpy(x.v, h->v, 24);
  memcpy(x.bg, h->bg, 24);
  memcpy(x.ba, h->ba, 24);
  memcpy(x.cov, h->cov, sizeof(x.cov));
  *iters = h->iter;
  // degeneracy test on the last iteration's nnt (odometry.cpp:244-254)
  const double* s = h->sums + 27;
  double L[6] = { s[0], s[1], s[2], s[3], s[4], s[5] }, ev[3], Q[9];
  eig3_sym(L, ev, Q);
  *not_degenerate = !(ev[0] < 14);
}

// upload the iterate and reset the per-point leaf cache (ctx stream)
// the iterate (x_curr after the IMU propagation, prior covariance, counters) to the device. The overlapped step does
// this BEFORE the deskew kernel: the copy does not depend on it and no longer sits between the deskew and the first
// IEKF launch
static int iekf_upload_iterate(vina_ctx* ctx, OdomHost* o, int num_max_iter)
{
  stage_iterate(o->x_curr, num_max_iter, ctx->h_iekf);
  int r = vn_check_cuda(ctx, cudaMemcpyAsync(ctx->d_iekf, ctx->h_iekf, sizeof(IekfDev), cudaMemcpyHostToDevice, ctx->stream),
                        "iterate upload");
  if (!r) ctx->iterate_uploaded = true;
  return r;
}

static int iekf_stage(vina_ctx* ctx, OdomHost* o, int which, int num_max_iter)
{
  int r = ctx->iterate_uploaded ? VINA_OK : iekf_upload_iterate(ctx, o, num_max_iter);
  ctx->iterate_uploaded = false;
  if (r) return r;
  ctx->iekf_which = which;
  const int n = ctx->n_pv[which];
  if (!(ctx->cache_is_reset && which == 0))
  {
    launch_fill_int(ctx->stream, ctx->d_cache, -1, n);  // vector<OctoTree*> octos(psize, nullptr), odometry.cpp:79
    ctx->launches += 1;
  }
  ctx->cache_is_reset = false;
  ctx->iekf_blocks = iekf_grid_blocks(n, ctx->sm_count);
  ctx->dbg_valid = false;
  return VINA_OK;
}

// enqueue the whole iteration loop on the ctx stream: upload of the iterate, cache reset, max_iter x k_iekf
// (each launch accumulates, and its last block solves and updates the device iterate; launches after
// convergence return at once), download of the result. No host synchronisation inside.
static int iekf_enqueue_device(vina_ctx* ctx, OdomHost* o, int which, int num_max_iter)
{
  int r = iekf_stage(ctx, o, which, num_max_iter);
  if (r) return r;
  IekfBatch bt;
  bt.mode = VN_IEKF_SOLVE | VN_IEKF_HANDOVER;  // the launch that ends the loop publishes the iterate itself
  bt.variant = 0;
  ++ctx->pub_seq;
  vn_iekf_fill_seq(ctx, &bt.s[0], false);
  if (ctx->profiling && (int)ctx->iekf_ev.size() < 2 * num_max_iter)
  {
    size_t old = ctx->iekf_ev.size();
    ctx->iekf_ev.resize(2 * (size_t)num_max_iter);
    for (size_t i = old; i < ctx->iekf_ev.size(); i++) cudaEventCreate(&ctx->iekf_ev[i]);
  }
  // the default: the whole loop as one persistent launch (needs 16-byte aligned pointVar rows for the bulk copies)
  if (ctx->iekf_loop && (ctx->cap_points & 1) == 0 && (reinterpret_cast<uintptr_t>(bt.s[0].pv_base) & 15) == 0)
  {
    IekfLoop lp;
    lp.q = bt.s[0];
    lp.bar = ctx->d_loop_bar;
    lp.partials = ctx->d_loop_partials;
    lp.status = ctx->d_status;
    lp.mode = VN_IEKF_HANDOVER;
    const int n = ctx->n_pv[which];
    int blocks = (n + 255) / 256;
    if (blocks > ctx->sm_count) blocks = ctx->sm_count;
    if (blocks < 1) blocks = 1;
    lp.chunk = iekf_loop_chunk(n, blocks);
    if (ctx->profiling) cudaEventRecord(ctx->iekf_ev[0], ctx->stream);
    int e = launch_iekf_loop(ctx->stream, lp, blocks);
    if (e) return vn_check_cuda(ctx, (cudaError_t)e, "k_iekf_loop launch");
    ctx->launches += 1;
    if (ctx->profiling) cudaEventRecord(ctx->iekf_ev[1], ctx->stream);
    ctx->iekf_looped = true;
    return VINA_OK;
  }
  ctx->iekf_looped = false;
  for (int it = 0; it < num_max_iter; it++)
  {
    if (ctx->profiling) cudaEventRecord(ctx->iekf_ev[2 * it], ctx->stream);
    int e = launch_iekf(ctx->stream, bt, 1, ctx->iekf_blocks, false, it >= 2 && !ctx->profiling);
    if (e) return vn_check_cuda(ctx, (cudaError_t)e, "k_iekf launch");
    ctx->launches += 1;
    if (ctx->profiling) cudaEventRecord(ctx->iekf_ev[2 * it + 1], ctx->stream);
  }
  return VINA_OK;
}

// VINA_SLAM::LioStateEstimation (odometry.cpp:64-255) with the whole iteration loop on the device: one
// host synchronisation per call, after the last iteration.
static int lio_state_estimation(vina_ctx* ctx, OdomHost* o, int which, int max_iter_override, int* not_degenerate)
{
  const int num_max_iter = max_iter_override > 0 ? max_iter_override : 20;
  int r = iekf_enqueue_device(ctx, o, which, num_max_iter);
  if (r) return r;
  r = vn_iterate_wait(ctx, ctx->stream);
  if (r) return r;
  if (ctx->n_down_pending && !ctx->n_down_mapped)
  {
    // the count's device-to-host copy was enqueued before the loop: it has landed (stream order)
    ctx->n_down = *ctx->h_n_down;
    ctx->n_down_pending = false;
  }
  unstage_iterate(ctx->h_pub, o->x_curr, &o->last_iters, not_degenerate);
  ctx->tm.iekf_iters = o->last_iters;
  float kernel_ms = 0;
  if (ctx->profiling && ctx->iekf_looped)
  {
    cudaEventSynchronize(ctx->iekf_ev[1]);
    cudaEventElapsedTime(&kernel_ms, ctx->iekf_ev[0], ctx->iekf_ev[1]);  // one launch = the whole loop
  }
  else if (ctx->profiling)
    cudaEventSynchronize(ctx->iekf_ev[2 * num_max_iter - 1]);
  if (ctx->profiling && !ctx->iekf_looped)
    for (int it = 0; it < o->last_iters; it++)
    {
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->iekf_ev[2 * it], ctx->iekf_ev[2 * it + 1]);
      kernel_ms += ms;
    }
  ctx->tm.iekf_kernel_ms = kernel_ms;
  return VINA_OK;
}

// local_mapping.cpp:425-451 and 489-546
// local_mapping.cpp:509-519, evaluated right after multi_margi with the window counters as they are there
static void journey_update(OdomHost* o)
{
  if ((o->win_base + o->win_count) % 10 != 0) return;
  const double d[3] = { o->x_curr.p[0] - o->last_pos[0], o->x_curr.p[1] - o->last_pos[1], o->x_curr.p[2] - o->last_pos[2] };
  const double spat = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (spat > 0.5)
  {
    o->jour += spat;
    memcpy(o->last_pos, o->x_curr.p, 24);
    o->release_flag = true;
  }
}

static int map_update(vina_ctx* ctx, OdomHost* o)
{
  vina_state& x = o->x_curr;
  double rot_var[9], tsl_var[9];
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++)
    {
      rot_var[i + 3 * j] = x.cov[i + 15 * j];
      tsl_var[i + 3 * j] = x.cov[(3 + i) + 15 * (3 + j)];
    }
  o->win_count++;
  vina_pose ps;
  memcpy(ps.R, x.R, 72);
  memcpy(ps.p, x.p, 24);
  o->x_buf.push_back(ps);
  if (o->if_BA)
  {
    o->xs_buf.push_back(x);
    if (o->win_count > 1)
    {
      // imu_pre_buf.push_back(new IMU_PRE(x_buf[win_count - 2].bg, .ba)); ->push_imu(imus) (local_mapping.cpp:437-441)
      const vina_state& prev = o->xs_buf[o->win_count - 2];
      o->imu_pre_buf.push_back(o->ba_imus_valid ? ba_imu_factor_new(prev.bg, prev.ba, o->ba_imus, o->scale_gravity, ctx->cfg)
                                                 : nullptr);
    }
    o->ba_imus_valid = false;
  }
  cudaEvent_t* ev = ctx->ev;
  if (ctx->profiling) cudaEventRecord(ev[4], ctx->stream);
  int r = vina_map_insert(ctx, o->win_count - 1, x.R, x.p, rot_var, tsl_var);  // pvec_update + cut_voxel_multi
  if (r) return r;
  if (ctx->profiling) cudaEventRecord(ev[5], ctx->stream);
  r = vina_map_recut(ctx, o->win_count, o->x_buf.data());
  if (r) return r;
  const bool full = o->win_count >= ctx->cfg.win_size;
  bool run_ba = o->if_BA && full && (int)o->imu_pre_buf.size() == o->win_count - 1;
  if (run_ba)
    for (ImuPre* f : o->imu_pre_buf) run_ba = run_ba && f != nullptr;
  if (full && (ctx->ba_capture || run_ba))
  {
    r = vn_ba_collect_enqueue(ctx);  // tras_opt: the factors damping_iter consumes (local_mapping.cpp:196-200)
    if (r) return r;
  }
  if (ctx->profiling) cudaEventRecord(ev[6], ctx->stream);
  if (run_ba)
  {
    // LI_BA_Optimizer::damping_iter (local_mapping.cpp:492-497): LM on the host around the device LiDAR factor
    r = ba_damping_iter(ctx, o->xs_buf, o->imu_pre_buf, o->imu_coef, &o->ba_last_iters);
    if (r) return r;
    o->ba_runs++;
    for (int i = 0; i < o->win_count; i++)
    {
      memcpy(o->x_buf[i].R, o->xs_buf[i].R, 72);
      memcpy(o->x_buf[i].p, o->xs_buf[i].p, 24);
    }
    r = vn_ba_writeback_enqueue(ctx);  // margi takes the re-evaluated pcr_add / eig back (octree.cpp:410-416)
    if (r) return r;
  }
  if (full)
  {
    // x_curr.R/p = x_buf.back() (local_mapping.cpp:501-502): the identity without BA
    memcpy(x.R, o->x_buf.back().R, 72);
    memcpy(x.p, o->x_buf.back().p, 24);
    ctx->map.jour = o->jour;  // multi_margi(surf_map_slide, jour, ...) (local_mapping.cpp:507)
    r = vina_map_margi(ctx, o->win_count, o->x_buf.data());
    if (r) return r;
    journey_update(o);
    r = vina_map_shift_window(ctx);
    if (r) return r;
    o->x_buf.erase(o->x_buf.begin());
    if (o->if_BA)
    {
      o->xs_buf.erase(o->xs_buf.begin());
      ba_imu_factor_delete(o->imu_pre_buf.front());
      o->imu_pre_buf.pop_front();
    }
    o->win_base += 1;
    o->win_count -= 1;
  }
  if (ctx->profiling) cudaEventRecord(ev[7], ctx->stream);
  return VINA_OK;
}

static void collect_timings(vina_ctx* ctx)
{
  if (!ctx->profiling) return;
  cudaEventSynchronize(ctx->ev[7]);
  cudaEvent_t* ev = ctx->ev;
  cudaEventElapsedTime(&ctx->tm.deskew_ms, ev[0], ev[1]);
  cudaEventElapsedTime(&ctx->tm.downsample_ms, ev[1], ev[2]);
  cudaEventElapsedTime(&ctx->tm.var_init_ms, ev[2], ev[3]);
  cudaEventElapsedTime(&ctx->tm.iekf_ms, ev[3], ev[4]);
  cudaEventElapsedTime(&ctx->tm.insert_ms, ev[4], ev[5]);
  cudaEventElapsedTime(&ctx->tm.recut_ms, ev[5], ev[6]);
  cudaEventElapsedTime(&ctx->tm.margi_ms, ev[6], ev[7]);
  cudaEventElapsedTime(&ctx->tm.total_ms, ev[0], ev[7]);
}

// the scan body of thd_odometry_localmapping (local_mapping.cpp:389-546), first part: a1 IMU propagation on
// the host, then deskew, down-sampling and var_init of the set the IEKF runs on - all enqueued, no sync
static int step_front(vina_ctx* ctx, OdomHost* o, double pcl_beg_time, double pcl_end_time, const vina_imu* imus, int m,
                      int iekf_on_full, int* which_out)
{
  o->pcl_beg_time = pcl_beg_time;
  o->pcl_end_time = pcl_end_time;
  int r = imu_propagate(ctx, o, imus, m);
  if (r) return r;
  cudaEvent_t* ev = ctx->ev;
  if (ctx->profiling) cudaEventRecord(ev[0], ctx->stream);
  ctx->front_was_fused = false;
  if (iekf_on_full && ctx->front_fused && ctx->iekf_loop && !ctx->batch_member && ctx->cfg.down_size >= 0.001 && ctx->n_scan > 0)
  {
    // the product schedule's two front launches (see odom_step_overlapped); the per-stage timers see the first as
    // "deskew" and the second as "downsample"
    r = vn_front_fused(ctx, o->imu_poses.data(), (int)o->imu_poses.size(), o->x_curr.R, o->x_curr.p);
    if (r) return r;
    if (ctx->profiling)
    {
      cudaEventRecord(ev[1], ctx->stream);  // (both launches are enqueued: the split is not resolved here)
      cudaEventRecord(ev[2], ctx->stream);
      cudaEventRecord(ev[3], ctx->stream);
    }
    ctx->front_was_fused = true;
    *which_out = 0;
    return VINA_OK;
  }
  r = vina_deskew(ctx, o->imu_poses.data(), (int)o->imu_poses.size(), o->x_curr.R, o->x_curr.p);
  if (r) return r;
  if (ctx->profiling) cudaEventRecord(ev[1], ctx->stream);
  r = vina_downsample(ctx);
  if (r) return r;
  if (ctx->profiling) cudaEventRecord(ev[2], ctx->stream);
  int which = 1;
  if (iekf_on_full)
  {
    r = vina_var_init(ctx, 0);
    if (r) return r;
    which = 0;
  }
  else
  {
    r = vn_finish_downsample(ctx);
    if (r) return r;
    r = vina_var_init(ctx, 1);
    if (r) return r;
  }
  if (ctx->profiling) cudaEventRecord(ev[3], ctx->stream);
  *which_out = which;
  return VINA_OK;
}

// ... last part, after the IEKF: degeneracy bookkeeping, var_init of the down-sampled set, map update
static int step_back(vina_ctx* ctx, OdomHost* o, int iekf_on_full, int ok, vina_state* x_out)
{
  if (ok)
  {
    if (o->degrade_cnt > 0) o->degrade_cnt--;
  }
  else
    o->degrade_cnt++;
  int r;
  if (iekf_on_full)
  {
    // the down-sampled count arrived with the IEKF readback: no extra sync in the common case
    r = vn_finish_downsample(ctx);
    if (r) return r;
    if (ctx->front_was_fused && !ctx->down_retried)
      ctx->n_pv[1] = ctx->n_down;  // k_down_emit_all has done the var_init of the emitted set
    else
    {
      r = vina_var_init(ctx, 1);
      if (r) return r;
    }
  }
  r = map_update(ctx, o);
  if (r) return r;
  if (x_out) *x_out = o->x_curr;
  return VINA_OK;
}

// The same scan body with two things taken off the critical path (vina_set_overlap, the default):
//  * down-sampling and the var_init of the map's point set do not feed the IEKF (VNC_lio runs on the full scan,
//    local_mapping.cpp:406-413): they run on a side stream, concurrently with the IEKF loop;
//  * the map update is enqueued right behind the loop, before its result has reached the host: the kernels read
//    the new pose and the posterior covariance blocks from the device iterate (LivePose, map_kernels.cu). The
//    host then picks the iterate up (mapped memory, no stream sync) while the map update is already running.
static int odom_step_overlapped(vina_ctx* ctx, OdomHost* o, double pcl_beg_time, double pcl_end_time,
                                const vina_imu* imus, int m, int max_iter, vina_state* x_out)
{
  const int l0 = ctx->launches;
  const bool tr = ctx->trace;
  double th[8] = { 0 };
  auto now_us = []() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
  };
  if (tr) th[0] = now_us();
  o->pcl_beg_time = pcl_beg_time;
  o->pcl_end_time = pcl_end_time;
  int r = imu_propagate(ctx, o, imus, m);
  if (r) return r;
  cudaStream_t A = ctx->stream, B = ctx->side_stream;
  if (tr) th[1] = now_us(), cudaEventRecord(ctx->tr_ev[0], A);
  const int num_max_iter = max_iter > 0 ? max_iter : 20;
  const bool fused = ctx->front_fused && ctx->iekf_loop && ctx->cfg.down_size >= 0.001 && ctx->n_scan > 0;
  if (fused)
  {
    // The persistent loop kernel fills every SM, so a side stream has nothing to run on next to it: the front is two
    // launches in stream order instead - deskew + var_init + cache reset + the accumulation pass of the down-sampling,
    // then the rest of the down-sampling with the var_init of the map's point set - and the loop follows. The
    // down-sampled count reaches the host through mapped memory while the loop runs.
    r = vn_settle_upload(ctx);
    if (r) return r;
    r = vn_front_fused(ctx, o->imu_poses.data(), (int)o->imu_poses.size(), o->x_curr.R, o->x_curr.p);
    if (r) return r;
    r = iekf_enqueue_device(ctx, o, 0, num_max_iter);
    if (r) return r;
    if (tr) th[2] = now_us(), cudaEventRecord(ctx->tr_ev[1], A);
    r = vn_finish_downsample(ctx);
    if (r) return r;
    if (ctx->down_retried)
    {
      r = vina_var_init(ctx, 1);  // (the "< 2000 points" retry re-ran the down-sampling with the separate kernels)
      if (r) return r;
    }
    else
      ctx->n_pv[1] = ctx->n_down;
    if (tr) th[3] = now_us();
    if (tr) cudaEventRecord(ctx->tr_ev[2], A);
  }
  else
  {
    // the default schedule. Deskew, var_init of the full scan and the leaf-cache reset are one kernel (the iterate goes
    // up first)
    r = iekf_upload_iterate(ctx, o, num_max_iter);
    if (r) return r;
    r = vn_deskew_var_init(ctx, o->imu_poses.data(), (int)o->imu_poses.size(), o->x_curr.R, o->x_curr.p);
    if (r) return r;
    r = vn_check_cuda(ctx, cudaEventRecord(ctx->ev_fork, A), "fork");
    if (r) return r;
    // the IEKF launches go out first: they are on the critical path, the side stream has slack
    r = iekf_enqueue_device(ctx, o, 0, num_max_iter);
    if (r) return r;
    if (tr) th[2] = now_us(), cudaEventRecord(ctx->tr_ev[1], A);
    cudaStreamWaitEvent(B, ctx->ev_fork, 0);
    // the map's point set on the side stream: down-sampling with the var_init of the emitted points in its last kernel.
    // The host needs the count (and the "< 2000 points" retry, local_mapping.cpp:396-403): it arrives through mapped
    // memory while the IEKF keeps running
    ctx->stream = B;
    ctx->down_fuse_var_init = ctx->cfg.down_size >= 0.001;
    r = vina_downsample(ctx);
    if (!r) r = vn_finish_downsample(ctx);
    if (!r)
    {
      if (ctx->down_fuse_var_init)
        ctx->n_pv[1] = ctx->n_down;
      else
        r = vina_var_init(ctx, 1);
    }
    ctx->down_fuse_var_init = false;
    ctx->stream = A;
    if (r) return r;
    if (tr) th[3] = now_us();
    cudaEventRecord(ctx->ev_join, B);
    cudaStreamWaitEvent(A, ctx->ev_join, 0);
    if (tr) cudaEventRecord(ctx->tr_ev[2], A);
    r = vn_mark_scan_read(ctx);  // (covers the side stream's readers of the scan buffer as well)
    if (r) return r;
  }
  if (o->if_BA)
  {
    // the LM loop of the BA needs the host between recut and margi: take the IEKF result first, then the map
    // update with host poses (the front of the step has still run fused / on the side stream)
    r = vn_iterate_wait(ctx, ctx->stream);
    if (r) return r;
    int ok = 0;
    unstage_iterate(ctx->h_pub, o->x_curr, &o->last_iters, &ok);
    ctx->tm.iekf_iters = o->last_iters;
    ctx->tm.iekf_kernel_ms = 0;
    if (ok)
    {
      if (o->degrade_cnt > 0) o->degrade_cnt--;
    }
    else
      o->degrade_cnt++;
    r = map_update(ctx, o);
    if (r) return r;
    if (x_out) *x_out = o->x_curr;
    ctx->tm.kernel_launches = ctx->launches - l0;
    return VINA_OK;
  }
  // local_mapping.cpp:425-451, 489-546 with if_BA == 0, pose of the new frame from the device
  o->win_count++;
  vina_pose ps;
  memset(&ps, 0, sizeof(ps));
  o->x_buf.push_back(ps);
  r = vn_map_insert_live(ctx, o->win_count - 1);
  if (r) return r;
  if (tr) cudaEventRecord(ctx->tr_ev[3], A);
  const bool margi = o->win_count >= ctx->cfg.win_size;
  if (margi && !ctx->ba_capture && ctx->split_overlap && ctx->side_stream)
  {
    // multi_recut and multi_margi as one enqueue: the subdivisions run on the side stream next to the
    // marginalisation of the leaves they do not touch
    ctx->map.jour = o->jour;
    r = vn_map_recut_margi_live(ctx, o->win_count, o->x_buf.data());
    if (r) return r;
    if (tr) cudaEventRecord(ctx->tr_ev[4], A);
    r = vina_map_shift_window(ctx);
    if (r) return r;
  }
  else
  {
    r = vn_map_recut_live(ctx, o->win_count, o->x_buf.data());
    if (r) return r;
    if (tr) cudaEventRecord(ctx->tr_ev[4], A);
    if (margi && ctx->ba_capture)
    {
      r = vn_ba_collect_enqueue(ctx);
      if (r) return r;
    }
    if (margi)
    {
      ctx->map.jour = o->jour;  // (the journey only advances after multi_margi: known before the pose is)
      r = vn_map_margi_live(ctx, o->win_count, o->x_buf.data());
      if (r) return r;
      r = vina_map_shift_window(ctx);
      if (r) return r;
    }
  }
  if (tr) th[4] = now_us(), cudaEventRecord(ctx->tr_ev[5], A);
  // the result of the loop (it landed while the map update was being enqueued)
  r = vn_iterate_wait(ctx, ctx->stream);
  if (r) return r;
  if (tr)
  {
    th[5] = now_us();
    cudaEventSynchronize(ctx->tr_ev[5]);
    for (int i = 1; i <= 5; i++)
    {
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->tr_ev[0], ctx->tr_ev[i]);
      ctx->tr_dev_us[i] += 1e3 * ms;
      ctx->tr_host_us[i] += th[i] - th[0];
    }
    ctx->tr_n++;
  }
  int ok = 0;
  unstage_iterate(ctx->h_pub, o->x_curr, &o->last_iters, &ok);
  ctx->tm.iekf_iters = o->last_iters;
  ctx->tm.iekf_kernel_ms = 0;
  if (ok)
  {
    if (o->degrade_cnt > 0) o->degrade_cnt--;
  }
  else
    o->degrade_cnt++;
  memcpy(o->x_buf.back().R, o->x_curr.R, 72);
  memcpy(o->x_buf.back().p, o->x_curr.p, 24);
  if (margi)
  {
    journey_update(o);
    o->x_buf.erase(o->x_buf.begin());
    o->win_base += 1;
    o->win_count -= 1;
  }
  if (x_out) *x_out = o->x_curr;
  ctx->tm.kernel_launches = ctx->launches - l0;
  return VINA_OK;
}

static int odom_step_resident(vina_ctx* ctx, OdomHost* o, double pcl_beg_time, double pcl_end_time,
                              const vina_imu* imus, int m, int iekf_on_full, int max_iter, vina_state* x_out)
{
  if (ctx->overlap && !ctx->profiling && iekf_on_full && ctx->side_stream)
    return odom_step_overlapped(ctx, o, pcl_beg_time, pcl_end_time, imus, m, max_iter, x_out);
  const int l0 = ctx->launches;
  int which = 1, ok = 0;
  int r = vn_settle_upload(ctx);
  if (r) return r;
  r = step_front(ctx, o, pcl_beg_time, pcl_end_time, imus, m, iekf_on_full, &which);
  if (r) return r;
  r = lio_state_estimation(ctx, o, which, max_iter, &ok);
  if (r) return r;
  r = step_back(ctx, o, iekf_on_full, ok, x_out);
  if (r) return r;
  collect_timings(ctx);
  ctx->tm.kernel_launches = ctx->launches - l0;
  return VINA_OK;
}

// ---------------------------------------------------------------------------
// Batch replay (BASELINE.json configs[4]: "batch replay of 8 independent synthetic sequences"): B contexts on one
// GPU advance one scan each per call. Everything per-sequence (deskew ... var_init, map update) runs on the
// sequence's own stream, concurrently; the IEKF iterations of ALL sequences run as ONE k_iekf launch per
// iteration (grid = blocks x B, every sequence with its own device iterate, convergence flag and solve), which
// is what turns the latency-bound single-scan launch into a bandwidth-shaped one.
struct vina_batch
{
  std::vector<vina_ctx*> c;
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> ready;
  cudaEvent_t done = nullptr;
  std::vector<cudaEvent_t> tev;  // [i], [i+1] bracket the i-th batched launch of the last step
  int device = 0;
  int iekf_launches = 0;
};

extern "C" int vina_batch_create(vina_ctx** ctxs, int n, vina_batch** out)
{
  if (!ctxs || !out || n < 1 || n > VN_MAX_BATCH) return VINA_E_ARG;
  for (int i = 0; i < n; i++)
    if (!ctxs[i] || ctxs[i]->device != ctxs[0]->device) return VINA_E_ARG;
  vina_batch* b = new vina_batch();
  b->c.assign(ctxs, ctxs + n);
  // (the sequences of a batch run their fronts concurrently on their own streams: they keep the plain kernels - several
  // cooperative launches at once would only queue behind each other)
  for (int i = 0; i < n; i++) ctxs[i]->batch_member = true;
  b->device = ctxs[0]->device;
  cudaSetDevice(b->device);
  if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess)
  {
    delete b;
    return VINA_E_CUDA;
  }
  b->ready.resize(n);
  for (int i = 0; i < n; i++) cudaEventCreateWithFlags(&b->ready[i], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming);
  *out = b;
  return VINA_OK;
}

extern "C" void vina_batch_destroy(vina_batch* b)
{
  if (!b) return;
  cudaSetDevice(b->device);
  cudaStreamSynchronize(b->stream);
  for (cudaEvent_t e : b->ready) cudaEventDestroy(e);
  cudaEventDestroy(b->done);
  for (cudaEvent_t e : b->tev) cudaEventDestroy(e);
  cudaStreamDestroy(b->stream);
  delete b;
}

extern "C" int vina_batch_step_resident(vina_batch* b, const void* const* d_xyzt, const int32_t* n,
                                        const double* pcl_beg_time, const double* pcl_end_time,
                                        const vina_imu* const* imus, const int32_t* m, int iekf_on_full, int max_iter,
                                        vina_state* x_out)
{
  if (!b || !d_xyzt || !n || !pcl_beg_time || !pcl_end_time || !imus || !m) return VINA_E_ARG;
  const int B = (int)b->c.size();
  const int num_max_iter = max_iter > 0 ? max_iter : 20;
  IekfBatch bt;
  bt.mode = VN_IEKF_SOLVE;
  bt.variant = 0;
  // per sequence, on its own stream: scan in, a1 (host), a2, f1, a3, iterate upload, cache reset
  for (int i = 0; i < B; i++)
  {
    vina_ctx* ctx = b->c[i];
    if (!d_xyzt[i] || !imus[i] || m[i] <= 0 || n[i] <= 0) return VINA_E_ARG;
    if (n[i] > ctx->cap_points) return vn_fail(ctx, VINA_E_CAPACITY, "scan of %d points > max_scan_points", n[i]);
    int r = vn_check_cuda(ctx, cudaMemcpyAsync(ctx->d_scan, d_xyzt[i], (size_t)n[i] * sizeof(float4),
                                                cudaMemcpyDeviceToDevice, ctx->stream), "device-to-device scan copy");
    if (r) return r;
    ctx->n_scan = n[i];
    OdomHost* o = odom(ctx);
    int which = 1;
    r = step_front(ctx, o, pcl_beg_time[i], pcl_end_time[i], imus[i], m[i], iekf_on_full, &which);
    if (r) return r;
    r = iekf_stage(ctx, o, which, num_max_iter);
    if (r) return r;
    vn_iekf_fill_seq(ctx, &bt.s[i], false);
    cudaEventRecord(b->ready[i], ctx->stream);
    cudaStreamWaitEvent(b->stream, b->ready[i], 0);
  }
  // the IEKF of all sequences: one launch per iteration; every sequence gets an equal share of the SMs
  int blocks = b->c[0]->sm_count / B;
  if (blocks < 1) blocks = 1;
  while ((int)b->tev.size() < num_max_iter + 1)
  {
    cudaEvent_t e;
    cudaEventCreate(&e);
    b->tev.push_back(e);
  }
  cudaEventRecord(b->tev[0], b->stream);
  for (int it = 0; it < num_max_iter; it++)
  {
    int e = launch_iekf(b->stream, bt, B, blocks, false);
    if (e) return vn_check_cuda(b->c[0], (cudaError_t)e, "batched k_iekf launch");
    cudaEventRecord(b->tev[it + 1], b->stream);
  }
  b->iekf_launches = num_max_iter;
  for (int i = 0; i < B; i++)
  {
    vina_ctx* ctx = b->c[i];
    ctx->launches += num_max_iter;
    int r = vn_iterate_publish(ctx, b->stream);
    if (r) return r;
  }
  cudaEventRecord(b->done, b->stream);
  for (int i = 0; i < B; i++) cudaStreamWaitEvent(b->c[i]->stream, b->done, 0);
  // per sequence: take the result back, then the map update (asynchronous, on the sequence's stream)
  for (int i = 0; i < B; i++)
  {
    vina_ctx* ctx = b->c[i];
    OdomHost* o = odom(ctx);
    int r = vn_iterate_wait(ctx, b->stream);
    if (r) return r;
    if (ctx->n_down_pending && !ctx->n_down_mapped)
    {
      ctx->n_down = *ctx->h_n_down;  // enqueued before the batched launches: has landed
      ctx->n_down_pending = false;
    }
    int ok = 0;
    unstage_iterate(ctx->h_pub, o->x_curr, &o->last_iters, &ok);
    ctx->tm.iekf_iters = o->last_iters;
    r = step_back(ctx, o, iekf_on_full, ok, x_out ? &x_out[i] : nullptr);
    if (r) return r;
  }
  return VINA_OK;
}

extern "C" int vina_batch_iekf_time(vina_batch* b, float* ms_per_launch, int cap, int32_t* launches)
{
  if (!b || !ms_per_launch || !launches) return VINA_E_ARG;
  *launches = b->iekf_launches;
  if (b->iekf_launches > 0) cudaEventSynchronize(b->tev[b->iekf_launches]);
  for (int it = 0; it < b->iekf_launches && it < cap; it++)
    cudaEventElapsedTime(&ms_per_launch[it], b->tev[it], b->tev[it + 1]);
  return VINA_OK;
}

extern "C" int vina_batch_sync(vina_batch* b)
{
  if (!b) return VINA_E_ARG;
  for (vina_ctx* ctx : b->c)
  {
    int r = vn_check_status(ctx);
    if (r) return r;
  }
  return VINA_OK;
}

// ---------------------------------------------------------------------------
extern "C" {

int vina_odom_set_state(vina_ctx* ctx, const vina_state* s)
{
  if (!ctx || !s) return VINA_E_ARG;
  odom(ctx)->x_curr = *s;
  return VINA_OK;
}
int vina_odom_get_state(vina_ctx* ctx, vina_state* s)
{
  if (!ctx || !s) return VINA_E_ARG;
  *s = odom(ctx)->x_curr;
  return VINA_OK;
}
int vina_odom_set_imu_anchor(vina_ctx* ctx, double last_pcl_end_time, const vina_imu* last_imu, double scale_gravity)
{
  if (!ctx || !last_imu) return VINA_E_ARG;
  OdomHost* o = odom(ctx);
  o->last_pcl_end_time = last_pcl_end_time;
  o->last_imu = *last_imu;
  o->scale_gravity = scale_gravity;
  return VINA_OK;
}

int vina_odom_bootstrap(vina_ctx* ctx, const float* xyzt, int n, const vina_state* x_known)
{
  if (!ctx || !xyzt || !x_known || n <= 0) return VINA_E_ARG;
  OdomHost* o = odom(ctx);
  o->x_curr = *x_known;
  int r = vina_scan_upload(ctx, xyzt, n);
  if (r) return r;
  r = vina_downsample(ctx);
  if (r) return r;
  r = vn_finish_downsample(ctx);
  if (r) return r;
  r = vina_var_init(ctx, 1);
  if (r) return r;
  return map_update(ctx, o);
}

int vina_odom_step(vina_ctx* ctx, const float* xyzt, int n, double pcl_beg_time, const vina_imu* imus, int m,
                   int iekf_on_full, int max_iter, vina_state* x_out)
{
  if (!ctx || !xyzt || !imus || n <= 0 || m <= 0) return VINA_E_ARG;
  // (in chunks, the compute stream not waiting yet: the overlapped step's fused deskew follows the chunks; every other
  // path lets the stream wait for the whole upload first)
  int r = vn_scan_upload_chunked(ctx, xyzt, n);
  if (r) return r;
  // pcl_end_time = pcl_beg_time + back().curvature (sync.cpp:40)
  return odom_step_resident(ctx, odom(ctx), pcl_beg_time, pcl_beg_time + (double)xyzt[4 * (size_t)(n - 1) + 3], imus, m,
                            iekf_on_full, max_iter, x_out);
}

int vina_odom_step_resident(vina_ctx* ctx, const void* d_xyzt, int n, double pcl_beg_time, double pcl_end_time,
                            const vina_imu* imus, int m, int iekf_on_full, int max_iter, vina_state* x_out)
{
  if (!ctx || !d_xyzt || !imus || m <= 0 || n <= 0) return VINA_E_ARG;
  if (n > ctx->cap_points) return vn_fail(ctx, VINA_E_CAPACITY, "scan of %d points > max_scan_points", n);
  int r = vn_check_cuda(ctx,
                        cudaMemcpyAsync(ctx->d_scan, d_xyzt, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice,
                                        ctx->stream),
                        "device-to-device scan copy");
  if (r) return r;
  ctx->n_scan = n;
  return odom_step_resident(ctx, odom(ctx), pcl_beg_time, pcl_end_time, imus, m, iekf_on_full, max_iter, x_out);
}

int vina_odom_step_prepared(vina_ctx* ctx, double pcl_beg_time, const vina_imu* imus, int m, int iekf_on_full,
                            int max_iter, vina_state* x_out)
{
  if (!ctx || !imus || m <= 0) return VINA_E_ARG;
  if (!ctx->front_valid) return vn_fail(ctx, VINA_E_ARG, "vina_odom_step_prepared without a scan from vina_scan_prepare");
  ctx->front_valid = false;
  // pcl_end_time = pcl_beg_time + back().curvature (sync.cpp:40)
  return odom_step_resident(ctx, odom(ctx), pcl_beg_time, pcl_beg_time + (double)ctx->front_t_last, imus, m, iekf_on_full,
                            max_iter, x_out);
}

int vina_odom_propagate(vina_ctx* ctx, double pcl_beg_time, double pcl_end_time, const vina_imu* imus, int m,
                        vina_imu_pose* poses_out, int cap)
{
  if (!ctx || !imus || m <= 0) return VINA_E_ARG;
  OdomHost* o = odom(ctx);
  o->pcl_beg_time = pcl_beg_time;
  o->pcl_end_time = pcl_end_time;
  int r = imu_propagate(ctx, o, imus, m);
  if (r) return r;
  int np = (int)o->imu_poses.size();
  if (poses_out && cap >= np) memcpy(poses_out, o->imu_poses.data(), (size_t)np * sizeof(vina_imu_pose));
  return np;
}

int vina_odom_iekf(vina_ctx* ctx, int which, int max_iter, int* iters_out, int* not_degenerate)
{
  if (!ctx || which < 0 || which > 1) return VINA_E_ARG;
  OdomHost* o = odom(ctx);
  int ok = 0;
  int r = lio_state_estimation(ctx, o, which, max_iter, &ok);
  if (r) return r;
  if (iters_out) *iters_out = o->last_iters;
  if (not_degenerate) *not_degenerate = ok;
  return VINA_OK;
}

int vina_odom_iekf_host(vina_ctx* ctx, int which, int max_iter, int* iters_out, int* not_degenerate)
{
  if (!ctx || which < 0 || which > 1) return VINA_E_ARG;
  OdomHost* o = odom(ctx);
  int ok = 0;
  int r = lio_state_estimation_host(ctx, o, which, max_iter, &ok);
  if (r) return r;
  if (iters_out) *iters_out = o->last_iters;
  if (not_degenerate) *not_degenerate = ok;
  return VINA_OK;
}

int vina_odom_iekf_host_begin(vina_ctx* ctx, int max_iter)
{
  if (!ctx) return VINA_E_ARG;
  OdomHost* o = odom(ctx);
  o->hi_prop = o->x_curr;
  inverse_lu<15>(o->x_curr.cov, o->hi_cov_inv);  // odometry.cpp:82
  o->hi_iter = 0;
  o->hi_max = max_iter > 0 ? max_iter : 20;
  o->hi_rematch = 0;
  o->hi_active = true;
  return VINA_OK;
}

// one pass of odometry.cpp:192-230 with the sums of this iteration (any source: one GPU, all-reduced shards)
int vina_odom_iekf_host_update(vina_ctx* ctx, const double sums34[34])
{
  if (!ctx || !sums34) return VINA_E_ARG;
  OdomHost* o = odom(ctx);
  if (!o->hi_active) return vn_fail(ctx, VINA_E_STATE, "vina_odom_iekf_host_update before vina_odom_iekf_host_begin");
  vina_state& x_curr = o->x_curr;
  double HTH[36], HTz[6], nnt[9];
  int32_t match_num = 0;
  vn_iekf_unpack(sums34, HTH, HTz, nnt, &match_num);
  double HTH15[225], tmp[225], K1[225], G6[90];
  memset(HTH15, 0, sizeof(HTH15));
  for (int j = 0; j < 6; j++)
    for (int i = 0; i < 6; i++) HTH15[i + 15 * j] = HTH[i + 6 * j];
  for (int i = 0; i < 225; i++) tmp[i] = HTH15[i] + o->hi_cov_inv[i];
  inverse_lu<15>(tmp, K1);
  mat_mul(15, 6, 6, K1, HTH, G6);
  double vec[15], sol[15], a[15], b[15];
  state_boxminus(o->hi_prop, x_curr, vec);
  mat_mul(15, 6, 1, K1, HTz, a);
  mat_mul(15, 6, 1, G6, vec, b);
  for (int i = 0; i < 15; i++) sol[i] = (a[i] + vec[i]) - b[i];
  state_boxplus(x_curr, sol);
  const bool conv = (norm3(sol) * 57.3 < 0.01) && (norm3(sol + 3) * 100 < 0.015);
  if (conv || ((o->hi_rematch == 0) && (o->hi_iter == o->hi_max - 2))) o->hi_rematch++;
  const bool fin = o->hi_rematch >= 2 || (o->hi_iter == o->hi_max - 1);
  o->hi_iter++;
  o->last_iters = o->hi_iter;
  if (fin)
  {
    double IG[225], nc[225];
    memset(IG, 0, sizeof(IG));
    for (int i = 0; i < 90; i++) IG[i] = -G6[i];
    for (int i = 0; i < 15; i++) IG[i + 15 * i] += 1.0;
    mat_mul(15, 15, 15, IG, x_curr.cov, nc);
    memcpy(x_curr.cov, nc, sizeof(nc));
    o->hi_active = false;
  }
  return fin ? 1 : 0;
}

// LioStateEstimation against a map sharded over GPUs, the whole loop on the devices (vn_shard_iekf_enqueue).
// No host synchronisation inside the loop. Phases: include/vina_b200.h.
int vina_odom_iekf_sharded_p2p(vina_ctx* ctx, int first, int count, int max_iter, int phase, int* iters_out,
                               int* not_degenerate)
{
  if (!ctx || phase < VINA_SHARD_IEKF_ALL || phase > VINA_SHARD_IEKF_FINISH) return VINA_E_ARG;
  OdomHost* o = odom(ctx);
  const int num_max_iter = max_iter > 0 ? max_iter : 20;
  int r;
  if (phase == VINA_SHARD_IEKF_ALL || phase == VINA_SHARD_IEKF_STAGE)
  {
    stage_iterate(o->x_curr, num_max_iter, ctx->h_iekf);
    r = vn_check_cuda(ctx, cudaMemcpyAsync(ctx->d_iekf, ctx->h_iekf, sizeof(IekfDev), cudaMemcpyHostToDevice, ctx->stream),
                      "iterate upload");
    if (r) return r;
  }
  if (phase == VINA_SHARD_IEKF_ALL)
    r = vn_shard_iekf_enqueue(ctx, first, count, num_max_iter, 0);
  else if (phase >= VINA_SHARD_IEKF_ROUTE && phase <= VINA_SHARD_IEKF_SOLVE)
    r = vn_shard_iekf_enqueue(ctx, first, count, 1, phase - VINA_SHARD_IEKF_ROUTE + 1);
  else
    r = VINA_OK;
  if (r) return r;
  if (phase == VINA_SHARD_IEKF_ALL || phase == VINA_SHARD_IEKF_FINISH)
  {
    r = vn_iterate_publish(ctx, ctx->stream);
    if (r) return r;
    r = vn_iterate_wait(ctx, ctx->stream);
    if (r) return r;
    int ok = 0;
    unstage_iterate(ctx->h_pub, o->x_curr, &o->last_iters, &ok);
    if (iters_out) *iters_out = o->last_iters;
    if (not_degenerate) *not_degenerate = ok;
    return vn_check_status(ctx);
  }
  return VINA_OK;
}

// LocalBA.if_BA / LocalBA.imu_coef (node.cpp:96, 247). Set before the first frame enters the window. With BA on,
// vina_odom_step runs LI_BA_Optimizer::damping_iter between recut and margi (local_mapping.cpp:492-497) once every
// pair of consecutive window frames has an IMU pre-integration factor (frames inserted by vina_odom_bootstrap /
// vina_odom_map_update have none); the step then uses the serial schedule (the LM loop needs the host).
int vina_odom_set_ba(vina_ctx* ctx, int on, double imu_coef)
{
  if (!ctx) return VINA_E_ARG;
  OdomHost* o = odom(ctx);
  if (o->win_count > 0 && (on != 0) != o->if_BA) return vn_fail(ctx, VINA_E_STATE, "vina_odom_set_ba after frames entered the window");
  o->if_BA = on != 0;
  if (imu_coef > 0) o->imu_coef = imu_coef;
  return VINA_OK;
}
int vina_odom_ba_stats(vina_ctx* ctx, int32_t* runs, int32_t* last_iters)
{
  if (!ctx || !runs || !last_iters) return VINA_E_ARG;
  *runs = odom(ctx)->ba_runs;
  *last_iters = odom(ctx)->ba_last_iters;
  return VINA_OK;
}

int vina_odom_map_update(vina_ctx* ctx)
{
  if (!ctx) return VINA_E_ARG;
  return map_update(ctx, odom(ctx));
}

int vina_odom_journey(vina_ctx* ctx, double* jour, int* release_flag)
{
  if (!ctx) return VINA_E_ARG;
  OdomHost* o = odom(ctx);
  if (jour) *jour = o->jour;
  if (release_flag) *release_flag = o->release_flag ? 1 : 0;
  return VINA_OK;
}

int vina_odom_idle(vina_ctx* ctx, int horizon, int64_t* roots_erased, int64_t* nodes_freed)
{
  if (!ctx) return VINA_E_ARG;
  if (roots_erased) *roots_erased = 0;
  if (nodes_freed) *nodes_freed = 0;
  OdomHost* o = odom(ctx);
  if (!o->release_flag) return VINA_OK;
  o->release_flag = false;
  return vina_map_prune(ctx, o->jour, horizon, roots_erased, nodes_freed);
}

int vina_odom_window(vina_ctx* ctx, int* win_count, int* mp, int cap)
{
  if (!ctx || !win_count) return VINA_E_ARG;
  *win_count = odom(ctx)->win_count;
  for (int i = 0; i < ctx->cfg.win_size && i < cap; i++) mp[i] = ctx->map.mp[i];
  return ctx->cfg.win_size;
}
}

// ---------------------------------------------------------------------------------------------------------------
// Start-up phase: what the reference does with the first scans before the per-scan loop
// above can run - VINA_SLAM::initialization (src/platform/ros2/node.cpp:293-366) and the loop's handling of its
// result (src/pipeline/local_mapping.cpp:362-388, then :489-546 on success). Per scan: IMU initialisation or
// IMU propagation + deskew (IMUEKF::process, imu_ekf.cpp:174-201), down-sampling at max(down_size, 0.5 m),
// var_init, the kd-tree IEKF against a local map of world points (lio_state_estimation_kdtree,
// odometry.cpp:267-439), the window bookkeeping, the frame's retained raw cloud (down_sampling_close + time sort);
// with the window full, Initialization::motion_init (initialization.cpp:158-367): up to ten rounds of "rebuild the
// map from re-deskewed frames, recut, gravity BA", gravity alignment after the first convergence.
// Where it runs: the point-sized work on the device (deskew, down-sampling, var_init, nearest neighbours + plane
// fit + normal equations of the kd-tree IEKF, the re-deskew / covariance / insert / recut of every motion_init
// round, the LiDAR factor of the BA); on the host the sequential 15-dim algebra (IMU, the 15 x 15 IEKF update, the
// LM loop, gravity alignment) and down_sampling_close, whose result depends on the order of a float running sum
// (point_utils.hpp:75-86) - ten scans, once per run.
namespace
{
// down_sampling_close (include/vina_slam/core/point_utils.hpp:47-113) on (x, y, z, t) rows; the voxel map iterates in
// the order of the reference's container (same key, same hash: types.hpp:13-41)
struct CloseKey
{
  int64_t x, y, z;
  bool operator==(const CloseKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct CloseHash
{
  size_t operator()(const CloseKey& s) const
  {
    using std::hash;
    const long long P = 1000033, N = 100000000000LL;
    return (size_t)((((hash<int64_t>()(s.z) * P) % N + hash<int64_t>()(s.y)) * P) % N + hash<int64_t>()(s.x));
  }
};
void down_sampling_close_host(std::vector<float>& pl, double voxel_size)
{
  if (voxel_size < 0.001) return;
  const size_t n = pl.size() / 4;
  std::unordered_map<CloseKey, std::vector<int>, CloseHash> feat_map;
  for (size_t i = 0; i < n; i++)
  {
    float loc[3];
    for (int j = 0; j < 3; j++)
    {
      loc[j] = pl[4 * i + j] / voxel_size;
      if (loc[j] < 0) loc[j] -= 1.0;
    }
    feat_map[CloseKey{ (int64_t)loc[0], (int64_t)loc[1], (int64_t)loc[2] }].push_back((int)i);
  }
  std::vector<float> out;
  out.reserve(4 * feat_map.size());
  for (auto it = feat_map.begin(); it != feat_map.end(); ++it)
  {
    const std::vector<int>& idx = it->second;
    float pb[3] = { pl[4 * (size_t)idx[0]], pl[4 * (size_t)idx[0] + 1], pl[4 * (size_t)idx[0] + 2] };
    const int plsize = (int)idx.size();
    for (int k = 1; k < plsize; k++)
      for (int j = 0; j < 3; j++) pb[j] += pl[4 * (size_t)idx[k] + j];
    for (int j = 0; j < 3; j++) pb[j] /= plsize;
    double ndis = 100;
    int mnum = 0;
    for (int k = 0; k < plsize; k++)
    {
      const float* pp = &pl[4 * (size_t)idx[k]];
      const double xx = pb[0] - pp[0], yy = pb[1] - pp[1], zz = pb[2] - pp[2];
      const double dis = xx * xx + yy * yy + zz * zz;
      if (dis < ndis)
      {
        mnum = k;
        ndis = dis;
      }
    }
    const float* q = &pl[4 * (size_t)idx[mnum]];
    out.insert(out.end(), q, q + 4);
  }
  pl.swap(out);
}
void sort_by_time(std::vector<float>& pl)
{
  struct P4
  {
    float v[4];
  };
  P4* b = reinterpret_cast<P4*>(pl.data());
  std::sort(b, b + pl.size() / 4, [](const P4& x, const P4& y) { return x.v[3] < y.v[3]; });
}

// IMUEKF::IMU_init (imu_ekf.cpp:147-172)
void imu_init(OdomHost* o, const vina_imu* imus, int m)
{
  for (int k = 0; k < m; k++)
  {
    if (o->imu_init_num != 0)
    {
      for (int j = 0; j < 3; j++)
      {
        o->mean_acc[j] += (imus[k].acc[j] - o->mean_acc[j]) / (double)o->imu_init_num;
        o->mean_gyr[j] += (imus[k].gyr[j] - o->mean_gyr[j]) / (double)o->imu_init_num;
      }
    }
    else
    {
      for (int j = 0; j < 3; j++) o->mean_acc[j] = imus[k].acc[j], o->mean_gyr[j] = imus[k].gyr[j];
      o->imu_init_num = 1;
    }
    o->imu_init_num++;
  }
  o->last_imu = imus[m - 1];
}

void state_set_zero(vina_state& x)  // IMUST::setZero (types.hpp:101-112)
{
  memset(&x, 0, sizeof(x));
  x.R[0] = x.R[4] = x.R[8] = 1;
  x.g[2] = -9.8;
  for (int i = 0; i < 15; i++) x.cov[i + 15 * i] = i < 9 ? 1e-4 : 1e-5;
}

// VINA_SLAM::lio_state_estimation_kdtree (odometry.cpp:267-439): point loop on the device, update on the host
int kdtree_iekf(vina_ctx* ctx, OdomHost* o)
{
  vina_state& x_curr = o->x_curr;
  int r;
  if (ctx->n_pv[1] <= 0) return VINA_OK;
  if (ctx->n_tree < 100)
  {
    r = vn_init_tree_push(ctx, x_curr.R, x_curr.p);
    return r;
  }
  const int num_max_iter = 4;
  const vina_state x_prop = x_curr;
  bool converged_once = false;
  double G[225], HTH15[225], cov_inv[225], K1[225], tmp[225];
  memset(G, 0, sizeof(G));
  memset(HTH15, 0, sizeof(HTH15));
  int rematch_num = 0;
  inverse_lu<15>(x_curr.cov, cov_inv);
  bool refind = true;
  for (int iterCount = 0; iterCount < num_max_iter; iterCount++)
  {
    double sums[28];
    r = vn_init_assoc(ctx, x_curr.R, x_curr.p, refind ? 1 : 0, sums);
    if (r) return r;
    double HTH[36], HTz[6];
    {
      int t = 0;
      for (int a = 0; a < 6; a++)
        for (int b = a; b < 6; b++, t++) HTH[a + 6 * b] = HTH[b + 6 * a] = sums[t];
      for (int a = 0; a < 6; a++) HTz[a] = sums[21 + a];
    }
    for (int j = 0; j < 6; j++)
      for (int i = 0; i < 6; i++) HTH15[i + 15 * j] = HTH[i + 6 * j];
    for (int i = 0; i < 225; i++) tmp[i] = HTH15[i] + cov_inv[i] / 1000;  // (H_T_H + cov_inv / 1000).inverse()
    inverse_lu<15>(tmp, K1);
    double G6[90];
    mat_mul(15, 6, 6, K1, HTH, G6);
    memcpy(G, G6, sizeof(G6));
    double vec[15], sol[15], a[15], b[15];
    state_boxminus(x_prop, x_curr, vec);
    mat_mul(15, 6, 1, K1, HTz, a);
    mat_mul(15, 6, 1, G6, vec, b);
    for (int i = 0; i < 15; i++) sol[i] = (a[i] + vec[i]) - b[i];
    state_boxplus(x_curr, sol);
    refind = false;
    if ((norm3(sol) * 57.3 < 0.01) && (norm3(sol + 3) * 100 < 0.015))
    {
      refind = true;
      converged_once = true;
      rematch_num++;
    }
    if (iterCount == num_max_iter - 2 && !converged_once) refind = true;
    if (rematch_num >= 2 || (iterCount == num_max_iter - 1))
    {
      double IG[225], nc[225];
      for (int i = 0; i < 225; i++) IG[i] = -G[i];
      for (int i = 0; i < 15; i++) IG[i + 15 * i] = 1.0 - G[i + 15 * i];
      mat_mul(15, 15, 15, IG, x_curr.cov, nc);
      memcpy(x_curr.cov, nc, sizeof(nc));
      break;
    }
  }
  // the scan joins the local map, which is thinned to a 0.5 m grid (odometry.cpp:429-438)
  r = vn_init_tree_push(ctx, x_curr.R, x_curr.p);
  if (r) return r;
  int n_out = 0;
  r = vn_downsample_cloud(ctx, ctx->d_tree[ctx->tree_cur], ctx->n_tree, 0.5, ctx->d_tree[1 - ctx->tree_cur], &n_out);
  if (r) return r;
  ctx->tree_cur = 1 - ctx->tree_cur;
  ctx->n_tree = n_out;
  return VINA_OK;
}

// Initialization::align_gravity (initialization.cpp:28-62)
void align_gravity(std::vector<vina_state>& xs)
{
  double g0[3] = { xs[0].g[0], xs[0].g[1], xs[0].g[2] };
  const double gn = norm3(g0);
  double n0[3] = { g0[0] / gn, g0[1] / gn, g0[2] / gn };
  double n1[3] = { 0, 0, 1 };
  if (n0[2] < 0) n1[2] = -1;
  double rv[3] = { n0[1] * n1[2] - n0[2] * n1[1], n0[2] * n1[0] - n0[0] * n1[2], n0[0] * n1[1] - n0[1] * n1[0] };
  const double rnorm = norm3(rv);
  for (int k = 0; k < 3; k++) rv[k] /= rnorm;
  // AngleAxisd(asin(rnorm), axis).toRotationMatrix()
  const double ang = std::asin(rnorm), s = std::sin(ang), c = std::cos(ang);
  double rot[9];  // column-major
  const double c1[3] = { (1 - c) * rv[0], (1 - c) * rv[1], (1 - c) * rv[2] }, sa[3] = { s * rv[0], s * rv[1], s * rv[2] };
  double tmpv;
  tmpv = c1[0] * rv[1];
  rot[0 + 3 * 1] = tmpv - sa[2];
  rot[1 + 3 * 0] = tmpv + sa[2];
  tmpv = c1[0] * rv[2];
  rot[0 + 3 * 2] = tmpv + sa[1];
  rot[2 + 3 * 0] = tmpv - sa[1];
  tmpv = c1[1] * rv[2];
  rot[1 + 3 * 2] = tmpv - sa[0];
  rot[2 + 3 * 1] = tmpv + sa[0];
  rot[0] = c1[0] * rv[0] + c;
  rot[4] = c1[1] * rv[1] + c;
  rot[8] = c1[2] * rv[2] + c;
  double g1[3];
  m3_vec(rot, g0, g1);
  const double p0[3] = { xs[0].p[0], xs[0].p[1], xs[0].p[2] };
  for (size_t i = 0; i < xs.size(); i++)
  {
    double d[3] = { xs[i].p[0] - p0[0], xs[i].p[1] - p0[1], xs[i].p[2] - p0[2] }, rd[3], Rn[9], vn[3];
    m3_vec(rot, d, rd);
    for (int k = 0; k < 3; k++) xs[i].p[k] = rd[k] + p0[k];
    m3_mul(rot, xs[i].R, Rn);
    memcpy(xs[i].R, Rn, 72);
    m3_vec(rot, xs[i].v, vn);
    memcpy(xs[i].v, vn, 24);
    memcpy(xs[i].g, g1, 24);
  }
}

// the pose table of Initialization::motion_blur (initialization.cpp:64-111): BACKWARD integration from the frame's
// end state xc (biases of the previous frame), latest pose first
void backward_poses(const vina_state& xc_in, const vina_state& xl, const std::deque<vina_imu>& imus, double pcl_beg_time,
                    double scale_gravity, std::vector<vina_imu_pose>& out)
{
  out.clear();
  double R_imu[9], pos[3], vel[3];
  memcpy(R_imu, xc_in.R, 72);
  memcpy(pos, xc_in.p, 24);
  memcpy(vel, xc_in.v, 24);
  for (size_t it = imus.size() - 1; it != 0; it--)
  {
    const vina_imu &head = imus[it - 1], &tail = imus[it];
    double w[3], a[3], acc[3];
    for (int k = 0; k < 3; k++)
    {
      w[k] = 0.5 * (head.gyr[k] + tail.gyr[k]) - xl.bg[k];
      a[k] = 0.5 * (head.acc[k] + tail.acc[k]) * scale_gravity - xl.ba[k];
    }
    const double dt = head.t - tail.t;
    double E[9], Rn[9], Ra[3];
    Exp_dt(w, dt, E);
    m3_vec(R_imu, a, Ra);
    for (int k = 0; k < 3; k++) acc[k] = Ra[k] + xc_in.g[k];
    for (int k = 0; k < 3; k++) pos[k] = pos[k] + vel[k] * dt + 0.5 * acc[k] * dt * dt;
    for (int k = 0; k < 3; k++) vel[k] = vel[k] + acc[k] * dt;
    m3_mul(R_imu, E, Rn);
    memcpy(R_imu, Rn, 72);
    vina_imu_pose ps;
    ps.t = head.t - pcl_beg_time;
    memcpy(ps.R, R_imu, 72);
    memcpy(ps.p, pos, 24);
    memcpy(ps.v, vel, 24);
    memcpy(ps.w, w, 24);
    memcpy(ps.a, acc, 24);
    out.push_back(ps);
  }
}

// Initialization::motion_init (initialization.cpp:158-367); 1 = converged, 0 = failed (the map is empty then)
int motion_init(vina_ctx* ctx, OdomHost* o, int* ok_out)
{
  *ok_out = 0;
  const int win_size = ctx->cfg.win_size;
  std::vector<vina_state>& xs = o->init_xs;
  int converge_flag = 0;
  MapView& M = ctx->map;
  const double min_eig_orig = M.min_eigen_value;
  double thre_orig[4];
  for (int k = 0; k < 4; k++) thre_orig[k] = M.thre[k];
  const int thread_num_orig = M.thread_num;
  M.min_eigen_value = 0.02;
  for (int k = 0; k < 4; k++) M.thre[k] = 1.0 / 4;
  M.thread_num = 0;  // motion_init calls cut_voxel / recut / tras_opt directly: none of the fan-out drivers' early-outs
  double converge_thre = 0.05;
  bool is_degrade = true;
  double eigvalue[3] = { 0, 0, 0 };
  std::vector<vina_imu_pose> poses;
  std::vector<vina_pose> xb(win_size);
  int r = VINA_OK;
  o->init_rounds = 0;
  for (int iterCnt = 0; iterCnt < 10 && r == VINA_OK; iterCnt++)
  {
    o->init_rounds++;
    if (converge_flag == 1)
    {
      M.min_eigen_value = min_eig_orig;
      for (int k = 0; k < 4; k++) M.thre[k] = thre_orig[k];
    }
    r = vn_map_clear(ctx);
    if (r) break;
    for (int i = 0; i < win_size && r == VINA_OK; i++)
    {
      const int l = i == 0 ? i : i - 1;
      backward_poses(xs[i], xs[l], o->vec_imus[i], o->beg_times[i], o->scale_gravity, poses);
      const std::vector<float>& pl = o->pl_origs[i];
      const int n = (int)(pl.size() / 4);
      // points at or before the earliest pose are not compensated and not pushed (initialization.cpp:139)
      const double t_first = poses.empty() ? 1e300 : poses.back().t;
      int n_skip = 0;
      while (n_skip < n && !((double)pl[4 * (size_t)n_skip + 3] > t_first)) n_skip++;
      r = vn_init_insert_frame(ctx, pl.data(), n, n_skip, poses.data(), (int)poses.size(), &xs[i], converge_flag, i);
    }
    if (r) break;
    for (int i = 0; i < win_size; i++)
    {
      memcpy(xb[i].R, xs[i].R, 72);
      memcpy(xb[i].p, xs[i].p, 24);
    }
    r = vina_map_recut(ctx, win_size, xb.data());
    if (r) break;
    int32_t n_fac = 0;
    r = vina_ba_collect(ctx, &n_fac);  // tras_opt of every root
    if (r) break;
    if (n_fac < 10) break;
    double resis[2] = { 0, 0 };
    int iters = 0;
    r = ba_damping_iter_ex(ctx, xs, o->imu_pre_buf, o->imu_coef, &iters, true, 3, resis);
    if (r) break;
    for (ImuPre* f : o->imu_pre_buf) ba_imu_factor_delete(f);
    o->imu_pre_buf.clear();
    for (int i = 1; i < win_size; i++)
      o->imu_pre_buf.push_back(ba_imu_factor_new(xs[i - 1].bg, xs[i - 1].ba, o->vec_imus[i], o->scale_gravity, ctx->cfg));
    if (std::fabs(resis[0] - resis[1]) / resis[0] < converge_thre && iterCnt >= 2)
    {
      // lambda_min of the sum of n n^T over the factors' normals (initialization.cpp:277-285)
      double nnt[9];
      r = vn_ba_normal_scatter(ctx, nnt);
      if (r) break;
      double L[6] = { nnt[0], nnt[1], nnt[2], nnt[4], nnt[5], nnt[8] }, Q[9];
      eig3_sym(L, eigvalue, Q);
      is_degrade = eigvalue[0] < 15;
      converge_thre = 0.01;
      if (converge_flag == 0)
      {
        align_gravity(xs);
        converge_flag = 1;
        continue;
      }
      break;
    }
  }
  M.min_eigen_value = min_eig_orig;
  for (int k = 0; k < 4; k++) M.thre[k] = thre_orig[k];
  M.thread_num = thread_num_orig;
  if (r) return r;
  o->x_curr = xs[win_size - 1];
  const double gnm = norm3(o->x_curr.g);
  if (is_degrade) converge_flag = 0;
  if (gnm < 9.6 || gnm > 10.0) converge_flag = 0;
  if (converge_flag == 0)
  {
    r = vn_map_clear(ctx);
    if (r) return r;
  }
  o->pl_origs.clear();
  o->vec_imus.clear();
  o->beg_times.clear();
  *ok_out = converge_flag;
  return VINA_OK;
}

// VINA_SLAM::system_reset (node.cpp:368-408)
int system_reset(vina_ctx* ctx, OdomHost* o, const vina_imu* imus, int m)
{
  int r = vn_map_clear(ctx);
  if (r) return r;
  state_set_zero(o->x_curr);
  o->x_curr.p[2] = 30;
  for (int k = 0; k < 3; k++) o->mean_acc[k] = 0;
  o->imu_init_num = 0;
  imu_init(o, imus, m);
  for (int k = 0; k < 3; k++) o->x_curr.g[k] = -o->mean_acc[k] * o->scale_gravity;
  for (ImuPre* f : o->imu_pre_buf) ba_imu_factor_delete(f);
  o->imu_pre_buf.clear();
  o->x_buf.clear();
  o->xs_buf.clear();
  o->init_xs.clear();
  ctx->n_tree = 0;
  o->win_base = 0;
  o->win_count = 0;
  return VINA_OK;
}
}  // namespace

int vina_odom_cold_start(vina_ctx* ctx)
{
  if (!ctx) return VINA_E_ARG;
  OdomHost* o = odom(ctx);
  int r = vn_init_ensure(ctx);
  if (r) return r;
  r = vn_map_clear(ctx);
  if (r) return r;
  state_set_zero(o->x_curr);
  o->in_init = true;
  o->imu_init_flag = false;
  o->imu_init_num = 0;
  for (int k = 0; k < 3; k++) o->mean_acc[k] = o->mean_gyr[k] = 0;
  for (ImuPre* f : o->imu_pre_buf) ba_imu_factor_delete(f);
  o->imu_pre_buf.clear();
  o->x_buf.clear();
  o->xs_buf.clear();
  o->init_xs.clear();
  o->pl_origs.clear();
  o->vec_imus.clear();
  o->beg_times.clear();
  o->win_base = o->win_count = 0;
  o->scale_gravity = 1.0;
  o->last_pcl_end_time = 0;
  ctx->n_tree = 0;
  ctx->tree_cur = 0;
  return VINA_OK;
}

int vina_odom_init_scan(vina_ctx* ctx, const float* xyzt, int n, double pcl_beg_time, const vina_imu* imus, int m,
                        vina_state* x_out, int32_t* status)
{
  if (!ctx || !xyzt || !imus || n <= 0 || m <= 0 || !status) return VINA_E_ARG;
  OdomHost* o = odom(ctx);
  if (!o->in_init) return vn_fail(ctx, VINA_E_STATE, "vina_odom_init_scan without vina_odom_cold_start");
  *status = 0;
  const int win_size = ctx->cfg.win_size;
  o->pcl_beg_time = pcl_beg_time;
  o->pcl_end_time = pcl_beg_time + (double)xyzt[4 * (size_t)(n - 1) + 3];  // sync.cpp:40
  int r;
  // ---- IMUEKF::process (imu_ekf.cpp:174-201)
  if (!o->imu_init_flag)
  {
    imu_init(o, imus, m);
    if (norm3(o->mean_acc) < 2) o->scale_gravity = 9.8;
    for (int k = 0; k < 3; k++) o->x_curr.g[k] = -o->mean_acc[k] * o->scale_gravity;
    if (o->imu_init_num > 30) o->imu_init_flag = true;  // min_init_num (ekf_imu.hpp:18)
    o->last_pcl_end_time = o->pcl_end_time;
    if (x_out) *x_out = o->x_curr;
    return VINA_OK;
  }
  r = vina_scan_upload(ctx, xyzt, n);
  if (r) return r;
  r = imu_propagate(ctx, o, imus, m);
  if (r) return r;
  r = vina_deskew(ctx, o->imu_poses.data(), (int)o->imu_poses.size(), o->x_curr.R, o->x_curr.p);
  if (r) return r;
  // ---- down_sampling_voxel(max(down_size, 0.5)), var_init, the kd-tree IEKF
  const double downkd = ctx->cfg.down_size >= 0.5 ? ctx->cfg.down_size : 0.5;
  int n_down = 0;
  r = vn_downsample_cloud(ctx, ctx->d_scan, ctx->n_scan, downkd, ctx->d_down, &n_down);
  if (r) return r;
  ctx->n_down = n_down;
  ctx->n_down_pending = false;
  r = vina_var_init(ctx, 1);
  if (r) return r;
  r = kdtree_iekf(ctx, o);
  if (r) return r;
  // ---- the window (node.cpp:322-331)
  o->win_count++;
  o->init_xs.push_back(o->x_curr);
  if (o->win_count > 1)
  {
    const vina_state& prev = o->init_xs[o->win_count - 2];
    o->imu_pre_buf.push_back(ba_imu_factor_new(prev.bg, prev.ba, o->ba_imus, o->scale_gravity, ctx->cfg));
  }
  // ---- the frame's retained raw cloud (node.cpp:333-345)
  std::vector<float> orig(xyzt, xyzt + 4 * (size_t)n);
  {
    std::vector<float> mid = orig;
    down_sampling_close_host(orig, ctx->cfg.down_size);
    if (orig.size() / 4 < 1000)
    {
      orig = mid;
      down_sampling_close_host(orig, ctx->cfg.down_size / 2);
    }
    sort_by_time(orig);
  }
  o->pl_origs.push_back(std::move(orig));
  o->beg_times.push_back(o->pcl_beg_time);
  o->vec_imus.push_back(o->ba_imus);
  if (x_out) *x_out = o->x_curr;
  if (o->win_count < win_size) return VINA_OK;
  // ---- motion_init and what the loop does with its result
  int ok = 0;
  r = motion_init(ctx, o, &ok);
  if (r) return r;
  if (!ok)
  {
    *status = -1;
    r = system_reset(ctx, o, imus, m);
    if (x_out) *x_out = o->x_curr;
    return r;
  }
  // success: the window tail of this scan (local_mapping.cpp:489-546) - BA on the factors motion_init left, margi,
  // window shift - and the hand-over to the per-scan loop
  *status = 1;
  o->in_init = false;
  o->x_buf.clear();
  o->xs_buf.clear();
  for (int i = 0; i < win_size; i++)
  {
    vina_pose ps;
    memcpy(ps.R, o->init_xs[i].R, 72);
    memcpy(ps.p, o->init_xs[i].p, 24);
    o->x_buf.push_back(ps);
    if (o->if_BA) o->xs_buf.push_back(o->init_xs[i]);
  }
  if (o->if_BA)
  {
    r = ba_damping_iter(ctx, o->xs_buf, o->imu_pre_buf, o->imu_coef, &o->ba_last_iters);
    if (r) return r;
    o->ba_runs++;
    for (int i = 0; i < win_size; i++)
    {
      memcpy(o->x_buf[i].R, o->xs_buf[i].R, 72);
      memcpy(o->x_buf[i].p, o->xs_buf[i].p, 24);
    }
  }
  r = vn_ba_writeback_enqueue(ctx);  // margi takes the factors' (re-evaluated) pcr_add / eig (octree.cpp:410-416)
  if (r) return r;
  memcpy(o->x_curr.R, o->x_buf.back().R, 72);
  memcpy(o->x_curr.p, o->x_buf.back().p, 24);
  ctx->map.jour = o->jour;
  r = vina_map_margi(ctx, o->win_count, o->x_buf.data());
  if (r) return r;
  journey_update(o);
  r = vina_map_shift_window(ctx);
  if (r) return r;
  o->x_buf.erase(o->x_buf.begin());
  if (o->if_BA) o->xs_buf.erase(o->xs_buf.begin());
  ba_imu_factor_delete(o->imu_pre_buf.front());
  o->imu_pre_buf.pop_front();
  if (!o->if_BA)
  {
    // (without BA the per-scan loop keeps no pre-integration factors)
    for (ImuPre* f : o->imu_pre_buf) ba_imu_factor_delete(f);
    o->imu_pre_buf.clear();
  }
  o->win_base += 1;
  o->win_count -= 1;
  o->init_xs.clear();
  if (x_out) *x_out = o->x_curr;
  return VINA_OK;
}

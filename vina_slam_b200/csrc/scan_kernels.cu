// Per-point scan kernels: deskew (a2), var_init (a3), down_sampling_voxel (f1).
// This translation unit is compiled with -fmad=false: every expression keeps
// the reference's operation order with one rounding per operation, so the
// results can be compared bit-for-bit with the strict CPU restatement.
#include <cstring>
#include "vn_kernels.cuh"

struct Cov2x
{
  double rot[9], tsl[9];
};

// ---------------------------------------------------------------------------
// a2 deskew: src/estimation/imu_ekf.cpp:114-144. The reference walks the
// time-sorted scan backwards together with the IMU pose table; for sorted
// input that is "pose k(i) = last k with t_k < curvature_i", found here by
// binary search. Points with curvature <= t_0 are left untouched (:124).
// Point 0 is re-compensated by every earlier pose as well, because the
// reference's inner loop breaks at begin() without leaving the outer loop
// (:139-142).
__device__ __forceinline__ void exp_so3_dt(const double* w, double dt, double* E)
{
  double nrm = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  for (int i = 0; i < 9; i++) E[i] = 0.0;
  E[0] = E[4] = E[8] = 1.0;
  if (nrm > 1e-7)  // math.hpp:29
  {
    double ax[3] = { w[0] / nrm, w[1] / nrm, w[2] / nrm };
    double K[9], KK[9], sK[9];
    hat3(ax, K);
    double r_ang = nrm * dt;
    double s = sin(r_ang), c1 = 1.0 - cos(r_ang);
    for (int i = 0; i < 9; i++) sK[i] = c1 * K[i];
    // ((1-cos) K) K, coefficient sums left to right
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 3; i++)
        KK[i + 3 * j] = sK[i] * K[3 * j] + sK[i + 3] * K[3 * j + 1] + sK[i + 6] * K[3 * j + 2];
    for (int i = 0; i < 9; i++) E[i] = (E[i] + s * K[i]) + KK[i];
  }
}

__device__ __forceinline__ void compensate(const DeskewPoses& P, int k, float curv, float* xyz)
{
  const vina_imu_pose& h = P.pose[k];
  double dt = (double)curv - h.t;
  double E[9], Ri[9];
  exp_so3_dt(h.w, dt, E);
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++)
      Ri[i + 3 * j] = h.R[i] * E[3 * j] + h.R[i + 3] * E[3 * j + 1] + h.R[i + 6] * E[3 * j + 2];
  double T[3];
  for (int i = 0; i < 3; i++) T[i] = ((h.p[i] + h.v[i] * dt) + ((0.5 * h.a[i]) * dt) * dt) - P.p_end[i];
  double Pi[3] = { (double)xyz[0], (double)xyz[1], (double)xyz[2] };
  double a[3], b[3], c[3], d[3];
  for (int i = 0; i < 3; i++)
    a[i] = ((P.ext_R[i] * Pi[0] + P.ext_R[i + 3] * Pi[1]) + P.ext_R[i + 6] * Pi[2]) + P.ext_t[i];
  for (int i = 0; i < 3; i++) b[i] = ((Ri[i] * a[0] + Ri[i + 3] * a[1]) + Ri[i + 6] * a[2]) + T[i];
  for (int i = 0; i < 3; i++)
    c[i] = ((P.R_end[3 * i] * b[0] + P.R_end[3 * i + 1] * b[1]) + P.R_end[3 * i + 2] * b[2]) - P.ext_t[i];
  for (int i = 0; i < 3; i++) d[i] = (P.ext_R[3 * i] * c[0] + P.ext_R[3 * i + 1] * c[1]) + P.ext_R[3 * i + 2] * c[2];
  xyz[0] = (float)d[0];
  xyz[1] = (float)d[1];
  xyz[2] = (float)d[2];
}

// one point of the deskew loop (imu_ekf.cpp:114-144); returns the point as the scan keeps it
__device__ __forceinline__ float4 deskew_point(const DeskewPoses& P, const float4* __restrict__ pts, int i, int n,
                                               int* __restrict__ status)
{
  float4 q = pts[i];
  // the contract of lidar_decoder.cpp:30: sorted by curvature
  if (i + 1 < n && pts[i + 1].w < q.w) atomicOr(status, VN_ST_UNSORTED);
  // k = last pose with t_k < curvature
  int lo = 0, hi = P.m;  // first index with t >= curv
  double cv = (double)q.w;
  while (lo < hi)
  {
    int mid = (lo + hi) >> 1;
    if (P.pose[mid].t < cv)
      lo = mid + 1;
    else
      hi = mid;
  }
  int k = lo - 1;
  if (k < 0) return q;  // points at or before the first pose are left untouched (imu_ekf.cpp:124)
  float xyz[3] = { q.x, q.y, q.z };
  compensate(P, k, q.w, xyz);
  if (i == 0)
    for (int kk = k - 1; kk >= 0; kk--) compensate(P, kk, q.w, xyz);
  return make_float4(xyz[0], xyz[1], xyz[2], q.w);
}

// the pose table into shared memory: the header, the m poses in use (of VINA_MAX_POSES slots - a scan has 20 to 40),
// the end pose and the extrinsic behind the array
__device__ __forceinline__ void stage_poses(DeskewPoses& P, const DeskewPoses* __restrict__ Pg)
{
  const int m = Pg->m;
  const int head = (int)(offsetof(DeskewPoses, pose) / 4) + m * (int)(sizeof(vina_imu_pose) / 4);
  const int tail0 = (int)(offsetof(DeskewPoses, R_end) / 4), words = (int)(sizeof(DeskewPoses) / 4);
  const int* src = reinterpret_cast<const int*>(Pg);
  int* dst = reinterpret_cast<int*>(&P);
  for (int i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
  for (int i = tail0 + threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) k_deskew(float4* __restrict__ pts, int n, const DeskewPoses* __restrict__ Pg,
                                                int* __restrict__ status)
{
  __shared__ DeskewPoses P;
  stage_poses(P, Pg);
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 r = deskew_point(P, pts, i, n, status);
  // (a thread reads its right-hand neighbour's time stamp, which deskewing never changes)
  pts[i] = r;
}

// ---------------------------------------------------------------------------
// a3 var_init: src/core/point_utils.cpp:3-52 (calcBodyVar + extrinsic).
// calcBodyVar (point_utils.cpp:3-34): pb may be mutated (z == 0 -> 1e-4), var = 3x3 column-major
__device__ __forceinline__ void calc_body_var(double* pb, float range_var, double dir_var, double* var)
{
  if (pb[2] == 0) pb[2] = 0.0001;
  float range = (float)sqrt(pb[0] * pb[0] + pb[1] * pb[1] + pb[2] * pb[2]);
  double rv = (double)range_var;
  double dv = dir_var;
  double d[3] = { pb[0], pb[1], pb[2] };
  {
    double z = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
    if (z > 0)
    {
      double s = sqrt(z);
      d[0] = d[0] / s;
      d[1] = d[1] / s;
      d[2] = d[2] / s;
    }
  }
  double dh[9];
  hat3(d, dh);
  double b1[3] = { 1.0, 1.0, -(d[0] + d[1]) / d[2] };
  {
    double z = (b1[0] * b1[0] + b1[1] * b1[1]) + b1[2] * b1[2];
    if (z > 0)
    {
      double s = sqrt(z);
      b1[0] = b1[0] / s;
      b1[1] = b1[1] / s;
      b1[2] = b1[2] / s;
    }
  }
  double b2[3] = { b1[1] * d[2] - b1[2] * d[1], b1[2] * d[0] - b1[0] * d[2], b1[0] * d[1] - b1[1] * d[0] };
  {
    double z = (b2[0] * b2[0] + b2[1] * b2[1]) + b2[2] * b2[2];
    if (z > 0)
    {
      double s = sqrt(z);
      b2[0] = b2[0] / s;
      b2[1] = b2[1] / s;
      b2[2] = b2[2] / s;
    }
  }
  // A = (range * direction_hat) * N, N = [b1 b2]
  double rh[9];
  for (int k = 0; k < 9; k++) rh[k] = (double)range * dh[k];
  double A[6];  // column-major 3x2
  for (int r = 0; r < 3; r++)
  {
    A[r] = (rh[r] * b1[0] + rh[r + 3] * b1[1]) + rh[r + 6] * b1[2];
    A[r + 3] = (rh[r] * b2[0] + rh[r + 3] * b2[1]) + rh[r + 6] * b2[2];
  }
  // var = (d*rv)*d^T + (A*diag(dv,dv))*A^T
  double AD[6];
  for (int r = 0; r < 3; r++)
  {
    AD[r] = A[r] * dv + A[r + 3] * 0.0;
    AD[r + 3] = A[r] * 0.0 + A[r + 3] * dv;
  }
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) var[r + 3 * c] = (d[r] * rv) * d[c] + (AD[r] * A[c] + AD[r + 3] * A[c + 3]);
}

__device__ __forceinline__ void var_init_point(const float4 q, const int i, const ScanView& out, const VarInitParams& prm)
{
  double pb[3] = { (double)q.x, (double)q.y, (double)q.z };
  double var[9];
  calc_body_var(pb, prm.range_var, prm.dir_var, var);
  // extrinsic: pnt = R_L pnt + t_L ; var = R_L var R_L^T
  double pn[3];
  for (int r = 0; r < 3; r++)
    pn[r] = ((prm.ext_R[r] * pb[0] + prm.ext_R[r + 3] * pb[1]) + prm.ext_R[r + 6] * pb[2]) + prm.ext_t[r];
  double T[9];
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++)
      T[r + 3 * c] = (prm.ext_R[r] * var[3 * c] + prm.ext_R[r + 3] * var[3 * c + 1]) + prm.ext_R[r + 6] * var[3 * c + 2];
  const int ui[6] = { 0, 0, 0, 1, 1, 2 }, uj[6] = { 0, 1, 2, 1, 2, 2 };
  out.p[0][i] = pn[0];
  out.p[1][i] = pn[1];
  out.p[2][i] = pn[2];
  for (int k = 0; k < 6; k++)
  {
    int r = ui[k], c = uj[k];
    out.v[k][i] = (T[r] * prm.ext_R[c] + T[r + 3] * prm.ext_R[c + 3]) + T[r + 6] * prm.ext_R[c + 6];
  }
}

__global__ void __launch_bounds__(256)
    k_var_init(const float4* __restrict__ pts, const int* __restrict__ n_ptr, int n_host, ScanView out, VarInitParams prm)
{
  vn_pdl_sync();
  int n = n_ptr ? *n_ptr : n_host;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  var_init_point(pts[i], i, out, prm);
}

// a2 + a3 of the full scan in one pass (the production schedule: VNC_lio runs on the un-downsampled scan,
// local_mapping.cpp:406-413): deskew, store the float point, var_init from that float value - exactly what the
// two kernels do one after the other - and reset the IEKF's per-point leaf cache (odometry.cpp:79)
// [first, last) = the part of the scan this launch covers: a scan that is still arriving from the host is processed
// chunk by chunk behind the chunks' copies (vn_deskew_var_init). The sortedness check of the contract looks at the
// right-hand neighbour, which for the last point of a chunk may not have landed yet: a chunk checks its first point
// against the point before it instead.
__global__ void __launch_bounds__(256, 3)
    k_deskew_var_init(float4* __restrict__ pts, int n, const DeskewPoses* __restrict__ Pg, int* __restrict__ status,
                      ScanView out, VarInitParams prm, int* __restrict__ cache, int first, int last)
{
  vn_pdl_sync();
  __shared__ DeskewPoses P;
  stage_poses(P, Pg);
  __syncthreads();
  int i = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= last) return;
  if (i == first && i > 0 && pts[i].w < pts[i - 1].w) atomicOr(status, VN_ST_UNSORTED);
  const float4 r = deskew_point(P, pts, i, last, status);
  pts[i] = r;
  var_init_point(r, i, out, prm);
  cache[i] = -1;
}

// ---------------------------------------------------------------------------
// f1 down_sampling_voxel: include/vina_slam/core/point_utils.hpp:7-44.
// The reference keeps an order-dependent fp32 running mean per voxel and emits
// voxels in unordered_map order; here each voxel's mean is the fp64 sum of its
// points (exact for float inputs, hence order-independent) divided by the count
// and rounded to float, and voxels are emitted in order of their first point.
// Same voxel set and counts as the reference; coordinates agree to fp32
// rounding ("tolerance parity only").
__device__ __forceinline__ void down_accum_point(const float4 q, const int i, double voxel_size, DownSlot* __restrict__ tab,
                                                 unsigned int mask, int* __restrict__ slot_of, int* __restrict__ status)
{
  const float data[3] = { q.x, q.y, q.z };
  long long kc[3];
  for (int j = 0; j < 3; j++)
  {
    float loc = (float)((double)data[j] / voxel_size);
    if (loc < 0) loc = loc - 1.0f;
    kc[j] = (long long)loc;
  }
  unsigned long long key;
  if (!pack_key(kc[0], kc[1], kc[2], &key))
  {
    atomicOr(status, VN_ST_KEY_RANGE);
    slot_of[i] = -1;
    return;
  }
  unsigned int h = hash_key(key) & mask;
  for (unsigned int probe = 0; probe <= mask; probe++)
  {
    unsigned long long old = tab[h].key;
    if (old == VN_EMPTY_KEY) old = atomicCAS(&tab[h].key, VN_EMPTY_KEY, key);
    if (old == VN_EMPTY_KEY || old == key)
    {
      atomicAdd(&tab[h].sum[0], (double)q.x);
      atomicAdd(&tab[h].sum[1], (double)q.y);
      atomicAdd(&tab[h].sum[2], (double)q.z);
      atomicAdd(&tab[h].cnt, 1);
      atomicMin(&tab[h].first, i);
      slot_of[i] = (int)h;
      return;
    }
    h = (h + 1) & mask;
  }
  atomicOr(status, VN_ST_DOWN_FULL);
  slot_of[i] = -1;
}

__global__ void __launch_bounds__(256)
    k_down_accum(const float4* __restrict__ pts, int n, double voxel_size, DownSlot* __restrict__ tab, unsigned int mask,
                 int* __restrict__ slot_of, int* __restrict__ status)
{
  vn_pdl_sync();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  down_accum_point(pts[i], i, voxel_size, tab, mask, slot_of, status);
}

// The front of the per-scan step in one pass over the raw scan: a2 + a3 of the full scan and the leaf-cache reset (as
// k_deskew_var_init) plus the accumulation pass of f1 on the deskewed point the thread still holds - the
// down-sampling no longer re-reads the scan, and what is left of it is one more launch (k_down_emit_all).
__global__ void __launch_bounds__(256, 3)
    k_deskew_var_init_down(float4* __restrict__ pts, int n, const DeskewPoses* __restrict__ Pg, int* __restrict__ status,
                           ScanView out, VarInitParams prm, int* __restrict__ cache, double voxel_size,
                           DownSlot* __restrict__ tab, unsigned int mask, int* __restrict__ slot_of)
{
  __shared__ DeskewPoses P;
  stage_poses(P, Pg);
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 r = deskew_point(P, pts, i, n, status);
  pts[i] = r;
  down_accum_point(r, i, voxel_size, tab, mask, slot_of, status);
  var_init_point(r, i, out, prm);
  cache[i] = -1;
}

// The rest of f1 (first-point flags, exclusive scan, emission of the voxel means in first-point order, table
// clean-up) and the var_init of the emitted set (a3 on the map's point set, point_utils.cpp:36-52) in ONE persistent
// launch: every block owns a contiguous chunk of the scan, counts its first points, meets the others at a grid
// barrier (cooperative launch), adds the counts of the blocks before it and emits. The total goes to mapped host
// memory with a sequence number, so that the host has the count without a copy or a stream synchronisation.
#define EMIT_T 1024
#define EMIT_KEEP 2  // rounds whose flags stay in registers (two rounds cover 148 x 2048 = 303 000 points)
#define EMIT_SPIN_LIMIT (1ll << 24)
__device__ __forceinline__ int block_sum_int(int v, int* wsum)
{
  // (all EMIT_T threads; returns the block total to everybody)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) wsum[w] = v;
  __syncthreads();
  int t = wsum[lane];
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

__global__ void __launch_bounds__(EMIT_T, 1) k_down_emit_all(const __grid_constant__ DownEmit a)
{
  __shared__ int wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = gridDim.x;
  const int first = blockIdx.x * a.chunk;
  const int end = min(a.n, first + a.chunk);
  DownSlot* __restrict__ tab = a.tab;
  // phase 1: first points of this chunk (the slot of a first point stays in a register for phase 2, -1 otherwise)
  int mine = 0;
  int keep[EMIT_KEEP];
#pragma unroll
  for (int k = 0; k < EMIT_KEEP; k++)
  {
    const int i = first + k * EMIT_T + tid;
    int s = -1;
    if (i < end)
    {
      s = a.slot_of[i];
      if (s >= 0 && tab[s].first != i) s = -1;
    }
    keep[k] = s;
    mine += s >= 0 ? 1 : 0;
  }
  for (int i = first + EMIT_KEEP * EMIT_T + tid; i < end; i += EMIT_T)
  {
    const int s = a.slot_of[i];
    mine += (s >= 0 && tab[s].first == i) ? 1 : 0;
  }
  const int total = block_sum_int(mine, wsum);
  if (tid == 0)
  {
    __stcg(a.counts + blockIdx.x, total);
    __threadfence();
    atomicAdd(a.bar, 1ull);
    long long spins = 0;
    while (*reinterpret_cast<volatile unsigned long long*>(a.bar) < (unsigned long long)nblk)
      if (++spins > EMIT_SPIN_LIMIT)
      {
        atomicOr(a.status, VN_ST_SPIN);
        break;
      }
    __threadfence();
  }
  __syncthreads();
  // the counts of the blocks before this one, and of all
  int before = 0, all = 0;
  for (int b = tid; b < nblk; b += EMIT_T)
  {
    const int c = __ldcg(a.counts + b);
    all += c;
    if (b < (int)blockIdx.x) before += c;
  }
  before = block_sum_int(before, wsum);
  all = block_sum_int(all, wsum);
  // phase 2: emit in first-point order, var_init of the emitted point, clean the slot for the next scan
  int run = before;
  int rk = 0;
  for (int r0 = first; r0 < end; r0 += EMIT_T, rk++)
  {
    const int i = r0 + tid;
    int s = -1;
    bool flag = false;
    if (rk < EMIT_KEEP)
    {
#pragma unroll
      for (int k = 0; k < EMIT_KEEP; k++)
        if (k == rk) s = keep[k];
      flag = s >= 0;
    }
    else if (i < end)
    {
      s = a.slot_of[i];
      flag = s >= 0 && tab[s].first == i;
    }
    const unsigned int bal = __ballot_sync(0xffffffffu, flag);
    __syncthreads();
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int t = wsum[lane], incl = t;
    for (int o = 1; o < 32; o <<= 1)
    {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    const int warp_off = __shfl_sync(0xffffffffu, incl - t, warp);
    const int round_total = __shfl_sync(0xffffffffu, incl, 31);
    if (flag)
    {
      const int ord = run + warp_off + __popc(bal & ((1u << lane) - 1u));
      DownSlot& sl = tab[s];
      const double c = (double)sl.cnt;
      const float4 m = make_float4((float)(sl.sum[0] / c), (float)(sl.sum[1] / c), (float)(sl.sum[2] / c), (float)sl.cnt);
      a.out[ord] = m;
      sl.key = VN_EMPTY_KEY;
      sl.sum[0] = sl.sum[1] = sl.sum[2] = 0.0;
      sl.cnt = 0;
      sl.first = 0x7fffffff;
      var_init_point(m, ord, a.pv, a.prm);
    }
    run += round_total;
  }
  if (tid == 0)
  {
    if (blockIdx.x == 0)
    {
      *a.n_out_dev = all;
      a.pub[1] = (unsigned long long)all;
      __threadfence_system();
      a.pub[0] = a.seq;
    }
    const unsigned long long t = atomicAdd(a.bar + 1, 1ull);
    if (t == (unsigned long long)nblk - 1ull)
    {
      a.bar[0] = 0ull;
      a.bar[1] = 0ull;
      __threadfence();
    }
  }
}

// three-kernel exclusive scan over int flags (n <= 1024*1024); with `tab` the flags are formed here (k_down_flag's
// job: flag[i] = 1 when point i is the first point of its voxel) and written to `flag_out` for the emission
__global__ void __launch_bounds__(1024) k_scan_block(const int* __restrict__ in, int* __restrict__ out, int n,
                                                     int* __restrict__ block_sums, const DownSlot* __restrict__ tab,
                                                     int* __restrict__ flag_out)
{
  vn_pdl_sync();
  __shared__ int warp_sums[32];
  int i = blockIdx.x * 1024 + threadIdx.x;
  int v = 0;
  if (i < n)
  {
    if (tab)
    {
      const int s = in[i];  // (`in` = slot_of)
      v = (s >= 0 && tab[s].first == i) ? 1 : 0;
      flag_out[i] = v;
    }
    else
      v = in[i];
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = v;
  for (int o = 1; o < 32; o <<= 1)
  {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[w] = x;
  __syncthreads();
  if (w == 0)
  {
    int s = warp_sums[lane];
    for (int o = 1; o < 32; o <<= 1)
    {
      int y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    warp_sums[lane] = s;
  }
  __syncthreads();
  int incl = x + (w > 0 ? warp_sums[w - 1] : 0);
  if (i < n) out[i] = incl - v;
  if (threadIdx.x == 1023) block_sums[blockIdx.x] = incl;
}
// (the total also goes to mapped host memory, then the sequence number the host polls: the host has the down-sampled
// count a few microseconds after this kernel, without a copy or a stream synchronisation - it can enqueue the map
// update while the emission and the IEKF loop are still running)
__global__ void __launch_bounds__(1024) k_scan_sums(int* __restrict__ block_sums, int nb, int* __restrict__ total,
                                                    volatile unsigned long long* __restrict__ pub, unsigned long long seq)
{
  vn_pdl_sync();
  __shared__ int warp_sums[32];
  int v = (threadIdx.x < nb) ? block_sums[threadIdx.x] : 0;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = v;
  for (int o = 1; o < 32; o <<= 1)
  {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[w] = x;
  __syncthreads();
  if (w == 0)
  {
    int s = warp_sums[lane];
    for (int o = 1; o < 32; o <<= 1)
    {
      int y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    warp_sums[lane] = s;
  }
  __syncthreads();
  int incl = x + (w > 0 ? warp_sums[w - 1] : 0);
  if (threadIdx.x < nb) block_sums[threadIdx.x] = incl - v;
  if (threadIdx.x == 1023)
  {
    *total = incl;
    if (pub)
    {
      pub[1] = (unsigned long long)(unsigned int)incl;
      __threadfence_system();
      pub[0] = seq;
    }
  }
}

// emit voxel means in first-point order and clean the table slot for the next scan
__global__ void __launch_bounds__(256)
    k_down_emit(int n, DownSlot* __restrict__ tab, const int* __restrict__ slot_of, const int* __restrict__ flag,
                const int* __restrict__ scan, const int* __restrict__ block_sums, float4* __restrict__ out, ScanView pv,
                VarInitParams prm)
{
  vn_pdl_sync();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!flag[i]) return;
  int ord = scan[i] + block_sums[i >> 10];
  DownSlot& s = tab[slot_of[i]];
  double c = (double)s.cnt;
  const float4 mean = make_float4((float)(s.sum[0] / c), (float)(s.sum[1] / c), (float)(s.sum[2] / c), (float)s.cnt);
  out[ord] = mean;
  // a3 on the emitted point right away (what k_var_init would do from the stored float4): one launch less per scan
  if (pv.p[0]) var_init_point(mean, ord, pv, prm);
  s.key = VN_EMPTY_KEY;
  s.sum[0] = s.sum[1] = s.sum[2] = 0.0;
  s.cnt = 0;
  s.first = 0x7fffffff;
}

__global__ void k_down_init(DownSlot* tab, unsigned int nslots)
{
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nslots) return;
  tab[i].key = VN_EMPTY_KEY;
  tab[i].sum[0] = tab[i].sum[1] = tab[i].sum[2] = 0.0;
  tab[i].cnt = 0;
  tab[i].first = 0x7fffffff;
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
// Start-up phase: the kernels behind vina_odom_init_scan.
//
// k_init_assoc: one iteration of VINA_SLAM::lio_state_estimation_kdtree (src/pipeline/odometry.cpp:332-392) for every
// point of the down-sampled scan: world point; when `refind`, the NMATCH = 5 nearest points of the local map
// (pcl::KdTreeFLANN::nearestKSearch - an exact search, here by brute force over the map tiled through shared
// memory: the map is a 0.5 m grid of a few 10^4 points and this runs for the first win_size scans only), the
// plane n.x + 1 = 0 through them by least squares (A.colPivHouseholderQr().solve(b), 5 x 3; normal equations here)
// and the acceptance test |n.a_k + 1| <= 0.1; then the point-to-plane terms HTH += j j^T, HTz -= j d, summed per
// block in a fixed order (28 partials per block, added up by k_init_sum).
#define INIT_NMATCH 5
#define INIT_THREADS 128
#define INIT_TILE 1024
#define INIT_NSUM 28  // 21 (HTH upper) + 6 (HTz) + 1 (valid)

__device__ __forceinline__ bool solve3_lu(const double* A, const double* b, double* x)
{
  // x = A^-1 b for a 3x3 (column-major) by LU with partial pivoting (Eigen's inverse() route for the 3 x 3 normal
  // equations is a closed form; the fit is toleranced)
  double m[3][4];
  for (int i = 0; i < 3; i++)
  {
    for (int j = 0; j < 3; j++) m[i][j] = A[i + 3 * j];
    m[i][3] = b[i];
  }
  for (int k = 0; k < 3; k++)
  {
    int piv = k;
    double best = fabs(m[k][k]);
    for (int i = k + 1; i < 3; i++)
      if (fabs(m[i][k]) > best)
      {
        best = fabs(m[i][k]);
        piv = i;
      }
    if (best == 0.0) return false;
    if (piv != k)
      for (int j = 0; j < 4; j++)
      {
        const double t = m[k][j];
        m[k][j] = m[piv][j];
        m[piv][j] = t;
      }
    for (int i = k + 1; i < 3; i++)
    {
      const double f = m[i][k] / m[k][k];
      for (int j = k; j < 4; j++) m[i][j] = m[i][j] - f * m[k][j];
    }
  }
  for (int i = 2; i >= 0; i--)
  {
    double sacc = m[i][3];
    for (int j = i + 1; j < 3; j++) sacc = sacc - m[i][j] * x[j];
    x[i] = sacc / m[i][i];
  }
  return true;
}

__global__ void __launch_bounds__(INIT_THREADS)
    k_init_assoc(ScanView pv, int n, PoseD x, const float4* __restrict__ tree, int n_tree, int refind,
                 double* __restrict__ ds, double* __restrict__ dir, double* __restrict__ partial)
{
  __shared__ float4 tile[INIT_TILE];
  __shared__ double red[INIT_THREADS / 32][INIT_NSUM];
  const int i = blockIdx.x * INIT_THREADS + threadIdx.x;
  const bool live = i < n;
  double pnt[3] = { 0, 0, 0 }, wld[3] = { 0, 0, 0 };
  if (live)
  {
    for (int k = 0; k < 3; k++) pnt[k] = pv.p[k][i];
    rot_trans(x.R, x.p, pnt, wld);
  }
  double d_i = -1.0, n_i[3] = { 0, 0, 0 };
  if (refind)
  {
    const float qx = (float)wld[0], qy = (float)wld[1], qz = (float)wld[2];
    float bd[INIT_NMATCH];
    int bi[INIT_NMATCH];
#pragma unroll
    for (int k = 0; k < INIT_NMATCH; k++) bd[k] = 3.0e38f, bi[k] = -1;
    for (int base = 0; base < n_tree; base += INIT_TILE)
    {
      __syncthreads();
      for (int t = threadIdx.x; t < INIT_TILE; t += INIT_THREADS)
        if (base + t < n_tree) tile[t] = tree[base + t];
      __syncthreads();
      const int m = min(INIT_TILE, n_tree - base);
      if (live)
        for (int t = 0; t < m; t++)
        {
          const float4 p = tile[t];
          const float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
          const float d2 = (dx * dx + dy * dy) + dz * dz;
          if (d2 < bd[INIT_NMATCH - 1])
          {
            // sorted insertion; candidates arrive in ascending index, so equal distances keep the smaller index first
            float cd = d2;
            int ci = base + t;
#pragma unroll
            for (int k = 0; k < INIT_NMATCH; k++)
              if (cd < bd[k])
              {
                const float td = bd[k];
                const int ti = bi[k];
                bd[k] = cd;
                bi[k] = ci;
                cd = td;
                ci = ti;
              }
          }
        }
    }
    if (live && bi[INIT_NMATCH - 1] >= 0)
    {
      double A[INIT_NMATCH][3];
#pragma unroll
      for (int k = 0; k < INIT_NMATCH; k++)
      {
        const float4 p = tree[bi[k]];
        A[k][0] = (double)p.x;
        A[k][1] = (double)p.y;
        A[k][2] = (double)p.z;
      }
      double AtA[9], Atb[3];
      for (int c = 0; c < 3; c++)
      {
        for (int r = 0; r < 3; r++)
        {
          double sacc = A[0][r] * A[0][c];
          for (int k = 1; k < INIT_NMATCH; k++) sacc = sacc + A[k][r] * A[k][c];
          AtA[r + 3 * c] = sacc;
        }
        double sb = A[0][c] * -1.0;
        for (int k = 1; k < INIT_NMATCH; k++) sb = sb + A[k][c] * -1.0;
        Atb[c] = sb;
      }
      double direct[3];
      bool ok = solve3_lu(AtA, Atb, direct);
      if (ok)
        for (int k = 0; k < INIT_NMATCH; k++)
        {
          const double v = (direct[0] * A[k][0] + direct[1] * A[k][1]) + direct[2] * A[k][2];
          if (!(fabs(v + 1.0) <= 0.1)) ok = false;
        }
      if (ok)
      {
        const double nn = sqrt((direct[0] * direct[0] + direct[1] * direct[1]) + direct[2] * direct[2]);
        d_i = 1.0 / nn;
        for (int k = 0; k < 3; k++) n_i[k] = direct[k] * d_i;
      }
    }
    if (live)
    {
      ds[i] = d_i;
      for (int k = 0; k < 3; k++) dir[3 * (size_t)i + k] = n_i[k];
    }
  }
  else if (live)
  {
    d_i = ds[i];
    for (int k = 0; k < 3; k++) n_i[k] = dir[3 * (size_t)i + k];
  }
  // point-to-plane terms (odometry.cpp:378-391): jac = [hat(pnt) R^T n ; n], HTH += jac jac^T, HTz += jac * (-pd2)
  double acc[INIT_NSUM];
#pragma unroll
  for (int k = 0; k < INIT_NSUM; k++) acc[k] = 0.0;
  if (live && d_i >= 0)
  {
    const double pd2 = ((n_i[0] * wld[0] + n_i[1] * wld[1]) + n_i[2] * wld[2]) + d_i;
    double m[3];
    rotT_vec(x.R, n_i, m);
    const double j[6] = { pnt[1] * m[2] - pnt[2] * m[1], pnt[2] * m[0] - pnt[0] * m[2], pnt[0] * m[1] - pnt[1] * m[0],
                          n_i[0], n_i[1], n_i[2] };
    int t = 0;
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
      for (int b = a; b < 6; b++) acc[t++] = j[a] * j[b];
#pragma unroll
    for (int a = 0; a < 6; a++) acc[21 + a] = j[a] * (-pd2);
    acc[27] = 1.0;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < INIT_NSUM; k++)
  {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < INIT_NSUM)
  {
    double v = 0.0;
    for (int w = 0; w < INIT_THREADS / 32; w++) v += red[w][threadIdx.x];
    partial[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = v;
  }
}

// the blocks' partials in block order -> out[28] (mapped host memory)
__global__ void __launch_bounds__(32 * INIT_NSUM) k_init_sum(const double* __restrict__ partial, int nblocks, double* __restrict__ out)
{
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double v = 0.0;
  for (int b = lane; b < nblocks; b += 32) v += partial[(size_t)k * nblocks + b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) out[k] = v;
}

// pl_tree->push_back(R pnt + p) for the whole scan (odometry.cpp:276-284, 429-436)
__global__ void __launch_bounds__(256) k_init_tree_append(ScanView pv, int n, PoseD x, float4* __restrict__ tree_tail)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double pnt[3] = { pv.p[0][i], pv.p[1][i], pv.p[2][i] };
  double w[3];
  rot_trans(x.R, x.p, pnt, w);
  tree_tail[i] = make_float4((float)w[0], (float)w[1], (float)w[2], 0.0f);
}

// Initialization::motion_blur (src/pipeline/initialization.cpp:64-156) + what motion_init does with every point
// right after (:222-241): the frame's retained raw points (time-sorted), compensated with the pose table of a
// BACKWARD integration from the frame's end state into the end IMU frame; then either the body covariance
// (calcBodyVar on that point, the z == 0 mutation included) and pvec_update with the frame's state, or - before
// the first convergence - the identity covariance and the plain world point. Output in the reference's order:
// the scan is walked from its last point to its first (index n - 1 - i), the points at or before the earliest pose
// are skipped, and point 0 is pushed once more for every further pose once the walk has reached begin()
// (:150-153: the inner loop breaks there, the outer loop goes on). The poses are ordered as the reference builds
// them: latest first.
__global__ void __launch_bounds__(256)
    k_init_redeskew(const float4* __restrict__ orig, int n, int n_skip, const DeskewPoses* __restrict__ Pg, PoseD x,
                    Cov2x cv, int converged, VarInitParams prm, ScanView out, InsertScratch sc)
{
  __shared__ DeskewPoses P;
  stage_poses(P, Pg);
  __syncthreads();
  const int i = n_skip + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = orig[i];
  const double cvt = (double)q.w;
  // first pose (latest first) with t < curvature
  int lo = 0, hi = P.m;
  while (lo < hi)
  {
    const int mid = (lo + hi) >> 1;
    if (P.pose[mid].t < cvt)
      hi = mid;
    else
      lo = mid + 1;
  }
  const int k0 = lo;
  const int reps = (i == 0) ? (P.m - k0) : 1;  // point 0 again for every later pose (all of them are earlier in time)
  for (int rep = 0; rep < reps; rep++)
  {
    const vina_imu_pose& h = P.pose[k0 + rep];
    const double dt = cvt - h.t;
    double E[9], Ri[9];
    exp_so3_dt(h.w, dt, E);
    for (int c = 0; c < 3; c++)
      for (int r = 0; r < 3; r++) Ri[r + 3 * c] = h.R[r] * E[3 * c] + h.R[r + 3] * E[3 * c + 1] + h.R[r + 6] * E[3 * c + 2];
    double T[3];
    for (int r = 0; r < 3; r++) T[r] = ((h.p[r] + h.v[r] * dt) + ((0.5 * h.a[r]) * dt) * dt) - P.p_end[r];
    const double Pi[3] = { (double)q.x, (double)q.y, (double)q.z };
    double a[3], b[3], pnt[3];
    for (int r = 0; r < 3; r++) a[r] = ((P.ext_R[r] * Pi[0] + P.ext_R[r + 3] * Pi[1]) + P.ext_R[r + 6] * Pi[2]) + P.ext_t[r];
    for (int r = 0; r < 3; r++) b[r] = ((Ri[r] * a[0] + Ri[r + 3] * a[1]) + Ri[r + 6] * a[2]) + T[r];
    for (int r = 0; r < 3; r++) pnt[r] = (P.R_end[3 * r] * b[0] + P.R_end[3 * r + 1] * b[1]) + P.R_end[3 * r + 2] * b[2];
    double v6[6] = { 1.0, 0.0, 0.0, 1.0, 0.0, 1.0 };
    double pw[3];
    if (converged)
    {
      double var[9], b6[6];
      calc_body_var(pnt, prm.range_var, prm.dir_var, var);
      b6[0] = var[0], b6[1] = var[3], b6[2] = var[6], b6[3] = var[4], b6[4] = var[7], b6[5] = var[8];
      world_var(x.R, pnt, b6, cv.rot, cv.tsl, v6);
    }
    rot_trans(x.R, x.p, pnt, pw);
    const int j = rep == 0 ? (n - 1 - i) : (n - n_skip) + (rep - 1);
    for (int r = 0; r < 3; r++)
    {
      out.p[r][j] = pnt[r];
      sc.pw[r][j] = pw[r];
    }
    for (int r = 0; r < 6; r++) sc.vw[r][j] = v6[r];
  }
}

int launch_init_assoc(cudaStream_t st, const ScanView& pv, int n, const PoseD& x, const float4* tree, int n_tree, int refind,
                      double* ds, double* dir, double* partial, double* out28)
{
  if (n <= 0) return 0;
  const int nb = (n + INIT_THREADS - 1) / INIT_THREADS;
  k_init_assoc<<<nb, INIT_THREADS, 0, st>>>(pv, n, x, tree, n_tree, refind, ds, dir, partial);
  k_init_sum<<<1, 32 * INIT_NSUM, 0, st>>>(partial, nb, out28);
  return 2;
}
int launch_init_tree_append(cudaStream_t st, const ScanView& pv, int n, const PoseD& x, float4* tree_tail)
{
  if (n <= 0) return 0;
  k_init_tree_append<<<(n + 255) / 256, 256, 0, st>>>(pv, n, x, tree_tail);
  return 1;
}
int launch_init_redeskew(cudaStream_t st, const float4* orig, int n, int n_skip, const DeskewPoses* d_poses, const PoseD& x,
                         const double* rot_var, const double* tsl_var, int converged, const VarInitParams& prm,
                         const ScanView& out, const InsertScratch& sc)
{
  if (n - n_skip <= 0) return 0;
  Cov2x cv;
  for (int k = 0; k < 9; k++) cv.rot[k] = rot_var[k], cv.tsl[k] = tsl_var[k];
  k_init_redeskew<<<(n - n_skip + 255) / 256, 256, 0, st>>>(orig, n, n_skip, d_poses, x, cv, converged, prm, out, sc);
  return 1;
}

void launch_deskew(cudaStream_t st, float4* pts, int n, const DeskewPoses* d_poses, int* status)
{
  if (n <= 0) return;
  k_deskew<<<(n + 255) / 256, 256, 0, st>>>(pts, n, d_poses, status);
}
void launch_deskew_var_init(cudaStream_t st, float4* pts, int n, const DeskewPoses* d_poses, int* status, ScanView out,
                            const VarInitParams& prm, int* cache, int first, int last)
{
  if (last > first)
    vn_launch(k_deskew_var_init, dim3((last - first + 255) / 256), dim3(256), 0, st, pts, n, d_poses, status, out, prm, cache,
              first, last);
}
void launch_deskew_var_init_down(cudaStream_t st, float4* pts, int n, const DeskewPoses* d_poses, int* status, ScanView out,
                                 const VarInitParams& prm, int* cache, double voxel_size, DownSlot* tab, unsigned int mask,
                                 int* slot_of)
{
  if (n > 0)
    k_deskew_var_init_down<<<(n + 255) / 256, 256, 0, st>>>(pts, n, d_poses, status, out, prm, cache, voxel_size, tab, mask,
                                                           slot_of);
}
int launch_down_emit_all(cudaStream_t st, DownEmit& a, int sm_count)
{
  int blocks = (a.n + EMIT_T - 1) / EMIT_T;
  if (blocks > sm_count) blocks = sm_count;
  if (blocks < 1) blocks = 1;
  a.chunk = (((a.n + blocks - 1) / blocks) + 31) & ~31;
  void* args[] = { &a };
  return (int)cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(k_down_emit_all), dim3(blocks), dim3(EMIT_T), args, 0, st);
}
void launch_var_init(cudaStream_t st, const float4* pts, const int* n_dev, int n_host, ScanView out,
                     const VarInitParams& prm)
{
  if (n_host <= 0) return;
  vn_launch(k_var_init, dim3((n_host + 255) / 256), dim3(256), 0, st, pts, n_dev, n_host, out, prm);
}
void launch_down_init(cudaStream_t st, DownSlot* tab, unsigned int nslots)
{
  k_down_init<<<(nslots + 255) / 256, 256, 0, st>>>(tab, nslots);
}
int launch_downsample(cudaStream_t st, const float4* pts, int n, double voxel_size, DownSlot* tab, unsigned int mask,
                      int* slot_of, int* flag, int* scan, int* block_sums, int* n_out_dev, float4* out, int* status,
                      unsigned long long* pub, unsigned long long seq, const ScanView* pv, const VarInitParams* prm)
{
  if (n <= 0) return 0;
  int nb = (n + 1023) / 1024;
  if (nb > 1024) return -1;
  vn_launch(k_down_accum, dim3((n + 255) / 256), dim3(256), 0, st, pts, n, voxel_size, tab, mask, slot_of, status);
  ScanView none;
  memset(&none, 0, sizeof(none));
  VarInitParams pz;
  memset(&pz, 0, sizeof(pz));
  vn_launch(k_scan_block, dim3(nb), dim3(1024), 0, st, slot_of, scan, n, block_sums, tab, flag);
  vn_launch(k_scan_sums, dim3(1), dim3(1024), 0, st, block_sums, nb, n_out_dev, pub, seq);
  vn_launch(k_down_emit, dim3((n + 255) / 256), dim3(256), 0, st, n, tab, slot_of, flag, scan, block_sums, out, pv ? *pv : none,
            prm ? *prm : pz);
  return 4;
}

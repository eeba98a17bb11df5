// The hot kernel: one IEKF iteration of VINA_SLAM::LioStateEstimation
// (src/pipeline/odometry.cpp:98-231) for one or several independent sequences:
//   * the point loop (odometry.cpp:111-148): world point, cached-leaf test (OctoTree::inside,
//     octree.cpp:732-737), voxel-hash lookup (match, voxel_map.cpp:241-266), octree descent and gate
//     (OctoTree::match, octree.cpp:551-595), residual / Jacobian and the H = J^T R^-1 J, b, n n^T sums
//     (odometry.cpp:136-146);
//   * optionally (VN_IEKF_SOLVE) the update itself (odometry.cpp:192-230): K = (blkdiag(H,0) + P^-1)^-1,
//     delta, boxplus, convergence / rematch logic and the final P = (I - G) P, done by one warp of the block
//     that finishes last - the iterate lives in device memory (IekfDev) and the host is not in the loop.
//
// Roofline: HBM-bound streaming of the pointVar SoA (72 B/pt + 8 B cache RMW) plus gathers of 16-B hash slots
// and 256-B leaf records; the dependent chain cache -> leaf record (or key -> slot -> root -> child -> leaf)
// makes it latency-bound unless enough points are in flight, so the kernel is shaped for occupancy: one
// persistent 768-thread block per SM, one point per thread per round, 80 registers. A cached leaf costs one
// gather (line 0 of its record: box, centre, normal, radius) before the fp32 gate and one more (line 1) for
// sigma_l; the next round's SoA rows are prefetched into L2 while the current round computes.
//
// Reduction: the 6x6 / 6 / 3x3 sums are a rank-1 update per point, C += a_i b_i^T with
//   a = [Rinv*j (6), n0, n1],  b = [j (6), r, 0],   j = [p x R^T n ; n]
// so that C(0:6,0:6) = H, C(0:6,6) = -b, C(6:8,3:6) = rows 0,1 of n n^T. The warp hands its 32 (a, b) pairs to
// the FP64 tensor pipe: 8 x mma.sync.m8n8k4.f64 (DMMA, 4 points each) per round, operands staged through a
// per-warp shared-memory tile - 37 TFLOP/s on B200 (scripts/fp64_pipes.cu), the same rate as the vector DFMA
// pipe but ~40 issue slots per round instead of the ~700 of a shuffle tree. n2*n2 and the match count are
// per-thread accumulators reduced once at the end. Everything is summed in a fixed order (DMMA chain, warps in
// order, blocks in order): results are deterministic run to run.
//
// Numerics: every decision-bearing expression (wld, key, child index, inside, the fp32 gate) uses the
// single-rounding helpers of vn_math.cuh in the order of SURVEY.md Appendix A, so keys and associations are
// bit-exact against the CPU restatement. sigma_l and the sums use FMA freely (tolerance 1e-4 rel):
// n^T var_world n is evaluated as (R^T n)^T var (R^T n) + (n x p)^T S_R (n x p) + n^T S_t n, which needs ~1/4
// of the flops of forming var_world.
#include <cstdio>
#include "vn_kernels.cuh"

#ifndef IEKF_THREADS
#define IEKF_THREADS 768  // 80 registers per thread; 1024 x 64 spills and is ~7 % slower (profiles/r01_iekf_block_sweep.txt)
#endif
#ifndef IEKF_BLOCKS_PER_SM
#define IEKF_BLOCKS_PER_SM 1
#endif
#define IEKF_WARPS (IEKF_THREADS / 32)
#define IEKF_ROWS 14                        // staged rows per warp: j0..j5, r, 0, Rinv*j0..Rinv*j5
#define IEKF_LD 36                          // row stride (doubles): 32 points + 4 -> conflict-free fragment loads
#define IEKF_TILE (IEKF_ROWS * IEKF_LD)     // doubles per warp
#define IEKF_SMEM (IEKF_WARPS * IEKF_TILE * sizeof(double))

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// where the k-th packed sum (21 HTH upper by rows, 6 HTz, 6 nnt upper, count) sits among the 66 per-warp
// values: 64 entries of C (row-major 8x8), then n2*n2, then the match count
__device__ const signed char IEKF_SRC[VN_IEKF_NACC] = { 0,  1,  2,  3,  4,  5,  9,  10, 11, 12, 13, 18,
                                                        19, 20, 21, 27, 28, 29, 36, 37, 45, 6,  14, 22,
                                                        30, 38, 46, 51, 52, 53, 60, 61, 64, 65 };
__device__ __forceinline__ int iekf_src(int k) { return IEKF_SRC[k]; }

// ---------------------------------------------------------------------------
// a7 on the device: one warp. The reference forms K = (blkdiag(H, 0) + P^-1)^-1 with two 15x15 inversions
// (odometry.cpp:82, 194) and then only uses K(:, 0:6). With E = [I6 0] the push-through identity gives
//   K(:, 0:6) = P(:, 0:6) (I6 + H P66)^-1,
// the same quantity from one well-conditioned 6x6 inversion and without P^-1 (the host variant in
// host/vina_pipeline.cpp keeps the reference's route; both agree to ~1e-10 relative).
// Lane j < 6 holds column j of a 6x6 matrix, lanes 6..11 the columns of the identity; Gauss-Jordan
// elimination with partial (row) pivoting turns the second half into the inverse.
__device__ __forceinline__ void warp_inverse6(double (&c)[6], int lane)
{
  const unsigned int F = 0xffffffffu;
#pragma unroll
  for (int k = 0; k < 6; k++)
  {
    int piv = k;
    double best = fabs(c[k]);
#pragma unroll
    for (int i = k + 1; i < 6; i++)
    {
      const double a = fabs(c[i]);
      if (a > best)
      {
        best = a;
        piv = i;
      }
    }
    piv = __shfl_sync(F, piv, k);
#pragma unroll
    for (int i = k + 1; i < 6; i++)
      if (i == piv)
      {
        const double t = c[k];
        c[k] = c[i];
        c[i] = t;
      }
    const double pkk = __shfl_sync(F, c[k], k);
    const double rk = c[k] / pkk;
    c[k] = rk;
#pragma unroll
    for (int i = 0; i < 6; i++)
      if (i != k)
      {
        const double f = __shfl_sync(F, c[i], k);
        c[i] = fma(-f, rk, c[i]);
      }
  }
}

// Log(R), math.hpp:43-48
__device__ void so3_log(const double* R, double* w)
{
  const double tr = (R[0] + R[4]) + R[8];
  const double theta = (tr > 3.0 - 1e-6) ? 0.0 : acos(0.5 * (tr - 1));
  const double K[3] = { R[5] - R[7], R[6] - R[2], R[1] - R[3] };
  const double f = (fabs(theta) < 0.001) ? 0.5 : (0.5 * theta / sin(theta));
  for (int i = 0; i < 3; i++) w[i] = f * K[i];
}

// The update of one iteration (odometry.cpp:192-230, types.hpp:67-86), by a whole thread block. fin = the 34
// packed sums, ws = IEKF_SOLVE_WS doubles of scratch, cov = the 15x15 prior covariance (replaced by the posterior
// when the loop ends), st = the first 42 doubles of IekfDev (x_curr: R p v bg ba, then x_prop), ctl = {iter, rematch,
// done, max_iter} - all in SHARED memory, so that nothing here waits on a global load. Every thread of the block
// calls it (it synchronises the block); needs >= 64 threads.
//
// A single warp walking through this took ~20 us per iteration on B200 (dependent fp64 operations cost 23-33 cycles
// each there, a division 123, and sin / cos / acos are long chains of them) - more than the point loop it follows.
// So the independent parts run side by side and the products are re-associated around the 6x6 core:
//   T = (I6 + H P66)^-1                              warp 0, Gauss-Jordan        | vec = x_prop (-) x_curr   warp 1
//   TH = T H (36 threads),  w = T HTz (6 threads)
//   u = w - TH vec(0:6)                              6 threads
//   solution = P(:, 0:6) u + vec  (15 threads)       = K HTz + vec - G vec(0:6) with K = P(:,0:6) T, G = K H
//   G(:, 0:6) = P(:, 0:6) TH      (90 threads, next to the solution)
//   x_curr (+)= solution: Exp() with sin, 1 - cos and the axis on different lanes | P - G P(0:6, :) into a spare
//   buffer (225 threads), committed only if this iteration ends the loop
#define IEKF_SOLVE_WS 536
__device__ __forceinline__ void iekf_solve_cta(int* ctl, const double* fin, double* ws, double* cov, double* st, int tid,
                                               int nthreads)
{
  double* HTH = ws;         // 6x6 column-major (symmetric)
  double* HTz = ws + 36;    // 6
  double* T6 = ws + 48;     // (I + H P66)^-1, column-major
  double* TH = ws + 96;     // T H, column-major
  double* wv = ws + 132;    // T HTz
  double* uv = ws + 138;    // w - TH vec(0:6)
  double* vec = ws + 144;   // x_prop (-) x_curr (15)
  double* sol = ws + 160;   // (15)
  double* G6 = ws + 176;    // G(:, 0:6), 15x6 column-major
  double* covn = ws + 272;  // candidate posterior (225)
  double* M9 = ws + 500;    // R^T Rp, later Exp(solution(0:3))
  double* tmp = ws + 512;   // hat(axis) (9)
  int* flg = reinterpret_cast<int*>(ws + 528);
  double *R = st, *p = st + 9;
  const double* Rp = st + 21;
  const int lane = tid & 31, warp = tid >> 5;
  const unsigned int F = 0xffffffffu;

  // ---- phase 0: unpack the sums; M = R^T Rp
  if (tid < 36)
  {
    const int i = tid % 6, j = tid / 6;
    const int a_ = i < j ? i : j, b_ = i < j ? j : i;
    HTH[tid] = fin[a_ * 6 - (a_ * (a_ - 1)) / 2 + (b_ - a_)];
  }
  else if (tid < 42)
    HTz[tid - 36] = fin[21 + (tid - 36)];
  else if (tid >= 64 - 9 && tid < 64)
  {
    const int e = tid - (64 - 9), i = e % 3, j = e / 3;  // M(i, j) = sum_k R(k, i) Rp(k, j)
    M9[e] = da(da(dm(R[3 * i], Rp[3 * j]), dm(R[3 * i + 1], Rp[3 * j + 1])), dm(R[3 * i + 2], Rp[3 * j + 2]));
  }
  __syncthreads();
  // ---- phase 1: warp 0 inverts, warp 1 forms x_prop (-) x_curr (types.hpp:77-86)
  if (warp == 0)
  {
    // lane j < 6 holds column j of I6 + H P66, lanes 6..11 the columns of the identity
    double c[6];
#pragma unroll
    for (int i = 0; i < 6; i++)
    {
      double x = (lane - 6 == i) ? 1.0 : 0.0;
      if (lane < 6)
      {
        x = (lane == i) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) x = fma(HTH[i + 6 * k], cov[k + 15 * lane], x);
      }
      c[i] = x;
    }
    warp_inverse6(c, lane);
    if (lane >= 6 && lane < 12)
#pragma unroll
      for (int i = 0; i < 6; i++) T6[i + 6 * (lane - 6)] = c[i];
  }
  else if (warp == 1)
  {
    if (lane == 0) so3_log(M9, vec);
    if (lane >= 3 && lane < 15) vec[lane] = st[21 + 6 + lane] - st[6 + lane];  // p, v, bg, ba: x_prop - x_curr
  }
  __syncthreads();
  // ---- phase 2: TH = T H, w = T HTz
  if (tid < 36)
  {
    const int a_ = tid % 6, b_ = tid / 6;
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) v = fma(T6[a_ + 6 * k], HTH[k + 6 * b_], v);
    TH[tid] = v;
  }
  else if (tid < 42)
  {
    const int a_ = tid - 36;
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) v = fma(T6[a_ + 6 * k], HTz[k], v);
    wv[a_] = v;
  }
  __syncthreads();
  // ---- phase 3: u = w - TH vec(0:6)
  if (tid < 6)
  {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) v = fma(TH[tid + 6 * k], vec[k], v);
    uv[tid] = wv[tid] - v;
  }
  __syncthreads();
  // ---- phase 4: solution = P(:, 0:6) u + vec;  G(:, 0:6) = P(:, 0:6) TH
  if (tid < 15)
  {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) v = fma(cov[tid + 15 * k], uv[k], v);
    sol[tid] = v + vec[tid];
  }
  else if (tid >= 32 && tid < 32 + 90)
  {
    const int e = tid - 32, i = e % 15, b_ = e / 15;
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) v = fma(cov[i + 15 * k], TH[k + 6 * b_], v);
    G6[e] = v;
  }
  __syncthreads();
  // ---- phase 5: x_curr (+)= solution (types.hpp:67-75, math.hpp:12-24) on warp 0; the candidate posterior
  // P - G P(0:6, :) (odometry.cpp:223: only the first 6 columns of G are non-zero) on the other warps
  if (warp == 0)
  {
    const double s0 = sol[0], s1 = sol[1], s2 = sol[2];
    const double nrm = sqrt(s0 * s0 + s1 * s1 + s2 * s2);
    double val = 0.0;
    if (lane == 0) val = sin(nrm);
    if (lane == 1) val = 1.0 - cos(nrm);
    if (lane >= 2 && lane < 5) val = sol[lane - 2] / nrm;
    if (lane == 5) val = sqrt(sol[3] * sol[3] + sol[4] * sol[4] + sol[5] * sol[5]);
    const double sn = __shfl_sync(F, val, 0), c1 = __shfl_sync(F, val, 1);
    const double ax[3] = { __shfl_sync(F, val, 2), __shfl_sync(F, val, 3), __shfl_sync(F, val, 4) };
    const double nt = __shfl_sync(F, val, 5);
    // E = I + sin K + (1 - cos) K K, K = hat(axis); lane e < 9 owns E(i, j)
    double Ee = 0.0;
    double* K = tmp;  // hat(axis), column-major, in shared memory (indexed by row / column below)
    if (lane < 9)
    {
      const double kv = lane == 1 ? ax[2] : lane == 2 ? -ax[1] : lane == 3 ? -ax[2] : lane == 5 ? ax[0] : lane == 6 ? ax[1]
                                                                                                      : lane == 7 ? -ax[0] : 0.0;
      K[lane] = kv;
    }
    __syncwarp();
    if (lane < 9)
    {
      const int i = lane % 3, j = lane / 3;
      double kk = 0.0;
      if (nrm >= 1e-9)
      {
        kk = da(da(dm(c1 * K[i], K[3 * j]), dm(c1 * K[i + 3], K[3 * j + 1])), dm(c1 * K[i + 6], K[3 * j + 2]));
        Ee = ((i == j ? 1.0 : 0.0) + sn * K[lane]) + kk;
      }
      else
        Ee = (i == j) ? 1.0 : 0.0;
      M9[lane] = Ee;
    }
    __syncwarp();
    double Rn = 0.0;
    if (lane < 9)
    {
      const int i = lane % 3, j = lane / 3;
      Rn = da(da(dm(R[i], M9[3 * j]), dm(R[i + 3], M9[3 * j + 1])), dm(R[i + 6], M9[3 * j + 2]));
    }
    __syncwarp();
    if (lane < 9) R[lane] = Rn;
    if (lane >= 9 && lane < 21) p[lane - 9] += sol[3 + (lane - 9)];  // p, v, bg, ba are contiguous behind R
    if (lane == 0)
    {
      const int iter = ctl[0], max_iter = ctl[3];
      const bool conv = (nrm * 57.3 < 0.01) && (nt * 100 < 0.015);
      int rematch = ctl[1];
      if (conv || (rematch == 0 && iter == max_iter - 2)) rematch++;
      const int fin_it = (rematch >= 2 || iter == max_iter - 1) ? 1 : 0;
      ctl[1] = rematch;
      ctl[0] = iter + 1;
      ctl[2] = fin_it;
      flg[0] = fin_it;
    }
  }
  else
  {
    for (int e = tid - 32; e < 225; e += nthreads - 32)
    {
      const int i = e % 15, j = e / 15;
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < 6; k++) v = fma(G6[i + 15 * k], cov[k + 15 * j], v);
      covn[e] = cov[e] - v;
    }
  }
  __syncthreads();
  if (flg[0])
    for (int e = tid; e < 225; e += nthreads) cov[e] = covn[e];
  __syncthreads();
}

// a7 with a whole block: stage the iterate in shared memory (`smem` >= IEKF_SOLVE_WS + 280 doubles), solve, write back
__device__ __forceinline__ void iekf_solve_block(IekfDev* dev, const double* fin, double* smem, int nthreads)
{
  double* ws = smem;
  double* s_cov = smem + IEKF_SOLVE_WS;
  double* s_st = smem + IEKF_SOLVE_WS + 232;
  int* s_ctl = reinterpret_cast<int*>(smem + IEKF_SOLVE_WS + 232 + 42);
  for (int i = threadIdx.x; i < 225; i += nthreads) s_cov[i] = dev->cov[i];
  if (threadIdx.x < 42) s_st[threadIdx.x] = reinterpret_cast<const double*>(dev)[threadIdx.x];
  if (threadIdx.x >= 64 && threadIdx.x < 68) s_ctl[threadIdx.x - 64] = (&dev->iter)[threadIdx.x - 64];
  __syncthreads();
  iekf_solve_cta(s_ctl, fin, ws, s_cov, s_st, threadIdx.x, nthreads);
  if (threadIdx.x < 21) reinterpret_cast<double*>(dev)[threadIdx.x] = s_st[threadIdx.x];
  if (threadIdx.x >= 64 && threadIdx.x < 67) (&dev->iter)[threadIdx.x - 64] = s_ctl[threadIdx.x - 64];
  if (reinterpret_cast<const int*>(ws + 528)[0])
    for (int i = threadIdx.x; i < 225; i += nthreads) dev->cov[i] = s_cov[i];
}

// ---------------------------------------------------------------------------
template <bool DEBUG>
__global__ void __launch_bounds__(IEKF_THREADS, IEKF_BLOCKS_PER_SM) k_iekf(const __grid_constant__ IekfBatch bt)
{
  vn_pdl_sync();  // (does nothing unless the launch carries the programmatic attribute: launch_iekf, `pdl`)
  const IekfSeq& q = bt.s[blockIdx.y];
  IekfDev* __restrict__ dev = q.dev;
  if ((bt.mode & (VN_IEKF_SOLVE | VN_IEKF_GATED)) && dev->done) return;  // converged earlier (uniform over the grid)

  extern __shared__ double smem[];
  __shared__ double cR[9], cp[3], crv[9], ctv[9];
  __shared__ double fin[VN_IEKF_NACC];
  __shared__ bool is_last;
  if (threadIdx.x < 9)
  {
    cR[threadIdx.x] = dev->R[threadIdx.x];
    crv[threadIdx.x] = dev->rot_var[threadIdx.x];
    ctv[threadIdx.x] = dev->tsl_var[threadIdx.x];
  }
  if (threadIdx.x < 3) cp[threadIdx.x] = dev->p[threadIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* Sw = smem + warp * IEKF_TILE;
  Sw[7 * IEKF_LD + lane] = 0.0;  // b7 = 0, never overwritten
  __syncthreads();

  const int n = q.n_ptr ? *q.n_ptr : q.n_host;
  const double* __restrict__ pv = q.pv_base;  // 9 contiguous arrays: p[3], v[6]
  const size_t pvs = (size_t)q.pv_stride;
  const NodeHot* __restrict__ hot = q.hot;
  int* __restrict__ cache = q.cache;

  // DMMA fragment coordinates: this lane holds C[g][2t], C[g][2t+1]; feeds A[g][t] and B[t][g]
  const int g = lane >> 2, t4 = lane & 3;
  const double* fa = Sw + (g < 6 ? 8 + g : g - 3) * IEKF_LD + t4;  // a = [Rinv j0..j5, n0 (= j3), n1 (= j4)]
  const double* fb = Sw + g * IEKF_LD + t4;                        // b = [j0..j5, r, 0]
  double c0 = 0.0, c1 = 0.0, s22 = 0.0;
  int cnt = 0;

  const int stride = gridDim.x * IEKF_THREADS;
  const int n_round = ((n + stride - 1) / stride) * stride;  // whole warps stay in the loop together
  for (int i = blockIdx.x * IEKF_THREADS + threadIdx.x; i < n_round; i += stride)
  {
    const bool live = i < n;
    const int ii = live ? i : 0;
    // the next round of this warp: pull its 9 SoA rows (2 x 128 B each) and the cache row towards L2 now
    if (!(bt.variant & 32))
    {
      const int nx = i - lane + stride;  // first point of the warp's next round
      if (nx < n && lane < 18)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pv + (size_t)(lane >> 1) * pvs + nx + 16 * (lane & 1)));
      else if (nx < n && lane == 18)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(cache + nx));
    }
    // the leaf this thread's NEXT point is cached on: pull both lines of its record towards L1 now, a whole round
    // ahead of their use (the record is two dependent gathers otherwise)
    if (!(bt.mode & VN_IEKF_NOCACHE) && i + stride < n)
    {
      const int nc = cache[i + stride];
      if (nc >= 0)
      {
        const char* rec = reinterpret_cast<const char*>(hot + nc);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(rec));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + 128));
      }
    }
    // independent loads first (memory-level parallelism): point, covariance, cached leaf
    const double* __restrict__ pi = pv + ii;
    const double pnt[3] = { __ldg(pi), __ldg(pi + pvs), __ldg(pi + 2 * pvs) };
    const double var6[6] = { __ldg(pi + 3 * pvs), __ldg(pi + 4 * pvs), __ldg(pi + 5 * pvs),
                             __ldg(pi + 6 * pvs), __ldg(pi + 7 * pvs), __ldg(pi + 8 * pvs) };
    const int cached = (live && !(bt.mode & VN_IEKF_NOCACHE)) ? cache[ii] : -1;
    double wld[3];
    rot_trans(cR, cp, pnt, wld);
    if (bt.variant & 16)  // experiment: streaming loads only
    {
      s22 += wld[0] + wld[1] + wld[2] + var6[0] + var6[1] + var6[2] + var6[3] + var6[4] + var6[5] + (double)cached;
      continue;
    }

    // ---- association. Cached leaf first (odometry.cpp:124-127): the map does not change during the call and
    // the cache only ever holds leaves that passed the gate, so a cached node is a plane leaf - if the point
    // is still inside its box neither the descent nor the flags are needed, and line 0 of the record (box,
    // centre, normal, radius) arrives in one gather.
    int node = -1;
    double2 l0, l1, l2;
    float radius = 0.0f;
    bool have = false;
    if (cached >= 0)
    {
      const double2* L = reinterpret_cast<const double2*>(hot + cached);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(L + 8));  // line 1 (sigma_l terms): on its way before the gate asks
      const double2 l3 = __ldg(L + 3), l4 = __ldg(L + 4);
      l0 = __ldg(L + 0);
      l1 = __ldg(L + 1);
      l2 = __ldg(L + 2);
      const double vc[3] = { l3.x, l3.y, l4.x };
      const float2 rq = *reinterpret_cast<const float2*>(&l4.y);  // (radius, quater_length)
      if (inside_box(wld, vc, rq.y))
      {
        node = cached;
        radius = rq.x;
        have = true;
      }
    }
    long long kc[3] = { 0, 0, 0 };
    if (DEBUG || (live && !have))
    {
#pragma unroll
      for (int k = 0; k < 3; k++) kc[k] = voxel_coord(wld[k], q.voxel_size);
    }
    if (live && !have)
    {
      unsigned long long key;
      if (pack_key(kc[0], kc[1], kc[2], &key))
      {
        unsigned int hh = hash_key(key) & q.hmask;
        for (unsigned int probe = 0; probe <= q.hmask; probe++)
        {
          const ulonglong2 s = __ldg(reinterpret_cast<const ulonglong2*>(q.slots + hh));
          if (s.x == key)
          {
            node = (int)(unsigned int)(s.y & 0xffffffffull);
            break;
          }
          if (s.x == VN_EMPTY_KEY) break;
          hh = (hh + 1) & q.hmask;
        }
      }
      // descend to the leaf (octree.cpp:584-591): one line per level
      int flags = 0;
      while (node >= 0)
      {
        const NodeHot* h = hot + node;
        flags = h->flags;
        if (!(flags & VN_FLAG_INTERIOR)) break;
        const double vc[3] = { h->vcenter[0], h->vcenter[1], h->vcenter[2] };
        node = h->children[child_index(wld, vc)];
      }
      if (node >= 0 && (flags & VN_FLAG_PLANE))
      {
        const double2* L = reinterpret_cast<const double2*>(hot + node);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(L + 8));
        l0 = __ldg(L + 0);
        l1 = __ldg(L + 1);
        l2 = __ldg(L + 2);
        radius = hot[node].radius;
        have = true;
      }
    }

    // contributions of this point (a, b), staged for the warp's DMMA chain; all zero unless the gate passes
    __syncwarp();  // the previous round's fragment loads are done
    int flag = 0;
    double sigma_l = 0.0;
    if (have && !(bt.variant & 4))
    {
      const double2* L = reinterpret_cast<const double2*>(hot + node);
      const double c[3] = { l0.x, l0.y, l1.x };
      const double nr[3] = { l1.y, l2.x, l2.y };
      const double d[3] = { ds(wld[0], c[0]), ds(wld[1], c[1]), ds(wld[2], c[2]) };
      const double dn = dot3(nr, d);
      const float dis_to_plane = (float)fabs(dn);
      const double e[3] = { ds(c[0], wld[0]), ds(c[1], wld[1]), ds(c[2], wld[2]) };
      const float dis_to_center = (float)dot3(e, e);
      const float range_dis = fs(dis_to_center, fm(dis_to_plane, dis_to_plane));
      if (range_dis <= fm(9.0f, radius))
      {
        // sigma_l = J plane_var J^T, J = [wld - center, -normal] = d^T A d - 2 d.(B n) + n^T C n (NodeHot, line 1)
        {
          const double2 l8 = __ldg(L + 8), l9 = __ldg(L + 9), l10 = __ldg(L + 10), l11 = __ldg(L + 11);
          const double qb2 = __ldg(reinterpret_cast<const double*>(L + 12)), qk = __ldg(reinterpret_cast<const double*>(L + 7) + 1);
          const double A0 = l8.x, A1 = l8.y, A2 = l9.x, A3 = l9.y, A4 = l10.x, A5 = l10.y;
          const double dAd = d[0] * (A0 * d[0] + 2.0 * (A1 * d[1] + A2 * d[2])) + d[1] * (A3 * d[1] + 2.0 * A4 * d[2]) +
                             d[2] * A5 * d[2];
          sigma_l = dAd - 2.0 * (d[0] * l11.x + d[1] * l11.y + d[2] * qb2) + qk;
        }
        // + n^T var_world n
        double m[3];
#pragma unroll
        for (int k = 0; k < 3; k++) m[k] = cR[3 * k] * nr[0] + cR[3 * k + 1] * nr[1] + cR[3 * k + 2] * nr[2];
        const double q1 = m[0] * (var6[0] * m[0] + 2.0 * (var6[1] * m[1] + var6[2] * m[2])) +
                          m[1] * (var6[3] * m[1] + 2.0 * var6[4] * m[2]) + m[2] * var6[5] * m[2];
        const double u[3] = { nr[1] * pnt[2] - nr[2] * pnt[1], nr[2] * pnt[0] - nr[0] * pnt[2],
                              nr[0] * pnt[1] - nr[1] * pnt[0] };  // hat(p)^T n = n x p
        double q2 = 0.0, q3 = 0.0;
#pragma unroll
        for (int a = 0; a < 3; a++)
        {
          q2 += u[a] * (crv[a] * u[0] + crv[a + 3] * u[1] + crv[a + 6] * u[2]);
          q3 += nr[a] * (ctv[a] * nr[0] + ctv[a + 3] * nr[1] + ctv[a + 6] * nr[2]);
        }
        sigma_l += q1 + q2 + q3;
        // dis_to_plane < 3 sqrt(sigma_l)  <=>  dis_to_plane^2 < 9 sigma_l (the fp32 value squares exactly in fp64)
        if ((double)dis_to_plane * (double)dis_to_plane < 9.0 * sigma_l)
        {
          flag = 1;
          if (node != cached && !(bt.mode & VN_IEKF_NOCACHE)) cache[ii] = node;  // oc = this (octree.cpp:571-575)
          const double Rinv = 1.0 / (0.0005 + sigma_l);
          // jac = [hat(p) R^T n ; n] = [p x m ; n]
          const double j0 = pnt[1] * m[2] - pnt[2] * m[1];
          const double j1 = pnt[2] * m[0] - pnt[0] * m[2];
          const double j2 = pnt[0] * m[1] - pnt[1] * m[0];
          Sw[0 * IEKF_LD + lane] = j0;
          Sw[1 * IEKF_LD + lane] = j1;
          Sw[2 * IEKF_LD + lane] = j2;
          Sw[3 * IEKF_LD + lane] = nr[0];
          Sw[4 * IEKF_LD + lane] = nr[1];
          Sw[5 * IEKF_LD + lane] = nr[2];
          Sw[6 * IEKF_LD + lane] = dn;
          Sw[8 * IEKF_LD + lane] = Rinv * j0;
          Sw[9 * IEKF_LD + lane] = Rinv * j1;
          Sw[10 * IEKF_LD + lane] = Rinv * j2;
          Sw[11 * IEKF_LD + lane] = Rinv * nr[0];
          Sw[12 * IEKF_LD + lane] = Rinv * nr[1];
          Sw[13 * IEKF_LD + lane] = Rinv * nr[2];
          s22 = fma(nr[2], nr[2], s22);
          cnt++;
        }
      }
    }
    if (DEBUG && live)
    {
      q.dbg.keys[3 * (size_t)i + 0] = kc[0];
      q.dbg.keys[3 * (size_t)i + 1] = kc[1];
      q.dbg.keys[3 * (size_t)i + 2] = kc[2];
      q.dbg.flags[i] = (unsigned char)flag;
      q.dbg.codes[i] = flag ? (hot[node].layer | (q.cold[node].path << 2)) : -1;
      q.dbg.sigma[i] = flag ? sigma_l : 0.0;
    }
    if (!flag)
    {
#pragma unroll
      for (int a = 0; a < 7; a++) Sw[a * IEKF_LD + lane] = 0.0;
#pragma unroll
      for (int a = 8; a < 14; a++) Sw[a * IEKF_LD + lane] = 0.0;
    }
    if (bt.variant & 8) continue;
    // let the FP64 tensor pipe sum the warp's 32 rank-1 updates
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 8; s++) dmma_8x8x4(c0, c1, fa[4 * s], fb[4 * s]);
  }

  // ---- block reduction across the warps in fixed order ------------------------------------------------
  // per warp: 64 entries of C, then n2*n2 and the count (shuffle tree over the lanes)
  double s22w = s22;
  int cntw = cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    s22w += __shfl_xor_sync(0xffffffffu, s22w, o);
    cntw += __shfl_xor_sync(0xffffffffu, cntw, o);
  }
  __syncwarp();
  Sw[g * 8 + 2 * t4] = c0;
  Sw[g * 8 + 2 * t4 + 1] = c1;
  if (lane == 0)
  {
    Sw[64] = s22w;
    Sw[65] = (double)cntw;
  }
  __syncthreads();
  const unsigned int nblk = gridDim.x;
  if (threadIdx.x < VN_IEKF_NACC)
  {
    const int src = iekf_src(threadIdx.x);
    double v = 0.0;
#pragma unroll 8
    for (int w = 0; w < IEKF_WARPS; w++) v += smem[w * IEKF_TILE + src];
    if (threadIdx.x >= 21 && threadIdx.x < 27) v = -v;  // HTz -= Rinv j r
    q.partials[(size_t)threadIdx.x * nblk + blockIdx.x] = v;  // [34][nblk]: the final pass reads it coalesced
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    unsigned int t = atomicAdd(q.ticket, 1u);
    is_last = (t == nblk - 1);
  }
  __syncthreads();
  if (!is_last) return;
  if (bt.variant & 1)
  {
    if (threadIdx.x == 0) *q.ticket = 0u;
    return;
  }
  // last block: one warp per column (34 columns over 32 warps), lanes stride over the per-block partials with
  // independent loads, then a fixed shuffle tree -> deterministic
  __threadfence();
  for (int k = warp; k < VN_IEKF_NACC; k += IEKF_WARPS)
  {
    const double* col = q.partials + (size_t)k * nblk;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0, v4 = 0.0;
    const unsigned int b = lane;
    if (b < nblk) v0 = __ldcg(col + b);
    if (b + 32 < nblk) v1 = __ldcg(col + b + 32);
    if (b + 64 < nblk) v2 = __ldcg(col + b + 64);
    if (b + 96 < nblk) v3 = __ldcg(col + b + 96);
    for (unsigned int bb = b + 128; bb < nblk; bb += 32) v4 += __ldcg(col + bb);
    double v = ((v0 + v1) + (v2 + v3)) + v4;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) fin[k] = v;
  }
  __syncthreads();
  if (threadIdx.x < VN_IEKF_NACC) dev->sums[threadIdx.x] = fin[threadIdx.x];
  if (threadIdx.x == 0) *q.ticket = 0u;
  if (bt.mode & VN_IEKF_SOLVE)
  {
    iekf_solve_block(dev, fin, smem, IEKF_THREADS);
    if ((bt.mode & VN_IEKF_HANDOVER) && reinterpret_cast<const int*>(smem + 528)[0])
    {
      // the loop has just finished: hand the iterate to the host (data, system-wide fence, then the flag)
      __syncthreads();
      __threadfence();
      const double* sd = reinterpret_cast<const double*>(dev);
      double* dd = reinterpret_cast<double*>(q.pub);
      for (int i = threadIdx.x; i < (int)(sizeof(IekfDev) / sizeof(double)); i += IEKF_THREADS) dd[i] = __ldcg(sd + i);
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(q.pub_flag) = q.pub_seq;
    }
  }
  if (bt.mode & VN_IEKF_PUBLISH)
  {
    // the sums go straight to mapped pinned host memory; the sequence number is written last so that the
    // host can poll for completion instead of paying a stream synchronisation per IEKF iteration
    if (threadIdx.x < VN_IEKF_NACC)
    {
      q.result[threadIdx.x] = fin[threadIdx.x];
      __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) reinterpret_cast<volatile unsigned long long*>(q.result)[40] = q.seq;
  }
}

// ---------------------------------------------------------------------------
// The whole iteration loop of LioStateEstimation (odometry.cpp:98-231) as ONE persistent launch (the default
// per-scan path). What changes against max_iter launches of k_iekf:
//   * the scan stays on chip: every block owns a contiguous chunk of the scan and pulls its 9 pointVar rows into
//     shared memory once, with 1-D TMA bulk copies (cp.async.bulk -> mbarrier, one barrier per round so that round 0
//     starts while round 1 is still in flight); iterations 1.. read the points from shared memory, and the per-point
//     leaf cache (odometry.cpp:79, 124-127) lives there too - after iteration 0 the only global traffic is the
//     gather of the cached leaves' records (L2);
//   * no launch per iteration: blocks meet at a grid barrier (arrival counter in global memory, cooperative launch
//     guarantees co-residency), then EVERY block adds the per-block partial sums in the same fixed order and
//     applies the update (a7) to its own shared-memory copy of the iterate - bitwise the same everywhere, so all
//     blocks take the same convergence decision and go straight on: one barrier per iteration, no broadcast;
//   * a cached leaf's two record lines are requested together (one L2 latency instead of two), the DMMA
//     reduction runs two independent accumulator chains, and the staging tile holds b = [j, r] and 1/(0.0005 +
//     sigma) once (the a-operand Rinv*j is formed in the fragment load: the same single product as before).
// Rounds beyond the resident store (scans larger than LOOP_RS * LOOP_T * gridDim.x points) stream from global
// memory with the global leaf cache, like k_iekf.
#define LOOP_T 832   // 26 warps: two rounds cover 240 000 points on 148 SMs
#define LOOP_RS 2    // resident rounds
#define LOOP_WARPS (LOOP_T / 32)
#define LOOP_ROWS 8  // j0..j5, r, Rinv
#define LOOP_TILE (LOOP_ROWS * IEKF_LD)
#define LOOP_SMEM ((size_t)(LOOP_RS * 9 * LOOP_T + LOOP_WARPS * LOOP_TILE) * sizeof(double) + (size_t)LOOP_RS * LOOP_T * sizeof(int))
#define LOOP_SPIN_LIMIT (1ll << 24)

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned int bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long* b)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned int parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LOOP_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LOOP_MBAR_DONE;\n"
      "bra LOOP_MBAR_WAIT;\n"
      "LOOP_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

#ifdef VINA_LOOP_TRACE
#define LT(k) \
  if (tid == 0 && tr_n < 64) tr[tr_n++] = (((unsigned long long)(k)) << 56) | (gtimer() & 0xffffffffffffffull)
__device__ __forceinline__ unsigned long long gtimer()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#else
#define LT(k)
#endif

__global__ void __launch_bounds__(LOOP_T, 1) k_iekf_loop(const __grid_constant__ IekfLoop a)
{
#ifdef VINA_LOOP_TRACE
  unsigned long long tr[64];
  int tr_n = 0;
#endif
  const IekfSeq& q = a.q;
  IekfDev* __restrict__ dev = q.dev;
  extern __shared__ __align__(128) double smem[];
  double* S = smem;                                   // [LOOP_RS][9][LOOP_T] resident pointVar rows
  double* tiles = smem + LOOP_RS * 9 * LOOP_T;        // [LOOP_WARPS][LOOP_TILE] staging; the solve's scratch afterwards
  int* scache = reinterpret_cast<int*>(tiles + LOOP_WARPS * LOOP_TILE);  // [LOOP_RS][LOOP_T] leaf cache
  __shared__ double s_st[42];    // x_curr (R p v bg ba), x_prop: the block's own copy of the iterate
  __shared__ double s_cov[232];  // prior covariance (posterior once the loop ends)
  __shared__ double crv[9], ctv[9];
  __shared__ double fin[VN_IEKF_NACC];
  __shared__ int ctl[4];  // iter, rematch, done, max_iter
  __shared__ __align__(8) unsigned long long mbar[LOOP_RS];
  __shared__ int s_abort;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  LT(0);
  const int nblk = gridDim.x;
  const int n = q.n_ptr ? *q.n_ptr : q.n_host;
  const int blk_base = blockIdx.x * a.chunk;
  const int blk_end = min(n, blk_base + a.chunk);
  const int blk_cnt = max(0, blk_end - blk_base);
  const int nrounds = (blk_cnt + LOOP_T - 1) / LOOP_T;
  const double* __restrict__ pv = q.pv_base;
  const size_t pvs = (size_t)q.pv_stride;
  const NodeHot* __restrict__ hot = q.hot;
  int* __restrict__ gcache = q.cache;

  if (tid == 0)
  {
    for (int r = 0; r < LOOP_RS; r++) mbar_init(&mbar[r], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_abort = 0;
    // the block's chunk, 9 rows per round, straight into shared memory (in flight while the iterate is staged)
    for (int r = 0; r < LOOP_RS; r++)
    {
      const int first = blk_base + r * LOOP_T;
      int cnt = min(LOOP_T, blk_end - first);
      if (cnt <= 0) break;
      cnt = (cnt + 1) & ~1;  // 16-byte granules (the row pitch is even and the chunk a multiple of 32)
      mbar_expect_tx(&mbar[r], 9u * (unsigned int)cnt * 8u);
      for (int k = 0; k < 9; k++) bulk_g2s(S + (r * 9 + k) * LOOP_T, pv + (size_t)k * pvs + first, (unsigned int)cnt * 8u, &mbar[r]);
    }
  }
  for (int i = tid; i < 225; i += LOOP_T) s_cov[i] = dev->cov[i];
  if (tid < 42) s_st[tid] = reinterpret_cast<const double*>(dev)[tid];
  if (tid < 9)
  {
    crv[tid] = dev->rot_var[tid];
    ctv[tid] = dev->tsl_var[tid];
  }
  if (tid < 4) ctl[tid] = (&dev->iter)[tid];
  for (int i = tid; i < LOOP_RS * LOOP_T; i += LOOP_T) scache[i] = -1;  // octos(psize, nullptr), odometry.cpp:79
  __syncthreads();
  // (the host stages an unfinished iterate before every launch: iekf_stage)

  double* Sw = tiles + warp * LOOP_TILE;
  const int g = lane >> 2, t4 = lane & 3;
  const double* fa = Sw + (g < 6 ? g : g - 3) * IEKF_LD + t4;  // a = [Rinv j0..j5, n0 (= j3), n1 (= j4)]
  const double* fr = Sw + 7 * IEKF_LD + t4;                    // Rinv per point
  const double* fb = Sw + (g < 7 ? g : 0) * IEKF_LD + t4;      // b = [j0..j5, r, 0]
  int par = 0;
  bool first_iter = true;

  for (int it_local = 0;; it_local++)
  {
    const double* cR = s_st;
    const double* cp = s_st + 9;
    LT(1);
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0, s22 = 0.0;
    int cnt = 0;
    for (int r = 0; r < nrounds; r++)
    {
      const bool res = r < LOOP_RS;
      const int off = r * LOOP_T + tid;
      const int idx = blk_base + off;
      const bool live = off < blk_cnt;
      if (res && first_iter) mbar_wait(&mbar[r], 0);
      const double* sp = S + (r * 9) * LOOP_T + tid;
      const double* gp = pv + (live ? idx : 0);
      int flag = 0;
      double st_j0 = 0, st_j1 = 0, st_j2 = 0, st_n0 = 0, st_n1 = 0, st_n2 = 0, st_r = 0, st_w = 0;
      if (live)
      {
        double pnt[3];
        if (res)
        {
          pnt[0] = sp[0];
          pnt[1] = sp[LOOP_T];
          pnt[2] = sp[2 * LOOP_T];
        }
        else
        {
          pnt[0] = __ldg(gp);
          pnt[1] = __ldg(gp + pvs);
          pnt[2] = __ldg(gp + 2 * pvs);
        }
        const int cached = res ? scache[off] : gcache[idx];
        double wld[3];
        rot_trans(cR, cp, pnt, wld);
        int node = -1;
        bool have = false;
        double2 l0, l1, l2, l8, l9, l10, l11;
        double qb2 = 0.0, qk = 0.0;
        float radius = 0.0f;
        if (cached >= 0)
        {
          // cached leaf (odometry.cpp:124-127): a plane leaf that passed the gate before; both lines of its record
          // are requested at once
          const double2* L = reinterpret_cast<const double2*>(hot + cached);
          const double2 l3 = __ldg(L + 3), l4 = __ldg(L + 4);
          l0 = __ldg(L + 0);
          l1 = __ldg(L + 1);
          l2 = __ldg(L + 2);
          l8 = __ldg(L + 8);
          l9 = __ldg(L + 9);
          l10 = __ldg(L + 10);
          l11 = __ldg(L + 11);
          qb2 = __ldg(reinterpret_cast<const double*>(L + 12));
          qk = __ldg(reinterpret_cast<const double*>(L + 7) + 1);
          const double vc[3] = { l3.x, l3.y, l4.x };
          const float2 rq = *reinterpret_cast<const float2*>(&l4.y);  // (radius, quater_length)
          if (inside_box(wld, vc, rq.y))
          {
            node = cached;
            radius = rq.x;
            have = true;
          }
        }
        if (!have)
        {
          long long kc[3];
#pragma unroll
          for (int k = 0; k < 3; k++) kc[k] = voxel_coord(wld[k], q.voxel_size);
          unsigned long long key;
          node = -1;
          if (pack_key(kc[0], kc[1], kc[2], &key))
          {
            unsigned int hh = hash_key(key) & q.hmask;
            for (unsigned int probe = 0; probe <= q.hmask; probe++)
            {
              const ulonglong2 s = __ldg(reinterpret_cast<const ulonglong2*>(q.slots + hh));
              if (s.x == key)
              {
                node = (int)(unsigned int)(s.y & 0xffffffffull);
                break;
              }
              if (s.x == VN_EMPTY_KEY) break;
              hh = (hh + 1) & q.hmask;
            }
          }
          int flags = 0;
          while (node >= 0)  // descend to the leaf (octree.cpp:584-591): one line per level
          {
            const NodeHot* h = hot + node;
            flags = h->flags;
            if (!(flags & VN_FLAG_INTERIOR)) break;
            const double vc[3] = { h->vcenter[0], h->vcenter[1], h->vcenter[2] };
            node = h->children[child_index(wld, vc)];
          }
          if (node >= 0 && (flags & VN_FLAG_PLANE))
          {
            const double2* L = reinterpret_cast<const double2*>(hot + node);
            l0 = __ldg(L + 0);
            l1 = __ldg(L + 1);
            l2 = __ldg(L + 2);
            l8 = __ldg(L + 8);
            l9 = __ldg(L + 9);
            l10 = __ldg(L + 10);
            l11 = __ldg(L + 11);
            qb2 = __ldg(reinterpret_cast<const double*>(L + 12));
            qk = __ldg(reinterpret_cast<const double*>(L + 7) + 1);
            radius = hot[node].radius;
            have = true;
          }
        }
        if (have)
        {
          const double c[3] = { l0.x, l0.y, l1.x };
          const double nr[3] = { l1.y, l2.x, l2.y };
          const double d[3] = { ds(wld[0], c[0]), ds(wld[1], c[1]), ds(wld[2], c[2]) };
          const double dn = dot3(nr, d);
          const float dis_to_plane = (float)fabs(dn);
          const double e[3] = { ds(c[0], wld[0]), ds(c[1], wld[1]), ds(c[2], wld[2]) };
          const float dis_to_center = (float)dot3(e, e);
          const float range_dis = fs(dis_to_center, fm(dis_to_plane, dis_to_plane));
          if (range_dis <= fm(9.0f, radius))
          {
            // sigma_l = J plane_var J^T, J = [wld - center, -normal] = d^T A d - 2 d.(B n) + n^T C n (NodeHot, line 1)
            const double A0 = l8.x, A1 = l8.y, A2 = l9.x, A3 = l9.y, A4 = l10.x, A5 = l10.y;
            const double dAd = d[0] * (A0 * d[0] + 2.0 * (A1 * d[1] + A2 * d[2])) + d[1] * (A3 * d[1] + 2.0 * A4 * d[2]) +
                               d[2] * A5 * d[2];
            double sigma_l = dAd - 2.0 * (d[0] * l11.x + d[1] * l11.y + d[2] * qb2) + qk;
            // + n^T var_world n
            double var6[6];
            if (res)
            {
#pragma unroll
              for (int k = 0; k < 6; k++) var6[k] = sp[(3 + k) * LOOP_T];
            }
            else
            {
#pragma unroll
              for (int k = 0; k < 6; k++) var6[k] = __ldg(gp + (3 + k) * pvs);
            }
            double m[3];
#pragma unroll
            for (int k = 0; k < 3; k++) m[k] = cR[3 * k] * nr[0] + cR[3 * k + 1] * nr[1] + cR[3 * k + 2] * nr[2];
            const double q1 = m[0] * (var6[0] * m[0] + 2.0 * (var6[1] * m[1] + var6[2] * m[2])) +
                              m[1] * (var6[3] * m[1] + 2.0 * var6[4] * m[2]) + m[2] * var6[5] * m[2];
            const double u[3] = { nr[1] * pnt[2] - nr[2] * pnt[1], nr[2] * pnt[0] - nr[0] * pnt[2],
                                  nr[0] * pnt[1] - nr[1] * pnt[0] };  // hat(p)^T n = n x p
            double q2 = 0.0, q3 = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++)
            {
              q2 += u[k] * (crv[k] * u[0] + crv[k + 3] * u[1] + crv[k + 6] * u[2]);
              q3 += nr[k] * (ctv[k] * nr[0] + ctv[k + 3] * nr[1] + ctv[k + 6] * nr[2]);
            }
            sigma_l += q1 + q2 + q3;
            // dis_to_plane < 3 sqrt(sigma_l)  <=>  dis_to_plane^2 < 9 sigma_l (the fp32 value squares exactly in fp64)
            if ((double)dis_to_plane * (double)dis_to_plane < 9.0 * sigma_l)
            {
              flag = 1;
              if (node != cached)  // oc = this (octree.cpp:571-575)
              {
                if (res)
                  scache[off] = node;
                else
                  gcache[idx] = node;
              }
              st_w = 1.0 / (0.0005 + sigma_l);
              // jac = [hat(p) R^T n ; n] = [p x m ; n]
              st_j0 = pnt[1] * m[2] - pnt[2] * m[1];
              st_j1 = pnt[2] * m[0] - pnt[0] * m[2];
              st_j2 = pnt[0] * m[1] - pnt[1] * m[0];
              st_n0 = nr[0];
              st_n1 = nr[1];
              st_n2 = nr[2];
              st_r = dn;
              s22 = fma(nr[2], nr[2], s22);
              cnt++;
            }
          }
        }
      }
      // the warp's 32 (a, b) pairs go through the FP64 tensor pipe; a point without a match contributes zeros
      __syncwarp();  // the previous round's fragment loads are done
      Sw[0 * IEKF_LD + lane] = st_j0;
      Sw[1 * IEKF_LD + lane] = st_j1;
      Sw[2 * IEKF_LD + lane] = st_j2;
      Sw[3 * IEKF_LD + lane] = st_n0;
      Sw[4 * IEKF_LD + lane] = st_n1;
      Sw[5 * IEKF_LD + lane] = st_n2;
      Sw[6 * IEKF_LD + lane] = st_r;
      Sw[7 * IEKF_LD + lane] = st_w;
      (void)flag;
      __syncwarp();
#pragma unroll
      for (int s = 0; s < 8; s += 2)
      {
        const double a0 = g < 6 ? fa[4 * s] * fr[4 * s] : fa[4 * s];
        const double b0 = g < 7 ? fb[4 * s] : 0.0;
        const double a1 = g < 6 ? fa[4 * s + 4] * fr[4 * s + 4] : fa[4 * s + 4];
        const double b1 = g < 7 ? fb[4 * s + 4] : 0.0;
        dmma_8x8x4(c0, c1, a0, b0);
        dmma_8x8x4(d0, d1, a1, b1);
      }
    }
    first_iter = false;
    LT(2);
    c0 += d0;
    c1 += d1;

    // ---- block reduction across the warps in fixed order
    double s22w = s22;
    int cntw = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      s22w += __shfl_xor_sync(0xffffffffu, s22w, o);
      cntw += __shfl_xor_sync(0xffffffffu, cntw, o);
    }
    __syncwarp();
    Sw[g * 8 + 2 * t4] = c0;
    Sw[g * 8 + 2 * t4 + 1] = c1;
    if (lane == 0)
    {
      Sw[64] = s22w;
      Sw[65] = (double)cntw;
    }
    __syncthreads();
    double* part = a.partials + (size_t)par * VN_IEKF_NACC * nblk;
    if (tid < VN_IEKF_NACC)
    {
      const int src = iekf_src(tid);
      double v = 0.0;
#pragma unroll 2
      for (int w = 0; w < LOOP_WARPS; w++) v += tiles[w * LOOP_TILE + src];
      if (tid >= 21 && tid < 27) v = -v;  // HTz -= Rinv j r
      __stcg(part + (size_t)tid * nblk + blockIdx.x, v);
    }
    __syncthreads();
    // ---- grid barrier: arrive, wait for everybody's partial sums
    LT(3);
    if (tid == 0)
    {
      __threadfence();
      atomicAdd(a.bar, 1ull);
      const unsigned long long target = (unsigned long long)(it_local + 1) * (unsigned long long)nblk;
      long long spins = 0;
      while (ld_acquire_u64(a.bar) < target)
      {
        if (++spins > LOOP_SPIN_LIMIT)
        {
          atomicOr(a.status, VN_ST_SPIN);
          s_abort = 1;
          break;
        }
      }
      __threadfence();
    }
    __syncthreads();
    LT(4);
    // every block: the same sums in the same order (a half-warp per column: 34 columns in one pass; lanes stride
    // over the blocks with independent loads, then a fixed shuffle tree)
    {
      const int hw = tid >> 4, hl = tid & 15;
      double v0 = 0.0, v1 = 0.0;
      if (hw < VN_IEKF_NACC)
      {
        const double* col = part + (size_t)hw * nblk;
#pragma unroll 5
        for (int b = hl; b < nblk; b += 32)
        {
          v0 += __ldcg(col + b);
          if (b + 16 < nblk) v1 += __ldcg(col + b + 16);
        }
      }
      double v = v0 + v1;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (hw < VN_IEKF_NACC && hl == 0) fin[hw] = v;
    }
    __syncthreads();
    LT(5);
    iekf_solve_cta(ctl, fin, tiles, s_cov, s_st, tid, LOOP_T);
    LT(6);
    par ^= 1;
    if (ctl[2] || s_abort) break;
  }

  // ---- the loop has finished: block 0 stores the iterate and hands it to the host
  if (blockIdx.x == 0)
  {
    double* dd = reinterpret_cast<double*>(dev);
    if (tid < 21) dd[tid] = s_st[tid];
    for (int i = tid; i < 225; i += LOOP_T) dev->cov[i] = s_cov[i];
    if (tid < VN_IEKF_NACC) dev->sums[tid] = fin[tid];
    if (tid < 3) (&dev->iter)[tid] = ctl[tid];
    if (a.mode & VN_IEKF_HANDOVER)
    {
      __threadfence();
      __syncthreads();
      const double* sd = reinterpret_cast<const double*>(dev);
      double* pd = reinterpret_cast<double*>(q.pub);
      for (int i = tid; i < (int)(sizeof(IekfDev) / sizeof(double)); i += LOOP_T) pd[i] = __ldcg(sd + i);
      __threadfence_system();
      __syncthreads();
      if (tid == 0) *reinterpret_cast<volatile unsigned long long*>(q.pub_flag) = q.pub_seq;
    }
  }
#ifdef VINA_LOOP_TRACE
  LT(7);
  if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == nblk - 1 || blockIdx.x == nblk / 2))
  {
    // phase stamps in ns since the block started: 1 iteration starts, 2 rounds done, 3 partials written, 4 barrier
    // passed, 5 sums added, 6 update done, 7 handed over
    char buf[8];
    (void)buf;
    printf("[loop trace] blk %d:", (int)blockIdx.x);
    for (int i = 1; i < tr_n; i++)
      printf(" %d:%llu", (int)(tr[i] >> 56), (tr[i] & 0xffffffffffffffull) - (tr[0] & 0xffffffffffffffull));
    printf("\n");
  }
#endif
  // the last block out leaves the barrier words at zero for the next launch
  if (tid == 0)
  {
    const unsigned long long t = atomicAdd(a.bar + 1, 1ull);
    if (t == (unsigned long long)nblk - 1ull)
    {
      a.bar[0] = 0ull;
      a.bar[1] = 0ull;
      __threadfence();
    }
  }
}

int iekf_loop_chunk(int n, int blocks)
{
  int c = (n + blocks - 1) / blocks;
  c = (c + 31) & ~31;
  return c < 32 ? 32 : c;
}

int launch_iekf_loop(cudaStream_t st, const IekfLoop& a, int blocks)
{
  static bool attr_set_dev[64] = { false };
  int dv = 0;
  cudaGetDevice(&dv);
  bool& attr_set = attr_set_dev[dv & 63];
  if (!attr_set)
  {
    cudaError_t e = cudaFuncSetAttribute(k_iekf_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LOOP_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  void* args[] = { const_cast<IekfLoop*>(&a) };
  return (int)cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(k_iekf_loop), dim3(blocks), dim3(LOOP_T), args, LOOP_SMEM, st);
}

// Hand the converged iterate to the host through mapped pinned memory: the data first, then (after a
// system-wide fence) the sequence number the host polls - a few microseconds instead of the wake-up latency of
// a stream synchronisation, and the host can enqueue the map update right away.
__global__ void __launch_bounds__(256) k_publish_iterate(const IekfDev* __restrict__ src, IekfDev* __restrict__ dst,
                                                         volatile unsigned long long* flag, unsigned long long seq)
{
  const double* s = reinterpret_cast<const double*>(src);
  double* d = reinterpret_cast<double*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(IekfDev) / sizeof(double)); i += blockDim.x) d[i] = s[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *flag = seq;
}

// ---- the IEKF against a map sharded over GPUs: the 34 sums of every shard travel through peer memory ----------
// (shard_kernels.cu has the record exchange). Every rank stores its sums into every peer's control block, then
// raises the peers' flags; every rank adds the rows in rank order - bitwise the same total everywhere - and applies
// the same update to its copy of the iterate: the ranks stay in lock step without a collective or the host.
__global__ void __launch_bounds__(64) k_p2p_sums_publish(ShardPeers peers, const IekfDev* __restrict__ dev, unsigned long long epoch)
{
  if (dev->done) return;
  if (threadIdx.x < VN_IEKF_NACC)
  {
    const double v = dev->sums[threadIdx.x];
    for (int w = 0; w < peers.world; w++)
      *reinterpret_cast<volatile double*>(&peers.ctrl[w]->sums[peers.rank][threadIdx.x]) = v;
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x < peers.world)
    *reinterpret_cast<volatile unsigned long long*>(&peers.ctrl[threadIdx.x]->sready[peers.rank]) = epoch;
}

#define P2P_SUMS_SPIN_LIMIT 20000000ll
__global__ void __launch_bounds__(256) k_p2p_sums_solve(ShardPeers peers, IekfDev* __restrict__ dev, unsigned long long epoch,
                                                        int* __restrict__ status)
{
  if (dev->done) return;
  __shared__ double smem[IEKF_SOLVE_WS + 288];
  __shared__ double fin[VN_IEKF_NACC];
  const ShardCtrl* me = peers.ctrl[peers.rank];
  if (threadIdx.x < peers.world)
  {
    long long spins = 0;
    while (*reinterpret_cast<const volatile unsigned long long*>(&me->sready[threadIdx.x]) != epoch)
      if (++spins > P2P_SUMS_SPIN_LIMIT)
      {
        atomicOr(status, VN_ST_SPIN);
        break;
      }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x < VN_IEKF_NACC)
  {
    double v = 0.0;
    for (int w = 0; w < peers.world; w++) v += *reinterpret_cast<const volatile double*>(&me->sums[w][threadIdx.x]);
    fin[threadIdx.x] = v;
    dev->sums[threadIdx.x] = v;
  }
  __syncthreads();
  iekf_solve_block(dev, fin, smem, 256);
}

void launch_p2p_sums_publish(cudaStream_t st, const ShardPeers& peers, IekfDev* dev, unsigned long long epoch)
{
  k_p2p_sums_publish<<<1, 64, 0, st>>>(peers, dev, epoch);
}
void launch_p2p_sums_solve(cudaStream_t st, const ShardPeers& peers, IekfDev* dev, unsigned long long epoch, int* status)
{
  k_p2p_sums_solve<<<1, 256, 0, st>>>(peers, dev, epoch, status);
}

void launch_publish_iterate(cudaStream_t st, const IekfDev* src, IekfDev* dst_mapped, unsigned long long* flag_mapped,
                            unsigned long long seq)
{
  k_publish_iterate<<<1, 256, 0, st>>>(src, dst_mapped, flag_mapped, seq);
}

__global__ void k_fill_int(int* p, int v, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int iekf_grid_blocks(int n, int sm_count)
{
  // one point per thread and round, one 1024-thread block per SM (the staging tiles take 126 KB of shared
  // memory); larger scans loop
  int need = (n + IEKF_THREADS - 1) / IEKF_THREADS;
  if (need < 1) need = 1;
  const int cap = sm_count * IEKF_BLOCKS_PER_SM;
  return need < cap ? need : cap;
}

int launch_iekf(cudaStream_t st, const IekfBatch& bt, int nseq, int blocks, bool debug, bool pdl)
{
  static bool attr_set_dev[64] = { false };
  int dv = 0;
  cudaGetDevice(&dv);
  bool& attr_set = attr_set_dev[dv & 63];
  if (!attr_set)
  {
    cudaError_t e = cudaFuncSetAttribute(k_iekf<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IEKF_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k_iekf<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IEKF_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid(blocks, nseq);
  if (debug)
    k_iekf<true><<<grid, IEKF_THREADS, IEKF_SMEM, st>>>(bt);
  else if (pdl)
    vn_launch(k_iekf<false>, grid, dim3(IEKF_THREADS), IEKF_SMEM, st, bt);
  else
    k_iekf<false><<<grid, IEKF_THREADS, IEKF_SMEM, st>>>(bt);  // (no programmatic launch for the first iterations: the next
                                                                // one's blocks would sit on the SMs through this one's tail -
                                                                // the reduction and the update in the last block - and keep
                                                                // the side stream's kernels out; from the third launch on the
                                                                // side stream is through and most launches are no-ops)
  return (int)cudaGetLastError();
}

void launch_fill_int(cudaStream_t st, int* p, int v, int n)
{
  if (n > 0) k_fill_int<<<(n + 255) / 256, 256, 0, st>>>(p, v, n);
}

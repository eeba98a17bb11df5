// The data-parallel part of the sliding-window BA: the LiDAR factor of
// LI_BA_Optimizer::damping_iter (src/mapping/optimizers.cpp:430-517),
//   * LidarFactor::acc_evaluate2        (src/mapping/factors.cpp:22-126): Hessian (6 win x 6 win), gradient, residual
//   * LidarFactor::evaluate_only_residual (factors.cpp:128-158): residual at candidate poses; overwrites the
//     factors' eig / pcr_add, which OctoTree::margi takes over afterwards (octree.cpp:410-422)
// on the factor store filled by k_ba_collect (map_kernels.cu) = the reference's `voxhess` container.
// The reference forks 5 threads over factor ranges and adds the partial 60x60 matrices; here a WARP owns a
// strided subset of the factors:
//   * lane i < win evaluates window frame i of the factor (A_uk 3x6, A_uk^T umumT, v_i x R_i^T u_k, the extra
//     terms of the diagonal block) and parks it in the warp's shared-memory slab,
//   * every lane owns up to two of the win (win + 1) / 2 upper 6x6 blocks and keeps them in registers across all
//     of the warp's factors (72 accumulators), forming H_ij = (A_i^T umumT) A_j + the pair terms,
//   * per-warp results go to a [warp][entry] partial buffer and k_ba_reduce adds them in warp order (fixed
//     order -> the same bits for the same factor order) and mirrors the lower blocks (factors.cpp:123-125).
// Roofline: a factor is 1 072 B and ~12 kflop -> compute/latency bound, not HBM: a few thousand factors take
// microseconds. FP64 FMA is allowed here (tolerance parity, 1e-9 of the largest entry); the eigenvalues of
// k_ba_residual come from the same single-rounding eig3_sym as recut / margi and are bit-exact.
#include "vn_kernels.cuh"

#define BA_WARPS 4
#define BA_THREADS (32 * BA_WARPS)
#define BA_FS 42                               // doubles per frame in the slab: A(18) T(18) w(3) n valid
#define BA_SLAB (VINA_MAX_WIN * (BA_FS + 36))  // + the diagonal block's extra terms (36) per frame
#define BA_NPAIR (VINA_MAX_WIN * (VINA_MAX_WIN + 1) / 2)
#define BA_ENTRIES (36 * BA_NPAIR + 6 * VINA_MAX_WIN + 1)  // blocks, gradient, residual

struct BaPoses
{
  PoseD x[VINA_MAX_WIN];
};

// Results go straight to mapped pinned host memory; the block that finishes last (atomic ticket) raises a sequence
// number there, after a system-wide fence - the host polls it instead of paying a copy and a stream synchronisation
// per evaluation (the LM loop makes four evaluations per iteration pair).
__device__ __forceinline__ void ba_signal_done(const BaDone& d)
{
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
  {
    const unsigned int t = atomicAdd(d.ticket, 1u);
    if (t == gridDim.x * gridDim.y - 1)
    {
      *d.ticket = 0u;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(d.flag) = d.seq;
    }
  }
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c)
{
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
// hat(v) entry (r, c): [0 -v2 v1; v2 0 -v0; -v1 v0 0]
__device__ __forceinline__ void hat33(const double* v, double (*H)[3])
{
  H[0][0] = 0.0, H[0][1] = -v[2], H[0][2] = v[1];
  H[1][0] = v[2], H[1][1] = 0.0, H[1][2] = -v[0];
  H[2][0] = -v[1], H[2][1] = v[0], H[2][2] = 0.0;
}

__global__ void __launch_bounds__(BA_THREADS)
    k_ba_hess(const BaFactor* __restrict__ fac, const int* __restrict__ n_ptr, BaPoses xs, int win, double* __restrict__ partial)
{
  __shared__ double slab_all[BA_WARPS][BA_SLAB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* slab = slab_all[warp];
  const int gw = blockIdx.x * BA_WARPS + warp, nw = gridDim.x * BA_WARPS;
  const int n = *n_ptr;
  const int npair = win * (win + 1) / 2;
  // the blocks this lane owns: pair q -> (i, j), i <= j, row by row
  int bi[2], bj[2];
#pragma unroll
  for (int s = 0; s < 2; s++)
  {
    int q = lane + 32 * s, i = 0;
    if (q < npair)
    {
      while (q >= win - i)
      {
        q -= win - i;
        i++;
      }
      bi[s] = i;
      bj[s] = i + q;
    }
    else
      bi[s] = bj[s] = -1;
  }
  double acc[2][36];
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int e = 0; e < 36; e++) acc[s][e] = 0.0;
  double jac[6] = { 0, 0, 0, 0, 0, 0 };
  double resid = 0.0;

  for (int a = gw; a < n; a += nw)
  {
    const BaFactor& f = fac[a];
    const double coe = f.coe;
    const double NN = (double)f.add.N;
    const double l0 = f.eig_value[0], l1 = f.eig_value[1], l2 = f.eig_value[2];
    double uk[3], u1[3], u2[3], vBar[3];
#pragma unroll
    for (int r = 0; r < 3; r++)
    {
      uk[r] = f.eig_vector[r];
      u1[r] = f.eig_vector[3 + r];
      u2[r] = f.eig_vector[6 + r];
      vBar[r] = f.add.v[r] / NN;
    }
    // umumT = sum_{m != 0} 2 / (l0 - lm) u_m u_m^T (factors.cpp:48-53)
    double um[3][3];
    {
      const double c1 = 2.0 / (l0 - l1), c2 = 2.0 / (l0 - l2);
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) um[r][c] = c1 * u1[r] * u1[c] + c2 * u2[r] * u2[c];
    }
    if (lane == 0) resid += coe * l0;
    __syncwarp();  // the previous factor's slab has been consumed

    if (lane < win)
    {
      double* F = slab + lane * BA_FS;
      double* D = slab + VINA_MAX_WIN * BA_FS + lane * 36;
      const Cluster& lc = f.local[lane];
      F[40] = lc.N != 0 ? 1.0 : 0.0;
      if (lc.N != 0)
      {
        const double ni = (double)lc.N;
        double P[3][3], vi[3], R[3][3], ti[3];
        P[0][0] = lc.P[0], P[1][0] = P[0][1] = lc.P[1], P[2][0] = P[0][2] = lc.P[2];
        P[1][1] = lc.P[3], P[2][1] = P[1][2] = lc.P[4], P[2][2] = lc.P[5];
#pragma unroll
        for (int r = 0; r < 3; r++)
        {
          vi[r] = lc.v[r];
          ti[r] = xs.x[lane].p[r] - vBar[r];
#pragma unroll
          for (int c = 0; c < 3; c++) R[r][c] = xs.x[lane].R[r + 3 * c];
        }
        double Rtu[3], PRtu[3], w[3];
#pragma unroll
        for (int r = 0; r < 3; r++) Rtu[r] = R[0][r] * uk[0] + R[1][r] * uk[1] + R[2][r] * uk[2];
#pragma unroll
        for (int r = 0; r < 3; r++) PRtu[r] = P[r][0] * Rtu[0] + P[r][1] * Rtu[1] + P[r][2] * Rtu[2];
        cross3(vi, Rtu, w);  // hat(v_i) R_i^T u_k
        const double ukt = uk[0] * ti[0] + uk[1] * ti[1] + uk[2] * ti[2];
        double hP[3][3], hv[3][3], hR[3][3], combo1[3][3], combo2[3];
        hat33(PRtu, hP);
        hat33(vi, hv);
        hat33(Rtu, hR);
#pragma unroll
        for (int r = 0; r < 3; r++)
        {
#pragma unroll
          for (int c = 0; c < 3; c++) combo1[r][c] = hP[r][c] + hv[r][c] * ukt;
          combo2[r] = R[r][0] * vi[0] + R[r][1] * vi[1] + R[r][2] * vi[2] + ni * ti[r];
        }
        // A(:, 0:3) = (R P + t v^T) hat(R^T u) - R combo1 ; A(:, 3:6) = combo2 u^T + (combo2 . u) I ; A /= NN
        double M1[3][3], A[3][6];
        const double c2u = combo2[0] * uk[0] + combo2[1] * uk[1] + combo2[2] * uk[2];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int c = 0; c < 3; c++) M1[r][c] = R[r][0] * P[0][c] + R[r][1] * P[1][c] + R[r][2] * P[2][c] + ti[r] * vi[c];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int c = 0; c < 3; c++)
          {
            const double x = M1[r][0] * hR[0][c] + M1[r][1] * hR[1][c] + M1[r][2] * hR[2][c];
            const double y = R[r][0] * combo1[0][c] + R[r][1] * combo1[1][c] + R[r][2] * combo1[2][c];
            A[r][c] = (x - y) / NN;
            A[r][3 + c] = (combo2[r] * uk[c] + (r == c ? c2u : 0.0)) / NN;
          }
        // gradient: coe A^T u_k (factors.cpp:85-86)
        double jj[6];
#pragma unroll
        for (int c = 0; c < 6; c++)
        {
          jj[c] = A[0][c] * uk[0] + A[1][c] * uk[1] + A[2][c] * uk[2];
          jac[c] += coe * jj[c];
        }
        // T = A^T umumT (6 x 3)
#pragma unroll
        for (int c = 0; c < 6; c++)
#pragma unroll
          for (int r = 0; r < 3; r++) F[18 + c + 6 * r] = A[0][c] * um[0][r] + A[1][c] * um[1][r] + A[2][c] * um[2][r];
#pragma unroll
        for (int r = 0; r < 3; r++)
        {
#pragma unroll
          for (int c = 0; c < 6; c++) F[r + 3 * c] = A[r][c];
          F[36 + r] = w[r];
        }
        F[39] = ni;
        // extra terms of the diagonal block (factors.cpp:88-95)
        double X[3][3], hj[3][3];
        hat33(jj, hj);
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int c = 0; c < 3; c++)
            X[r][c] = combo1[r][c] - (hR[r][0] * P[0][c] + hR[r][1] * P[1][c] + hR[r][2] * P[2][c]);
        const double k1 = 2.0 / NN, k2 = 2.0 / NN / NN, kh = 2.0 / NN * (1.0 - ni / NN), k33 = 2.0 / NN * (ni - ni * ni / NN);
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int c = 0; c < 3; c++)
          {
            D[r + 6 * c] = k1 * (X[r][0] * hR[0][c] + X[r][1] * hR[1][c] + X[r][2] * hR[2][c]) - k2 * w[r] * w[c] - 0.5 * hj[r][c];
            D[r + 6 * (3 + c)] = kh * w[r] * uk[c];
            D[3 + r + 6 * c] = kh * w[c] * uk[r];
            D[3 + r + 6 * (3 + c)] = k33 * uk[r] * uk[c];
          }
      }
    }
    __syncwarp();

#pragma unroll
    for (int s = 0; s < 2; s++)
    {
      const int i = bi[s], j = bj[s];
      if (i < 0) continue;
      const double* Fi = slab + i * BA_FS;
      const double* Fj = slab + j * BA_FS;
      if (Fi[40] == 0.0 || Fj[40] == 0.0) continue;
      double T[18], Aj[18];
#pragma unroll
      for (int e = 0; e < 18; e++)
      {
        T[e] = Fi[18 + e];
        Aj[e] = Fj[e];
      }
      if (i == j)
      {
        const double* D = slab + VINA_MAX_WIN * BA_FS + i * 36;
#pragma unroll
        for (int c = 0; c < 6; c++)
#pragma unroll
          for (int r = 0; r < 6; r++)
          {
            const double h = T[r] * Aj[3 * c] + T[r + 6] * Aj[1 + 3 * c] + T[r + 12] * Aj[2 + 3 * c] + D[r + 6 * c];
            acc[s][r + 6 * c] += coe * h;
          }
      }
      else
      {
        const double wi[3] = { Fi[36], Fi[37], Fi[38] }, wj[3] = { Fj[36], Fj[37], Fj[38] };
        const double ni = Fi[39], nj = Fj[39];
        const double k00 = -2.0 / NN / NN, k03 = -2.0 * nj / NN / NN, k30 = -2.0 * ni / NN / NN, k33 = -2.0 * ni * nj / NN / NN;
#pragma unroll
        for (int c = 0; c < 6; c++)
#pragma unroll
          for (int r = 0; r < 6; r++)
          {
            double h = T[r] * Aj[3 * c] + T[r + 6] * Aj[1 + 3 * c] + T[r + 12] * Aj[2 + 3 * c];
            if (r < 3 && c < 3) h += k00 * wi[r] * wj[c];
            if (r < 3 && c >= 3) h += k03 * wi[r] * uk[c - 3];
            if (r >= 3 && c < 3) h += k30 * uk[r - 3] * wj[c];
            if (r >= 3 && c >= 3) h += k33 * uk[r - 3] * uk[c - 3];
            acc[s][r + 6 * c] += coe * h;
          }
      }
    }
  }
  // the block's partial: [36 * pair + e], then the gradient, then the residual. The four warps add their
  // registers into one shared-memory image in warp order, the block writes it out coalesced.
  __shared__ double blk[BA_ENTRIES];
  __syncthreads();
  for (int w = 0; w < BA_WARPS; w++)
  {
    if (warp == w)
    {
#pragma unroll
      for (int s = 0; s < 2; s++)
      {
        const int q = lane + 32 * s;
        if (q < npair)
#pragma unroll
          for (int e = 0; e < 36; e++) blk[36 * q + e] = (w == 0 ? 0.0 : blk[36 * q + e]) + acc[s][e];
      }
      if (lane < win)
#pragma unroll
        for (int c = 0; c < 6; c++)
        {
          const int e = 36 * BA_NPAIR + 6 * lane + c;
          blk[e] = (w == 0 ? 0.0 : blk[e]) + jac[c];
        }
      if (lane == 0)
      {
        const int e = 36 * BA_NPAIR + 6 * VINA_MAX_WIN;
        blk[e] = (w == 0 ? 0.0 : blk[e]) + resid;
      }
    }
    __syncthreads();
  }
  double* out = partial + (size_t)blockIdx.x * BA_ENTRIES;
  for (int e = threadIdx.x; e < BA_ENTRIES; e += BA_THREADS) out[e] = blk[e];
}

// sum of the warps' partials; Hess (6 win)^2 column-major with the lower blocks mirrored (factors.cpp:123-125),
// JacT, residual. A block owns 32 consecutive entries; its 8 warps each add a fixed, interleaved subset of the
// partial rows (coalesced 256-byte loads), then the 8 slice sums are added in slice order: fixed order, so the
// same bits for the same factor order.
__global__ void __launch_bounds__(256)
    k_ba_reduce(const double* __restrict__ partial, int nwarps, int win, double* __restrict__ Hess, double* __restrict__ JacT,
                double* __restrict__ residual, BaDone done)
{
  __shared__ double part[8][33];
  const int el = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + el;
  const int npair = win * (win + 1) / 2;
  const int dim = 6 * win;
  double s = 0.0;
  if (e < BA_ENTRIES)
  {
    double s0 = 0.0, s1 = 0.0;
    int w = slice;
    for (; w + 8 < nwarps; w += 16)
    {
      s0 += partial[(size_t)w * BA_ENTRIES + e];
      s1 += partial[(size_t)(w + 8) * BA_ENTRIES + e];
    }
    if (w < nwarps) s0 += partial[(size_t)w * BA_ENTRIES + e];
    s = s0 + s1;
  }
  part[slice][el] = s;
  __syncthreads();
  if (slice == 0 && e < BA_ENTRIES)
  {
    s = ((part[0][el] + part[1][el]) + (part[2][el] + part[3][el])) + ((part[4][el] + part[5][el]) + (part[6][el] + part[7][el]));
    if (e < 36 * BA_NPAIR)
    {
      int q = e / 36, i = 0;
      if (q < npair)
      {
        const int r = (e % 36) % 6, c = (e % 36) / 6;
        while (q >= win - i)
        {
          q -= win - i;
          i++;
        }
        const int j = i + q;
        Hess[(6 * i + r) + (size_t)dim * (6 * j + c)] = s;
        if (i != j) Hess[(6 * j + c) + (size_t)dim * (6 * i + r)] = s;
      }
    }
    else if (e < 36 * BA_NPAIR + 6 * VINA_MAX_WIN)
    {
      if ((e - 36 * BA_NPAIR) / 6 < win) JacT[e - 36 * BA_NPAIR] = s;
    }
    else
      *residual = s;
  }
  ba_signal_done(done);
}

// LidarFactor::evaluate_only_residual: thread per factor. The factor's eig / pcr_add are overwritten like in the
// reference's container; lam0[a] = lambda_0 of factor a; block sums of coe * lambda_0 go to `partial`.
__global__ void __launch_bounds__(128)
    k_ba_residual(BaFactor* __restrict__ fac, const int* __restrict__ n_ptr, BaPoses xs, int win, double* __restrict__ partial,
                  double* __restrict__ lam0, BaDone done)
{
  __shared__ double red[4];
  const int n = *n_ptr;
  double mine = 0.0;
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x)
  {
    BaFactor& f = fac[a];
    Cluster sig = f.fix;
    for (int i = 0; i < win; i++)
      if (f.local[i].N != 0)
      {
        Cluster w;
        cluster_transform(w, f.local[i], xs.x[i].R, xs.x[i].p);
        cluster_add(sig, w);
      }
    double L[6], ev[3], Q[9];
    cluster_cov(sig, L);
    eig3_sym(L, ev, Q);
    for (int k = 0; k < 3; k++) f.eig_value[k] = ev[k];
    for (int k = 0; k < 9; k++) f.eig_vector[k] = Q[k];
    f.add = sig;
    if (lam0) lam0[a] = ev[0];
    mine += f.coe * ev[0];
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) partial[blockIdx.x] = ((red[0] + red[1]) + red[2]) + red[3];
  ba_signal_done(done);
}

// OctoTree::margi takes the factors' re-evaluated pcr_add / eig back into the leaves (octree.cpp:410-416)
__global__ void __launch_bounds__(128) k_ba_writeback(MapView M, const BaFactor* __restrict__ fac, const int* __restrict__ n_ptr)
{
  const int n = *n_ptr;
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x)
  {
    const BaFactor& f = fac[a];
    NodeCold& c = M.cold[f.node];
    c.pcr_add = f.add;
    for (int k = 0; k < 3; k++) c.eig_value[k] = f.eig_value[k];
    for (int k = 0; k < 9; k++) c.eig_vector[k] = f.eig_vector[k];
  }
}
int launch_ba_writeback(cudaStream_t st, const MapView& map, const BaFactor* fac, const int* n_dev, int sm_count)
{
  k_ba_writeback<<<sm_count * 2, 128, 0, st>>>(map, fac, n_dev);
  return 1;
}

int ba_hess_warps(int sm_count) { return sm_count * 2; }  // partial rows = blocks (2 per SM: 255 registers per thread)
size_t ba_partial_doubles(int sm_count) { return (size_t)ba_hess_warps(sm_count) * BA_ENTRIES; }

int launch_ba_hess(cudaStream_t st, const BaFactor* fac, const int* n_dev, const PoseD* h_xs, int win, int sm_count,
                   double* partial, double* d_out, const BaDone& done)
{
  BaPoses xs;
  memset(&xs, 0, sizeof(xs));
  for (int i = 0; i < win && i < VINA_MAX_WIN; i++) xs.x[i] = h_xs[i];
  const int blocks = ba_hess_warps(sm_count);
  k_ba_hess<<<blocks, BA_THREADS, 0, st>>>(fac, n_dev, xs, win, partial);
  const int dim = 6 * win;
  k_ba_reduce<<<(BA_ENTRIES + 31) / 32, 256, 0, st>>>(partial, blocks, win, d_out, d_out + (size_t)dim * dim,
                                                      d_out + (size_t)dim * dim + dim, done);
  return 2;
}

int launch_ba_residual(cudaStream_t st, BaFactor* fac, const int* n_dev, const PoseD* h_xs, int win, int sm_count,
                       double* partial, double* lam0, const BaDone& done)
{
  BaPoses xs;
  memset(&xs, 0, sizeof(xs));
  for (int i = 0; i < win && i < VINA_MAX_WIN; i++) xs.x[i] = h_xs[i];
  k_ba_residual<<<sm_count * 2, 128, 0, st>>>(fac, n_dev, xs, win, partial, lam0, done);
  return 1;
}

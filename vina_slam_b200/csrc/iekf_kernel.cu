// The hot kernel: one IEKF iteration of VINA_SLAM::LioStateEstimation's point
// loop (src/pipeline/odometry.cpp:111-148): world point, cached-leaf test
// (OctoTree::inside, octree.cpp:732-737), voxel-hash lookup (match,
// voxel_map.cpp:241-266), octree descent and gate (OctoTree::match,
// octree.cpp:551-595), residual / Jacobian and the 6x6 H = J^T R^-1 J, b, n n^T
// reduction (odometry.cpp:136-146).
//
// Roofline: HBM-bound streaming of the pointVar SoA (72 B/pt + 8 B cache RMW)
// plus gathers of 16-B hash slots and 256-B leaf records; the dependent chain
// cache -> leaf record (or key -> slot -> root -> child -> leaf) makes it
// latency-bound unless enough points are in flight, so the kernel is shaped for
// occupancy: ONE point per thread, and the 34 sums never live in registers as
// accumulators - each warp reduces its 32 points at once with a transposed
// butterfly (reduce-scatter over lanes: 36 double shuffles instead of 170) that
// leaves two finished sums per lane. No dense contraction -> no tensor cores.
//
// Numerics: every decision-bearing expression (wld, key, child index, inside,
// the fp32 gate) uses the single-rounding helpers of vn_math.cuh in the order of
// SURVEY.md Appendix A, so keys and associations are bit-exact against the CPU
// restatement. sigma_l and the sums use FMA freely (tolerance 1e-4 rel):
// n^T var_world n is evaluated as (R^T n)^T var (R^T n) + (n x p)^T S_R (n x p)
// + n^T S_t n, which needs ~1/4 of the flops of forming var_world.
// The reduction order is fixed (butterfly, warps in order, blocks in order):
// results are deterministic run to run.
#include "vn_kernels.cuh"

#define IEKF_THREADS 1024
#define IEKF_WARPS (IEKF_THREADS / 32)
#ifndef IEKF_MIN_BLOCKS
#define IEKF_MIN_BLOCKS 1
#endif

// one reduce-scatter stage: N values per lane -> ceil(N/2); lanes with the `off` bit clear keep the
// lower half of the index range, the others the upper half
template <int N, int OFF>
__device__ __forceinline__ void bfly_stage(const double* in, double* out, bool bit)
{
  constexpr int H = (N + 1) / 2;
#pragma unroll
  for (int i = 0; i < H; i++)
  {
    const double lo = in[i];
    const double hi = (i + H < N) ? in[i + H] : 0.0;
    const double send = bit ? lo : hi;
    const double keep = bit ? hi : lo;
    out[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}

// the 34 per-point contributions, evaluated on demand (k is a compile-time constant after unrolling) so
// that they never occupy 34 registers at once: 21 x HTH upper triangle (Rinv j_a j_b), 6 x HTz
// (-Rinv j_a r), 6 x n n^T upper triangle, 1 x match count
__device__ __forceinline__ double xval(int k, const double* ra, const double* jac, const double* nn, double dotn,
                                       double cntv)
{
  constexpr int HA[21] = { 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 5 };
  constexpr int HB[21] = { 0, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 2, 3, 4, 5, 3, 4, 5, 4, 5, 5 };
  constexpr int NA[6] = { 0, 0, 0, 1, 1, 2 }, NB[6] = { 0, 1, 2, 1, 2, 2 };
  if (k < 21) return ra[HA[k]] * jac[HB[k]];
  if (k < 27) return -ra[k - 21] * dotn;
  if (k < 33) return nn[NA[k - 27]] * nn[NB[k - 27]];
  if (k < 34) return cntv;
  return 0.0;
}

// Warp reduce-scatter of the values [K0, K0 + N) of every lane down to ONE finished sum per lane.
// Sizes per stage: N -> ceil(N/2) -> ... (5 stages, offsets 16, 8, 4, 2, 1); N <= 32.
template <int K0, int N>
__device__ __forceinline__ double bfly_reduce(int lane, const double* ra, const double* jac, const double* nn,
                                              double dotn, double cntv)
{
  constexpr int N1 = (N + 1) / 2, N2 = (N1 + 1) / 2, N3 = (N2 + 1) / 2, N4 = (N3 + 1) / 2;
  static_assert((N4 + 1) / 2 == 1, "five halvings must reach one value");
  double y1[N1], y2[N2], y3[N3], y4[N4], y5[1];
  {
    const bool bit = (lane & 16) != 0;
#pragma unroll
    for (int k = 0; k < N1; k++)
    {
      const double lo = xval(K0 + k, ra, jac, nn, dotn, cntv);
      const double hi = (k + N1 < N) ? xval(K0 + k + N1, ra, jac, nn, dotn, cntv) : 0.0;
      y1[k] = (bit ? hi : lo) + __shfl_xor_sync(0xffffffffu, bit ? lo : hi, 16);
    }
  }
  bfly_stage<N1, 8>(y1, y2, (lane & 8) != 0);
  bfly_stage<N2, 4>(y2, y3, (lane & 4) != 0);
  bfly_stage<N3, 2>(y3, y4, (lane & 2) != 0);
  bfly_stage<N4, 1>(y4, y5, (lane & 1) != 0);
  return y5[0];
}

// which value (relative to K0) a lane ends up owning after bfly_reduce<K0, N>; -1 = padding
template <int N>
__device__ __forceinline__ int bfly_owner(int lane)
{
  constexpr int N1 = (N + 1) / 2, N2 = (N1 + 1) / 2, N3 = (N2 + 1) / 2, N4 = (N3 + 1) / 2, N5 = (N4 + 1) / 2;
  int k = 0;
  bool ok = true;
  k = (lane & 1) ? N5 + k : k;
  ok = ok && k < N4;
  k = (lane & 2) ? N4 + k : k;
  ok = ok && k < N3;
  k = (lane & 4) ? N3 + k : k;
  ok = ok && k < N2;
  k = (lane & 8) ? N2 + k : k;
  ok = ok && k < N1;
  k = (lane & 16) ? N1 + k : k;
  ok = ok && k < N;
  return ok ? k : -1;
}

template <bool DEBUG>
__global__ void __launch_bounds__(IEKF_THREADS, IEKF_MIN_BLOCKS)
    k_iekf(ScanView scan, const int* __restrict__ n_ptr, int n_host, int* __restrict__ cache,
           const HashSlot* __restrict__ slots, unsigned int hmask, const NodeHot* __restrict__ hot,
           const NodeCold* __restrict__ cold, IekfParams prm, double* __restrict__ partials,
           unsigned int* __restrict__ ticket, double* __restrict__ result, IekfDebug dbg)
{
  const int n = n_ptr ? *n_ptr : n_host;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double t0 = 0.0, t1 = 0.0;  // the two finished sums this lane owns: one of the 21 HTH terms, one of the 13 others

  const int stride = gridDim.x * IEKF_THREADS;
  const int n_round = ((n + stride - 1) / stride) * stride;  // whole warps stay in the loop together
  for (int i = blockIdx.x * IEKF_THREADS + threadIdx.x; i < n_round; i += stride)
  {
    const bool live = i < n;
    const int ii = live ? i : 0;
    // independent loads first (memory-level parallelism): point, covariance, cached leaf
    const double pnt[3] = { __ldg(scan.p[0] + ii), __ldg(scan.p[1] + ii), __ldg(scan.p[2] + ii) };
    const double var6[6] = { __ldg(scan.v[0] + ii), __ldg(scan.v[1] + ii), __ldg(scan.v[2] + ii),
                             __ldg(scan.v[3] + ii), __ldg(scan.v[4] + ii), __ldg(scan.v[5] + ii) };
    const int cached = live ? cache[ii] : -1;
    double wld[3];
    rot_trans(prm.R, prm.p, pnt, wld);
    if (prm.variant & 16)  // experiment: streaming loads only
    {
      t0 += wld[0] + wld[1] + wld[2] + var6[0] + var6[1] + var6[2] + var6[3] + var6[4] + var6[5] + (double)cached;
      continue;
    }

    int node = -1;
    if (cached >= 0)
    {
      const NodeHot* h = hot + cached;
      const double vc[3] = { h->vcenter[0], h->vcenter[1], h->vcenter[2] };
      if (inside_box(wld, vc, h->ql)) node = cached;
    }
    long long kc[3] = { 0, 0, 0 };
    if (DEBUG || (live && node < 0))
    {
#pragma unroll
      for (int k = 0; k < 3; k++) kc[k] = voxel_coord(wld[k], prm.voxel_size);
    }
    if (live && node < 0)
    {
      unsigned long long key;
      if (pack_key(kc[0], kc[1], kc[2], &key))
      {
        unsigned int hh = hash_key(key) & hmask;
        for (unsigned int probe = 0; probe <= hmask; probe++)
        {
          const ulonglong2 s = __ldg(reinterpret_cast<const ulonglong2*>(slots + hh));
          if (s.x == key)
          {
            node = (int)(unsigned int)(s.y & 0xffffffffull);
            break;
          }
          if (s.x == VN_EMPTY_KEY) break;
          hh = (hh + 1) & hmask;
        }
      }
    }
    // descend to the leaf (octree.cpp:584-591)
    int flags = 0;
    while (node >= 0)
    {
      const NodeHot* h = hot + node;
      flags = h->flags;
      if (!(flags & VN_FLAG_INTERIOR)) break;
      const double vc[3] = { h->vcenter[0], h->vcenter[1], h->vcenter[2] };
      node = h->children[child_index(wld, vc)];
    }

    // contributions of this point; all zero unless the gate passes
    double ra[6] = { 0, 0, 0, 0, 0, 0 }, jac[6] = { 0, 0, 0, 0, 0, 0 }, nn[3] = { 0, 0, 0 };
    double dotn = 0.0, cntv = 0.0;
    int flag = 0;
    double sigma_l = 0.0;
    if (node >= 0 && (flags & VN_FLAG_PLANE) && !(prm.variant & 4))
    {
      const NodeHot* h = hot + node;
      const double c[3] = { h->center[0], h->center[1], h->center[2] };
      const double nr[3] = { h->normal[0], h->normal[1], h->normal[2] };
      const double d[3] = { ds(wld[0], c[0]), ds(wld[1], c[1]), ds(wld[2], c[2]) };
      const double dn = dot3(nr, d);
      const float dis_to_plane = (float)fabs(dn);
      const double e[3] = { ds(c[0], wld[0]), ds(c[1], wld[1]), ds(c[2], wld[2]) };
      const float dis_to_center = (float)dot3(e, e);
      const float range_dis = fs(dis_to_center, fm(dis_to_plane, dis_to_plane));
      if (range_dis <= fm(9.0f, h->radius))
      {
        // sigma_l = J plane_var J^T, J = [wld - center, -normal] = d^T A d - 2 d.(B n) + n^T C n (NodeHot)
        {
          const double A0 = h->qA[0], A1 = h->qA[1], A2 = h->qA[2], A3 = h->qA[3], A4 = h->qA[4], A5 = h->qA[5];
          const double dAd = d[0] * (A0 * d[0] + 2.0 * (A1 * d[1] + A2 * d[2])) + d[1] * (A3 * d[1] + 2.0 * A4 * d[2]) +
                             d[2] * A5 * d[2];
          sigma_l = dAd - 2.0 * (d[0] * h->qb[0] + d[1] * h->qb[1] + d[2] * h->qb[2]) + h->qk;
        }
        // + n^T var_world n
        double m[3];
#pragma unroll
        for (int k = 0; k < 3; k++) m[k] = prm.R[3 * k] * nr[0] + prm.R[3 * k + 1] * nr[1] + prm.R[3 * k + 2] * nr[2];
        const double q1 = m[0] * (var6[0] * m[0] + 2.0 * (var6[1] * m[1] + var6[2] * m[2])) +
                          m[1] * (var6[3] * m[1] + 2.0 * var6[4] * m[2]) + m[2] * var6[5] * m[2];
        const double u[3] = { nr[1] * pnt[2] - nr[2] * pnt[1], nr[2] * pnt[0] - nr[0] * pnt[2],
                              nr[0] * pnt[1] - nr[1] * pnt[0] };  // hat(p)^T n = n x p
        double q2 = 0.0, q3 = 0.0;
#pragma unroll
        for (int a = 0; a < 3; a++)
        {
          q2 += u[a] * (prm.rot_var[a] * u[0] + prm.rot_var[a + 3] * u[1] + prm.rot_var[a + 6] * u[2]);
          q3 += nr[a] * (prm.tsl_var[a] * nr[0] + prm.tsl_var[a + 3] * nr[1] + prm.tsl_var[a + 6] * nr[2]);
        }
        sigma_l += q1 + q2 + q3;
        // dis_to_plane < 3 sqrt(sigma_l)  <=>  dis_to_plane^2 < 9 sigma_l (the fp32 value squares exactly in fp64)
        if ((double)dis_to_plane * (double)dis_to_plane < 9.0 * sigma_l)
        {
          flag = 1;
          cache[ii] = node;  // oc = this (octree.cpp:571-575)
          const double Rinv = 1.0 / (0.0005 + sigma_l);
          // jac = [hat(p) R^T n ; n] = [p x m ; n]
          jac[0] = pnt[1] * m[2] - pnt[2] * m[1];
          jac[1] = pnt[2] * m[0] - pnt[0] * m[2];
          jac[2] = pnt[0] * m[1] - pnt[1] * m[0];
          jac[3] = nr[0];
          jac[4] = nr[1];
          jac[5] = nr[2];
#pragma unroll
          for (int a = 0; a < 6; a++) ra[a] = Rinv * jac[a];
          nn[0] = nr[0];
          nn[1] = nr[1];
          nn[2] = nr[2];
          dotn = dn;
          cntv = 1.0;
        }
      }
    }
    if (DEBUG && live)
    {
      dbg.keys[3 * (size_t)i + 0] = kc[0];
      dbg.keys[3 * (size_t)i + 1] = kc[1];
      dbg.keys[3 * (size_t)i + 2] = kc[2];
      dbg.flags[i] = (unsigned char)flag;
      dbg.codes[i] = flag ? (hot[node].layer | (cold[node].path << 2)) : -1;
      dbg.sigma[i] = flag ? sigma_l : 0.0;
    }
    // warp reduce-scatter: the 21 HTH terms (21 -> 11 -> 6 -> 3 -> 2 -> 1) and the 13 others (13 -> 7 -> 4 -> 2 -> 1 -> 1)
    if (!(prm.variant & 8))
    {
      t0 += bfly_reduce<0, 21>(lane, ra, jac, nn, dotn, cntv);
      t1 += bfly_reduce<21, 13>(lane, ra, jac, nn, dotn, cntv);
    }
    else
      t0 += ra[0] + dotn + cntv + nn[0] + jac[1];
  }

  // block reduction across the warps in fixed order
  __shared__ double sm[IEKF_WARPS][VN_IEKF_NACC];
  __shared__ bool is_last;
  const int k0 = bfly_owner<21>(lane), k1 = bfly_owner<13>(lane);
  if (k0 >= 0) sm[warp][k0] = t0;
  if (k1 >= 0) sm[warp][21 + k1] = t1;
  __syncthreads();
  const unsigned int nblk = gridDim.x;
  if (threadIdx.x < VN_IEKF_NACC)
  {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < IEKF_WARPS; w++) v += sm[w][threadIdx.x];
    partials[(size_t)threadIdx.x * nblk + blockIdx.x] = v;  // [34][nblk]: the final pass reads it coalesced
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == nblk - 1);
  }
  __syncthreads();
  if (!is_last) return;
  if (prm.variant & 1)
  {
    if (threadIdx.x == 0) *ticket = 0u;
    return;
  }
  // last block: one warp per column (34 columns over 32 warps), lanes stride over the <= 148 x IEKF_MIN_BLOCKS
  // per-block partials with independent loads, then a fixed shuffle tree -> deterministic
  __threadfence();
  __shared__ double fin[VN_IEKF_NACC];
  for (int k = warp; k < VN_IEKF_NACC; k += IEKF_WARPS)
  {
    const double* col = partials + (size_t)k * nblk;
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
  // the sums go straight to mapped pinned host memory; the sequence number is written last so that the
  // host can poll for completion instead of paying a stream synchronisation per IEKF iteration
  double* out = (prm.variant & 2) ? partials : result;
  if (threadIdx.x < VN_IEKF_NACC)
  {
    out[threadIdx.x] = fin[threadIdx.x];
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    *ticket = 0u;
    reinterpret_cast<volatile unsigned long long*>(out)[40] = prm.seq;
  }
}

__global__ void k_fill_int(int* p, int v, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int iekf_grid_blocks(int n, int sm_count)
{
  // one point per thread: every point of the scan is in flight at once (a B200 holds 148 x 4 x 256
  // = 151 552 threads of this kernel per wave; larger scans take a second wave)
  int need = (n + IEKF_THREADS - 1) / IEKF_THREADS;
  int cap = sm_count * IEKF_MIN_BLOCKS;  // one resident wave; larger scans loop
  if (need < 1) need = 1;
  return need < cap ? need : cap;
}

void launch_iekf(cudaStream_t st, const ScanView& scan, const int* n_dev, int n_host, int* cache, const MapView& map,
                 const IekfParams& prm, double* partials, unsigned int* ticket, double* result, int blocks,
                 const IekfDebug* dbg)
{
  if (dbg)
    k_iekf<true><<<blocks, IEKF_THREADS, 0, st>>>(scan, n_dev, n_host, cache, map.slots, map.hmask, map.hot, map.cold,
                                                   prm, partials, ticket, result, *dbg);
  else
  {
    IekfDebug none = { nullptr, nullptr, nullptr, nullptr };
    k_iekf<false><<<blocks, IEKF_THREADS, 0, st>>>(scan, n_dev, n_host, cache, map.slots, map.hmask, map.hot, map.cold,
                                                    prm, partials, ticket, result, none);
  }
}

void launch_fill_int(cudaStream_t st, int* p, int v, int n)
{
  if (n > 0) k_fill_int<<<(n + 255) / 256, 256, 0, st>>>(p, v, n);
}

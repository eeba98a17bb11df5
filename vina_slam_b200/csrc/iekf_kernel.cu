// The hot kernel: one IEKF iteration of VINA_SLAM::LioStateEstimation's point
// loop (src/pipeline/odometry.cpp:111-148): world point, cached-leaf test
// (OctoTree::inside, octree.cpp:732-737), voxel-hash lookup (match,
// voxel_map.cpp:241-266), octree descent and gate (OctoTree::match,
// octree.cpp:551-595), residual / Jacobian and the 6x6 H = J^T R^-1 J, b, n n^T
// reduction (odometry.cpp:136-146).
//
// Roofline: HBM-bound streaming of the pointVar SoA (72 B/pt + 8 B cache RMW)
// plus L2-resident gathers of 16-B hash slots and 256-B leaf records; fp64 FMA
// pipe is the secondary limiter (DESIGN.md "Kernels"). No dense contraction ->
// no tensor cores.
//
// Numerics: every decision-bearing expression (wld, key, child index, inside,
// the fp32 gate) uses the single-rounding helpers of vn_math.cuh in the order of
// SURVEY.md Appendix A, so keys and associations are bit-exact against the CPU
// restatement. sigma_l and the sums use FMA freely (tolerance 1e-4 rel):
// n^T var_world n is evaluated as (R^T n)^T var (R^T n) + (n x p)^T S_R (n x p)
// + n^T S_t n, which needs ~1/4 of the flops of forming var_world.
#include "vn_kernels.cuh"

#define IEKF_THREADS 256

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <bool DEBUG>
__global__ void __launch_bounds__(IEKF_THREADS, 2)
    k_iekf(ScanView scan, const int* __restrict__ n_ptr, int n_host, int* __restrict__ cache,
           const HashSlot* __restrict__ slots, unsigned int hmask, const NodeHot* __restrict__ hot,
           const NodeCold* __restrict__ cold, IekfParams prm, double* __restrict__ partials,
           unsigned int* __restrict__ ticket, double* __restrict__ result, IekfDebug dbg)
{
  const int n = n_ptr ? *n_ptr : n_host;
  double acc[VN_IEKF_NACC];
#pragma unroll
  for (int k = 0; k < VN_IEKF_NACC; k++) acc[k] = 0.0;

  const int stride = gridDim.x * IEKF_THREADS;
  for (int i = blockIdx.x * IEKF_THREADS + threadIdx.x; i < n; i += stride)
  {
    const double pnt[3] = { __ldg(scan.p[0] + i), __ldg(scan.p[1] + i), __ldg(scan.p[2] + i) };
    double wld[3];
    rot_trans(prm.R, prm.p, pnt, wld);

    int node = -1;
    const int cached = cache[i];
    if (cached >= 0)
    {
      const NodeHot* h = hot + cached;
      const double vc[3] = { h->vcenter[0], h->vcenter[1], h->vcenter[2] };
      if (inside_box(wld, vc, h->ql)) node = cached;
    }
    long long kc[3];
    if (DEBUG || node < 0)
    {
#pragma unroll
      for (int k = 0; k < 3; k++) kc[k] = voxel_coord(wld[k], prm.voxel_size);
    }
    if (node < 0)
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
      node = cold[node].children[child_index(wld, vc)];
    }

    int flag = 0;
    double sigma_l = 0.0;
    if (node >= 0 && (flags & VN_FLAG_PLANE))
    {
      const NodeHot* h = hot + node;
      const double c[3] = { h->center[0], h->center[1], h->center[2] };
      const double nr[3] = { h->normal[0], h->normal[1], h->normal[2] };
      const double d[3] = { ds(wld[0], c[0]), ds(wld[1], c[1]), ds(wld[2], c[2]) };
      const double dotn = dot3(nr, d);
      const float dis_to_plane = (float)fabs(dotn);
      const double e[3] = { ds(c[0], wld[0]), ds(c[1], wld[1]), ds(c[2], wld[2]) };
      const float dis_to_center = (float)dot3(e, e);
      const float range_dis = fs(dis_to_center, fm(dis_to_plane, dis_to_plane));
      if (range_dis <= fm(9.0f, h->radius))
      {
        // sigma_l = J plane_var J^T, J = [wld - center, -normal]
        const double J[6] = { d[0], d[1], d[2], -nr[0], -nr[1], -nr[2] };
        double s = 0.0;
        int q = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
        {
          double row = 0.5 * J[a] * h->pvar[q];  // diagonal counted once
          q++;
#pragma unroll
          for (int b = a + 1; b < 6; b++, q++) row += J[b] * h->pvar[q];
          s += J[a] * row;
        }
        sigma_l = 2.0 * s;
        // + n^T var_world n
        const double var6[6] = { __ldg(scan.v[0] + i), __ldg(scan.v[1] + i), __ldg(scan.v[2] + i),
                                 __ldg(scan.v[3] + i), __ldg(scan.v[4] + i), __ldg(scan.v[5] + i) };
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
        if ((double)dis_to_plane < 3.0 * sqrt(sigma_l))
        {
          flag = 1;
          cache[i] = node;  // oc = this (octree.cpp:571-575)
          const double Rinv = 1.0 / (0.0005 + sigma_l);
          // jac = [hat(p) R^T n ; n] = [p x m ; n]
          const double jac[6] = { pnt[1] * m[2] - pnt[2] * m[1], pnt[2] * m[0] - pnt[0] * m[2],
                                  pnt[0] * m[1] - pnt[1] * m[0], nr[0], nr[1], nr[2] };
          int t = 0;
#pragma unroll
          for (int a = 0; a < 6; a++)
          {
            const double ra = Rinv * jac[a];
#pragma unroll
            for (int b = a; b < 6; b++, t++) acc[t] += ra * jac[b];
            acc[21 + a] -= ra * dotn;
          }
          acc[27] += nr[0] * nr[0];
          acc[28] += nr[0] * nr[1];
          acc[29] += nr[0] * nr[2];
          acc[30] += nr[1] * nr[1];
          acc[31] += nr[1] * nr[2];
          acc[32] += nr[2] * nr[2];
          acc[33] += 1.0;
        }
      }
    }
    if (DEBUG)
    {
      dbg.keys[3 * (size_t)i + 0] = kc[0];
      dbg.keys[3 * (size_t)i + 1] = kc[1];
      dbg.keys[3 * (size_t)i + 2] = kc[2];
      dbg.flags[i] = (unsigned char)flag;
      dbg.codes[i] = flag ? (hot[node].layer | (cold[node].path << 2)) : -1;
      dbg.sigma[i] = flag ? sigma_l : 0.0;
    }
  }

  // block reduction: warp shuffle tree, then across the 8 warps in fixed order
  __shared__ double sm[IEKF_THREADS / 32][VN_IEKF_NACC];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < VN_IEKF_NACC; k++)
  {
    double v = warp_sum(acc[k]);
    if (lane == 0) sm[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < VN_IEKF_NACC)
  {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < IEKF_THREADS / 32; w++) v += sm[w][threadIdx.x];
    partials[(size_t)blockIdx.x * VN_IEKF_NACC + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0)
  {
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  // last block: sum the per-block partials in block order (deterministic)
  __threadfence();
  __shared__ double fin[7][VN_IEKF_NACC];
  if (threadIdx.x < 7 * VN_IEKF_NACC)
  {
    const int k = threadIdx.x % VN_IEKF_NACC, seg = threadIdx.x / VN_IEKF_NACC;
    double v = 0.0;
    for (unsigned int b = seg; b < gridDim.x; b += 7) v += __ldcg(partials + (size_t)b * VN_IEKF_NACC + k);
    fin[seg][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < VN_IEKF_NACC)
  {
    double v = 0.0;
#pragma unroll
    for (int s = 0; s < 7; s++) v += fin[s][threadIdx.x];
    result[threadIdx.x] = v;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

__global__ void k_fill_int(int* p, int v, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int iekf_grid_blocks(int n, int sm_count)
{
  int need = (n + IEKF_THREADS - 1) / IEKF_THREADS;
  int cap = sm_count * 2;  // two resident CTAs per SM (launch bounds), one wave
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

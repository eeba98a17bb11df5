// Scan front end on the device: what happens to a raw scan between the sensor driver and IMUEKF::process.
//   * the keep rule every decoder handler applies (src/sensor/lidar_pointcloud_decoder.cpp:68-71, 91-95, ...):
//     point i stays iff i % point_filter_num == 0 and x*x + y*y + z*z > blind (float products, compared as double);
//   * pcl_handler (src/sensor/lidar_decoder.cpp:16-34): sort by time offset (`curvature`), drop the tail beyond
//     0.11 s (the two-point dummy scan of :16-27 for an empty cloud is the host's job, vn_ctx.cu).
// The reference sorts with std::sort on the CPU (order of equal stamps unspecified); here it is a STABLE LSD radix
// sort over (key, index) pairs, 8 bits per pass: dropped points get the largest key and end up behind the kept
// ones, so filter, sort and cut are one sort. Output order = ascending time, ties in arrival order.
// Compiled -fmad=false (the keep rule is a decision: one rounding per operation like the strict oracle).
#include <cstdio>
#include "vn_kernels.cuh"

#define SORT_THREADS 256
#define SORT_ROUNDS 8
#define SORT_TILE (SORT_THREADS * SORT_ROUNDS)
#define KEY_DROPPED 0xFFFFFFFFu

// time offset -> unsigned key with the same order (handles negative offsets as well)
__device__ __forceinline__ unsigned int time_key(float t)
{
  const unsigned int u = __float_as_uint(t);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// keys + identity permutation; counters[0] = points kept by the decoder rule, counters[1] = kept and within 0.11 s
__global__ void __launch_bounds__(256) k_front_keys(const float4* __restrict__ raw, int n, int point_filter_num, double blind,
                                                    unsigned int* __restrict__ key, int* __restrict__ idx,
                                                    int* __restrict__ counters)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int kept = 0, inwin = 0;
  if (i < n)
  {
    const float4 p = raw[i];
    const float r2 = p.x * p.x + p.y * p.y + p.z * p.z;
    kept = ((i % point_filter_num) == 0 && (double)r2 > blind) ? 1 : 0;
    // `while (back().curvature > 0.11) pop_back()` on the sorted cloud == drop every stamp > 0.11 (float -> double)
    inwin = (kept && !((double)p.w > 0.11)) ? 1 : 0;
    key[i] = inwin ? time_key(p.w) : KEY_DROPPED;
    idx[i] = i;
  }
  const unsigned int bk = __ballot_sync(0xffffffffu, kept), bw = __ballot_sync(0xffffffffu, inwin);
  if ((threadIdx.x & 31) == 0)
  {
    if (bk) atomicAdd(&counters[0], __popc(bk));
    if (bw) atomicAdd(&counters[1], __popc(bw));
  }
}

// per-tile digit histogram, digit-major: hist[d * nb + tile]
__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const unsigned int* __restrict__ key, int n, int shift, int nb,
                                                            int* __restrict__ hist)
{
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * SORT_TILE;
  for (int r = 0; r < SORT_ROUNDS; r++)
  {
    const int i = base + r * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(key[i] >> shift) & 255u], 1);
  }
  __syncthreads();
  hist[threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of the 256 * nb counters in place (one block; the array is small: 256 ints per 2048 points)
__global__ void __launch_bounds__(1024) k_sort_scan(int* __restrict__ hist, int total)
{
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < total; base += 1024)
  {
    const int i = base + threadIdx.x;
    const int v = (i < total) ? hist[i] : 0;
    int x = v;
    for (int o = 1; o < 32; o <<= 1)
    {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    if (w == 0)
    {
      int s = warp_sums[lane];
      for (int o = 1; o < 32; o <<= 1)
      {
        const int y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      warp_sums[lane] = s;
    }
    __syncthreads();
    const int incl = x + (w > 0 ? warp_sums[w - 1] : 0) + carry;
    if (i < total) hist[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
}

// stable scatter of one tile: the tile is walked in index order (8 rounds of 256 consecutive items, warp after
// warp), every warp ranks its items among equal digits with match_any and advances the digit's output cursor
__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const unsigned int* __restrict__ key_in,
                                                               const int* __restrict__ idx_in, int n, int shift, int nb,
                                                               const int* __restrict__ offs_g,
                                                               unsigned int* __restrict__ key_out, int* __restrict__ idx_out)
{
  __shared__ int offs[256];
  offs[threadIdx.x] = offs_g[threadIdx.x * nb + blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int base = blockIdx.x * SORT_TILE;
  for (int r = 0; r < SORT_ROUNDS; r++)
  {
    const int i = base + r * SORT_THREADS + threadIdx.x;
    const bool valid = i < n;
    const unsigned int k = valid ? key_in[i] : 0u;
    const int id = valid ? idx_in[i] : 0;
    const unsigned int digit = (k >> shift) & 255u;
    for (int w = 0; w < SORT_THREADS / 32; w++)
    {
      if (warp == w)
      {
        // lanes past the end get values no digit can take, each its own
        const unsigned int peers = __match_any_sync(0xffffffffu, valid ? digit : 256u + (unsigned int)lane);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int leader = __ffs(peers) - 1;
        int pos = 0;
        if (valid) pos = offs[digit];
        __syncwarp();
        if (valid && lane == leader) offs[digit] = pos + __popc(peers);
        __syncwarp();
        if (valid)
        {
          key_out[pos + rank] = k;
          idx_out[pos + rank] = id;
        }
      }
      __syncthreads();
    }
  }
}

// the sorted scan: out[j] = raw[idx[j]] for the n_keep points in front; result[0] = time offset of the last one
__global__ void __launch_bounds__(256) k_front_gather(const float4* __restrict__ raw, const int* __restrict__ idx,
                                                      const int* __restrict__ counters, float4* __restrict__ out,
                                                      float* __restrict__ t_last)
{
  const int n_keep = counters[1];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_keep) return;
  const float4 p = raw[idx[j]];
  out[j] = p;
  if (j == n_keep - 1) *t_last = p.w;
}

// ---------------------------------------------------------------------------
// The fast path: the stamps of a scan lie in [0, 0.11] s, so ONE bucket pass by time (1024 buckets of ~0.1 ms, a
// monotone function of the stamp) leaves a few hundred points per bucket, and each bucket is finished by a bitonic
// sort in shared memory over the 64-bit pairs (time key, arrival index) - unique values, so any sort gives the stable
// order. Three launches and 56 B of traffic per point instead of thirteen launches; the filter and the 0.11 s cut are
// still "dropped points never enter a bucket". A bucket that does not fit shared memory (every point carrying the same
// stamp, say) raises counters[3] and the caller falls back to the radix passes below.
#define FB_BUCKETS 1024
#define FB_CAP 4096  // pairs per bucket the shared-memory sort takes
#define FB_TILE 512  // points per block of the scatter pass
__device__ __forceinline__ int time_bucket(float t)
{
  const float x = t * (float)(FB_BUCKETS / 0.11);  // monotone in t (one rounding), like the conversion below
  int b = (int)x;
  if (!(x > 0.0f)) b = 0;
  return b > FB_BUCKETS - 1 ? FB_BUCKETS - 1 : b;
}

__global__ void __launch_bounds__(256) k_front_count(const float4* __restrict__ raw, int n, int point_filter_num, double blind,
                                                     int* __restrict__ bkt_of, int* __restrict__ counts,
                                                     int* __restrict__ counters)
{
  vn_pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int kept = 0, inwin = 0;
  if (i < n)
  {
    const float4 p = raw[i];
    const float r2 = p.x * p.x + p.y * p.y + p.z * p.z;
    kept = ((i % point_filter_num) == 0 && (double)r2 > blind) ? 1 : 0;
    inwin = (kept && !((double)p.w > 0.11)) ? 1 : 0;
    const int b = inwin ? time_bucket(p.w) : -1;
    bkt_of[i] = b;
    if (inwin) atomicAdd(&counts[b], 1);
  }
  const unsigned int bk = __ballot_sync(0xffffffffu, kept), bw = __ballot_sync(0xffffffffu, inwin);
  if ((threadIdx.x & 31) == 0)
  {
    if (bk) atomicAdd(&counters[0], __popc(bk));
    if (bw) atomicAdd(&counters[1], __popc(bw));
  }
}

// exclusive scan of the FB_BUCKETS counts into shared memory (256 threads, 4 consecutive counts each)
__device__ __forceinline__ void bucket_offsets(const int* __restrict__ counts, int* offs, int* wsum)
{
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int4 c = reinterpret_cast<const int4*>(counts)[t];
  const int mine = c.x + c.y + c.z + c.w;
  int x = mine;
  for (int o = 1; o < 32; o <<= 1)
  {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[w] = x;
  __syncthreads();
  int before = x - mine;
  for (int k = 0; k < w; k++) before += wsum[k];
  offs[4 * t] = before;
  offs[4 * t + 1] = before + c.x;
  offs[4 * t + 2] = before + c.x + c.y;
  offs[4 * t + 3] = before + c.x + c.y + c.z;
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_front_scatter(const float4* __restrict__ raw, int n, const int* __restrict__ bkt_of,
                                                       const int* __restrict__ counts, int* __restrict__ cursors,
                                                       unsigned long long* __restrict__ pairs, int* __restrict__ offs_g)
{
  vn_pdl_sync();
  __shared__ int offs[FB_BUCKETS];
  __shared__ int wsum[8];
  bucket_offsets(counts, offs, wsum);
  if (blockIdx.x == 0)  // the sort kernel's blocks read their first output position from here
    for (int k = threadIdx.x; k < FB_BUCKETS; k += 256) offs_g[k] = offs[k];
  const int base = blockIdx.x * FB_TILE;
  for (int r = 0; r < FB_TILE / 256; r++)
  {
    const int i = base + r * 256 + threadIdx.x;
    if (i >= n) break;
    const int b = bkt_of[i];
    if (b < 0) continue;
    const int pos = offs[b] + atomicAdd(&cursors[b], 1);  // (any order inside the bucket: the pair sort below fixes it)
    pairs[pos] = ((unsigned long long)time_key(raw[i].w) << 32) | (unsigned int)i;
  }
}

// counters: [0] kept by the decoder rule, [1] kept and within 0.11 s, [2] time offset of the last point (float bits),
// [3] a bucket overflowed, [4] ticket of the finishing blocks. The block that finishes last copies [0..3] to mapped
// host memory and then writes the sequence number the host polls: no copy, no stream synchronisation.
__global__ void __launch_bounds__(256) k_front_sort(const float4* __restrict__ raw, const unsigned long long* __restrict__ pairs,
                                                    const int* __restrict__ counts, const int* __restrict__ offs_g,
                                                    int* __restrict__ counters, float4* __restrict__ out,
                                                    volatile unsigned long long* __restrict__ pub, unsigned long long seq)
{
  vn_pdl_sync();
  __shared__ unsigned long long s[FB_CAP];
  const int b = blockIdx.x, t = threadIdx.x;
  const int nb = counts[b];
  if (nb > FB_CAP)
  {
    if (t == 0) counters[3] = 1;
  }
  else if (nb > 0)
  {
    const int off = offs_g[b];
    int P = 32;
    while (P < nb) P <<= 1;
    for (int i = t; i < P; i += 256) s[i] = i < nb ? pairs[off + i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1)
      {
        for (int i = t; i < P; i += 256)
        {
          const int x = i ^ j;
          if (x > i)
          {
            const unsigned long long a = s[i], c = s[x];
            if ((a > c) == ((i & k) == 0))
            {
              s[i] = c;
              s[x] = a;
            }
          }
        }
        __syncthreads();
      }
    const int n_keep = counters[1];
    for (int i = t; i < nb; i += 256)
    {
      const float4 p = raw[(int)(unsigned int)(s[i] & 0xffffffffull)];
      out[off + i] = p;
      if (off + i == n_keep - 1) counters[2] = __float_as_int(p.w);
    }
  }
  __syncthreads();
  if (t == 0)
  {
    __threadfence();
    if (atomicAdd(&counters[4], 1) == (int)gridDim.x - 1)
    {
      __threadfence();
      for (int k = 0; k < 4; k++) pub[1 + k] = (unsigned long long)(unsigned int)__ldcg(counters + k);
      __threadfence_system();
      pub[0] = seq;
    }
  }
}

// work = 3 * FB_BUCKETS + 8 ints: bucket counts, cursors, first positions, counters (one memset clears what must be zero)
int launch_front_prepare_buckets(cudaStream_t st, const float4* raw, int n, int point_filter_num, double blind, int* bkt_of,
                                 int* work, unsigned long long* pairs, float4* out, unsigned long long* pub_mapped,
                                 unsigned long long seq)
{
  const int nb = (n + FB_TILE - 1) / FB_TILE;
  int* counts = work;
  int* cursors = work + FB_BUCKETS;
  int* counters = work + 2 * FB_BUCKETS;
  int* offs_g = work + 2 * FB_BUCKETS + 8;
  cudaMemsetAsync(work, 0, (2 * FB_BUCKETS + 8) * sizeof(int), st);
  vn_launch(k_front_count, dim3((n + 255) / 256), dim3(256), 0, st, raw, n, point_filter_num, blind, bkt_of, counts, counters);
  vn_launch(k_front_scatter, dim3(nb), dim3(256), 0, st, raw, n, bkt_of, counts, cursors, pairs, offs_g);
  vn_launch(k_front_sort, dim3(FB_BUCKETS), dim3(256), 0, st, raw, pairs, counts, offs_g, counters, out, pub_mapped, seq);
  return 3;
}

int launch_front_prepare(cudaStream_t st, const float4* raw, int n, int point_filter_num, double blind, unsigned int* key[2],
                         int* idx[2], int* hist, int* counters, float4* out, float* t_last)
{
  const int nb = (n + SORT_TILE - 1) / SORT_TILE;
  cudaMemsetAsync(counters, 0, 2 * sizeof(int), st);
  k_front_keys<<<(n + 255) / 256, 256, 0, st>>>(raw, n, point_filter_num, blind, key[0], idx[0], counters);
  int launches = 1, cur = 0;
  for (int shift = 0; shift < 32; shift += 8, cur ^= 1)
  {
    k_sort_hist<<<nb, SORT_THREADS, 0, st>>>(key[cur], n, shift, nb, hist);
    k_sort_scan<<<1, 1024, 0, st>>>(hist, 256 * nb);
    k_sort_scatter<<<nb, SORT_THREADS, 0, st>>>(key[cur], idx[cur], n, shift, nb, hist, key[cur ^ 1], idx[cur ^ 1]);
    launches += 3;
  }
  // four passes: the result is back in buffer 0
  k_front_gather<<<(n + 255) / 256, 256, 0, st>>>(raw, idx[0], counters, out, t_last);
  return launches + 1;
}

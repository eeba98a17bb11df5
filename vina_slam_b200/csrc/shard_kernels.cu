// Voxel-hash-range sharding of the map across GPUs (BASELINE.json configs[4]).
//
// A root voxel - and with it every leaf, cluster sum and plane below it - belongs to exactly one rank:
//   owner(key) = floor(hash(key) * world / 2^32)          (contiguous ranges of the 32-bit voxel hash)
// because a leaf's accumulators depend only on the points of its own root voxel (voxel_map.cpp:86,
// 108-134). Building the map therefore needs ONE exchange per scan: every rank turns its share of the scan
// into world-frame point records (pvec_update, point_utils.cpp:54-65), partitions them by owner, and the
// records travel to their owners (all-to-all over NVLink); insert / recut / margi then run locally and
// unchanged on what arrived.
//
// The partition is STABLE (records of one owner keep the scan order) and ranks hold ascending slices of the
// scan, so the concatenation an owner receives is in ascending scan order - the order the reference's
// push_back(i) loop uses (voxel_map.cpp:86). Every cluster sum of the sharded map is therefore bit-identical
// to the single-GPU map's, which tests/test_gpu_parity.py asserts on the union of the shards.
//
// Compiled with -fmad=false like map_kernels.cu: pw / vw must be the bits k_insert_root would compute.
#include "vn_kernels.cuh"

#define SH_THREADS 256
#define SH_WARPS (SH_THREADS / 32)

struct ShardCov
{
  double rot[9], tsl[9];
};

// pass 1: owner of every point of the slice + per-block histogram
__global__ void __launch_bounds__(SH_THREADS)
    k_shard_count(ScanView scan, int first, int count, PoseD x, double voxel_size, int world,
                  unsigned char* __restrict__ owner, int* __restrict__ hist, int* __restrict__ status)
{
  __shared__ int h[VN_MAX_WORLD];
  if (threadIdx.x < VN_MAX_WORLD) h[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * SH_THREADS + threadIdx.x;
  if (i < count)
  {
    const int s = first + i;
    const double pnt[3] = { scan.p[0][s], scan.p[1][s], scan.p[2][s] };
    double pw[3];
    rot_trans(x.R, x.p, pnt, pw);
    long long kc[3];
    for (int k = 0; k < 3; k++) kc[k] = voxel_coord(pw[k], voxel_size);
    unsigned long long key;
    int ow = 0;
    if (pack_key(kc[0], kc[1], kc[2], &key))
      ow = shard_owner(key, world);
    else
      atomicOr(status, VN_ST_KEY_RANGE);
    owner[i] = (unsigned char)ow;
    atomicAdd(&h[ow], 1);
  }
  __syncthreads();
  if (threadIdx.x < world) hist[blockIdx.x * world + threadIdx.x] = h[threadIdx.x];
}

// pass 2 (one block, warp w = owner w): exclusive scan of the histogram over the blocks; segment starts
__global__ void __launch_bounds__(32 * VN_MAX_WORLD)
    k_shard_offsets(int* __restrict__ hist, int nblk, int world, int* __restrict__ counts, int* __restrict__ starts)
{
  __shared__ int tot[VN_MAX_WORLD];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (w < world)
  {
    int run = 0;
    for (int b0 = 0; b0 < nblk; b0 += 32)
    {
      const int b = b0 + lane;
      const int v = b < nblk ? hist[b * world + w] : 0;
      int xs = v;
      for (int o = 1; o < 32; o <<= 1)
      {
        const int y = __shfl_up_sync(0xffffffffu, xs, o);
        if (lane >= o) xs += y;
      }
      if (b < nblk) hist[b * world + w] = run + xs - v;
      run += __shfl_sync(0xffffffffu, xs, 31);
    }
    if (lane == 0) tot[w] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    int s = 0;
    for (int k = 0; k < world; k++)
    {
      counts[k] = tot[k];
      starts[k] = s;
      s += tot[k];
    }
    starts[world] = s;
  }
}

// pass 3: stable scatter of the records into the owner segments. QUERY == false: map-build records, 13 doubles
// (pvec_update: body point, world covariance, world point, scan index). QUERY == true: association records,
// 10 doubles (body point, body covariance, scan index) - the owner evaluates the gate itself.
template <bool QUERY>
__global__ void __launch_bounds__(SH_THREADS)
    k_shard_scatter(ScanView scan, int first, int count, PoseD x, ShardCov cv, int world,
                    const unsigned char* __restrict__ owner, const int* __restrict__ hist,
                    const int* __restrict__ starts, double* __restrict__ out, long long gidx_base)
{
  __shared__ int wcnt[SH_WARPS][VN_MAX_WORLD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * SH_THREADS + threadIdx.x;
  const int ow = i < count ? (int)owner[i] : -1;
  int myrank = 0;
  for (int w = 0; w < world; w++)
  {
    const unsigned int m = __ballot_sync(0xffffffffu, ow == w);
    if (ow == w) myrank = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wcnt[warp][w] = __popc(m);
  }
  __syncthreads();
  if (ow < 0) return;
  int base = 0;
  for (int ww = 0; ww < warp; ww++) base += wcnt[ww][ow];
  const size_t pos = (size_t)starts[ow] + hist[blockIdx.x * world + ow] + base + myrank;
  const int s = first + i;
  const double pnt[3] = { scan.p[0][s], scan.p[1][s], scan.p[2][s] };
  double var6[6];
  for (int k = 0; k < 6; k++) var6[k] = scan.v[k][s];
  if (QUERY)
  {
    double* r = out + pos * VINA_SHARD_QUERY_DOUBLES;
    for (int k = 0; k < 3; k++) r[k] = pnt[k];
    for (int k = 0; k < 6; k++) r[3 + k] = var6[k];
    reinterpret_cast<long long*>(r)[9] = gidx_base + s;
    return;
  }
  double pw[3], vw[6];
  rot_trans(x.R, x.p, pnt, pw);
  world_var(x.R, pnt, var6, cv.rot, cv.tsl, vw);
  double* r = out + pos * VINA_SHARD_RECORD_DOUBLES;
  for (int k = 0; k < 3; k++) r[k] = pnt[k];
  for (int k = 0; k < 6; k++) r[3 + k] = vw[k];
  for (int k = 0; k < 3; k++) r[9 + k] = pw[k];
  reinterpret_cast<long long*>(r)[12] = gidx_base + s;
}

// received association records -> the pointVar SoA the accumulate kernel reads
__global__ void __launch_bounds__(SH_THREADS) k_shard_unpack_query(const double* __restrict__ rec, int n, ScanView scan)
{
  const int i = blockIdx.x * SH_THREADS + threadIdx.x;
  if (i >= n) return;
  const double* r = rec + (size_t)i * VINA_SHARD_QUERY_DOUBLES;
  for (int k = 0; k < 3; k++) scan.p[k][i] = r[k];
  for (int k = 0; k < 6; k++) scan.v[k][i] = r[3 + k];
}

// received records -> the SoA buffers the insert kernels read (body point, world point, world covariance)
__global__ void __launch_bounds__(SH_THREADS)
    k_shard_unpack(const double* __restrict__ rec, int n, ScanView scan, InsertScratch sc)
{
  const int i = blockIdx.x * SH_THREADS + threadIdx.x;
  if (i >= n) return;
  const double* r = rec + (size_t)i * VINA_SHARD_RECORD_DOUBLES;
  for (int k = 0; k < 3; k++) scan.p[k][i] = r[k];
  for (int k = 0; k < 6; k++) sc.vw[k][i] = r[3 + k];
  for (int k = 0; k < 3; k++) sc.pw[k][i] = r[9 + k];
}

// ---------------------------------------------------------------------------
// Fused route + exchange over peer memory (NVLink). Instead of staging the records for an NCCL all-to-all, the
// scatter kernel stores every record straight into its owner's inbox (a peer pointer: CUDA IPC between the
// per-GPU processes), at the position it has in the owner's scan-ordered receive sequence:
//   position = sum_{r < me} counts_r[owner] + (stable position among my records for that owner).
// The counts table is exchanged the same way: every rank stores its row into every peer's control block and
// raises that peer's `ready` flag; after its records a rank raises the peers' `done` flags. Flags carry the scan
// epoch, so nothing is ever reset. All of it is enqueued without a host synchronisation.
#define P2P_SPIN_LIMIT 20000000ll  // ~ seconds of polling a local flag: a lost peer is reported (VN_ST_SPIN), never a hang

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long* p)
{
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

// my counts row -> every peer's table, then the peers' ready flags (one warp; lane = peer)
__global__ void __launch_bounds__(32) k_p2p_publish(ShardPeers peers, const int* __restrict__ counts, unsigned long long epoch,
                                                    int chan, const IekfDev* __restrict__ gate)
{
  if (gate && gate->done) return;  // the sharded IEKF has converged (on every rank alike)
  const int w = threadIdx.x;
  if (w < peers.world)
  {
    ShardChan* pc = &peers.ctrl[w]->ch[chan];
    for (int k = 0; k < peers.world; k++) *reinterpret_cast<volatile int*>(&pc->counts[peers.rank][k]) = counts[k];
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&pc->ready[peers.rank]) = epoch;
  }
}

// wait for every rank's row in MY table and derive my base offset at every owner. Only this single warp ever
// spins, so a waiting rank cannot starve the kernels it waits for.
__global__ void __launch_bounds__(32)
    k_p2p_base(ShardPeers peers, unsigned long long epoch, long long* __restrict__ base, int* __restrict__ status, int chan,
               const IekfDev* __restrict__ gate)
{
  if (gate && gate->done) return;
  const int w = threadIdx.x;
  const ShardChan* me = &peers.ctrl[peers.rank]->ch[chan];
  if (w < peers.world)
  {
    long long spins = 0;
    while (ld_flag(&me->ready[w]) != epoch)
      if (++spins > P2P_SPIN_LIMIT)
      {
        atomicOr(status, VN_ST_SPIN);
        break;
      }
  }
  __syncwarp();
  __threadfence_system();
  if (w < peers.world)
  {
    long long b = 0;
    for (int r = 0; r < peers.rank; r++) b += *reinterpret_cast<const volatile int*>(&me->counts[r][w]);
    base[w] = b;
  }
}

__global__ void __launch_bounds__(SH_THREADS)
    k_p2p_scatter(ScanView scan, int first, int count, PoseD x, ShardCov cv, ShardPeers peers,
                  const unsigned char* __restrict__ owner, const int* __restrict__ hist,
                  const long long* __restrict__ base, long long gidx_base, long long inbox_cap, int* __restrict__ status)
{
  __shared__ int wcnt[SH_WARPS][VN_MAX_WORLD];
  const int world = peers.world;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * SH_THREADS + threadIdx.x;
  const int ow = i < count ? (int)owner[i] : -1;
  int myrank = 0;
  for (int w = 0; w < world; w++)
  {
    const unsigned int m = __ballot_sync(0xffffffffu, ow == w);
    if (ow == w) myrank = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wcnt[warp][w] = __popc(m);
  }
  __syncthreads();
  if (ow >= 0)
  {
    int inblk = 0;
    for (int ww = 0; ww < warp; ww++) inblk += wcnt[ww][ow];
    const long long pos = base[ow] + hist[blockIdx.x * world + ow] + inblk + myrank;
    if (pos >= inbox_cap)
      atomicOr(status, VN_ST_WIN_FULL);
    else
    {
      const int s = first + i;
      const double pnt[3] = { scan.p[0][s], scan.p[1][s], scan.p[2][s] };
      double var6[6], pw[3], vw[6];
      for (int k = 0; k < 6; k++) var6[k] = scan.v[k][s];
      rot_trans(x.R, x.p, pnt, pw);
      world_var(x.R, pnt, var6, cv.rot, cv.tsl, vw);
      double* r = peers.inbox[ow] + pos * VINA_SHARD_RECORD_DOUBLES;  // peer store over NVLink
      for (int k = 0; k < 3; k++) r[k] = pnt[k];
      for (int k = 0; k < 6; k++) r[3 + k] = vw[k];
      for (int k = 0; k < 3; k++) r[9 + k] = pw[k];
      reinterpret_cast<long long*>(r)[12] = gidx_base + s;
    }
  }
  __threadfence_system();
}

// my records have landed everywhere: raise the peers' done flags
__global__ void __launch_bounds__(32) k_p2p_signal(ShardPeers peers, unsigned long long epoch, int chan,
                                                   const IekfDev* __restrict__ gate)
{
  if (gate && gate->done) return;
  __threadfence_system();
  if (threadIdx.x < peers.world)
    *reinterpret_cast<volatile unsigned long long*>(&peers.ctrl[threadIdx.x]->ch[chan].done[peers.rank]) = epoch;
}

// wait for every rank's records in MY inbox; the number of records received
// (clamped to the inbox capacity: the senders count every record they route, including the ones an overflowing
// inbox dropped - the receiver must neither read past the records that exist nor report success)
__global__ void __launch_bounds__(32) k_p2p_wait(ShardPeers peers, unsigned long long epoch, int* __restrict__ n_recv,
                                                 int* __restrict__ status, int chan, const IekfDev* __restrict__ gate,
                                                 int cap)
{
  if (gate && gate->done)
  {
    if (threadIdx.x == 0) *n_recv = 0;
    return;
  }
  const ShardChan* me = &peers.ctrl[peers.rank]->ch[chan];
  int mine = 0;
  if (threadIdx.x < peers.world)
  {
    long long spins = 0;
    while (ld_flag(&me->done[threadIdx.x]) != epoch)
      if (++spins > P2P_SPIN_LIMIT)
      {
        atomicOr(status, VN_ST_SPIN);
        break;
      }
    __threadfence_system();
    mine = *reinterpret_cast<const volatile int*>(&me->counts[threadIdx.x][peers.rank]);
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if (threadIdx.x == 0)
  {
    if (mine > cap)
    {
      atomicOr(status, VN_ST_WIN_FULL);
      mine = cap;
    }
    *n_recv = mine;
  }
}

__global__ void __launch_bounds__(SH_THREADS)
    k_shard_unpack_n(const double* __restrict__ rec, const int* __restrict__ n_ptr, ScanView scan, InsertScratch sc)
{
  const int i = blockIdx.x * SH_THREADS + threadIdx.x;
  if (i >= *n_ptr) return;
  const double* r = rec + (size_t)i * VINA_SHARD_RECORD_DOUBLES;
  for (int k = 0; k < 3; k++) scan.p[k][i] = r[k];
  for (int k = 0; k < 6; k++) sc.vw[k][i] = r[3 + k];
  for (int k = 0; k < 3; k++) sc.pw[k][i] = r[9 + k];
}

int launch_shard_route_p2p(cudaStream_t st, const ScanView& scan, int first, int count, const PoseD& x,
                           const double* rot_var, const double* tsl_var, double voxel_size, const ShardPeers& peers,
                           unsigned char* owner, int* hist, int* counts, int* starts, long long* base,
                           unsigned long long epoch, long long gidx_base, long long inbox_cap, int* status, int phase)
{
  ShardCov cv;
  for (int k = 0; k < 9; k++) cv.rot[k] = rot_var[k], cv.tsl[k] = tsl_var[k];
  const int nblk = count <= 0 ? 1 : (count + SH_THREADS - 1) / SH_THREADS;
  int launches = 0;
  if (phase != 2)
  {
    // phase A never waits: owners, counts, my row to every peer
    k_shard_count<<<nblk, SH_THREADS, 0, st>>>(scan, first, count, x, voxel_size, peers.world, owner, hist, status);
    k_shard_offsets<<<1, 32 * VN_MAX_WORLD, 0, st>>>(hist, nblk, peers.world, counts, starts);
    k_p2p_publish<<<1, 32, 0, st>>>(peers, counts, epoch, VN_CHAN_BUILD, nullptr);
    launches += 3;
  }
  if (phase != 1)
  {
    // phase B waits for the peers' rows (one spinning warp), then stores the records and signals
    k_p2p_base<<<1, 32, 0, st>>>(peers, epoch, base, status, VN_CHAN_BUILD, nullptr);
    k_p2p_scatter<<<nblk, SH_THREADS, 0, st>>>(scan, first, count, x, cv, peers, owner, hist, base, gidx_base, inbox_cap,
                                               status);
    k_p2p_signal<<<1, 32, 0, st>>>(peers, epoch, VN_CHAN_BUILD, nullptr);
    launches += 3;
  }
  return launches;
}

int launch_shard_recv_p2p(cudaStream_t st, const ShardPeers& peers, unsigned long long epoch, int* n_recv, int cap,
                          const ScanView& scan, const InsertScratch& sc, int* status)
{
  k_p2p_wait<<<1, 32, 0, st>>>(peers, epoch, n_recv, status, VN_CHAN_BUILD, nullptr, cap);
  k_shard_unpack_n<<<(cap + SH_THREADS - 1) / SH_THREADS, SH_THREADS, 0, st>>>(peers.inbox[peers.rank], n_recv, scan, sc);
  return 2;
}

// ---- the association query over the same fused exchange (pose and convergence flag from the device iterate) ----
__global__ void __launch_bounds__(SH_THREADS)
    k_shard_count_q(ScanView scan, int first, int count, const IekfDev* __restrict__ it, double voxel_size, int world,
                    unsigned char* __restrict__ owner, int* __restrict__ hist, int* __restrict__ status)
{
  if (it->done) return;
  __shared__ int h[VN_MAX_WORLD];
  __shared__ double R[9], p[3];
  if (threadIdx.x < VN_MAX_WORLD) h[threadIdx.x] = 0;
  if (threadIdx.x < 9) R[threadIdx.x] = it->R[threadIdx.x];
  if (threadIdx.x < 3) p[threadIdx.x] = it->p[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * SH_THREADS + threadIdx.x;
  if (i < count)
  {
    const int s = first + i;
    const double pnt[3] = { scan.p[0][s], scan.p[1][s], scan.p[2][s] };
    double pw[3];
    rot_trans(R, p, pnt, pw);
    long long kc[3];
    for (int k = 0; k < 3; k++) kc[k] = voxel_coord(pw[k], voxel_size);
    unsigned long long key;
    int ow = 0;
    if (pack_key(kc[0], kc[1], kc[2], &key)) ow = shard_owner(key, world);  // (a key out of range matches nothing anywhere)
    owner[i] = (unsigned char)ow;
    atomicAdd(&h[ow], 1);
  }
  __syncthreads();
  if (threadIdx.x < world) hist[blockIdx.x * world + threadIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(SH_THREADS)
    k_p2p_scatter_q(ScanView scan, int first, int count, ShardPeers peers, const unsigned char* __restrict__ owner,
                    const int* __restrict__ hist, const long long* __restrict__ base, long long gidx_base, long long inbox_cap,
                    int* __restrict__ status, const IekfDev* __restrict__ it)
{
  if (it->done) return;
  __shared__ int wcnt[SH_WARPS][VN_MAX_WORLD];
  const int world = peers.world;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * SH_THREADS + threadIdx.x;
  const int ow = i < count ? (int)owner[i] : -1;
  int myrank = 0;
  for (int w = 0; w < world; w++)
  {
    const unsigned int m = __ballot_sync(0xffffffffu, ow == w);
    if (ow == w) myrank = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wcnt[warp][w] = __popc(m);
  }
  __syncthreads();
  if (ow >= 0)
  {
    int inblk = 0;
    for (int ww = 0; ww < warp; ww++) inblk += wcnt[ww][ow];
    const long long pos = base[ow] + hist[blockIdx.x * world + ow] + inblk + myrank;
    if (pos >= inbox_cap)
      atomicOr(status, VN_ST_WIN_FULL);
    else
    {
      const int s = first + i;
      // the query region follows the map-build region in the owner's inbox
      double* r = peers.inbox[ow] + inbox_cap * VINA_SHARD_RECORD_DOUBLES + pos * VINA_SHARD_QUERY_DOUBLES;
      for (int k = 0; k < 3; k++) r[k] = scan.p[k][s];
      for (int k = 0; k < 6; k++) r[3 + k] = scan.v[k][s];
      reinterpret_cast<long long*>(r)[9] = gidx_base + s;
    }
  }
  __threadfence_system();
}

__global__ void __launch_bounds__(SH_THREADS)
    k_shard_unpack_query_n(const double* __restrict__ rec, const int* __restrict__ n_ptr, ScanView scan)
{
  const int i = blockIdx.x * SH_THREADS + threadIdx.x;
  if (i >= *n_ptr) return;
  const double* r = rec + (size_t)i * VINA_SHARD_QUERY_DOUBLES;
  for (int k = 0; k < 3; k++) scan.p[k][i] = r[k];
  for (int k = 0; k < 6; k++) scan.v[k][i] = r[3 + k];
}

// one iteration's routing of the association queries: phase 1 never waits (owners, counts, my row to the peers),
// phase 2 waits for the peers' rows, stores the queries into the owners' inboxes and signals, phase 3 waits for
// the peers' queries and unpacks them into `recv_set`. phase 0 = all three.
int launch_shard_query_p2p(cudaStream_t st, const ScanView& scan, int first, int count, const IekfDev* it,
                           double voxel_size, const ShardPeers& peers, unsigned char* owner, int* hist, int* counts,
                           int* starts, long long* base, unsigned long long epoch, long long inbox_cap, int* n_recv,
                           const ScanView& recv_set, int* status, int phase)
{
  const int nblk = count <= 0 ? 1 : (count + SH_THREADS - 1) / SH_THREADS;
  int launches = 0;
  if (phase == 0 || phase == 1)
  {
    k_shard_count_q<<<nblk, SH_THREADS, 0, st>>>(scan, first, count, it, voxel_size, peers.world, owner, hist, status);
    k_shard_offsets<<<1, 32 * VN_MAX_WORLD, 0, st>>>(hist, nblk, peers.world, counts, starts);
    k_p2p_publish<<<1, 32, 0, st>>>(peers, counts, epoch, VN_CHAN_QUERY, it);
    launches += 3;
  }
  if (phase == 0 || phase == 2)
  {
    k_p2p_base<<<1, 32, 0, st>>>(peers, epoch, base, status, VN_CHAN_QUERY, it);
    k_p2p_scatter_q<<<nblk, SH_THREADS, 0, st>>>(scan, first, count, peers, owner, hist, base, 0, inbox_cap, status, it);
    k_p2p_signal<<<1, 32, 0, st>>>(peers, epoch, VN_CHAN_QUERY, it);
    launches += 3;
  }
  if (phase == 0 || phase == 3)
  {
    k_p2p_wait<<<1, 32, 0, st>>>(peers, epoch, n_recv, status, VN_CHAN_QUERY, it, (int)inbox_cap);
    const int cap = (int)inbox_cap;
    k_shard_unpack_query_n<<<(cap + SH_THREADS - 1) / SH_THREADS, SH_THREADS, 0, st>>>(
        peers.inbox[peers.rank] + inbox_cap * VINA_SHARD_RECORD_DOUBLES, n_recv, recv_set);
    launches += 2;
  }
  return launches;
}

int launch_shard_route(cudaStream_t st, const ScanView& scan, int first, int count, const PoseD& x,
                       const double* rot_var, const double* tsl_var, double voxel_size, int world,
                       unsigned char* owner, int* hist, int* counts, int* starts, double* out, long long gidx_base,
                       int* status, bool query)
{
  ShardCov cv;
  for (int k = 0; k < 9; k++) cv.rot[k] = rot_var[k], cv.tsl[k] = tsl_var[k];
  const int nblk = count <= 0 ? 1 : (count + SH_THREADS - 1) / SH_THREADS;
  k_shard_count<<<nblk, SH_THREADS, 0, st>>>(scan, first, count, x, voxel_size, world, owner, hist, status);
  k_shard_offsets<<<1, 32 * VN_MAX_WORLD, 0, st>>>(hist, nblk, world, counts, starts);
  if (query)
    k_shard_scatter<true><<<nblk, SH_THREADS, 0, st>>>(scan, first, count, x, cv, world, owner, hist, starts, out, gidx_base);
  else
    k_shard_scatter<false><<<nblk, SH_THREADS, 0, st>>>(scan, first, count, x, cv, world, owner, hist, starts, out, gidx_base);
  return 3;
}

int launch_shard_unpack(cudaStream_t st, const double* rec, int n, const ScanView& scan, const InsertScratch& sc)
{
  if (n <= 0) return 0;
  k_shard_unpack<<<(n + SH_THREADS - 1) / SH_THREADS, SH_THREADS, 0, st>>>(rec, n, scan, sc);
  return 1;
}

int launch_shard_unpack_query(cudaStream_t st, const double* rec, int n, const ScanView& scan)
{
  if (n <= 0) return 0;
  k_shard_unpack_query<<<(n + SH_THREADS - 1) / SH_THREADS, SH_THREADS, 0, st>>>(rec, n, scan);
  return 1;
}

// Voxel-map kernels: pvec_update + cut_voxel_multi (a8-a10), multi_recut (a11),
// multi_margi (a12), export. Compiled with -fmad=false (see scan_kernels.cu).
//
// Parallel decomposition. The reference forks 5 host threads over ROOT voxels
// (voxel_map.cpp:94-134, local_mapping.cpp:17-84, 144-201). Here every octree
// NODE is one unit of work: recut / margi walk the trees of surf_map_slide
// layer by layer (one launch per layer, one thread per node, the next layer's
// node list is built with an atomic append), and the per-leaf accumulation of
// an insert is done by one warp per touched leaf. Inside a leaf the reference's
// sequential order is kept for the cluster sums (P, v, N: one lane per scalar,
// points in ascending index order), so those sums - and through them every
// eigen-decomposition and plane decision - are reproduced bit for bit; only
// the order in which independent nodes are visited differs.
#include <cstdio>
#include "vn_kernels.cuh"

#define SPIN_LIMIT 4000000

struct PoseBuf
{
  PoseD x[VINA_MAX_WIN];
};
struct Cov2
{
  double rot[9], tsl[9];
};
// The newest frame's pose may still be on its way to the host when the map update is enqueued (the IEKF loop
// runs on the device): kernels then read frame `idx` from the device iterate instead of the host's copy - the
// same doubles the host receives a moment later (k_publish_iterate), so nothing changes numerically.
struct LivePose
{
  const IekfDev* dev;  // nullptr: every pose comes from the host
  int idx;
};
__device__ __forceinline__ void pose_sel(const PoseBuf& xb, const LivePose& lv, int i, double (&R)[9], double (&p)[3])
{
  if (lv.dev && i == lv.idx)
  {
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = lv.dev->R[k];
#pragma unroll
    for (int k = 0; k < 3; k++) p[k] = lv.dev->p[k];
  }
  else
  {
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = xb.x[i].R[k];
#pragma unroll
    for (int k = 0; k < 3; k++) p[k] = xb.x[i].p[k];
  }
}

__device__ __forceinline__ int ld_volatile(const int* p) { return *((const volatile int*)p); }

// The map kernels are latency chains over a nearly empty machine: a node's NodeCold record (1.6 KB, 13 lines) is read
// field by field along a dependent computation, every first touch of a line a DRAM round trip. Pulling the lines
// towards L1 as soon as the node id is known turns all but the first into cache hits.
__device__ __forceinline__ void prefetch_line(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#define VN_COLD_LINES ((int)((sizeof(NodeCold) + 127) / 128) + 1)  // (+1: records are not line-aligned)
__device__ __forceinline__ void prefetch_cold(const NodeCold* c, int part, int parts)
{
  const char* b = reinterpret_cast<const char*>(c);
  for (int l = part; l < VN_COLD_LINES; l += parts) prefetch_line(b + 128 * l);
}

// VINA_SPLIT_TRACE build (debugging): cycle stamps along the work of one unit (a leaf's group / block), collected
// per kernel and summarised by vn_split_trace_dump (mean and maximum of every phase, the slowest unit)
#ifdef VINA_SPLIT_TRACE
#define KT_SLOTS 8192
#define KT_PH 24
__device__ long long g_kt[3][KT_SLOTS][KT_PH];
__device__ int g_kt_n[3];
#define KT_DECL      \
  long long kt_[KT_PH]; \
  int kt_k = 0
#define KT() \
  do \
  { \
    if (kt_k < KT_PH) kt_[kt_k++] = clock64(); \
  } while (0)
#define KT_COMMIT(which) \
  do \
  { \
    const int s_ = atomicAdd(&g_kt_n[which], 1); \
    if (s_ < KT_SLOTS) \
      for (int i_ = 0; i_ < KT_PH; i_++) g_kt[which][s_][i_] = i_ < kt_k ? kt_[i_] : 0; \
  } while (0)
#else
#define KT_DECL \
  do \
  { \
  } while (0)
#define KT() \
  do \
  { \
  } while (0)
#define KT_COMMIT(which) \
  do \
  { \
  } while (0)
#endif

__device__ int alloc_node(const MapView& M)
{
  // ids released by the pruning first (records are zero, like never-used pool memory)
  if (ld_volatile(M.free_count) > 0)
  {
    const int k = atomicSub(M.free_count, 1);
    if (k > 0) return M.free_nodes[k - 1];
    atomicAdd(M.free_count, 1);
  }
  int id = atomicAdd(M.node_count, 1);
  if (id >= M.max_nodes)
  {
    atomicOr(M.status, VN_ST_NODES_FULL);
    return -1;
  }
  return id;
}

// leaves[leafnum] = new OctoTree(layer+1) (octree.cpp:217-224); caller owns the parent
__device__ int make_child(const MapView& M, int parent, int ci)
{
  int id = alloc_node(M);
  if (id < 0) return -1;
  const NodeHot& ph = M.hot[parent];
  NodeHot& h = M.hot[id];
  const int xyz[3] = { (ci >> 2) & 1, (ci >> 1) & 1, ci & 1 };
  for (int k = 0; k < 3; k++) h.vcenter[k] = ph.vcenter[k] + (double)((float)(2 * xyz[k] - 1) * ph.ql);
  h.ql = ph.ql / 2;
  h.layer = ph.layer + 1;
  h.flags = 0;
  for (int k = 0; k < 8; k++) h.children[k] = -1;
  NodeCold& c = M.cold[id];
  c.rootkey = M.cold[parent].rootkey;
  c.root = M.cold[parent].root;
  c.path = M.cold[parent].path | (ci << (3 * ph.layer));
  c.fix_head = c.fix_tail = -1;
  for (int k = 0; k < 8; k++) c.children[k] = -1;
  return id;
}

// ---------------------------------------------------------------------------
// insert, phase 1: pvec_update (point_utils.cpp:54-65) + voxel key + root find/create
// (voxel_map.cpp:53-87).
__global__ void __launch_bounds__(256)
    k_insert_root(MapView M, ScanView scan, const int* __restrict__ n_ptr, int n_host, InsertScratch sc, PoseD x_host,
                  Cov2 cv_host, int pre, const IekfDev* __restrict__ live)
{
  vn_pdl_sync();
  // pose and posterior covariance blocks of pvec_update: from the host, or straight from the device iterate
  __shared__ PoseD x;
  __shared__ Cov2 cv;
  if (threadIdx.x < 9)
  {
    const int t = threadIdx.x, a = t % 3, b = t / 3;
    x.R[t] = live ? live->R[t] : x_host.R[t];
    cv.rot[t] = live ? live->cov[a + 15 * b] : cv_host.rot[t];
    cv.tsl[t] = live ? live->cov[(3 + a) + 15 * (3 + b)] : cv_host.tsl[t];
    if (t < 3) x.p[t] = live ? live->p[t] : x_host.p[t];
  }
  // (the counters of the NEXT insert: ping-pong, no launch spent on zeroing)
  if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 36) sc.counters_alt[threadIdx.x - 32] = 0;
  __syncthreads();
  int n = n_ptr ? *n_ptr : n_host;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double pw[3];
  if (pre)
  {
    // sharded map: the sender already ran pvec_update (shard_kernels.cu), sc.pw / sc.vw hold the result
    for (int k = 0; k < 3; k++) pw[k] = sc.pw[k][i];
  }
  else
  {
    double pnt[3] = { scan.p[0][i], scan.p[1][i], scan.p[2][i] };
    double var6[6], vw[6];
    for (int k = 0; k < 6; k++) var6[k] = scan.v[k][i];
    rot_trans(x.R, x.p, pnt, pw);
    world_var(x.R, pnt, var6, cv.rot, cv.tsl, vw);
    for (int k = 0; k < 3; k++) sc.pw[k][i] = pw[k];
    for (int k = 0; k < 6; k++) sc.vw[k][i] = vw[k];
  }

  long long kc[3];
  for (int k = 0; k < 3; k++) kc[k] = voxel_coord(pw[k], M.voxel_size);
  unsigned long long key;
  sc.root_of[i] = -1;
  if (!pack_key(kc[0], kc[1], kc[2], &key))
  {
    atomicOr(M.status, VN_ST_KEY_RANGE);
    return;
  }
  unsigned int h = hash_key(key) & M.hmask;
  int root = -1;
  bool created = false;
  for (unsigned int probe = 0;; probe++)
  {
    if (probe > M.hmask)
    {
      atomicOr(M.status, VN_ST_HASH_FULL);
      return;
    }
    unsigned long long old = *((volatile unsigned long long*)&M.slots[h].key);
    if (old == VN_EMPTY_KEY) old = atomicCAS(&M.slots[h].key, VN_EMPTY_KEY, key);
    if (old == VN_EMPTY_KEY)
    {
      // ot = new OctoTree(0, wdsize) (voxel_map.cpp:77-83)
      int id = alloc_node(M);
      if (id >= 0)
      {
        NodeHot& nh = M.hot[id];
        for (int k = 0; k < 3; k++) nh.vcenter[k] = (0.5 + (double)kc[k]) * M.voxel_size;
        nh.ql = (float)(M.voxel_size / 4.0);
        nh.layer = 0;
        nh.flags = 0;
        for (int k = 0; k < 8; k++) nh.children[k] = -1;
        NodeCold& nc = M.cold[id];
        nc.rootkey = key;
        nc.root = id;
        nc.path = 0;
        nc.fix_head = nc.fix_tail = -1;
        for (int k = 0; k < 8; k++) nc.children[k] = -1;
        atomicAdd(M.root_count, 1);
        __threadfence();
      }
      atomicExch(&M.slots[h].root, id);
      root = id;
      created = true;
      break;
    }
    if (old == key)
    {
      int r = ld_volatile(&M.slots[h].root);
      int spins = 0;
      while (r == -2)
      {
        if (++spins > SPIN_LIMIT)
        {
          atomicOr(M.status, VN_ST_SPIN);
          return;
        }
        r = ld_volatile(&M.slots[h].root);
      }
      __threadfence();
      root = r;
      break;
    }
    h = (h + 1) & M.hmask;
  }
  if (root < 0) return;
  NodeCold& rc = M.cold[root];
  if (!created) rc.isexist = 1;  // voxel_map.cpp:70
  if (atomicExch(&rc.in_slide, 1) == 0)  // feat_tem_map[position] = ot (voxel_map.cpp:71-72, 83)
  {
    int pos = atomicAdd(&M.slide_count[M.slide_cur], 1);
    M.slide_list[M.slide_cur][pos] = root;
  }
  if (atomicExch(&rc.touch_stamp, sc.stamp) != sc.stamp) atomicAdd(&sc.counters[0], 1);
  sc.root_of[i] = root;
}

// insert, phase 2: OctoTree::allocate descent (octree.cpp:203-228), creating children lazily
__global__ void __launch_bounds__(256)
    k_insert_leaf(MapView M, const int* __restrict__ n_ptr, int n_host, InsertScratch sc)
{
  vn_pdl_sync();
  int n = n_ptr ? *n_ptr : n_host;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sc.leaf_of[i] = -1;
  if (sc.counters[0] < M.thread_num) return;  // voxel_map.cpp:96-97
  int node = sc.root_of[i];
  if (node < 0) return;
  double pw[3] = { sc.pw[0][i], sc.pw[1][i], sc.pw[2][i] };
  while (M.hot[node].flags & VN_FLAG_INTERIOR)
  {
    int ci = child_index(pw, M.hot[node].vcenter);
    int* slot = &M.cold[node].children[ci];
    int ch = ld_volatile(slot);
    if (ch == -1)
    {
      int old = atomicCAS(slot, -1, -2);
      if (old == -1)
      {
        int id = make_child(M, node, ci);
        if (id >= 0) M.hot[node].children[ci] = id;  // mirror read by k_iekf
        __threadfence();
        atomicExch(slot, id < 0 ? -3 : id);
        ch = id < 0 ? -3 : id;
      }
      else
        ch = old;
    }
    int spins = 0;
    while (ch == -2)
    {
      if (++spins > SPIN_LIMIT)
      {
        atomicOr(M.status, VN_ST_SPIN);
        return;
      }
      ch = ld_volatile(slot);
    }
    if (ch < 0) return;
    __threadfence();
    node = ch;
  }
  sc.leaf_of[i] = node;
  int r = atomicAdd(&M.cold[node].pend_cnt, 1);
  sc.rank_of[i] = r;
  if (r == 0)
  {
    int t = atomicAdd(&sc.counters[1], 1);
    sc.touched[t] = node;
  }
}

__global__ void __launch_bounds__(128) k_insert_alloc(MapView M, InsertScratch sc)
{
  vn_pdl_sync();
  int nt = sc.counters[1];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nt; j += gridDim.x * blockDim.x)
  {
    NodeCold& c = M.cold[sc.touched[j]];
    c.pend_off = atomicAdd(&sc.counters[2], c.pend_cnt);
  }
}

__global__ void __launch_bounds__(256)
    k_insert_scatter(MapView M, const int* __restrict__ n_ptr, int n_host, InsertScratch sc)
{
  vn_pdl_sync();
  int n = n_ptr ? *n_ptr : n_host;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int leaf = sc.leaf_of[i];
  if (leaf < 0) return;
  sc.idx[M.cold[leaf].pend_off + sc.rank_of[i]] = i;
}

// ---------------------------------------------------------------------------
// Warp-cooperative leaf accumulation (used by the insert and by the subdivision).
// One warp owns one leaf; points are consumed strictly in the reference's order.
//   lanes  0.. 8 own cluster A (P[6] lower triangle, v[3])    exact single-rounding ops
//   lanes  9..17 own cluster B, lanes 18..26 own cluster C
//   every lane   owns cov_add entries `lane` and `lane + 32` of the 45 packed ones:
//                entry(r,c) += W_r V W_c^T with W = [Bi; I3] (Bf_var, octree.cpp:83-92)
// Points are staged 32 at a time in shared memory (one parallel gather), then applied one by one.
struct LaneRole
{
  int er[2], ec[2];       // cov entries (row, col), -1 = none
  float coefR[2][3], coefC[2][3];
  int compR[2][3], compC[2][3];
  int ck, ci, cj;         // cluster scalar: P[ck] = sum s[ci]*s[cj] (ck<6) or v[ck-6] = sum s[ci]
};

__device__ __forceinline__ double sel4(const double* v, int comp)
{
  return comp == 0 ? v[0] : (comp == 1 ? v[1] : (comp == 2 ? v[2] : 1.0));
}

__device__ void role_init(int lane, LaneRole& L)
{
  // W[r][j] = coef * vec[comp] (comp 3 = the constant 1): Bi rows [2x 0 0; y x 0; z 0 x; 0 2y 0; 0 z y; 0 0 2z], then I3
  const float A[9][3] = { { 2, 0, 0 }, { 1, 1, 0 }, { 1, 0, 1 }, { 0, 2, 0 }, { 0, 1, 1 }, { 0, 0, 2 },
                          { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
  const int Cc[9][3] = { { 0, 0, 0 }, { 1, 0, 0 }, { 2, 0, 0 }, { 0, 1, 0 }, { 0, 2, 1 }, { 0, 0, 2 },
                         { 3, 3, 3 }, { 3, 3, 3 }, { 3, 3, 3 } };
  for (int q = 0; q < 2; q++)
  {
    int e = lane + 32 * q, r = 0;
    L.er[q] = L.ec[q] = -1;
    if (e < 45)
    {
      while (e >= 9 - r)
      {
        e -= 9 - r;
        r++;
      }
      L.er[q] = r;
      L.ec[q] = r + e;
    }
    const int rr = L.er[q] < 0 ? 0 : L.er[q], cc = L.ec[q] < 0 ? 0 : L.ec[q];
    for (int j = 0; j < 3; j++)
    {
      L.coefR[q][j] = A[rr][j];
      L.compR[q][j] = Cc[rr][j];
      L.coefC[q][j] = A[cc][j];
      L.compC[q][j] = Cc[cc][j];
    }
  }
  const int li[6] = { 0, 1, 2, 1, 2, 2 }, lj[6] = { 0, 0, 0, 1, 1, 2 };
  L.ck = lane % 9;
  L.ci = L.ck < 6 ? li[L.ck] : L.ck - 6;
  L.cj = L.ck < 6 ? lj[L.ck] : 0;
}

__device__ __forceinline__ double cov_term(const LaneRole& L, int q, const double* V, const double* vec)
{
  double wr[3], wc[3];
#pragma unroll
  for (int j = 0; j < 3; j++)
  {
    wr[j] = (double)L.coefR[q][j] * sel4(vec, L.compR[q][j]);
    wc[j] = (double)L.coefC[q][j] * sel4(vec, L.compC[q][j]);
  }
  const double t0 = V[0] * wc[0] + V[3] * wc[1] + V[6] * wc[2];
  const double t1 = V[1] * wc[0] + V[4] * wc[1] + V[7] * wc[2];
  const double t2 = V[2] * wc[0] + V[5] * wc[1] + V[8] * wc[2];
  return wr[0] * t0 + wr[1] * t1 + wr[2] * t2;
}

// PointCluster::push for the scalar this lane owns (types.hpp:137-142)
__device__ __forceinline__ double cluster_term(const LaneRole& L, double cl, const double* s)
{
  const double a_i = L.ci == 0 ? s[0] : (L.ci == 1 ? s[1] : s[2]);
  const double a_j = L.cj == 0 ? s[0] : (L.cj == 1 ? s[1] : s[2]);
  return L.ck < 6 ? da(cl, dm(a_i, a_j)) : da(cl, a_i);
}
__device__ __forceinline__ double cluster_get(const Cluster& c, int ck) { return ck < 6 ? c.P[ck] : c.v[ck - 6]; }
__device__ __forceinline__ void cluster_set(Cluster& c, int ck, double v)
{
  if (ck < 6)
    c.P[ck] = v;
  else
    c.v[ck - 6] = v;
}

#define PT_STRIDE 13   // staged point: p[3] (cluster B/C input), v[6], pw[3] (cluster A input and Bf_var vector)
#define RED_STRIDE 47  // staged Bf_var contribution of one point: 45 packed entries

// The 45 packed upper-triangle entries of Bf_var(pv, vec) (octree.cpp:83-92) for one point:
// [[Bi V Bi^T, Bi V], [., V]], Bi = d(xx,xy,xz,yy,yz,zz)/dp. Fully unrolled, Bi's sparsity used.
__device__ __forceinline__ void bf_var_terms(const double* v, const double* p, double* o)
{
  const double x = p[0], y = p[1], z = p[2];
  const double r0[3] = { v[0], v[1], v[2] }, r1[3] = { v[1], v[3], v[4] }, r2[3] = { v[2], v[4], v[5] };
  double U[6][3];
#pragma unroll
  for (int j = 0; j < 3; j++)
  {
    U[0][j] = 2.0 * x * r0[j];
    U[1][j] = y * r0[j] + x * r1[j];
    U[2][j] = z * r0[j] + x * r2[j];
    U[3][j] = 2.0 * y * r1[j];
    U[4][j] = z * r1[j] + y * r2[j];
    U[5][j] = 2.0 * z * r2[j];
  }
  int t = 0;
#pragma unroll
  for (int r = 0; r < 6; r++)
  {
    const double c6[6] = { 2.0 * x * U[r][0],           y * U[r][0] + x * U[r][1], z * U[r][0] + x * U[r][2],
                           2.0 * y * U[r][1],           z * U[r][1] + y * U[r][2], 2.0 * z * U[r][2] };
#pragma unroll
    for (int c = r; c < 6; c++) o[t++] = c6[c];
#pragma unroll
    for (int c = 0; c < 3; c++) o[t++] = U[r][c];
  }
  o[t++] = v[0];
  o[t++] = v[1];
  o[t++] = v[2];
  o[t++] = v[3];
  o[t++] = v[4];
  o[t++] = v[5];
}

// insert, phase 3: OctoTree::push for every point of a leaf (octree.cpp:151-177), one warp per touched
// leaf, points in ascending index order (= voxel_map.cpp:86 push_back(i)).
// cluster A = pcr_add (world point, lanes 0..8), cluster B = pcrs_local[slot] (body point, lanes 9..17):
// sequential exact sums. cov_add: every lane computes the 45 Bf_var terms of one staged point, the warp
// then sums the columns (lane owns packed entries `lane` and `lane + 32`).
#define ACC_WARPS 3
#define ACC_PT_STRIDE 19
#define ACC_RED_ROWS 8
__global__ void __launch_bounds__(32 * ACC_WARPS) k_insert_accum(MapView M, ScanView scan, InsertScratch sc, int win_ord)
{
  vn_pdl_sync();
  // shared memory per warp bounds the number of resident warps of this kernel: the Bf_var terms are staged
  // 8 rows at a time (3 KB); per row the 9 + 9 products of PointCluster::push (types.hpp:137-142) for the world
  // and the body point are formed by the row's lane, so that the sequential chains only add (4.8 KB)
  __shared__ double pt_s[ACC_WARPS][32][ACC_PT_STRIDE];
  __shared__ double red_s[ACC_WARPS][ACC_RED_ROWS][RED_STRIDE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nt = sc.counters[1];
  const int mord = M.mp[win_ord];
  LaneRole L;
  role_init(lane, L);
  double(*pt)[ACC_PT_STRIDE] = pt_s[warp];
  double(*red)[RED_STRIDE] = red_s[warp];
  const bool has2 = lane + 32 < 45;

  for (int j = blockIdx.x * ACC_WARPS + warp; j < nt; j += gridDim.x * ACC_WARPS)
  {
    KT_DECL;
    KT();
    const int leaf = sc.touched[j];
    NodeCold& c = M.cold[leaf];
    // everything this leaf's pass needs from its records goes out NOW, together: these loads are independent once the
    // leaf is known, and one L2 / DRAM round trip costs about as much as the whole accumulation of a small leaf
    const int cnt = c.pend_cnt;
    const int poff = c.pend_off;
    const int old_cnt = c.win_cnt[mord];
    const int old_off = c.win_off[mord];
    const int leaf_layer = M.hot[leaf].layer;
    double cl = 0.0;
    if (lane < 9)
      cl = cluster_get(c.pcr_add, L.ck);
    else if (lane < 18)
      cl = cluster_get(c.pcrs_local[mord], L.ck);
    double cv0 = c.cov_add[lane];
    double cv1 = has2 ? c.cov_add[lane + 32] : 0.0;
    const int n_add0 = c.pcr_add.N, n_loc0 = c.pcrs_local[mord].N;
    KT();
    int* idx = sc.idx + poff;
    // ascending point order
    if (cnt <= 32)
    {
      const int v = lane < cnt ? idx[lane] : 0x7fffffff;
      int rank = 0;
      for (int o = 0; o < 32; o++)
      {
        int u = __shfl_sync(0xffffffffu, v, o);
        rank += (u < v) ? 1 : 0;
      }
      __syncwarp();
      if (lane < cnt) idx[rank] = v;
    }
    else if (cnt <= ACC_RED_ROWS * RED_STRIDE * 2)
    {
      // rank sort through shared memory (indices are distinct): rank = number of smaller ones
      int* sidx = reinterpret_cast<int*>(&red[0][0]);
      for (int a = lane; a < cnt; a += 32) sidx[a] = idx[a];
      __syncwarp();
      for (int a = lane; a < cnt; a += 32)
      {
        const int v = sidx[a];
        int rank = 0;
        for (int b = 0; b < cnt; b++) rank += (sidx[b] < v) ? 1 : 0;
        idx[rank] = v;
      }
    }
    else if (lane == 0)
    {
      for (int a = 1; a < cnt; a++)
      {
        int v = idx[a], b = a - 1;
        while (b >= 0 && idx[b] > v)
        {
          idx[b + 1] = idx[b];
          b--;
        }
        idx[b + 1] = v;
      }
    }
    __syncwarp();
    KT();
    const bool store = leaf_layer < M.max_layer;
    int woff = 0;
    if (store)
    {
      if (lane == 0) woff = atomicAdd(&M.win_cursor[mord], cnt + old_cnt);
      woff = __shfl_sync(0xffffffffu, woff, 0);
      if ((long long)woff + cnt + old_cnt > M.win_cap)
      {
        if (lane == 0)
        {
          atomicOr(M.status, VN_ST_WIN_FULL);
          c.pend_cnt = 0;
        }
        continue;
      }
      PointRec* pool = M.win_pool[mord];
      for (int a = lane; a < old_cnt; a += 32) pool[woff + a] = pool[old_off + a];
    }
    KT();
    for (int base = 0; base < cnt; base += 32)
    {
      const int a = base + lane;
      PointRec pr;
      double pw[3];
      if (a < cnt)
      {
        const int i = idx[a];
        for (int k = 0; k < 3; k++) pr.p[k] = scan.p[k][i];
        for (int k = 0; k < 6; k++) pr.v[k] = sc.vw[k][i];
        for (int k = 0; k < 3; k++) pw[k] = sc.pw[k][i];
        const double* s3[2] = { pw, pr.p };
#pragma unroll
        for (int w = 0; w < 2; w++)
        {
          const double* q = s3[w];
          double* v = pt[lane] + 9 * w;
          v[0] = dm(q[0], q[0]);
          v[1] = dm(q[1], q[0]);
          v[2] = dm(q[2], q[0]);
          v[3] = dm(q[1], q[1]);
          v[4] = dm(q[2], q[1]);
          v[5] = dm(q[2], q[2]);
          v[6] = q[0];
          v[7] = q[1];
          v[8] = q[2];
        }
        if (store) M.win_pool[mord][woff + old_cnt + a] = pr;
      }
      const int m = min(32, cnt - base);
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int part = 0; part < 32 / ACC_RED_ROWS; part++)
      {
        if (part * ACC_RED_ROWS >= m) break;  // (uniform)
        if ((lane / ACC_RED_ROWS) == part && a < cnt)
        {
          double o[45];
          bf_var_terms(pr.v, pw, o);
#pragma unroll
          for (int e = 0; e < 45; e++) red[lane % ACC_RED_ROWS][e] = o[e];
        }
        __syncwarp();
        const int mh = min(ACC_RED_ROWS, m - ACC_RED_ROWS * part);
        if (mh == ACC_RED_ROWS)
        {
          // (straight-line: a counted loop costs ~20 cycles of branch per row on top of the 10-cycle add)
#pragma unroll
          for (int r = 0; r < ACC_RED_ROWS; r++)
          {
            s0 += red[r][lane];
            if (has2) s1 += red[r][lane + 32];
          }
        }
        else
          for (int r = 0; r < mh; r++)
          {
            s0 += red[r][lane];
            if (has2) s1 += red[r][lane + 32];
          }
        __syncwarp();
      }
      // lanes 0..8: pcr_add (world point), lanes 9..17: pcrs_local[slot] (body point); L.ck = lane % 9
      if (lane < 18)
      {
        const int col = lane;  // 0..8 world terms, 9..17 body terms
        if (m == 32)
        {
          // a full batch, straight-line: the 32 loads go out ahead of the chain of dependent adds
          double v[32];
#pragma unroll
          for (int r = 0; r < 32; r++) v[r] = pt[r][col];
#pragma unroll
          for (int r = 0; r < 32; r++) cl = da(cl, v[r]);
        }
        else
        {
          int r = 0;
          for (; r + 8 <= m; r += 8)
          {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = pt[r + u][col];
#pragma unroll
            for (int u = 0; u < 8; u++) cl = da(cl, v[u]);
          }
          for (; r < m; r++) cl = da(cl, pt[r][col]);
        }
      }
      cv0 += s0;
      cv1 += s1;
      __syncwarp();
    }
    KT();
    if (lane < 9)
      cluster_set(c.pcr_add, L.ck, cl);
    else if (lane < 18)
      cluster_set(c.pcrs_local[mord], L.ck, cl);
    c.cov_add[lane] = cv0;
    if (has2) c.cov_add[lane + 32] = cv1;
    if (lane == 0)
    {
      c.pcr_add.N = n_add0 + cnt;
      c.pcrs_local[mord].N = n_loc0 + cnt;
      c.has_sw = 1;   // sw acquired (octree.cpp:154-164); recycled windows are empty
      c.isexist = 1;  // octree.cpp:165-166
      if (store)
      {
        c.win_off[mord] = woff;
        c.win_cnt[mord] = cnt + old_cnt;
      }
      c.pend_cnt = 0;
      KT();
      KT_COMMIT(2);
    }
    __syncwarp();
  }
}


// push_fix into a child (octree.cpp:179-188)
__device__ __forceinline__ void child_push_fix(NodeCold& k, const PointRec& pr)
{
  cluster_push(k.pcr_fix, pr.p);
  cluster_push(k.pcr_add, pr.p);
  bf_var_add(k.cov_add, pr.v, pr.p);
}

// append one segment of `cnt` fixed points to a node's point_fix chain; returns pool offset or -1. Called by
// the one thread that owns the node.
__device__ int fix_append(const MapView& M, NodeCold& c, int cnt)
{
  int off = atomicAdd(M.fix_cursor, cnt);
  if ((long long)off + cnt > M.fix_cap)
  {
    atomicOr(M.status, VN_ST_FIX_FULL);
    return -1;
  }
  if (c.fix_tail >= 0 && M.fix_segs[c.fix_tail].n < VN_FIXSEG_PER_BLOCK)
  {
    FixSeg& b = M.fix_segs[c.fix_tail];
    const int e = b.n;
    b.off[e] = off;
    b.cnt[e] = cnt;
    b.n = e + 1;
  }
  else
  {
    int sid = -1;
    if (ld_volatile(M.free_seg_count) > 0)
    {
      const int k = atomicSub(M.free_seg_count, 1);
      if (k > 0)
        sid = M.free_segs[k - 1];
      else
        atomicAdd(M.free_seg_count, 1);
    }
    if (sid < 0) sid = atomicAdd(M.fixseg_cursor, 1);
    if (sid >= M.fixseg_cap)
    {
      atomicOr(M.status, VN_ST_FIX_FULL);
      return -1;
    }
    FixSeg& b = M.fix_segs[sid];
    b.off[0] = off;
    b.cnt[0] = cnt;
    b.n = 1;
    b.next = -1;
    if (c.fix_tail >= 0)
      M.fix_segs[c.fix_tail].next = sid;
    else
      c.fix_head = sid;
    c.fix_tail = sid;
  }
  c.fix_count += cnt;
  return off;
}

__device__ __forceinline__ const int* layer_nodes(const MapView& M, const LayerLists& LL, int layer, int* count)
{
  if (layer == 0)
  {
    *count = M.slide_count[M.slide_cur];
    return M.slide_list[M.slide_cur];
  }
  *count = LL.count[layer];
  return LL.list[layer];
}

// The leaf branch of OctoTree::recut (octree.cpp:335-393) for node n, then tras_opt's BA-factor marking of the same
// node (octree.cpp:498-521, local_mapping.cpp:196-200). Returns true when the leaf has to be subdivided.
__device__ bool recut_leaf(const MapView& M, NodeHot& h, NodeCold& c)
{
  c.opt_state = -1;
  if ((double)c.pcr_add.N <= M.min_point[h.layer])
  {
    h.flags &= ~VN_FLAG_PLANE;
    return false;
  }
  if (!c.isexist || !c.has_sw) return false;
  double L[6], ev[3], Q[9];
  cluster_cov(c.pcr_add, L);
  eig3_sym(L, ev, Q);
  for (int k = 0; k < 3; k++) c.eig_value[k] = ev[k];
  for (int k = 0; k < 9; k++) c.eig_vector[k] = Q[k];
  const bool is_plane = (ev[0] < M.min_eigen_value) && ((ev[0] / ev[2]) < M.thre[h.layer]);  // octree.cpp:198-201
  if (is_plane)
  {
    h.flags |= VN_FLAG_PLANE;
    if (!(ev[0] / ev[1] > 0.12)) c.opt_state = 1;  // tras_opt: this leaf is a BA factor
    return false;
  }
  h.flags &= ~VN_FLAG_PLANE;
  return h.layer < M.max_layer;
}

// recut_leaf for a child k_split has just created, from what the block still holds: the child's pcr_add (shared
// memory), its layer, "has points in the window" - no global read on the way to the eigen-solver, and the flags are a
// plain store (a new node's flags are 0). Same arithmetic and decisions as recut_leaf.
__device__ bool recut_child(const MapView& M, NodeHot& h, NodeCold& c, const Cluster& add, bool in_window, int layer)
{
  c.opt_state = -1;
  if ((double)add.N <= M.min_point[layer]) return false;  // (flags stay 0: not a plane)
  if (!in_window) return false;
  double L[6], ev[3], Q[9];
  cluster_cov(add, L);
  eig3_sym(L, ev, Q);
  for (int k = 0; k < 3; k++) c.eig_value[k] = ev[k];
  for (int k = 0; k < 9; k++) c.eig_vector[k] = Q[k];
  const bool is_plane = (ev[0] < M.min_eigen_value) && ((ev[0] / ev[2]) < M.thre[layer]);  // octree.cpp:198-201
  if (is_plane)
  {
    h.flags = VN_FLAG_PLANE;
    if (!(ev[0] / ev[1] > 0.12)) c.opt_state = 1;  // tras_opt: this leaf is a BA factor
    return false;
  }
  return layer < M.max_layer;
}

// multi_recut, step 1: the nodes below the roots of surf_map_slide, listed per layer (layer 0 is the slide list
// itself). EIGHT threads per root, one per first-level child; a thread walks that child's subtree through the
// children mirrors of NodeHot (one 128-byte line per node, the loads of a level issued together). The trees are at
// most max_layer <= 3 deep. One atomic per warp and layer reserves the warp's list positions. Thread 0 also clears
// the counters the NEXT multi_recut will use (ping-pong, so that no launch is spent on zeroing).
__device__ __forceinline__ int warp_reserve(int* counter, int mine, int lane)
{
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  if (lane == 31 && total > 0) base = atomicAdd(counter, total);
  base = __shfl_sync(0xffffffffu, base, 31);
  return base + incl - mine;
}

__global__ void __launch_bounds__(128) k_recut_collect(MapView M, LayerLists LL)
{
  vn_pdl_sync();
  if (blockIdx.x == 0 && threadIdx.x < 8) LL.count_alt[threadIdx.x] = 0;
  const int nroots = M.slide_count[M.slide_cur];
  if (nroots + M.slide_others < M.thread_num) return;  // local_mapping.cpp:150-154
  const int* roots = M.slide_list[M.slide_cur];
  const int lane = threadIdx.x & 31;
  const int total = 8 * nroots;
  const int stride = gridDim.x * blockDim.x;
  for (int t0 = blockIdx.x * blockDim.x + threadIdx.x - lane; t0 < total; t0 += stride)  // warp-uniform trip count
  {
    const int t = t0 + lane;
    int n1 = -1;
    if (t < total)
    {
      const NodeHot& h0 = M.hot[roots[t >> 3]];
      if (h0.flags & VN_FLAG_INTERIOR) n1 = h0.children[t & 7];
    }
    // level 2: the children of n1, level 3: theirs
    int c2[8], m2 = 0, m3 = 0;
    unsigned int int2 = 0;  // which level-2 children are interior
#pragma unroll
    for (int b = 0; b < 8; b++) c2[b] = -1;
    if (n1 >= 0)
    {
      const NodeHot& h1 = M.hot[n1];
      if (h1.flags & VN_FLAG_INTERIOR)
      {
#pragma unroll
        for (int b = 0; b < 8; b++) c2[b] = h1.children[b];
        int f2[8];
#pragma unroll
        for (int b = 0; b < 8; b++) f2[b] = c2[b] >= 0 ? M.hot[c2[b]].flags : 0;
#pragma unroll
        for (int b = 0; b < 8; b++)
        {
          if (c2[b] >= 0) m2++;
          if (f2[b] & VN_FLAG_INTERIOR) int2 |= 1u << b;
        }
      }
    }
    for (int b = 0; b < 8; b++)
      if (int2 & (1u << b))
        for (int d = 0; d < 8; d++) m3 += M.hot[c2[b]].children[d] >= 0 ? 1 : 0;
    int p1 = warp_reserve(&LL.count[1], n1 >= 0 ? 1 : 0, lane);
    int p2 = warp_reserve(&LL.count[2], m2, lane);
    int p3 = __any_sync(0xffffffffu, m3 > 0) ? warp_reserve(&LL.count[3], m3, lane) : 0;
    if (n1 >= 0) LL.list[1][p1] = n1;
#pragma unroll
    for (int b = 0; b < 8; b++)
      if (c2[b] >= 0) LL.list[2][p2++] = c2[b];
    for (int b = 0; b < 8; b++)
      if (int2 & (1u << b))
        for (int d = 0; d < 8; d++)
        {
          const int n3 = M.hot[c2[b]].children[d];
          if (n3 >= 0) LL.list[3][p3++] = n3;
        }
  }
}

// multi_recut, step 2: OctoTree::recut for every leaf that exists at this point, all layers in ONE launch
// (blockIdx.y = layer): the subtrees below different nodes are independent, only the children a subdivision creates
// have to wait for it - those are handled by k_split itself. Leaves that must be subdivided go to the split list.
__global__ void __launch_bounds__(128) k_recut_all(MapView M, LayerLists LL)
{
  vn_pdl_sync();
  // the lists as they are before any subdivision of this multi_recut (k_split appends the children it creates)
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 4) LL.snap[threadIdx.x] = LL.count[threadIdx.x];
  if (M.slide_count[M.slide_cur] + M.slide_others < M.thread_num) return;  // local_mapping.cpp:150-154
  int nn;
  const int* nodes = layer_nodes(M, LL, blockIdx.y, &nn);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nn; j += gridDim.x * blockDim.x)
  {
    const int n = nodes[j];
    NodeHot& h = M.hot[n];
    {
      // what recut_leaf touches: pcr_add (front of the record), the eigen-decomposition and the flags (back)
      const char* b = reinterpret_cast<const char*>(&M.cold[n]);
      prefetch_line(b);
      prefetch_line(b + offsetof(NodeCold, eig_value));
      prefetch_line(b + offsetof(NodeCold, last_num));
    }
    if (h.flags & VN_FLAG_INTERIOR) continue;
    if (recut_leaf(M, h, M.cold[n]))
    {
      h.flags |= VN_FLAG_SPLIT_PENDING;
      LL.split[atomicAdd(&LL.count[4], 1)] = n;
    }
  }
}

// tras_opt with the container (octree.cpp:498-521, local_mapping.cpp:196-200): every leaf k_recut_layer marked as a
// BA factor is copied into the factor store (the reference copies the same fields into LidarFactor's vectors).
// blockIdx.y = layer. The order of the factors is the order of arrival (atomic cursor).
__global__ void __launch_bounds__(128) k_ba_collect(MapView M, LayerLists LL, BaFactor* __restrict__ out, int* __restrict__ count,
                                                    int cap)
{
  if (M.slide_count[M.slide_cur] + M.slide_others < M.thread_num) return;  // multi_recut's early-out
  int nn;
  const int* nodes = layer_nodes(M, LL, blockIdx.y, &nn);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nn; j += gridDim.x * blockDim.x)
  {
    const int n = nodes[j];
    if (M.hot[n].flags & VN_FLAG_INTERIOR) continue;
    const NodeCold& c = M.cold[n];
    if (c.opt_state < 0) continue;
    const int a = atomicAdd(count, 1);
    if (a >= cap)
    {
      atomicOr(M.status, VN_ST_NODES_FULL);
      continue;
    }
    BaFactor& f = out[a];
    for (int i = 0; i < M.win_size; i++) f.local[i] = c.pcrs_local[M.mp[i]];
    f.fix = c.pcr_fix;
    f.add = c.pcr_add;
    for (int k = 0; k < 3; k++) f.eig_value[k] = c.eig_value[k];
    for (int k = 0; k < 9; k++) f.eig_vector[k] = c.eig_vector[k];
    f.coe = 1.0;
    f.node = n;
  }
}

// The subdivision branch of OctoTree::recut (octree.cpp:375-387): fix_divide (:257-277), subdivide per
// window frame (:279-300), release of the parent's SlideWindow (:384-387). One 256-thread block per
// splitting leaf. Source classes in the reference's order: class 0 = point_fix, class 1+si =
// sw->points[mp[si]]. All classes (and all segments of the point_fix chain) form ONE row stream that is
// consumed 256 rows at a time (one row per thread) - a leaf's ~20 short lists cost two or three batches
// instead of one batch each. In steady state only a handful of leaves split per scan, so what matters in
// this kernel is the latency of one block, not throughput.
// Inside a batch the rows of each child are listed in stream order (stable compaction) and
//   thread t < 72  owns the cluster scalar s = t % 9 of child k = t / 9 (pcr_add and pcr_fix /
//                  pcrs_local[slot]) and applies that child's rows sequentially (exact sums, reference order;
//                  the per-class cluster is flushed whenever the class of the next row changes),
//   every thread   owns up to three (child, cov_add entry) pairs and sums that child's staged Bf_var terms.
#define VN_SPLIT_STAMP(k) \
  do \
  { \
    if (threadIdx.x == 0) KT(); \
  } while (0)
#define SPLIT_THREADS 256
#define SPLIT_BATCH 256  // rows per batch = threads: every thread stages one row
#define SPLIT_WARPS (SPLIT_BATCH / 32)
#define SPLIT_PAIRS ((360 + SPLIT_THREADS - 1) / SPLIT_THREADS)
#define SPLIT_SMEM ((SPLIT_BATCH * (19 + RED_STRIDE)) * sizeof(double))
#define SPLIT_MAXSEG 128
struct SplitSeg
{
  const PointRec* src;
  int start;  // stream index of the first row
  int cnt;
  int cls;
};

// number of rows of child kk among the first i rows of the batch (i in 0..SPLIT_BATCH)
__device__ __forceinline__ int split_rows_before(const unsigned int (*bm)[8], int kk, int i)
{
  int r = 0;
  const int w = i >> 5;
  for (int ww = 0; ww < w; ww++) r += __popc(bm[ww][kk]);
  if (i & 31) r += __popc(bm[w][kk] & ((1u << (i & 31)) - 1u));
  return r;
}

// The work queue of k_split: LL.split holds the leaves to subdivide, LL.count[4] = pushed, [5] = claimed, [6] =
// completed. k_recut_all pushes the leaves it finds; a block that subdivides a leaf judges the children it creates
// and pushes those that have to be subdivided themselves - all levels in ONE launch, a child's subdivision starts
// as soon as its parent's block is through instead of a launch later. Slots hold -1 until their item is written
// (the claim of a position and the write of the item are two steps) and are reset by the consumer.
// Returns the node to subdivide, or -1 when every pushed item has been completed (nothing can be pushed any more).
#define SPLIT_POP_LIMIT 4000000
__device__ int split_pop(const MapView& M, const LayerLists& LL, int* index)
{
  for (int spins = 0; spins < SPLIT_POP_LIMIT; spins++)
  {
    const int done = ld_volatile(&LL.count[6]);
    __threadfence();  // `done` is read before `tail`: done == tail then means nothing was in flight at that moment
    const int tail = ld_volatile(&LL.count[4]);
    const int head = ld_volatile(&LL.count[5]);
    if (head < tail)
    {
      if (atomicCAS(&LL.count[5], head, head + 1) != head) continue;
      int v = ld_volatile(&LL.split[head]);
      for (int w = 0; v < 0 && w < SPLIT_POP_LIMIT; w++) v = ld_volatile(&LL.split[head]);
      if (v < 0) break;
      LL.split[head] = -1;
      *index = head;
      return v;
    }
    if (done == tail) return -1;
    __nanosleep(100);
  }
  atomicOr(M.status, VN_ST_SPIN);
  return -1;
}
__device__ __forceinline__ void split_push(const LayerLists& LL, int node)
{
  __threadfence();  // the node's record (written by this block, behind a barrier) before the item
  const int pos = atomicAdd(&LL.count[4], 1);
  *((volatile int*)&LL.split[pos]) = node;
}

__global__ void __launch_bounds__(SPLIT_THREADS) k_split(MapView M, LayerLists LL, int win_count, PoseBuf xb, LivePose lv)
{
  vn_pdl_sync();
  extern __shared__ double split_smem[];
  double(*val)[19] = reinterpret_cast<double(*)[19]>(split_smem);  // per row: the 9 push() terms of the world point, then of the stored point
  double(*red)[RED_STRIDE] = reinterpret_cast<double(*)[RED_STRIDE]>(split_smem + SPLIT_BATCH * 19);
  __shared__ SplitSeg segs[SPLIT_MAXSEG];
  __shared__ int cbase[9];  // first slot of every child in the child-major row order of the batch
  __shared__ int cnt[11][8];
  __shared__ int off[11][8];
  __shared__ int fill[11][8];
  __shared__ int cls_first[11];
  __shared__ int kid[8];
  __shared__ unsigned int bm[SPLIT_WARPS][8];  // per staging warp and child: which rows go to that child
  __shared__ int clsrow[SPLIT_BATCH];
  __shared__ int nseg, total, nfix, fixtot;
  __shared__ int wcnt_s[VINA_MAX_WIN], woff_s[VINA_MAX_WIN];
  __shared__ int s_item, s_index;
  __shared__ double s_add[8][9];  // the children's pcr_add as the chains leave it (P lower triangle, v), for the judge
  __shared__ int s_addN[8], s_made[8];
  if (M.slide_count[M.slide_cur] + M.slide_others < M.thread_num) return;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  LaneRole L;
  // the 72 (child, cluster scalar) chains belong to the LAST 72 threads: those own one (child, cov entry) pair
  // each, the first 104 threads own two - the sequential chains no longer sit on the busiest threads
  const int ct = t - (SPLIT_THREADS - 72);
  const bool chain = ct >= 0;
  role_init(chain ? ct : 0, L);  // L.ck = ct % 9
  const int my_k = chain ? ct / 9 : 0;  // cluster chain of this thread
  for (;;)
  {
    __syncthreads();  // the previous leaf's shared state is no longer in use
    if (t == 0) s_item = split_pop(M, LL, &s_index);
    __syncthreads();
    const int n = s_item;
    if (n < 0) return;
    __threadfence();  // acquire: the leaf may have been written by another block of this launch
    KT_DECL;
    NodeCold& c = M.cold[n];
    NodeHot& h = M.hot[n];
    const int h_layer = h.layer;
    const bool store = (h_layer + 1) < M.max_layer;
    const int flags0 = h.flags;
    const double vc[3] = { h.vcenter[0], h.vcenter[1], h.vcenter[2] };
    const bool has_fix = c.pcr_fix.N != 0;
    __syncthreads();
    if (t < 88)
    {
      (&cnt[0][0])[t] = 0;
      (&off[0][0])[t] = -1;
      (&fill[0][0])[t] = 0;
    }
    __syncthreads();
    VN_SPLIT_STAMP(0);
    // the row stream: thread 0 walks the point_fix chain ONCE (dependent loads), threads 1..win_count fetch
    // the window frames' lists meanwhile; everything below works from this table. A chain has at most
    // max_points folds + 1 inherited segment (a leaf stops folding at pcr_fix.N >= max_points, octree.cpp:448),
    // so SPLIT_MAXSEG covers it; a longer one is reported, never truncated silently.
    if (t == 0)
    {
      int ns = 0, tot = 0;
      bool overflow = false;
      if (has_fix)
        for (int sg = c.fix_head; sg >= 0 && !overflow;)
        {
          const FixSeg& blk = M.fix_segs[sg];  // one 128-byte line: 15 segments per dependent load
          const int nb = blk.n;
          for (int e = 0; e < nb; e++)
          {
            const int scnt = blk.cnt[e];
            if (scnt <= 0) continue;
            if (ns >= SPLIT_MAXSEG - VINA_MAX_WIN)
            {
              overflow = true;
              break;
            }
            segs[ns].src = M.fix_pool + blk.off[e];
            segs[ns].start = tot;
            segs[ns].cnt = scnt;
            segs[ns].cls = 0;
            tot += scnt;
            ns++;
          }
          sg = blk.next;
        }
      if (overflow) atomicOr(M.status, VN_ST_FIX_FULL);
      nfix = ns;
      fixtot = tot;
    }
    else if (t <= win_count)
    {
      const int slot = M.mp[t - 1];
      wcnt_s[t - 1] = c.win_cnt[slot];
      woff_s[t - 1] = c.win_off[slot];
    }
    __syncthreads();
    if (t == 0)
    {
      int ns = nfix, tot = fixtot;
      for (int q = 0; q < 11; q++) cls_first[q] = -1;
      if (ns > 0) cls_first[0] = 0;
      for (int si = 0; si < win_count; si++)
        if (wcnt_s[si] > 0)
        {
          cls_first[1 + si] = tot;
          segs[ns].src = M.win_pool[M.mp[si]] + woff_s[si];
          segs[ns].start = tot;
          segs[ns].cnt = wcnt_s[si];
          segs[ns].cls = 1 + si;
          tot += wcnt_s[si];
          ns++;
        }
      nseg = ns;
      total = tot;
    }
    __syncthreads();
    VN_SPLIT_STAMP(1);
    // pass 1: how many points of every class go to every child
    {
      const int ns = nseg, tot = total;
      for (int g = t; g < tot; g += SPLIT_THREADS)
      {
        int sgi = 0, hi = ns - 1;  // last segment that starts at or before g
        while (sgi < hi)
        {
          const int mid = (sgi + hi + 1) >> 1;
          if (segs[mid].start <= g)
            sgi = mid;
          else
            hi = mid - 1;
        }
        const int cls = segs[sgi].cls;
        const PointRec* src = segs[sgi].src + (g - segs[sgi].start);
        double pw[3];
        if (cls == 0)
        {
          pw[0] = src->p[0];
          pw[1] = src->p[1];
          pw[2] = src->p[2];
        }
        else
        {
          double xr[9], xp[3];
          pose_sel(xb, lv, cls - 1, xr, xp);
          rot_trans(xr, xp, src->p, pw);
        }
        atomicAdd(&cnt[cls][child_index(pw, vc)], 1);
      }
    }
    __syncthreads();
    VN_SPLIT_STAMP(2);
    // children (thread k owns child k), then their storage in parallel over (child, class)
    if (t < 8)
    {
      const int k = t;
      int tot = 0;
      for (int cls = 0; cls <= win_count; cls++) tot += cnt[cls][k];
      int id = -1;
      if (tot > 0)
      {
        s_made[k] = c.children[k] < 0 ? 1 : 0;
        id = c.children[k] >= 0 ? c.children[k] : make_child(M, n, k);
        c.children[k] = id;
        if (id >= 0)
        {
          NodeCold& kc = M.cold[id];
          if (cnt[0][k] > 0 && store) off[0][k] = fix_append(M, kc, cnt[0][k]);
          if (tot > cnt[0][k])
          {
            kc.has_sw = 1;
            kc.isexist = 1;
          }
        }
      }
      kid[k] = id;
    }
    __syncthreads();
    VN_SPLIT_STAMP(3);
    if (t < 8 * win_count && store)
    {
      const int k = t & 7, si = t >> 3;
      const int m = cnt[1 + si][k];
      if (m > 0 && kid[k] >= 0)
      {
        const int slot = M.mp[si];
        NodeCold& kc = M.cold[kid[k]];
        const int woff = atomicAdd(&M.win_cursor[slot], m);
        if ((long long)woff + m > M.win_cap)
          atomicOr(M.status, VN_ST_WIN_FULL);
        else
        {
          off[1 + si][k] = woff;
          kc.win_off[slot] = woff;
          kc.win_cnt[slot] = m;
        }
      }
    }
    __syncthreads();
    VN_SPLIT_STAMP(4);

    // running sums (children are new: they start from zero)
    double clA = 0.0, clB = 0.0;  // thread t < 72: scalar L.ck of child my_k
    int cur_cls = -1;             // class clB currently accumulates
    double cv[SPLIT_PAIRS];  // pairs p = t + SPLIT_THREADS q < 360: child p / 45, entry p % 45
    for (int q = 0; q < SPLIT_PAIRS; q++) cv[q] = 0.0;

    // pass 2: the row stream, 256 rows at a time. The row of the NEXT batch is fetched while this one is processed
    // (a heavy leaf is twenty batches: one L2 round trip each on the block's critical path otherwise); the segment of a
    // row by bisection (a leaf has up to ~50 segments: a linear walk per row and batch was ~1000 cycles)
    {
      const int ns = nseg, tot = total;
      auto fetch_row = [&](int g, PointRec& out, int& out_cls) {
        if (g >= tot) return;
        int lo = 0, hi = ns - 1;
        while (lo < hi)
        {
          const int mid = (lo + hi + 1) >> 1;
          if (segs[mid].start <= g)
            lo = mid;
          else
            hi = mid - 1;
        }
        out_cls = segs[lo].cls;
        out = segs[lo].src[g - segs[lo].start];
      };
      PointRec pr_next;
      int cls_next = 0;
      if (t < SPLIT_BATCH) fetch_row(t, pr_next, cls_next);
      for (int base = 0; base < tot; base += SPLIT_BATCH)
      {
        // phase 1 (rows in stream order): load, world position, child index
        int kk = -1, cls = 0;
        PointRec pr;
        double pw[3] = { 0.0, 0.0, 0.0 };
        if (t < SPLIT_BATCH)
        {
          const int g = base + t;
          pr = pr_next;
          cls = cls_next;
          fetch_row(g + SPLIT_BATCH, pr_next, cls_next);
          if (g < tot)
          {
            if (cls == 0)
            {
              pw[0] = pr.p[0];
              pw[1] = pr.p[1];
              pw[2] = pr.p[2];
            }
            else
            {
              double xr[9], xp[3];
              pose_sel(xb, lv, cls - 1, xr, xp);
              rot_trans(xr, xp, pr.p, pw);
            }
            kk = child_index(pw, vc);
          }
#pragma unroll
          for (int k = 0; k < 8; k++)
          {
            const unsigned mask = __ballot_sync(0xffffffffu, kk == k);
            if (lane == 0) bm[warp][k] = mask;
          }
        }
        __syncthreads();
        VN_SPLIT_STAMP(51);
        // phase 2: every row goes to its slot in child-major order (stable: stream order inside a child), so
        // that the sequential consumers below read contiguous rows; the products of PointCluster::push
        // (types.hpp:137-142) are formed here, in parallel - the chains only add
        if (t < 8)
        {
          int o = 0;
          for (int k = 0; k < t; k++) o += split_rows_before(bm, k, SPLIT_BATCH);
          cbase[t] = o;
          if (t == 7) cbase[8] = o + split_rows_before(bm, 7, SPLIT_BATCH);
        }
        if (t < SPLIT_BATCH && kk >= 0)
        {
          const int cs = cls_first[cls] - base;  // batch-local index of the first row of this class
          const int before = split_rows_before(bm, kk, t);                       // ... inside its child
          const int before_cls = before - split_rows_before(bm, kk, cs > 0 ? cs : 0);  // ... and inside its class
          int slot = before;
          for (int k = 0; k < kk; k++) slot += split_rows_before(bm, k, SPLIT_BATCH);
          const double* s3[2] = { pw, pr.p };
#pragma unroll
          for (int w = 0; w < 2; w++)
          {
            const double* q = s3[w];
            double* v = val[slot] + 9 * w;
            v[0] = dm(q[0], q[0]);
            v[1] = dm(q[1], q[0]);
            v[2] = dm(q[2], q[0]);
            v[3] = dm(q[1], q[1]);
            v[4] = dm(q[2], q[1]);
            v[5] = dm(q[2], q[2]);
            v[6] = q[0];
            v[7] = q[1];
            v[8] = q[2];
          }
          clsrow[slot] = cls;
          double o[45];
          bf_var_terms(pr.v, pw, o);
#pragma unroll
          for (int e = 0; e < 45; e++) red[slot][e] = o[e];
          if (off[cls][kk] >= 0)
          {
            const int dst = off[cls][kk] + fill[cls][kk] + before_cls;
            if (cls == 0)
              M.fix_pool[dst] = pr;
            else
              M.win_pool[M.mp[cls - 1]][dst] = pr;
          }
        }
        __syncthreads();
        VN_SPLIT_STAMP(52);
        // phase 3: push_fix (octree.cpp:179-188) / push (octree.cpp:151-177) of this batch's rows. The adds are
        // sequential (reference order, exact sums); the shared-memory loads are not: 8 rows are fetched ahead of
        // the chain so that a row costs one dependent add instead of a load-to-use latency
        if (chain)
        {
          const int r0 = cbase[my_k], r1 = cbase[my_k + 1];
          for (int rb = r0; rb < r1; rb += 8)
          {
            double va[8], vb[8];
            int rcs[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
            {
              const int r = rb + u < r1 ? rb + u : r1 - 1;
              va[u] = val[r][L.ck];
              vb[u] = val[r][9 + L.ck];
              rcs[u] = clsrow[r];
            }
            if (rb + 8 <= r1 && rcs[0] == cur_cls && rcs[7] == cur_cls)
            {
              // eight rows of the class in progress (a child's rows come in class order): two straight-line add chains
#pragma unroll
              for (int u = 0; u < 8; u++)
              {
                clA = da(clA, va[u]);
                clB = da(clB, vb[u]);
              }
              continue;
            }
#pragma unroll
            for (int u = 0; u < 8; u++)
            {
              if (rb + u >= r1) break;
              const int rc = rcs[u];
              if (rc != cur_cls)
              {
                // the per-class cluster (pcr_fix or pcrs_local[slot]) of the finished class is complete
                if (cur_cls >= 0 && kid[my_k] >= 0)
                {
                  NodeCold& kc = M.cold[kid[my_k]];
                  Cluster& dst = cur_cls == 0 ? kc.pcr_fix : kc.pcrs_local[M.mp[cur_cls - 1]];
                  cluster_set(dst, L.ck, clB);  // (the counts go in behind the loop: a read-modify-write here puts a
                                                // global round trip per class on the chain)
                }
                cur_cls = rc;
                clB = 0.0;
              }
              clA = da(clA, va[u]);
              clB = da(clB, vb[u]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < SPLIT_PAIRS; q++)
        {
          const int p = t + SPLIT_THREADS * q;
          if (p < 360)
          {
            const int k = p / 45, e = p % 45;
            const int r0 = cbase[k], r1 = cbase[k + 1];
            // (cov_add is toleranced, 1e-12: four interleaved partial sums break the add chain)
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int r = r0;
            for (; r + 3 < r1; r += 4)
            {
              s0 += red[r][e];
              s1 += red[r + 1][e];
              s2 += red[r + 2][e];
              s3 += red[r + 3][e];
            }
            for (; r < r1; r++) s0 += red[r][e];
            cv[q] += (s0 + s1) + (s2 + s3);
          }
        }
        if (t < SPLIT_BATCH && kk >= 0) atomicAdd(&fill[cls][kk], 1);
        __syncthreads();
        VN_SPLIT_STAMP(5);
      }
    }
    // last class of every chain
    if (chain && cur_cls >= 0 && kid[my_k] >= 0)
    {
      NodeCold& kc = M.cold[kid[my_k]];
      Cluster& dst = cur_cls == 0 ? kc.pcr_fix : kc.pcrs_local[M.mp[cur_cls - 1]];
      cluster_set(dst, L.ck, clB);
    }
    // children's pcr_add / cov_add
    if (chain && kid[my_k] >= 0)
    {
      NodeCold& kc = M.cold[kid[my_k]];
      cluster_set(kc.pcr_add, L.ck, clA);
      s_add[my_k][L.ck] = clA;
    }
    // the point counts of the clusters the chains filled: one thread per (class, child) and per child, all in flight
    // together (PointCluster::push counts every point: N += 1 per row)
    if (t >= 128 && t < 128 + 8 * (win_count + 1))  // (not thread 0: it has the parent's bookkeeping below)
    {
      const int k = (t - 128) & 7, cls = (t - 128) >> 3;
      if (cnt[cls][k] > 0 && kid[k] >= 0)
      {
        NodeCold& kc = M.cold[kid[k]];
        Cluster& dst = cls == 0 ? kc.pcr_fix : kc.pcrs_local[M.mp[cls - 1]];
        dst.N += cnt[cls][k];
      }
    }
    else if (t >= 232 && t < 240 && kid[t - 232] >= 0)
    {
      int tot = 0;
      for (int cls = 0; cls <= win_count; cls++) tot += cnt[cls][t - 232];
      M.cold[kid[t - 232]].pcr_add.N += tot;
      s_addN[t - 232] = tot;
    }
#pragma unroll
    for (int q = 0; q < SPLIT_PAIRS; q++)
    {
      const int p = t + SPLIT_THREADS * q;
      if (p < 360 && kid[p / 45] >= 0) M.cold[kid[p / 45]].cov_add[p % 45] = cv[q];
    }
    // PVec().swap(point_fix); sw->clear(); sws.push_back(sw); sw = nullptr; octo_state = 1 (spread over threads:
    // one thread doing the ~120 stores and 8 loads in a row is 3 us of this block's latency)
    if (t == 0)
    {
      if (has_fix)
      {
        c.fix_head = c.fix_tail = -1;
        c.fix_count = 0;
      }
      c.has_sw = 0;
      // (one store, from the value read when the leaf was claimed - nobody else writes this node's flags meanwhile: a
      // concurrent margi pass skips both states, and thread 0 does not wait for a load here)
      h.flags = (flags0 | VN_FLAG_INTERIOR) & ~VN_FLAG_SPLIT_PENDING;
    }
    else if (t >= 32 && t < 32 + M.win_size)
    {
      c.win_cnt[t - 32] = 0;
      cluster_clear(c.pcrs_local[t - 32]);
    }
    else if (t >= 64 && t < 72)
      h.children[t - 64] = c.children[t - 64];  // mirror read by k_iekf
    VN_SPLIT_STAMP(6);
    // leaves[i]->recut(...) of the new children (octree.cpp:388-392), one thread per child: their sums are
    // complete (written by this block), nobody else knows them yet. A child that has to be subdivided itself
    // goes into the queue; every child joins its layer's node list for the multi_margi of this scan.
    __syncthreads();
    const int child_layer = h_layer + 1;
    if (t < 8 && kid[t] >= 0)
    {
      NodeHot& kh = M.hot[kid[t]];
      bool again;
      if (s_made[t])
      {
        // a child this block created: everything the judge needs is still on chip
        Cluster add;
        for (int q = 0; q < 6; q++) add.P[q] = s_add[t][q];
        for (int q = 0; q < 3; q++) add.v[q] = s_add[t][6 + q];
        add.N = s_addN[t];
        again = recut_child(M, kh, M.cold[kid[t]], add, s_addN[t] > cnt[0][t], child_layer);
      }
      else
        again = !(kh.flags & VN_FLAG_INTERIOR) && recut_leaf(M, kh, M.cold[kid[t]]);
      if (again) split_push(LL, kid[t]);
    }
    else if (t >= 32 && t < 40 && kid[t - 32] >= 0 && child_layer <= 3)  // (next to the judge: the children join their layer's list)
      LL.list[child_layer][atomicAdd(&LL.count[child_layer], 1)] = kid[t - 32];
    __syncthreads();
    if (t == 0)
    {
      KT();
#ifdef VINA_SPLIT_TRACE
      for (int z = kt_k; z < KT_PH - 1; z++) kt_[z] = 0;
      kt_[KT_PH - 1] = total + 1000000ll * h.layer;  // (rows of the leaf + 10^6 x its layer, shown as the tag)
      kt_k = KT_PH;
#endif
      KT_COMMIT(0);
      __threadfence();
      atomicAdd(&LL.count[6], 1);  // this leaf is complete (its pushes are in)
    }
  }
}
#ifdef VINA_SPLIT_TRACE
void vn_split_trace_dump()
{
  static long long h[3][KT_SLOTS][KT_PH];
  int n[3];
  cudaMemcpyFromSymbol(h, g_kt, sizeof(h));
  cudaMemcpyFromSymbol(n, g_kt_n, sizeof(n));
  const char* names[3] = { "k_split (per leaf: block thread 0)", "k_margi_leaves (per leaf: lane 0 of the group)", "k_insert_accum (per leaf: lane 0)" };
  for (int w = 0; w < 3; w++)
  {
    const int cnt = n[w] < KT_SLOTS ? n[w] : KT_SLOTS;
    if (cnt == 0) continue;
    double mean[KT_PH] = { 0 }, mx[KT_PH] = { 0 };
    int worst = 0;
    long long worst_t = 0;
    int np = 0;
    for (int s = 0; s < cnt; s++)
    {
      int k = 0;
      while (k < KT_PH - 1 && h[w][s][k] != 0) k++;
      if (k > np) np = k;
      for (int i = 1; i < k; i++)
      {
        const double d = (double)(h[w][s][i] - h[w][s][i - 1]);
        mean[i] += d / cnt;
        if (d > mx[i]) mx[i] = d;
      }
      if (k > 1 && h[w][s][k - 1] - h[w][s][0] > worst_t) worst_t = h[w][s][k - 1] - h[w][s][0], worst = s;
    }
    fprintf(stderr, "[vina ktrace] %s: %d units (all scans), phases in cycles\n  mean:", names[w], n[w]);
    for (int i = 1; i < np; i++) fprintf(stderr, " %.0f", mean[i]);
    fprintf(stderr, "\n  max: ");
    for (int i = 1; i < np; i++) fprintf(stderr, " %.0f", mx[i]);
    fprintf(stderr, "\n  slowest unit (%lld cycles, tag %lld):", worst_t, h[w][worst][KT_PH - 1]);
    for (int i = 1; i < np && h[w][worst][i] != 0; i++) fprintf(stderr, " %lld", h[w][worst][i] - h[w][worst][i - 1]);
    fprintf(stderr, "\n");
    if (w == 0)
    {
      // the last 40 units (the last scans): tag, total, phases
      for (int s = cnt > 40 ? cnt - 40 : 0; s < cnt; s++)
      {
        int k = 0;
        while (k < KT_PH - 1 && h[w][s][k] != 0) k++;
        fprintf(stderr, "    unit %d tag %lld start %lld total %lld:", s, h[w][s][KT_PH - 1], h[w][s][0] % 100000000ll,
                k > 1 ? h[w][s][k - 1] - h[w][s][0] : 0ll);
        for (int i = 1; i < k; i++) fprintf(stderr, " %lld", h[w][s][i] - h[w][s][i - 1]);
        fprintf(stderr, "\n");
      }
    }
  }
}
#endif



// OctoTree::plane_update (octree.cpp:302-333)
__device__ void plane_update(NodeHot& h, NodeCold& c)
{
  const double N = (double)c.pcr_add.N;
  double center[3] = { c.pcr_add.v[0] / N, c.pcr_add.v[1] / N, c.pcr_add.v[2] / N };
  const double nv = 1.0 / N;
  double u[3][3];
  for (int k = 0; k < 3; k++)
    for (int r = 0; r < 3; r++) u[k][r] = c.eig_vector[r + 3 * k];
  double uc[3][9];
  for (int r = 0; r < 3; r++)
    for (int q = 0; q < 9; q++) uc[r][q] = 0.0;
  const int l = 0;
  for (int k = 1; k < 3; k++)
  {
    double ukl[3][3];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) ukl[a][b] = u[k][a] * u[l][b];
    double f[9];
    f[0] = ukl[0][0];
    f[1] = ukl[1][0] + ukl[0][1];
    f[2] = ukl[2][0] + ukl[0][2];
    f[3] = ukl[1][1];
    f[4] = ukl[1][2] + ukl[2][1];
    f[5] = ukl[2][2];
    double dk = (u[k][0] * center[0] + u[k][1] * center[1]) + u[k][2] * center[2];
    double dl = (u[l][0] * center[0] + u[l][1] * center[1]) + u[l][2] * center[2];
    for (int a = 0; a < 3; a++) f[6 + a] = -(dk * u[l][a] + dl * u[k][a]);
    double coef = nv / (c.eig_value[l] - c.eig_value[k]);
    for (int r = 0; r < 3; r++)
    {
      double cu = coef * u[k][r];
      for (int q = 0; q < 9; q++) uc[r][q] += cu * f[q];
    }
  }
  double Jc[3][9];
  for (int r = 0; r < 3; r++)
    for (int q = 0; q < 9; q++)
    {
      double s = 0.0;
      for (int t = 0; t < 9; t++) s += uc[r][t] * c.cov_add[sN(9, t, q)];
      Jc[r][q] = s;
    }
  // plane_var = [[Jc u_c^T, nv Jc(:,6:9)],[.^T, nv^2 cov_add(6:9,6:9)]], stored as the upper triangle
  for (int a = 0; a < 3; a++)
    for (int b = a; b < 3; b++)
    {
      double s = 0.0;
      for (int t = 0; t < 9; t++) s += Jc[a][t] * uc[b][t];
      c.plane_var[sN(6, a, b)] = s;
    }
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) c.plane_var[sN(6, a, 3 + b)] = nv * Jc[a][6 + b];
  for (int a = 0; a < 3; a++)
    for (int b = a; b < 3; b++) c.plane_var[sN(6, 3 + a, 3 + b)] = (nv * nv) * c.cov_add[sN(9, 6 + a, 6 + b)];
  for (int a = 0; a < 3; a++)
  {
    h.center[a] = center[a];
    h.normal[a] = u[0][a];
  }
  h.radius = (float)c.eig_value[2];
  // per-plane part of sigma_l (see NodeHot): A, B n, n^T C n
  const double* pv = c.plane_var;
  const double* nrm = u[0];
  for (int a = 0, q = 0; a < 3; a++)
    for (int b = a; b < 3; b++, q++) h.qA[q] = pv[sN(6, a, b)];
  for (int a = 0; a < 3; a++)
    h.qb[a] = pv[sN(6, a, 3)] * nrm[0] + pv[sN(6, a, 4)] * nrm[1] + pv[sN(6, a, 5)] * nrm[2];
  double qk = 0.0;
  for (int a = 0; a < 3; a++)
    qk += nrm[a] * (pv[sN(6, 3 + a, 3)] * nrm[0] + pv[sN(6, 3 + a, 4)] * nrm[1] + pv[sN(6, 3 + a, 5)] * nrm[2]);
  h.qk = qk;
}

// OctoTree::margi, leaf branch (octree.cpp:397-484, mgsize = 1) + plane_update (octree.cpp:302-333), EIGHT lanes per
// leaf. A thread per leaf is one long dependent chain (ten cluster transforms, the eigen-solver, the 3x9 * 9x9
// products of plane_update) on a machine that is nearly empty (~10^4 leaves): the kernel's time is that chain's
// latency. Here the lanes of a group share it:
//   * lane g transforms the clusters of window frames g and g + 8 (PointCluster::transform) into shared memory; every
//     lane then adds them up in frame order - the same additions in the same order as the reference, so pcr_add
//     stays bit-exact - and runs the eigen-solver on the sum (replicated: its result is needed by all);
//   * plane_update: the 27 entries of u_c, the 27 of J_c = u_c cov_add and the blocks of plane_var are spread
//     over the lanes (every entry is one lane's sum in the order it always had);
//   * lane 0 does the bookkeeping (fold into pcr_fix or subtract, slot clearing, isexist) and owns the copy job.
#define MG 8  // lanes per leaf
struct MargiShared
{
  double w[VINA_MAX_WIN][10];  // transformed clusters: P[6], v[3], N
  double cov[45];              // cov_add
  double f[2][9];              // plane_update: the row vectors of k = 1, 2
  double uc[27], Jc[27];
};

// One pass of a warp: the four leaves nodes[j0 .. j0 + 3], one per 8-lane group. The groups take different branches
// (BA-factor leaf or not, plane or not, update due or not); the phases are separated by FULL-warp barriers so that
// the groups reconverge after every phase and run the common ones in lock step - without them four diverged groups
// execute one after the other and a warp takes the sum of its leaves' times instead of the longest.
// mode 0: every leaf of the list. mode 1 (runs next to k_split): the leaves multi_recut has not queued for
// subdivision. mode 2 (after k_split): what mode 1 left out - the children the subdivisions created (list positions
// >= n_old) and any queued leaf that was not subdivided after all.
__device__ void margi_warp_pass(const MapView& M, const int* __restrict__ nodes, int nn, int j0, int win_count,
                                const PoseBuf& xb, const LivePose& lv, int lane, MargiShared* sh4,
                                const PointRec*& job_src, int& job_off, int& job_np, int mode, int n_old)
{
  const int g = lane & (MG - 1), grp = lane / MG;
  MargiShared& sh = sh4[grp];
  const int j = j0 + grp;
  KT_DECL;
  KT();
  int n = 0;
  bool active = false;
  if (j < nn)
  {
    n = nodes[j];
    const int fl = M.hot[n].flags;
    active = !(fl & VN_FLAG_INTERIOR);
    if (mode == 1) active = active && !(fl & VN_FLAG_SPLIT_PENDING);
    if (mode == 2 && j < n_old)
    {
      active = active && (fl & VN_FLAG_SPLIT_PENDING);
      if (active && g == 0) M.hot[n].flags = fl & ~VN_FLAG_SPLIT_PENDING;  // (queued, but not subdivided after all)
    }
  }
  NodeHot& h = M.hot[n];
  NodeCold& c = M.cold[n];
  if (active)
  {
    prefetch_cold(&c, g, MG);
    active = c.isexist && c.has_sw;
  }
  KT();
  const int s0 = M.mp[0];
  bool is_plane = false, factor = false, have_eig = false;
  Cluster fix, add, world0;
  cluster_clear(world0);
  cluster_clear(fix);
  cluster_clear(add);
  double ev[3] = { 0, 0, 0 }, Q[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
  int n_loc0 = 0;
  // ---- phase A: the frames' clusters into the world frame (PointCluster::transform)
  if (active)
  {
    is_plane = (h.flags & VN_FLAG_PLANE) != 0;
    factor = c.opt_state >= 0;
    fix = c.pcr_fix;
    if (factor)
    {
      add = c.pcr_add;
      const Cluster loc0 = c.pcrs_local[s0];
      n_loc0 = loc0.N;
      if (loc0.N != 0)
      {
        double xr[9], xp[3];
        pose_sel(xb, lv, 0, xr, xp);
        cluster_transform(world0, loc0, xr, xp);
      }
    }
    else
    {
      for (int i = g; i < win_count; i += MG)
      {
        const Cluster loc = c.pcrs_local[M.mp[i]];
        double* w = sh.w[i];
        w[9] = (double)loc.N;
        if (loc.N != 0)
        {
          Cluster t;
          double xr[9], xp[3];
          pose_sel(xb, lv, i, xr, xp);
          cluster_transform(t, loc, xr, xp);
#pragma unroll
          for (int k = 0; k < 6; k++) w[k] = t.P[k];
#pragma unroll
          for (int k = 0; k < 3; k++) w[6 + k] = t.v[k];
        }
      }
    }
  }
  __syncwarp();
  KT();
  // ---- phase B: pcr_add = pcr_fix + the frames in order (the reference's additions, octree.cpp:425-431), then the
  // eigen-decomposition of a plane (:432-438)
  if (active && !factor)
  {
    add = fix;
    n_loc0 = (int)sh.w[0][9];
    for (int i = 0; i < win_count; i++)
    {
      const double* w = sh.w[i];
      if (w[9] != 0.0)
      {
        Cluster t;
#pragma unroll
        for (int k = 0; k < 6; k++) t.P[k] = w[k];
#pragma unroll
        for (int k = 0; k < 3; k++) t.v[k] = w[6 + k];
        t.N = (int)w[9];
        if (i == 0) world0 = t;
        cluster_add(add, t);
      }
    }
  }
  __syncwarp();
  KT();
  if (active && !factor && is_plane)
  {
    double L[6];
    cluster_cov(add, L);
    eig3_sym(L, ev, Q);
    have_eig = true;
  }
  __syncwarp();
  KT();
  // ---- phase C: plane_update (octree.cpp:302-333, 441-446)
  const int last_num = active ? c.last_num : 0;
  const bool update = active && fix.N < M.max_points && is_plane && (add.N - last_num >= 5 || last_num <= 10);
  double nv = 0.0, coef0 = 0.0, coef1 = 0.0;
  double center[3] = { 0, 0, 0 };
  if (update)
  {
    if (!have_eig)
    {
#pragma unroll
      for (int k = 0; k < 3; k++) ev[k] = c.eig_value[k];
#pragma unroll
      for (int k = 0; k < 9; k++) Q[k] = c.eig_vector[k];
    }
    for (int e = g; e < 45; e += MG) sh.cov[e] = c.cov_add[e];
    const double N = (double)add.N;
    center[0] = add.v[0] / N;
    center[1] = add.v[1] / N;
    center[2] = add.v[2] / N;
    nv = 1.0 / N;
    if (g < 2)
    {
      // lanes 0 and 1 build the row vectors of k = 1, 2 (l = 0; columns picked by selects: Q stays in registers)
      const double uk[3] = { g == 0 ? Q[3] : Q[6], g == 0 ? Q[4] : Q[7], g == 0 ? Q[5] : Q[8] };
      const double* ul = Q;
      double* f = sh.f[g];
      f[0] = uk[0] * ul[0];
      f[1] = uk[1] * ul[0] + uk[0] * ul[1];
      f[2] = uk[2] * ul[0] + uk[0] * ul[2];
      f[3] = uk[1] * ul[1];
      f[4] = uk[1] * ul[2] + uk[2] * ul[1];
      f[5] = uk[2] * ul[2];
      const double dk = (uk[0] * center[0] + uk[1] * center[1]) + uk[2] * center[2];
      const double dl = (ul[0] * center[0] + ul[1] * center[1]) + ul[2] * center[2];
      for (int a = 0; a < 3; a++) f[6 + a] = -(dk * ul[a] + dl * uk[a]);
    }
    coef0 = nv / (ev[0] - ev[1]);
    coef1 = nv / (ev[0] - ev[2]);
  }
  __syncwarp();
  if (update)
    for (int o = g; o < 27; o += MG)
    {
      const int r = o / 9, q = o - 9 * r;
      const double u1r = r == 0 ? Q[3] : (r == 1 ? Q[4] : Q[5]);  // (no dynamically indexed register arrays)
      const double u2r = r == 0 ? Q[6] : (r == 1 ? Q[7] : Q[8]);
      double s = 0.0;
      s += (coef0 * u1r) * sh.f[0][q];
      s += (coef1 * u2r) * sh.f[1][q];
      sh.uc[o] = s;
    }
  __syncwarp();
  if (update)
    for (int o = g; o < 27; o += MG)
    {
      const int r = o / 9, q = o - 9 * r;
      // (plane_var is toleranced, 1e-7: three interleaved fused partial sums instead of one chain of nine
      // multiply-add pairs - a dependent fp64 operation costs ~25-30 cycles here)
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int t = 0; t < 3; t++)
      {
        s0 = fma(sh.uc[9 * r + t], sh.cov[sN(9, t, q)], s0);
        s1 = fma(sh.uc[9 * r + 3 + t], sh.cov[sN(9, 3 + t, q)], s1);
        s2 = fma(sh.uc[9 * r + 6 + t], sh.cov[sN(9, 6 + t, q)], s2);
      }
      sh.Jc[o] = (s0 + s1) + s2;
    }
  __syncwarp();
  if (update)
  {
    // plane_var = [[Jc u_c^T, nv Jc(:,6:9)],[.^T, nv^2 cov_add(6:9,6:9)]] (upper triangle) and the per-plane part
    // of sigma_l for the IEKF (NodeHot: A, B n, n^T C n)
    const double* nrm = Q;  // u[0]
    if (g < 6)
    {
      const int a = g < 3 ? 0 : (g < 5 ? 1 : 2), b = g < 3 ? g : (g < 5 ? g - 2 : 2);  // (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int t = 0; t < 3; t++)
      {
        s0 = fma(sh.Jc[9 * a + t], sh.uc[9 * b + t], s0);
        s1 = fma(sh.Jc[9 * a + 3 + t], sh.uc[9 * b + 3 + t], s1);
        s2 = fma(sh.Jc[9 * a + 6 + t], sh.uc[9 * b + 6 + t], s2);
      }
      const double s = (s0 + s1) + s2;
      c.plane_var[sN(6, a, b)] = s;
      h.qA[g] = s;
      c.plane_var[sN(6, 3 + a, 3 + b)] = (nv * nv) * sh.cov[sN(9, 6 + a, 6 + b)];
    }
    for (int o = g; o < 9; o += MG) c.plane_var[sN(6, o / 3, 3 + o % 3)] = nv * sh.Jc[9 * (o / 3) + 6 + o % 3];
    if (g < 3)
    {
      const int a = g;
      h.qb[a] = (nv * sh.Jc[9 * a + 6]) * nrm[0] + (nv * sh.Jc[9 * a + 7]) * nrm[1] + (nv * sh.Jc[9 * a + 8]) * nrm[2];
    }
    if (g == 6)
    {
      double qk = 0.0;
      for (int a = 0; a < 3; a++)
        qk += nrm[a] * (((nv * nv) * sh.cov[sN(9, 6 + a, 6)]) * nrm[0] + ((nv * nv) * sh.cov[sN(9, 6 + a, 7)]) * nrm[1] +
                        ((nv * nv) * sh.cov[sN(9, 6 + a, 8)]) * nrm[2]);
      h.qk = qk;
    }
    if (g == 7)
    {
      for (int a = 0; a < 3; a++)
      {
        h.center[a] = center[a];
        h.normal[a] = nrm[a];
      }
      h.radius = (float)ev[2];
      c.last_num = add.N;
    }
  }
  __syncwarp();  // (the next pass reuses `sh`)
  KT();
  // ---- bookkeeping (octree.cpp:448-481): lane 0 of the group
  if (!active || g != 0) return;
  if (factor)
    c.opt_state = -1;
  else if (have_eig)
  {
    for (int k = 0; k < 3; k++) c.eig_value[k] = ev[k];
    for (int k = 0; k < 9; k++) c.eig_vector[k] = Q[k];
  }
  if (fix.N < M.max_points)
  {
    if (!factor) c.pcr_add = add;
    if (world0.N != 0)
    {
      cluster_add(fix, world0);
      c.pcr_fix = fix;
      const int np = c.win_cnt[s0];
      if (np > 0)
      {
        int off = fix_append(M, c, np);
        if (off >= 0)
        {
          job_src = M.win_pool[s0] + c.win_off[s0];
          job_off = off;
          job_np = np;
        }
      }
    }
  }
  else
  {
    if (world0.N != 0) cluster_sub(add, world0);
    if (!factor || world0.N != 0) c.pcr_add = add;
    c.fix_head = c.fix_tail = -1;
    c.fix_count = 0;
  }
  if (n_loc0 != 0)
  {
    cluster_clear(c.pcrs_local[s0]);
    c.win_cnt[s0] = 0;
  }
  c.isexist = (fix.N >= add.N) ? 0 : 1;
  KT();
  KT_COMMIT(1);
}

// OctoTree::margi, leaf branch, for every leaf under surf_map_slide (blockIdx.y = layer)
__global__ void __launch_bounds__(128, 4) k_margi_leaves(MapView M, LayerLists LL, int win_count, PoseBuf xb, LivePose lv, int mode)
{
  vn_pdl_sync();
  __shared__ MargiShared sh_all[128 / MG];
  // (the slide list the compaction fills after this kernel starts empty)
  if (mode != 2 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) M.slide_count[1 - M.slide_cur] = 0;
  if (M.slide_count[M.slide_cur] + M.slide_others < M.thread_num) return;  // local_mapping.cpp:26-28
  // (the deepest layer first: it holds most of the leaves, and the grid's rows are scheduled in order - 592 blocks are
  // resident at a time)
  const int layer = M.max_layer - (int)blockIdx.y;
  int nn;
  const int* nodes = layer_nodes(M, LL, layer, &nn);
  int n_old = nn;  // layer 0 is the slide list: subdivisions add nothing to it
  if (layer > 0 && mode != 0)
  {
    n_old = LL.snap[layer];
    if (mode == 1) nn = n_old;  // (k_split may be appending behind it right now)
  }
  const int lane = threadIdx.x & 31;
  MargiShared* sh4 = sh_all + (threadIdx.x >> 5) * (32 / MG);
  const int gpw = 32 / MG;  // leaves per warp and pass
  const int stride = gridDim.x * (blockDim.x / 32) * gpw;
  for (int j0 = (blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) * gpw; j0 < nn; j0 += stride)  // warp-uniform
  {
    const PointRec* job_src = nullptr;
    int job_off = 0, job_np = 0;
    margi_warp_pass(M, nodes, nn, j0, win_count, xb, lv, lane, sh4, job_src, job_off, job_np, mode, n_old);
    __syncwarp();
    // the warp's copy jobs as ONE stream of points, 32 per pass (points go to the world frame of x_buf[0]): a
    // leaf folds only a handful of points per scan, so walking the jobs one after the other would leave most
    // lanes idle and pay one load latency per leaf instead of one per 32 points
    int incl = job_np;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - job_np;
    for (int base = 0; base < total; base += 32)
    {
      const int idx = base + lane;
      // owner = the last lane whose first point is at or before idx (empty jobs share their successor's start)
      int own = 0;
#pragma unroll
      for (int st = 16; st > 0; st >>= 1)
      {
        const int cand = own + st;
        const int e = __shfl_sync(0xffffffffu, excl, cand & 31);
        if (e <= idx) own = cand;
      }
      const PointRec* src = reinterpret_cast<const PointRec*>(
          __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(job_src), own));
      const int off = __shfl_sync(0xffffffffu, job_off, own);
      const int a = idx - __shfl_sync(0xffffffffu, excl, own);
      if (idx < total)
      {
        PointRec pr = src[a];
        double pw[3], xr[9], xp[3];
        pose_sel(xb, lv, 0, xr, xp);
        rot_trans(xr, xp, pr.p, pw);
        for (int k = 0; k < 3; k++) pr.p[k] = pw[k];
        M.fix_pool[off + a] = pr;
      }
    }
  }
}

// OctoTree::clear_slwd of one node (octree.cpp:739-756): the SlideWindow goes back to the pool
__device__ __forceinline__ void clear_slwd_node(const MapView& M, NodeCold& c)
{
  if (!c.has_sw) return;
  for (int s = 0; s < M.win_size; s++)
  {
    c.win_cnt[s] = 0;
    cluster_clear(c.pcrs_local[s]);
  }
  c.has_sw = 0;
}

// The rest of multi_margi in one launch, EIGHT threads per root of surf_map_slide, one per first-level child (the
// trees are at most 3 levels deep, a thread walks its child's subtree):
//  * OctoTree::margi, interior branch (octree.cpp:485-494): isexist = OR over the children, bottom-up;
//  * the erase loop (local_mapping.cpp:67-78): a root whose isexist is false leaves surf_map_slide and every node
//    below it gives its SlideWindow back (OctoTree::clear_slwd); the surviving roots go to the other slide list
//    (the caller flips slide_cur), whose counter k_margi_leaves zeroed;
//  * iter->second->jour = jour (local_mapping.cpp:36).
__global__ void __launch_bounds__(128) k_margi_finish(MapView M)
{
  vn_pdl_sync();
  const int cur = M.slide_cur;
  const int nroots = M.slide_count[cur];
  const bool early_out = nroots + M.slide_others < M.thread_num;  // multi_margi returned before its erase loop
  const int lane = threadIdx.x & 31;
  const int total = 8 * nroots;
  const int stride = gridDim.x * blockDim.x;
  for (int t0 = blockIdx.x * blockDim.x + threadIdx.x - lane; t0 < total; t0 += stride)  // warp-uniform trip count
  {
    const int t = t0 + lane;
    const bool live = t < total;
    const int a = t & 7;
    const int root = live ? M.slide_list[cur][t >> 3] : -1;
    int keep = 1;
    if (!early_out)
    {
      int n1 = -1, root_interior = 0, ex = 0;
      int c2[8];
      unsigned int int2 = 0;
#pragma unroll
      for (int b = 0; b < 8; b++) c2[b] = -1;
      if (live)
      {
        const NodeHot& h0 = M.hot[root];
        root_interior = (h0.flags & VN_FLAG_INTERIOR) ? 1 : 0;
        if (root_interior) n1 = h0.children[a];
      }
      if (n1 >= 0)
      {
        const NodeHot& h1 = M.hot[n1];
        if (h1.flags & VN_FLAG_INTERIOR)
        {
#pragma unroll
          for (int b = 0; b < 8; b++) c2[b] = h1.children[b];
          int f2[8], e2[8];
#pragma unroll
          for (int b = 0; b < 8; b++)
          {
            f2[b] = c2[b] >= 0 ? M.hot[c2[b]].flags : 0;
            e2[b] = c2[b] >= 0 ? M.cold[c2[b]].isexist : 0;
          }
          int ex1 = 0;
#pragma unroll
          for (int b = 0; b < 8; b++)
          {
            if (f2[b] & VN_FLAG_INTERIOR)
            {
              int2 |= 1u << b;
              int ex2 = 0;
              for (int d = 0; d < 8; d++)
              {
                const int n3 = M.hot[c2[b]].children[d];
                if (n3 >= 0) ex2 |= M.cold[n3].isexist;
              }
              M.cold[c2[b]].isexist = ex2;
              e2[b] = ex2;
            }
            ex1 |= e2[b];
          }
          M.cold[n1].isexist = ex1;
          ex = ex1;
        }
        else
          ex = M.cold[n1].isexist;
      }
      // the root: OR over its eight children's threads (a leaf root keeps what margi_leaf wrote)
      ex |= __shfl_xor_sync(0xffffffffu, ex, 1);
      ex |= __shfl_xor_sync(0xffffffffu, ex, 2);
      ex |= __shfl_xor_sync(0xffffffffu, ex, 4);
      if (live)
      {
        NodeCold& rc = M.cold[root];
        keep = root_interior ? ex : rc.isexist;
        if (a == 0)
        {
          rc.jour = M.jour;
          if (root_interior) rc.isexist = ex;
        }
        if (!keep)
        {
          if (a == 0) clear_slwd_node(M, rc);
          if (n1 >= 0)
          {
            clear_slwd_node(M, M.cold[n1]);
            for (int b = 0; b < 8; b++)
              if (c2[b] >= 0)
              {
                clear_slwd_node(M, M.cold[c2[b]]);
                if (int2 & (1u << b))
                  for (int d = 0; d < 8; d++)
                  {
                    const int n3 = M.hot[c2[b]].children[d];
                    if (n3 >= 0) clear_slwd_node(M, M.cold[n3]);
                  }
              }
          }
        }
      }
    }
    // compaction of the slide list: one atomic per warp
    const int mine = (live && a == 0 && keep) ? 1 : 0;
    const int pos = warp_reserve(&M.slide_count[1 - cur], mine, lane);
    if (mine) M.slide_list[1 - cur][pos] = root;
    if (live && a == 0 && !keep) M.cold[root].in_slide = 0;
  }
}

__global__ void k_zero_ints(int* p, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0;
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_export(MapView M, vina_node_record* out, long long cap, long long* count)
{
  const int nn = min(*M.node_count, M.max_nodes);
  if (blockIdx.x == 0 && threadIdx.x == 0) *count = nn;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nn && n < cap; n += gridDim.x * blockDim.x)
  {
    const NodeHot& h = M.hot[n];
    const NodeCold& c = M.cold[n];
    vina_node_record& r = out[n];
    if (h.layer < 0)
    {
      r.layer = -1;  // a record on the free stack (map pruning); vina_map_export drops these
      continue;
    }
    long long k3[3];
    unpack_key(c.rootkey, k3);
    for (int k = 0; k < 3; k++) r.key[k] = k3[k];
    r.code = h.layer | (c.path << 2);
    r.layer = h.layer;
    r.octo_state = (h.flags & VN_FLAG_INTERIOR) ? 1 : 0;
    r.isexist = c.isexist;
    r.has_sw = c.has_sw;
    r.is_plane = (h.flags & VN_FLAG_PLANE) ? 1 : 0;
    r.last_num = c.last_num;
    r.opt_state = c.opt_state >= 0 ? 1 : 0;
    r.N_add = c.pcr_add.N;
    r.N_fix = c.pcr_fix.N;
    r.n_point_fix = c.fix_count;
    r.n_win_points = 0;
    for (int i = 0; i < 16; i++) r.N_local[i] = 0;
    if (c.has_sw)
      for (int i = 0; i < M.win_size; i++)
      {
        r.N_local[i] = c.pcrs_local[M.mp[i]].N;
        r.n_win_points += c.win_cnt[M.mp[i]];
      }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
      {
        r.P_add[i + 3 * j] = c.pcr_add.P[s3(i, j)];
        r.P_fix[i + 3 * j] = c.pcr_fix.P[s3(i, j)];
      }
    for (int i = 0; i < 3; i++)
    {
      r.v_add[i] = c.pcr_add.v[i];
      r.v_fix[i] = c.pcr_fix.v[i];
      r.eig_value[i] = c.eig_value[i];
      r.center[i] = h.center[i];
      r.normal[i] = h.normal[i];
      r.voxel_center[i] = h.vcenter[i];
    }
    for (int i = 0; i < 9; i++) r.eig_vector[i] = c.eig_vector[i];
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++)
        r.plane_var[i + 6 * j] = c.plane_var[sN(6, i, j)];
    r.radius = (double)h.radius;
    for (int i = 0; i < 9; i++)
      for (int j = 0; j < 9; j++) r.cov_add[i + 9 * j] = c.cov_add[sN(9, i, j)];
    r.quater_length = (double)h.ql;
  }
}

__global__ void k_map_init(MapView M, unsigned int nslots)
{
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nslots)
  {
    M.slots[i].key = VN_EMPTY_KEY;
    M.slots[i].root = -2;
    M.slots[i].pad = 0;
  }
}

// ---------------------------------------------------------------------------
// Map pruning: the `else if (release_flag)` branch of the idle path of thd_odometry_localmapping
// (local_mapping.cpp:317-341). A root voxel whose last multi_margi is `horizon` metres of travel or more behind
// (700 in the reference) is erased from surf_map with its whole subtree (OctoTree::tras_ptr, octree.cpp:597-608).
// On the device the subtree's records are zeroed (the state of never-used pool memory, which the allocators rely
// on) and their ids pushed onto the free stack, the hash table is rebuilt from the surviving roots, and the
// fixed-point pool is compacted: every live chain is copied, in order, into one contiguous segment.
// counters: [0] roots erased [1] nodes freed [2] fixed points kept (cursor of the compacted pool).

// phase 1: mark the stale roots. Roots still in surf_map_slide are kept (they cannot be stale with the reference's
// horizon: their stamp is at most one marginalisation old; the oracle and the reference harness skip them too).
__global__ void __launch_bounds__(256) k_prune_mark(MapView M, double jour, int horizon, int* __restrict__ counters)
{
  const int nn = min(*M.node_count, M.max_nodes);
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += gridDim.x * blockDim.x)
  {
    NodeHot& h = M.hot[n];
    if (h.layer != 0 || (h.flags & VN_FLAG_DEAD)) continue;
    const NodeCold& c = M.cold[n];
    if (c.root != n || c.in_slide) continue;
    const int dis = (int)(jour - c.jour);  // int dis = jour - iter->second->jour (local_mapping.cpp:323)
    if (dis < horizon) continue;
    h.flags |= VN_FLAG_DEAD;
    atomicAdd(&counters[0], 1);
  }
}

__device__ __forceinline__ void push_free_seg(const MapView& M, int sid)
{
  const int k = atomicAdd(M.free_seg_count, 1);
  M.free_segs[k] = sid;
}

// phase 2, one thread per node: a node under a marked root gives its chain blocks and its id back and is zeroed;
// a surviving node with fixed points moves them to `tmp` (contiguous) and keeps one chain block.
__global__ void __launch_bounds__(128) k_prune_sweep(MapView M, PointRec* __restrict__ tmp, int* __restrict__ counters)
{
  const int nn = min(*M.node_count, M.max_nodes);
  constexpr int HOT_WORDS = sizeof(NodeHot) / 4, COLD_WORDS = sizeof(NodeCold) / 4;
  constexpr int FLAGS_WORD = offsetof(NodeHot, flags) / 4, LAYER_WORD = offsetof(NodeHot, layer) / 4;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += gridDim.x * blockDim.x)
  {
    NodeHot& h = M.hot[n];
    NodeCold& c = M.cold[n];
    if (h.layer < 0) continue;  // already on the free stack
    const int root = c.root;
    if (M.hot[root].flags & VN_FLAG_DEAD)
    {
      for (int sg = c.fix_head; sg >= 0;)
      {
        const int nx = M.fix_segs[sg].next;
        push_free_seg(M, sg);
        sg = nx;
      }
      // the other threads of this subtree read hot[root].flags: that word is never written with anything but
      // DEAD here
      int* hw = reinterpret_cast<int*>(&h);
      for (int w = 0; w < HOT_WORDS; w++)
        if (w != FLAGS_WORD) hw[w] = (w == LAYER_WORD) ? -1 : 0;
      if (n != root) h.flags = VN_FLAG_DEAD;
      int* cw = reinterpret_cast<int*>(&c);
      for (int w = 0; w < COLD_WORDS; w++) cw[w] = 0;
      const int k = atomicAdd(M.free_count, 1);
      M.free_nodes[k] = n;
      atomicAdd(&counters[1], 1);
      continue;
    }
    if (!tmp) continue;
    if (c.fix_head < 0 || c.fix_count <= 0)
    {
      // (an empty chain keeps no block)
      for (int sg = c.fix_head; sg >= 0;)
      {
        const int nx = M.fix_segs[sg].next;
        push_free_seg(M, sg);
        sg = nx;
      }
      c.fix_head = c.fix_tail = -1;
      continue;
    }
    const int base = atomicAdd(&counters[2], c.fix_count);
    int w = base;
    const int head = c.fix_head;
    for (int sg = head; sg >= 0;)
    {
      const FixSeg& blk = M.fix_segs[sg];
      for (int e = 0; e < blk.n; e++)
        for (int a = 0; a < blk.cnt[e]; a++) tmp[w++] = M.fix_pool[blk.off[e] + a];
      const int nx = blk.next;
      if (sg != head) push_free_seg(M, sg);
      sg = nx;
    }
    FixSeg& hb = M.fix_segs[head];
    hb.off[0] = base;
    hb.cnt[0] = w - base;
    hb.n = 1;
    hb.next = -1;
    c.fix_tail = head;
  }
}

// phase 3 (after the table has been cleared): the surviving roots go back into the hash table
__global__ void __launch_bounds__(256) k_prune_rehash(MapView M)
{
  const int nn = min(*M.node_count, M.max_nodes);
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += gridDim.x * blockDim.x)
  {
    if (M.hot[n].layer != 0 || M.cold[n].root != n) continue;
    const unsigned long long key = M.cold[n].rootkey;
    unsigned int h = hash_key(key) & M.hmask;
    for (unsigned int probe = 0; probe <= M.hmask; probe++)
    {
      if (atomicCAS(&M.slots[h].key, VN_EMPTY_KEY, key) == VN_EMPTY_KEY)
      {
        M.slots[h].root = n;
        break;
      }
      h = (h + 1) & M.hmask;
    }
  }
}

// phase 4 (after the compacted points are back in the pool): root count and the pool's cursor
__global__ void k_prune_finish(MapView M, const int* __restrict__ counters, int compacted)
{
  if (blockIdx.x == 0 && threadIdx.x == 0)
  {
    atomicSub(M.root_count, counters[0]);
    if (compacted) *M.fix_cursor = counters[2];
  }
}

void launch_map_prune_mark(cudaStream_t st, const MapView& map, double jour, int horizon, int* d_counters)
{
  cudaMemsetAsync(d_counters, 0, 4 * sizeof(int), st);
  k_prune_mark<<<592, 256, 0, st>>>(map, jour, horizon, d_counters);
}

void launch_map_prune_sweep(cudaStream_t st, const MapView& map, unsigned int nslots, PointRec* d_tmp, int* d_counters)
{
  k_prune_sweep<<<1184, 128, 0, st>>>(map, d_tmp, d_counters);
  k_map_init<<<(nslots + 255) / 256, 256, 0, st>>>(map, nslots);
  k_prune_rehash<<<592, 256, 0, st>>>(map);
}

void launch_map_prune_finish(cudaStream_t st, const MapView& map, const int* d_counters, int compacted)
{
  k_prune_finish<<<1, 32, 0, st>>>(map, d_counters, compacted);
}

// ---------------------------------------------------------------------------
static int grid_for(int n, int block) { return n <= 0 ? 1 : (n + block - 1) / block; }

int launch_map_insert_roots(cudaStream_t st, const MapView& map, const ScanView& scan, const int* n_dev, int n_host,
                            InsertScratch& sc, const PoseD& x, const double* rot_var, const double* tsl_var,
                            int pre, const IekfDev* live)
{
  Cov2 cv;
  for (int k = 0; k < 9; k++) cv.rot[k] = rot_var[k], cv.tsl[k] = tsl_var[k];
  if (n_host <= 0)
  {
    k_zero_ints<<<1, 32, 0, st>>>(sc.counters, 4);  // (nothing to insert: the callers still read zero counts)
    return 1;
  }
  // this insert's counters are the set the previous insert cleared (both sets start out zero)
  int* t = sc.counters;
  sc.counters = sc.counters_alt;
  sc.counters_alt = t;
  vn_launch(k_insert_root, dim3(grid_for(n_host, 256)), dim3(256), 0, st, map, scan, n_dev, n_host, sc, x, cv, pre, live);
  return 1;
}

// early != nullptr: the node lists of the multi_recut that follows are collected on early->side right behind
// k_insert_leaf (they only depend on the tree's structure, which is final once the leaves exist), next to the
// accumulation of this insert; launch_map_recut(..., collected = true) then only waits for early->done
int launch_map_insert_leaves(cudaStream_t st, const MapView& map, const ScanView& scan, const int* n_dev, int n_host,
                             const InsertScratch& sc, int win_ord, const EarlyCollect* early)
{
  if (n_host <= 0) return 0;
  vn_launch(k_insert_leaf, dim3(grid_for(n_host, 256)), dim3(256), 0, st, map, n_dev, n_host, sc);
  int extra = 0;
  if (early)
  {
    LayerLists& LL = *early->LL;
    int* t = LL.count;  // this multi_recut's counters are the set the previous one cleared
    LL.count = LL.count_alt;
    LL.count_alt = t;
    cudaEventRecord(early->fork, st);
    cudaStreamWaitEvent(early->side, early->fork, 0);
    vn_launch(k_recut_collect, dim3(592), dim3(128), 0, early->side, map, LL);
    cudaEventRecord(early->done, early->side);
    extra = 1;
  }
  int tg = grid_for(n_host, 128);
  if (tg > 1184) tg = 1184;
  vn_launch(k_insert_alloc, dim3(tg), dim3(128), 0, st, map, sc);
  vn_launch(k_insert_scatter, dim3(grid_for(n_host, 256)), dim3(256), 0, st, map, n_dev, n_host, sc);
  int ag = grid_for(n_host, ACC_WARPS);  // at most one warp per point's leaf
  if (ag > 148 * 16) ag = 148 * 16;
  vn_launch(k_insert_accum, dim3(ag), dim3(32 * ACC_WARPS), 0, st, map, scan, sc, win_ord);
  return 4 + extra;
}

int launch_map_insert(cudaStream_t st, const MapView& map, const ScanView& scan, const int* n_dev, int n_host,
                      InsertScratch& sc, int win_ord, const PoseD& x, const double* rot_var,
                      const double* tsl_var, const IekfDev* live, const EarlyCollect* early)
{
  if (n_host <= 0) return 0;
  int k = launch_map_insert_roots(st, map, scan, n_dev, n_host, sc, x, rot_var, tsl_var, 0, live);
  return k + launch_map_insert_leaves(st, map, scan, n_dev, n_host, sc, win_ord, early);
}

static PoseBuf make_posebuf(const PoseD* xbuf, int win_count)
{
  PoseBuf b;
  memset(&b, 0, sizeof(b));
  for (int i = 0; i < win_count && i < VINA_MAX_WIN; i++) b.x[i] = xbuf[i];
  return b;
}

int launch_map_recut(cudaStream_t st, const MapView& map, LayerLists& LL, int win_count, const PoseD* h_xbuf,
                     const IekfDev* live, const EarlyCollect* collected)
{
  const LivePose lv = { live, win_count - 1 };
  static bool attr_set[64] = { false };
  int dv = 0;
  cudaGetDevice(&dv);
  if (!attr_set[dv & 63])
  {
    cudaFuncSetAttribute(k_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPLIT_SMEM);
    attr_set[dv & 63] = true;
  }
  PoseBuf b = make_posebuf(h_xbuf, win_count);
  int launches = 1;
  if (collected)
    cudaStreamWaitEvent(st, collected->done, 0);  // (the lists were collected next to the insert's accumulation)
  else
  {
    // this call's counters are the set the previous call cleared (both sets start out zero)
    int* t = LL.count;
    LL.count = LL.count_alt;
    LL.count_alt = t;
    vn_launch(k_recut_collect, dim3(592), dim3(128), 0, st, map, LL);
    launches = 2;
  }
  vn_launch(k_recut_all, dim3(296, map.max_layer + 1), dim3(128), 0, st, map, LL);
  // every subdivision of this multi_recut, all levels, through the kernel's work queue (two blocks per SM: all
  // 296 are resident; a block that finds the queue empty but work in flight polls)
  if (map.max_layer > 0)
  {
    vn_launch(k_split, dim3(296), dim3(SPLIT_THREADS), SPLIT_SMEM, st, map, LL, win_count, b, lv);
    launches++;
  }
  return launches;
}

int launch_map_margi(cudaStream_t st, const MapView& map, const LayerLists& LL, int win_count, const PoseD* h_xbuf,
                     const IekfDev* live)
{
  const LivePose lv = { live, win_count - 1 };
  PoseBuf b = make_posebuf(h_xbuf, win_count);
  vn_launch(k_margi_leaves, dim3(592, map.max_layer + 1), dim3(128), 0, st, map, LL, win_count, b, lv, 0);
  vn_launch(k_margi_finish, dim3(592), dim3(128), 0, st, map);
  return 2;
}

// multi_recut followed by multi_margi of the same frame set (local_mapping.cpp:451, 507 with if_BA == 0), with the
// subdivisions taken off the critical path: a scan subdivides a dozen leaves, two levels deep, and that is one block's
// latency per level (~50 us on an otherwise idle GPU) - while the marginalisation of the ~10^4 leaves that are NOT
// being subdivided does not depend on it. So k_split runs on `side`, k_margi_leaves(mode 1) next to it on `st`, and a
// second, small k_margi_leaves(mode 2) picks up the children once k_split is through. Per-leaf arithmetic is
// unchanged; only the order in which independent leaves draw from the fixed-point pool differs.
int launch_map_recut_margi(cudaStream_t st, cudaStream_t side, cudaEvent_t ev_fork, cudaEvent_t ev_join, const MapView& map,
                           LayerLists& LL, int win_count, const PoseD* h_xbuf, const IekfDev* live)
{
  const LivePose lv = { live, win_count - 1 };
  static bool attr_set[64] = { false };
  int dv = 0;
  cudaGetDevice(&dv);
  if (!attr_set[dv & 63])
  {
    cudaFuncSetAttribute(k_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPLIT_SMEM);
    attr_set[dv & 63] = true;
  }
  PoseBuf b = make_posebuf(h_xbuf, win_count);
  int* t = LL.count;
  LL.count = LL.count_alt;
  LL.count_alt = t;
  vn_launch(k_recut_collect, dim3(592), dim3(128), 0, st, map, LL);
  vn_launch(k_recut_all, dim3(296, map.max_layer + 1), dim3(128), 0, st, map, LL);
  int launches = 2;
  if (map.max_layer > 0)
  {
    cudaEventRecord(ev_fork, st);
    cudaStreamWaitEvent(side, ev_fork, 0);
    vn_launch(k_split, dim3(296), dim3(SPLIT_THREADS), SPLIT_SMEM, side, map, LL, win_count, b, lv);
    cudaEventRecord(ev_join, side);
    vn_launch(k_margi_leaves, dim3(592, map.max_layer + 1), dim3(128), 0, st, map, LL, win_count, b, lv, 1);
    cudaStreamWaitEvent(st, ev_join, 0);
    vn_launch(k_margi_leaves, dim3(148, map.max_layer + 1), dim3(128), 0, st, map, LL, win_count, b, lv, 2);
    launches += 3;
  }
  else
  {
    vn_launch(k_margi_leaves, dim3(592, map.max_layer + 1), dim3(128), 0, st, map, LL, win_count, b, lv, 0);
    launches += 1;
  }
  vn_launch(k_margi_finish, dim3(592), dim3(128), 0, st, map);
  return launches + 1;
}

int launch_ba_collect(cudaStream_t st, const MapView& map, const LayerLists& LL, BaFactor* out, int* count, int cap)
{
  k_zero_ints<<<1, 32, 0, st>>>(count, 1);
  k_ba_collect<<<dim3(296, map.max_layer + 1), 128, 0, st>>>(map, LL, out, count, cap);
  return 2;
}

int launch_map_export(cudaStream_t st, const MapView& map, vina_node_record* d_out, long long cap, long long* d_count)
{
  k_export<<<592, 128, 0, st>>>(map, d_out, cap, d_count);
  return 1;
}

void launch_map_init(cudaStream_t st, const MapView& map, unsigned int nslots)
{
  k_map_init<<<(nslots + 255) / 256, 256, 0, st>>>(map, nslots);
}

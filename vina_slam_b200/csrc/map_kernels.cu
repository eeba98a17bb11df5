// Voxel-map kernels: pvec_update + cut_voxel_multi (a8-a10), multi_recut (a11),
// multi_margi (a12), export. Compiled with -fmad=false (see scan_kernels.cu).
//
// Parallel decomposition mirrors the reference's own task parallelism
// (fork-join over ROOT voxels, voxel_map.cpp:94-134, local_mapping.cpp:17-84,
// 144-201) scaled from 5 host threads to one GPU thread per root/leaf: the
// work inside one root keeps the reference's sequential order, so every
// cluster sum (P, v, N) and therefore every plane decision is reproduced
// bit-for-bit; only the order in which independent roots are visited differs.
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

__device__ __forceinline__ int ld_volatile(const int* p) { return *((const volatile int*)p); }

__device__ int alloc_node(const MapView& M)
{
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
  NodeCold& c = M.cold[id];
  c.rootkey = M.cold[parent].rootkey;
  c.path = M.cold[parent].path | (ci << (3 * ph.layer));
  c.fix_head = c.fix_tail = -1;
  for (int k = 0; k < 8; k++) c.children[k] = -1;
  return id;
}

// ---------------------------------------------------------------------------
// insert, phase 1: pvec_update (point_utils.cpp:54-65) + voxel key + root find/create
// (voxel_map.cpp:53-87).
__global__ void __launch_bounds__(256)
    k_insert_root(MapView M, ScanView scan, const int* __restrict__ n_ptr, int n_host, InsertScratch sc, PoseD x,
                  Cov2 cv)
{
  int n = n_ptr ? *n_ptr : n_host;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double pnt[3] = { scan.p[0][i], scan.p[1][i], scan.p[2][i] };
  double var6[6];
  for (int k = 0; k < 6; k++) var6[k] = scan.v[k][i];
  double pw[3], vw[6];
  rot_trans(x.R, x.p, pnt, pw);
  world_var(x.R, pnt, var6, cv.rot, cv.tsl, vw);
  for (int k = 0; k < 3; k++) sc.pw[k][i] = pw[k];
  for (int k = 0; k < 6; k++) sc.vw[k][i] = vw[k];

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
        NodeCold& nc = M.cold[id];
        nc.rootkey = key;
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
  int n = n_ptr ? *n_ptr : n_host;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int leaf = sc.leaf_of[i];
  if (leaf < 0) return;
  sc.idx[M.cold[leaf].pend_off + sc.rank_of[i]] = i;
}

// insert, phase 3: OctoTree::push for every point of the leaf in ascending point
// order (octree.cpp:151-177; order = voxel_map.cpp:86 push_back(i))
__global__ void __launch_bounds__(64) k_insert_accum(MapView M, ScanView scan, InsertScratch sc, int win_ord)
{
  int nt = sc.counters[1];
  const int mord = M.mp[win_ord];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nt; j += gridDim.x * blockDim.x)
  {
    int leaf = sc.touched[j];
    NodeCold& c = M.cold[leaf];
    const int cnt = c.pend_cnt;
    int* idx = sc.idx + c.pend_off;
    for (int a = 1; a < cnt; a++)  // arrival order is nearly sorted already
    {
      int v = idx[a], b = a - 1;
      while (b >= 0 && idx[b] > v)
      {
        idx[b + 1] = idx[b];
        b--;
      }
      idx[b + 1] = v;
    }
    const bool store = M.hot[leaf].layer < M.max_layer;
    int woff = 0;
    const int old_cnt = c.win_cnt[mord];
    if (store)
    {
      woff = atomicAdd(&M.win_cursor[mord], cnt + old_cnt);
      if ((long long)woff + cnt + old_cnt > M.win_cap)
      {
        atomicOr(M.status, VN_ST_WIN_FULL);
        c.pend_cnt = 0;
        continue;
      }
      PointRec* pool = M.win_pool[mord];
      for (int a = 0; a < old_cnt; a++) pool[woff + a] = pool[c.win_off[mord] + a];
    }
    c.has_sw = 1;   // sw acquired (octree.cpp:154-164); recycled windows are empty
    c.isexist = 1;  // octree.cpp:165-166
    Cluster add = c.pcr_add, loc = c.pcrs_local[mord];
    double cov[45];
    for (int k = 0; k < 45; k++) cov[k] = c.cov_add[k];
    for (int a = 0; a < cnt; a++)
    {
      int i = idx[a];
      PointRec pr;
      for (int k = 0; k < 3; k++) pr.p[k] = scan.p[k][i];
      for (int k = 0; k < 6; k++) pr.v[k] = sc.vw[k][i];
      double pw[3] = { sc.pw[0][i], sc.pw[1][i], sc.pw[2][i] };
      if (store) M.win_pool[mord][woff + old_cnt + a] = pr;
      cluster_push(loc, pr.p);
      cluster_push(add, pw);
      bf_var_add(cov, pr.v, pw);
    }
    c.pcr_add = add;
    c.pcrs_local[mord] = loc;
    for (int k = 0; k < 45; k++) c.cov_add[k] = cov[k];
    if (store)
    {
      c.win_off[mord] = woff;
      c.win_cnt[mord] = cnt + old_cnt;
    }
    c.pend_cnt = 0;
  }
}

// ---------------------------------------------------------------------------
// push_fix into a child (octree.cpp:179-188)
__device__ __forceinline__ void child_push_fix(NodeCold& k, const PointRec& pr)
{
  cluster_push(k.pcr_fix, pr.p);
  cluster_push(k.pcr_add, pr.p);
  bf_var_add(k.cov_add, pr.v, pr.p);
}

// append one segment of `cnt` fixed points to a node's point_fix chain; returns pool offset or -1
__device__ int fix_append(const MapView& M, NodeCold& c, int cnt)
{
  int off = atomicAdd(M.fix_cursor, cnt);
  int sid = atomicAdd(M.fixseg_cursor, 1);
  if ((long long)off + cnt > M.fix_cap || sid >= M.fixseg_cap)
  {
    atomicOr(M.status, VN_ST_FIX_FULL);
    return -1;
  }
  FixSeg& s = M.fix_segs[sid];
  s.off = off;
  s.cnt = cnt;
  s.next = -1;
  if (c.fix_tail >= 0)
    M.fix_segs[c.fix_tail].next = sid;
  else
    c.fix_head = sid;
  c.fix_tail = sid;
  c.fix_count += cnt;
  return off;
}

// the subdivision branch of OctoTree::recut (octree.cpp:375-387): fix_divide (:257-277),
// subdivide per window frame (:279-300), release of the parent's SlideWindow
__device__ void split_leaf(const MapView& M, int n, int win_count, const PoseBuf& xb)
{
  NodeCold& c = M.cold[n];
  NodeHot& h = M.hot[n];
  const int child_layer = h.layer + 1;
  const bool store = child_layer < M.max_layer;
  int cnt8[8], off8[8], fill8[8];

  if (c.pcr_fix.N != 0)
  {
    for (int k = 0; k < 8; k++) cnt8[k] = 0, fill8[k] = 0, off8[k] = -1;
    for (int s = c.fix_head; s >= 0; s = M.fix_segs[s].next)
    {
      const FixSeg seg = M.fix_segs[s];
      for (int a = 0; a < seg.cnt; a++) cnt8[child_index(M.fix_pool[seg.off + a].p, h.vcenter)]++;
    }
    for (int k = 0; k < 8; k++)
      if (cnt8[k] > 0)
      {
        if (c.children[k] < 0) c.children[k] = make_child(M, n, k);
        if (c.children[k] >= 0 && store) off8[k] = fix_append(M, M.cold[c.children[k]], cnt8[k]);
      }
    for (int s = c.fix_head; s >= 0; s = M.fix_segs[s].next)
    {
      const FixSeg seg = M.fix_segs[s];
      for (int a = 0; a < seg.cnt; a++)
      {
        PointRec pr = M.fix_pool[seg.off + a];
        int k = child_index(pr.p, h.vcenter);
        int kid = c.children[k];
        if (kid < 0) continue;
        if (store && off8[k] >= 0) M.fix_pool[off8[k] + fill8[k]++] = pr;
        child_push_fix(M.cold[kid], pr);
      }
    }
    c.fix_head = c.fix_tail = -1;  // PVec().swap(point_fix)
    c.fix_count = 0;
  }

  for (int si = 0; si < win_count; si++)
  {
    const int slot = M.mp[si];
    const int np = c.win_cnt[slot];
    if (np == 0) continue;
    const PointRec* src = M.win_pool[slot] + c.win_off[slot];
    const PoseD& x = xb.x[si];
    for (int k = 0; k < 8; k++) cnt8[k] = 0, fill8[k] = 0, off8[k] = -1;
    for (int a = 0; a < np; a++)
    {
      double pw[3];
      rot_trans(x.R, x.p, src[a].p, pw);
      cnt8[child_index(pw, h.vcenter)]++;
    }
    for (int k = 0; k < 8; k++)
      if (cnt8[k] > 0)
      {
        if (c.children[k] < 0) c.children[k] = make_child(M, n, k);
        int kid = c.children[k];
        if (kid < 0) continue;
        NodeCold& kc = M.cold[kid];
        kc.has_sw = 1;
        kc.isexist = 1;
        if (store)
        {
          int woff = atomicAdd(&M.win_cursor[slot], cnt8[k]);
          if ((long long)woff + cnt8[k] > M.win_cap)
          {
            atomicOr(M.status, VN_ST_WIN_FULL);
            continue;
          }
          off8[k] = woff;
          kc.win_off[slot] = woff;
          kc.win_cnt[slot] = cnt8[k];
        }
      }
    for (int a = 0; a < np; a++)
    {
      PointRec pr = src[a];
      double pw[3];
      rot_trans(x.R, x.p, pr.p, pw);
      int k = child_index(pw, h.vcenter);
      int kid = c.children[k];
      if (kid < 0) continue;
      NodeCold& kc = M.cold[kid];
      if (store && off8[k] >= 0) M.win_pool[slot][off8[k] + fill8[k]++] = pr;
      cluster_push(kc.pcrs_local[slot], pr.p);
      cluster_push(kc.pcr_add, pw);
      bf_var_add(kc.cov_add, pr.v, pw);
    }
  }
  // sw->clear(); sws.push_back(sw); sw = nullptr; octo_state = 1 (octree.cpp:384-387)
  for (int s = 0; s < M.win_size; s++)
  {
    c.win_cnt[s] = 0;
    cluster_clear(c.pcrs_local[s]);
  }
  c.has_sw = 0;
  h.flags |= VN_FLAG_INTERIOR;
}

// OctoTree::recut over one root (octree.cpp:335-393) followed by tras_opt's
// BA-factor marking (octree.cpp:498-521, local_mapping.cpp:196-200)
__global__ void __launch_bounds__(64) k_recut(MapView M, int cur, int win_count, PoseBuf xb)
{
  const int nroots = M.slide_count[cur];
  if (nroots < M.thread_num) return;  // local_mapping.cpp:150-154
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nroots; j += gridDim.x * blockDim.x)
  {
    int stack[40];
    int sp = 0;
    const int root = M.slide_list[cur][j];
    stack[sp++] = root;
    while (sp > 0)
    {
      int n = stack[--sp];
      NodeHot& h = M.hot[n];
      NodeCold& c = M.cold[n];
      if (!(h.flags & VN_FLAG_INTERIOR))
      {
        c.opt_state = -1;
        if ((double)c.pcr_add.N <= M.min_point[h.layer])
        {
          h.flags &= ~VN_FLAG_PLANE;
          continue;
        }
        if (!c.isexist || !c.has_sw) continue;
        double L[6], ev[3], Q[9];
        cluster_cov(c.pcr_add, L);
        eig3_sym(L, ev, Q);
        for (int k = 0; k < 3; k++) c.eig_value[k] = ev[k];
        for (int k = 0; k < 9; k++) c.eig_vector[k] = Q[k];
        bool is_plane = (ev[0] < M.min_eigen_value) && ((ev[0] / ev[2]) < M.thre[h.layer]);  // octree.cpp:198-201
        if (is_plane)
          h.flags |= VN_FLAG_PLANE;
        else
          h.flags &= ~VN_FLAG_PLANE;
        if (is_plane || h.layer >= M.max_layer) continue;
        split_leaf(M, n, win_count, xb);
      }
      for (int k = 7; k >= 0; k--)
        if (c.children[k] >= 0) stack[sp++] = c.children[k];
    }
    // tras_opt: which leaves are BA factors
    sp = 0;
    stack[sp++] = root;
    while (sp > 0)
    {
      int n = stack[--sp];
      NodeHot& h = M.hot[n];
      NodeCold& c = M.cold[n];
      if (!(h.flags & VN_FLAG_INTERIOR))
      {
        if (c.isexist && (h.flags & VN_FLAG_PLANE) && c.has_sw)
          if (!(c.eig_value[0] / c.eig_value[1] > 0.12)) c.opt_state = 1;
      }
      else
        for (int k = 7; k >= 0; k--)
          if (c.children[k] >= 0) stack[sp++] = c.children[k];
    }
  }
}

// ---------------------------------------------------------------------------
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
      h.pvar[sN(6, a, b)] = s;
    }
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) h.pvar[sN(6, a, 3 + b)] = nv * Jc[a][6 + b];
  for (int a = 0; a < 3; a++)
    for (int b = a; b < 3; b++) h.pvar[sN(6, 3 + a, 3 + b)] = (nv * nv) * c.cov_add[sN(9, 6 + a, 6 + b)];
  for (int a = 0; a < 3; a++)
  {
    h.center[a] = center[a];
    h.normal[a] = u[0][a];
  }
  h.radius = (float)c.eig_value[2];
}

// leaf branch of OctoTree::margi (octree.cpp:397-484), mgsize = 1
__device__ void margi_leaf(const MapView& M, int n, int win_count, const PoseBuf& xb)
{
  NodeHot& h = M.hot[n];
  NodeCold& c = M.cold[n];
  if (!c.isexist || !c.has_sw) return;
  const int s0 = M.mp[0];
  Cluster world0;
  cluster_clear(world0);
  const bool is_plane = (h.flags & VN_FLAG_PLANE) != 0;
  if (c.opt_state >= 0)
  {
    c.opt_state = -1;
    if (c.pcrs_local[s0].N != 0) cluster_transform(world0, c.pcrs_local[s0], xb.x[0].R, xb.x[0].p);
  }
  else
  {
    Cluster add = c.pcr_fix;
    for (int i = 0; i < win_count; i++)
    {
      const Cluster& loc = c.pcrs_local[M.mp[i]];
      if (loc.N != 0)
      {
        Cluster w;
        cluster_transform(w, loc, xb.x[i].R, xb.x[i].p);
        if (i == 0) world0 = w;
        cluster_add(add, w);
      }
    }
    c.pcr_add = add;
    if (is_plane)
    {
      double L[6], ev[3], Q[9];
      cluster_cov(c.pcr_add, L);
      eig3_sym(L, ev, Q);
      for (int k = 0; k < 3; k++) c.eig_value[k] = ev[k];
      for (int k = 0; k < 9; k++) c.eig_vector[k] = Q[k];
    }
  }

  if (c.pcr_fix.N < M.max_points && is_plane)
    if (c.pcr_add.N - c.last_num >= 5 || c.last_num <= 10)
    {
      plane_update(h, c);
      c.last_num = c.pcr_add.N;
    }

  if (c.pcr_fix.N < M.max_points)
  {
    if (world0.N != 0)
    {
      cluster_add(c.pcr_fix, world0);
      const int np = c.win_cnt[s0];
      if (np > 0)
      {
        int off = fix_append(M, c, np);
        if (off >= 0)
        {
          const PointRec* src = M.win_pool[s0] + c.win_off[s0];
          for (int a = 0; a < np; a++)
          {
            PointRec pr = src[a];
            double pw[3];
            rot_trans(xb.x[0].R, xb.x[0].p, pr.p, pw);
            for (int k = 0; k < 3; k++) pr.p[k] = pw[k];
            M.fix_pool[off + a] = pr;
          }
        }
      }
    }
  }
  else
  {
    if (world0.N != 0) cluster_sub(c.pcr_add, world0);
    c.fix_head = c.fix_tail = -1;
    c.fix_count = 0;
  }

  if (c.pcrs_local[s0].N != 0)
  {
    cluster_clear(c.pcrs_local[s0]);
    c.win_cnt[s0] = 0;
  }
  c.isexist = (c.pcr_fix.N >= c.pcr_add.N) ? 0 : 1;
}

__global__ void __launch_bounds__(64) k_margi(MapView M, int cur, int win_count, PoseBuf xb)
{
  const int nroots = M.slide_count[cur];
  if (nroots < M.thread_num) return;  // local_mapping.cpp:26-28
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nroots; j += gridDim.x * blockDim.x)
  {
    int stack[40];
    int sp = 0;
    stack[sp++] = M.slide_list[cur][j] << 1;
    while (sp > 0)
    {
      int e = stack[sp - 1];
      int n = e >> 1;
      NodeHot& h = M.hot[n];
      NodeCold& c = M.cold[n];
      if (!(h.flags & VN_FLAG_INTERIOR))
      {
        margi_leaf(M, n, win_count, xb);
        sp--;
      }
      else if (!(e & 1))
      {
        stack[sp - 1] = e | 1;
        for (int k = 7; k >= 0; k--)
          if (c.children[k] >= 0) stack[sp++] = c.children[k] << 1;
      }
      else
      {
        int ex = 0;
        for (int k = 0; k < 8; k++)
          if (c.children[k] >= 0) ex |= M.cold[c.children[k]].isexist;
        c.isexist = ex;
        sp--;
      }
    }
  }
}

// erase loop of multi_margi (local_mapping.cpp:67-78): roots without live window data
// leave surf_map_slide and give their SlideWindows back (OctoTree::clear_slwd, octree.cpp:739-756)
__global__ void __launch_bounds__(64) k_slide_compact(MapView M, int cur)
{
  const int nroots = M.slide_count[cur];
  if (nroots < M.thread_num)
  {
    // early-out of multi_margi: the slide map is left as is -> copy the list over
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nroots; j += gridDim.x * blockDim.x)
    {
      int pos = atomicAdd(&M.slide_count[1 - cur], 1);
      M.slide_list[1 - cur][pos] = M.slide_list[cur][j];
    }
    return;
  }
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nroots; j += gridDim.x * blockDim.x)
  {
    const int root = M.slide_list[cur][j];
    if (M.cold[root].isexist)
    {
      int pos = atomicAdd(&M.slide_count[1 - cur], 1);
      M.slide_list[1 - cur][pos] = root;
      continue;
    }
    M.cold[root].in_slide = 0;
    int stack[40];
    int sp = 0;
    stack[sp++] = root;
    while (sp > 0)
    {
      int n = stack[--sp];
      NodeCold& c = M.cold[n];
      if (M.hot[n].flags & VN_FLAG_INTERIOR)
        for (int k = 0; k < 8; k++)
          if (c.children[k] >= 0) stack[sp++] = c.children[k];
      if (c.has_sw)
      {
        for (int s = 0; s < M.win_size; s++)
        {
          c.win_cnt[s] = 0;
          cluster_clear(c.pcrs_local[s]);
        }
        c.has_sw = 0;
      }
    }
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
      for (int j = 0; j < 6; j++) r.plane_var[i + 6 * j] = h.pvar[sN(6, i, j)];
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
static int grid_for(int n, int block) { return n <= 0 ? 1 : (n + block - 1) / block; }

int launch_map_insert(cudaStream_t st, const MapView& map, const ScanView& scan, const int* n_dev, int n_host,
                      const InsertScratch& sc, int win_ord, const PoseD& x, const double* rot_var,
                      const double* tsl_var)
{
  Cov2 cv;
  for (int k = 0; k < 9; k++) cv.rot[k] = rot_var[k], cv.tsl[k] = tsl_var[k];
  if (n_host <= 0) return 0;
  k_zero_ints<<<1, 32, 0, st>>>(sc.counters, 3);
  k_insert_root<<<grid_for(n_host, 256), 256, 0, st>>>(map, scan, n_dev, n_host, sc, x, cv);
  k_insert_leaf<<<grid_for(n_host, 256), 256, 0, st>>>(map, n_dev, n_host, sc);
  int tg = grid_for(n_host, 128);
  if (tg > 1184) tg = 1184;
  k_insert_alloc<<<tg, 128, 0, st>>>(map, sc);
  k_insert_scatter<<<grid_for(n_host, 256), 256, 0, st>>>(map, n_dev, n_host, sc);
  k_insert_accum<<<1184, 64, 0, st>>>(map, scan, sc, win_ord);
  return 6;
}

static PoseBuf make_posebuf(const PoseD* xbuf, int win_count)
{
  PoseBuf b;
  memset(&b, 0, sizeof(b));
  for (int i = 0; i < win_count && i < VINA_MAX_WIN; i++) b.x[i] = xbuf[i];
  return b;
}

int launch_map_recut(cudaStream_t st, const MapView& map, int win_count, const PoseD* h_xbuf)
{
  PoseBuf b = make_posebuf(h_xbuf, win_count);
  k_recut<<<1184, 64, 0, st>>>(map, map.slide_cur, win_count, b);
  return 1;
}

int launch_map_margi(cudaStream_t st, const MapView& map, int win_count, const PoseD* h_xbuf)
{
  const int cur_list = map.slide_cur;
  PoseBuf b = make_posebuf(h_xbuf, win_count);
  k_margi<<<1184, 64, 0, st>>>(map, cur_list, win_count, b);
  k_zero_ints<<<1, 32, 0, st>>>(map.slide_count + (1 - cur_list), 1);
  k_slide_compact<<<1184, 64, 0, st>>>(map, cur_list);
  return 3;
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

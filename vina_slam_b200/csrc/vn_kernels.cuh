// Launcher prototypes and small parameter blocks shared by the kernel TUs and the context.
#pragma once
#include <cstdlib>
#include <cstddef>
#include <cuda_runtime.h>
#include "vn_math.cuh"
#include "vn_types.cuh"

struct VarInitParams
{
  float range_var;  // (float)dept_err * (float)dept_err   (point_utils.cpp:11)
  double dir_var;   // sin(deg2rad((float)beam_err))^2      (point_utils.cpp:14)
  double ext_R[9], ext_t[3];
};

struct DownSlot
{
  unsigned long long key;
  double sum[3];
  int cnt;
  int first;
};


// Programmatic dependent launch (sm_90+): the kernels of the per-scan step are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and begin with vn_pdl_sync(): "my dependents may be scheduled"
// right away, then "wait until everything before me in the stream has completed and is visible". The next kernel's
// launch processing and block scheduling then overlap the tail of the running one instead of following its
// completion - the step is a chain of ~20 short dependent kernels, and the gaps between them are a tenth of it.
// Nothing may touch global memory before vn_pdl_sync(). Launched without the attribute both instructions do nothing.
#ifdef __CUDACC__
__device__ __forceinline__ void vn_pdl_sync()
{
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif
inline bool vn_pdl_enabled()
{
  static const bool on = [] {
    const char* e = getenv("VINA_PDL");
    return !(e && atoi(e) == 0);
  }();
  return on;
}
template <typename... KA, typename... A>
inline cudaError_t vn_launch(void (*k)(KA...), dim3 g, dim3 b, size_t smem, cudaStream_t st, A&&... a)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = g;
  cfg.blockDim = b;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = vn_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k, static_cast<KA>(a)...);
}

#define VN_IEKF_NACC 34  // 21 (HTH upper) + 6 (HTz) + 6 (nnt upper) + 1 (count)
#define VN_MAX_BATCH 16  // sequences per batched k_iekf launch

struct IekfDebug
{
  long long* keys;  // n x 3
  int* codes;
  unsigned char* flags;
  double* sigma;
};

// Device-resident iterate of VINA_SLAM::LioStateEstimation (odometry.cpp:64-255): the state the IEKF loop
// reads and updates without going back to the host. k_iekf takes R/p and the prior covariance blocks from here;
// with VN_IEKF_SOLVE its last block also performs the a7 update (15x15 solve, boxplus, convergence logic).
struct IekfDev
{
  double R[9], p[3], v[3], bg[3], ba[3];       // x_curr (column-major R)
  double Rp[9], pp[3], vp[3], bgp[3], bap[3];  // x_prop
  double cov[225];                             // prior covariance; posterior once `done`
  double rot_var[9], tsl_var[9];               // prior blocks, fixed over the iterations (odometry.cpp:105-106)
  double sums[40];                             // packed sums of the last iteration (nnt = [27..32], count = [33])
  int iter, rematch, done, max_iter;
};

#define VN_IEKF_PUBLISH 1  // write the 34 sums + sequence number to mapped host memory (low-level ABI, tests)
#define VN_IEKF_SOLVE 2    // last block runs the IEKF update on the device state
#define VN_IEKF_GATED 4    // return at once when the device iterate says the loop has finished (sharded loop)
#define VN_IEKF_HANDOVER 16  // with VN_IEKF_SOLVE: publish the iterate to the host from the launch that ends the loop
#define VN_IEKF_NOCACHE 8  // no per-point leaf cache: every point looks its voxel up (queries that crossed GPUs)

// one sequence of a (batched) k_iekf launch
struct IekfSeq
{
  const double* pv_base;  // pointVar SoA: 9 contiguous arrays p[3], v[6] of pv_stride doubles each
  long long pv_stride;
  const int* n_ptr;  // device-side point count (or null -> n_host)
  int n_host;
  unsigned int hmask;
  int* cache;
  const HashSlot* slots;
  const NodeHot* hot;
  const NodeCold* cold;
  IekfDev* dev;
  double* partials;      // [34][gridDim.x]
  unsigned int* ticket;
  double* result;        // mapped pinned host memory (VN_IEKF_PUBLISH)
  double voxel_size;
  unsigned long long seq;  // launch sequence number, written after the sums so that the host can poll for them
  // VN_IEKF_HANDOVER: the block that finishes the loop copies the iterate to mapped host memory (pub) and then
  // writes pub_seq to pub_flag - the host has the pose without a separate publishing kernel
  IekfDev* pub;
  unsigned long long* pub_flag;
  unsigned long long pub_seq;
  IekfDebug dbg;
};

struct IekfBatch
{
  int mode;     // VN_IEKF_*
  int variant;  // experiment switches (0 = product path), see vina_iekf_time_kernel
  IekfSeq s[VN_MAX_BATCH];
};

// the whole iteration loop as one persistent cooperative launch (k_iekf_loop)
struct IekfLoop
{
  IekfSeq q;
  unsigned long long* bar;  // [0] arrivals at the grid barrier, [1] blocks that have left; both zero between launches
  double* partials;         // [2][VN_IEKF_NACC][gridDim.x], ping-pong over the iterations
  int* status;
  int chunk;                // points per block, a multiple of 32
  int mode;                 // VN_IEKF_HANDOVER
};

// scan_kernels.cu
void launch_deskew(cudaStream_t st, float4* pts, int n, const DeskewPoses* d_poses, int* status);
void launch_var_init(cudaStream_t st, const float4* pts, const int* n_dev, int n_host, ScanView out,
                     const VarInitParams& prm);
// deskew + var_init of the full scan + leaf-cache reset in one pass
void launch_deskew_var_init(cudaStream_t st, float4* pts, int n, const DeskewPoses* d_poses, int* status, ScanView out,
                            const VarInitParams& prm, int* cache, int first, int last);
// the fused front of the per-scan step: deskew + var_init + cache reset + the accumulation pass of the down-sampling,
// then ONE cooperative launch for the rest of the down-sampling and the var_init of the emitted set
struct DownEmit
{
  int n;
  int chunk;  // points per block (set by the launcher)
  DownSlot* tab;
  const int* slot_of;
  float4* out;
  int* n_out_dev;
  ScanView pv;
  VarInitParams prm;
  int* counts;                       // [gridDim.x]
  unsigned long long* bar;           // [0] arrivals, [1] exits; zero between launches
  volatile unsigned long long* pub;  // mapped host memory: [0] sequence number (written last), [1] count
  unsigned long long seq;
  int* status;
};
void launch_deskew_var_init_down(cudaStream_t st, float4* pts, int n, const DeskewPoses* d_poses, int* status, ScanView out,
                                 const VarInitParams& prm, int* cache, double voxel_size, DownSlot* tab, unsigned int mask,
                                 int* slot_of);
int launch_down_emit_all(cudaStream_t st, DownEmit& a, int sm_count);
void launch_down_init(cudaStream_t st, DownSlot* tab, unsigned int nslots);
int launch_downsample(cudaStream_t st, const float4* pts, int n, double voxel_size, DownSlot* tab, unsigned int mask,
                      int* slot_of, int* flag, int* scan, int* block_sums, int* n_out_dev, float4* out, int* status,
                      unsigned long long* pub = nullptr, unsigned long long seq = 0, const ScanView* pv = nullptr,
                      const VarInitParams* prm = nullptr);

// start-up phase (scan_kernels.cu): kd-tree IEKF association / sums, local-map append, motion_init's re-deskew
struct InsertScratch;
int launch_init_assoc(cudaStream_t st, const ScanView& pv, int n, const PoseD& x, const float4* tree, int n_tree, int refind,
                      double* ds, double* dir, double* partial, double* out28);
int launch_init_tree_append(cudaStream_t st, const ScanView& pv, int n, const PoseD& x, float4* tree_tail);
int launch_init_redeskew(cudaStream_t st, const float4* orig, int n, int n_skip, const DeskewPoses* d_poses, const PoseD& x,
                         const double* rot_var, const double* tsl_var, int converged, const VarInitParams& prm,
                         const ScanView& out, const InsertScratch& sc);

// iekf_kernel.cu
int iekf_grid_blocks(int n, int sm_count);
// grid = (blocks, nseq); every sequence gets `blocks` persistent 1024-thread blocks
int launch_iekf(cudaStream_t st, const IekfBatch& bt, int nseq, int blocks, bool debug, bool pdl = false);
int iekf_loop_chunk(int n, int blocks);
int launch_iekf_loop(cudaStream_t st, const IekfLoop& a, int blocks);
void launch_fill_int(cudaStream_t st, int* p, int v, int n);
static_assert(sizeof(IekfDev) % sizeof(double) == 0, "IekfDev is copied as doubles");
static_assert(offsetof(IekfDev, Rp) == 21 * sizeof(double) && offsetof(IekfDev, cov) == 42 * sizeof(double),
              "k_iekf stages x_curr / x_prop as the first 42 doubles of IekfDev");
void launch_publish_iterate(cudaStream_t st, const IekfDev* src, IekfDev* dst_mapped, unsigned long long* flag_mapped,
                            unsigned long long seq);

// front_kernels.cu: decoder keep rule + pcl_handler (filter, stable radix sort by time offset, cut at 0.11 s).
// counters: [0] kept by the decoder rule, [1] of those within 0.11 s (= points written to `out`, time-sorted)
int launch_front_prepare_buckets(cudaStream_t st, const float4* raw, int n, int point_filter_num, double blind, int* bkt_of,
                                 int* work, unsigned long long* pairs, float4* out, unsigned long long* pub_mapped,
                                 unsigned long long seq);
int launch_front_prepare(cudaStream_t st, const float4* raw, int n, int point_filter_num, double blind2,
                         unsigned int* key[2], int* idx[2], int* hist, int* counters, float4* out, float* t_last);

// map_kernels.cu
struct LayerLists;
struct InsertScratch
{
  double* pw[3];   // world points of the down-sampled scan
  double* vw[6];   // world covariance (symmetric)
  int* root_of;    // per point
  int* leaf_of;    // per point
  int* rank_of;    // per point: arrival rank within its leaf
  int* touched;    // leaves touched by this scan
  int* counters;   // [0]=g_size (distinct roots) [1]=n_touched [2]=idx cursor
  int* counters_alt;  // the set the next insert uses (cleared by this one)
  int* idx;        // point indices grouped by leaf
  int stamp;       // scan stamp for distinct-root counting
};
struct EarlyCollect  // k_recut_collect of the following multi_recut on `side`, next to the insert's accumulation
{
  LayerLists* LL;
  cudaStream_t side;
  cudaEvent_t fork, done;
};
int launch_map_insert(cudaStream_t st, const MapView& map, const ScanView& scan, const int* n_dev, int n_host,
                      InsertScratch& sc, int win_ord, const PoseD& x, const double* rot_var,
                      const double* tsl_var, const IekfDev* live = nullptr, const EarlyCollect* early = nullptr);
// the two halves of an insert, for the sharded map: (1) key + root find/create on points whose world
// position / covariance are already in sc.pw / sc.vw (pre != 0) or come from pvec_update; (2) the rest.
// Between them the caller may overwrite sc.counters[0] (distinct roots) with the all-reduced count.
int launch_map_insert_roots(cudaStream_t st, const MapView& map, const ScanView& scan, const int* n_dev, int n_host,
                            InsertScratch& sc, const PoseD& x, const double* rot_var, const double* tsl_var,
                            int pre, const IekfDev* live = nullptr);
int launch_map_insert_leaves(cudaStream_t st, const MapView& map, const ScanView& scan, const int* n_dev, int n_host,
                             const InsertScratch& sc, int win_ord, const EarlyCollect* early = nullptr);
// shard_kernels.cu
int launch_shard_route(cudaStream_t st, const ScanView& scan, int first, int count, const PoseD& x,
                       const double* rot_var, const double* tsl_var, double voxel_size, int world,
                       unsigned char* owner, int* hist, int* counts, int* starts, double* out, long long gidx_base,
                       int* status, bool query);
int launch_shard_unpack_query(cudaStream_t st, const double* rec, int n, const ScanView& scan);
int launch_shard_route_p2p(cudaStream_t st, const ScanView& scan, int first, int count, const PoseD& x,
                           const double* rot_var, const double* tsl_var, double voxel_size, const ShardPeers& peers,
                           unsigned char* owner, int* hist, int* counts, int* starts, long long* base,
                           unsigned long long epoch, long long gidx_base, long long inbox_cap, int* status, int phase);
int launch_shard_query_p2p(cudaStream_t st, const ScanView& scan, int first, int count, const IekfDev* it,
                           double voxel_size, const ShardPeers& peers, unsigned char* owner, int* hist, int* counts,
                           int* starts, long long* base, unsigned long long epoch, long long inbox_cap, int* n_recv,
                           const ScanView& recv_set, int* status, int phase);
// the 34 sums of every shard -> every rank, summed in rank order, then the IEKF update (iekf_kernel.cu)
void launch_p2p_sums_publish(cudaStream_t st, const ShardPeers& peers, IekfDev* dev, unsigned long long epoch);
void launch_p2p_sums_solve(cudaStream_t st, const ShardPeers& peers, IekfDev* dev, unsigned long long epoch, int* status);
int launch_shard_recv_p2p(cudaStream_t st, const ShardPeers& peers, unsigned long long epoch, int* n_recv, int cap,
                          const ScanView& scan, const InsertScratch& sc, int* status);
int launch_shard_unpack(cudaStream_t st, const double* rec, int n, const ScanView& scan, const InsertScratch& sc);
// live != nullptr: the pose of frame win_count - 1 (and, for the insert, the posterior covariance blocks) are read
// from the device iterate instead of the host arguments (the IEKF result need not have reached the host yet)
int launch_map_recut(cudaStream_t st, const MapView& map, LayerLists& LL, int win_count, const PoseD* h_xbuf,
                     const IekfDev* live = nullptr, const EarlyCollect* collected = nullptr);
// margi + erase loop; the surviving roots land in slide_list[1 - map.slide_cur] (caller flips slide_cur)
int launch_map_recut_margi(cudaStream_t st, cudaStream_t side, cudaEvent_t ev_fork, cudaEvent_t ev_join, const MapView& map,
                           LayerLists& LL, int win_count, const PoseD* h_xbuf, const IekfDev* live);
int launch_map_margi(cudaStream_t st, const MapView& map, const LayerLists& LL, int win_count, const PoseD* h_xbuf,
                     const IekfDev* live = nullptr);
// BA LiDAR factor (map_kernels.cu: collect; ba_kernels.cu: Hessian / residual). d_out = Hess (6 win)^2, JacT, residual
struct BaDone  // completion signal of a BA evaluation (ba_kernels.cu)
{
  unsigned int* ticket;        // device counter, zero between launches
  unsigned long long* flag;    // mapped pinned host memory
  unsigned long long seq;
};
int launch_ba_collect(cudaStream_t st, const MapView& map, const LayerLists& LL, BaFactor* out, int* count, int cap);
int ba_hess_warps(int sm_count);
size_t ba_partial_doubles(int sm_count);
int launch_ba_hess(cudaStream_t st, const BaFactor* fac, const int* n_dev, const PoseD* h_xs, int win, int sm_count,
                   double* partial, double* d_out, const BaDone& done);
int launch_ba_writeback(cudaStream_t st, const MapView& map, const BaFactor* fac, const int* n_dev, int sm_count);
int launch_ba_residual(cudaStream_t st, BaFactor* fac, const int* n_dev, const PoseD* h_xs, int win, int sm_count,
                       double* partial, double* lam0, const BaDone& done);
int launch_map_export(cudaStream_t st, const MapView& map, vina_node_record* d_out, long long cap, long long* d_count);
// map pruning (local_mapping.cpp:317-341): mark stale roots -> [sync, read counters] -> sweep + hash rebuild ->
// [copy the compacted fixed points back] -> finish. d_counters: 4 ints (roots erased, nodes freed, points kept).
void launch_map_prune_mark(cudaStream_t st, const MapView& map, double jour, int horizon, int* d_counters);
void launch_map_prune_sweep(cudaStream_t st, const MapView& map, unsigned int nslots, PointRec* d_tmp, int* d_counters);
void launch_map_prune_finish(cudaStream_t st, const MapView& map, const int* d_counters, int compacted);
void launch_map_init(cudaStream_t st, const MapView& map, unsigned int nslots);

// per-layer node lists of the roots in surf_map_slide, rebuilt by every multi_recut and reused by the
// multi_margi of the same scan (layer 0 is the slide list itself)
struct LayerLists
{
  int* list[4];
  int* split;  // leaves to subdivide: the lists of the subdivision rounds, one behind the other
  int* count;  // [0..3] nodes per layer, [4..7] leaves to subdivide per round
  int* count_alt;  // the set the next multi_recut uses (cleared by this one: no launch spent on zeroing)
  int* snap;       // [4] nodes per layer before the subdivisions of this multi_recut (written by k_recut_all)
};

// Device data model of the voxel map and scan buffers (DESIGN.md "Data layout in HBM").
// Mirrors what the hot path needs from the reference's structures:
//   unordered_map<VOXEL_LOC,OctoTree*>  -> open-addressing table of 16-B slots (packed key -> root id)
//   OctoTree (octree.hpp:24-45)          -> NodeHot (256 B, what match() reads) + NodeCold (accumulators)
//   SlideWindow::points[slot]            -> one bump arena per window slot, (offset,count) per leaf
//   OctoTree::point_fix                  -> chained segments in a fixed-point pool
#pragma once
#include <cstdint>
#include "../../include/vina_b200.h"

#define VN_FLAG_PLANE 1     // Plane::is_plane
#define VN_FLAG_INTERIOR 2  // octo_state == 1
#define VN_FLAG_SPLIT_PENDING 8  // a leaf multi_recut has queued for subdivision (cleared when k_split turns it interior)
#define VN_FLAG_DEAD 4      // released by the map pruning (k_prune_*): the record is zero, layer == -1, id on the free stack
#define VN_KEY_BIAS (1 << 20)
#define VN_KEY_MASK ((1u << 21) - 1)
#define VN_EMPTY_KEY 0ull
#define VN_MAX_WORLD VINA_MAX_WORLD

// status bits written by kernels into Ctx::d_status
#define VN_ST_HASH_FULL 1
#define VN_ST_NODES_FULL 2
#define VN_ST_WIN_FULL 4
#define VN_ST_FIX_FULL 8
#define VN_ST_KEY_RANGE 16
#define VN_ST_SPIN 32
#define VN_ST_UNSORTED 64
#define VN_ST_DOWN_FULL 128

struct __align__(16) HashSlot
{
  unsigned long long key;  // packed voxel key, 0 = empty
  int root;                // node id, -2 while the creator is still initialising it
  int pad;
};

// what OctoTree::match / inside read (octree.cpp:551-595, 732-737); one 256-B record = two 128-B lines.
// Line 0 is what the descent, inside() and the first (fp32) gate test need: plane centre / normal / radius,
// voxel_center, quater_length, flags and the children mirror - a cached leaf costs ONE gather before the gate.
// Line 1 holds the sigma_l terms. sigma_l = J plane_var J^T with J = [w - c, -n] (octree.cpp:564-567) is
// hoisted per plane: with plane_var = [[A, B], [B^T, C]],
//   J plane_var J^T = d^T A d - 2 d.(B n) + n^T C n,  d = w - c,
// so the leaf stores A (6), B n (3) and n^T C n (1) instead of the 21 entries; the full plane_var lives in
// NodeCold (export / parity only).
struct __align__(128) NodeHot
{
  // ---- line 0
  double center[3];   // plane.center                     (double2 #0, #1.x)
  double normal[3];   // plane.normal                     (#1.y, #2)
  double vcenter[3];  // voxel_center                     (#3, #4.x)
  float radius;       // plane.radius                     (#4.y low)
  float ql;           // quater_length                    (#4.y high)
  int flags;          // VN_FLAG_*                        (#5.x low)
  int layer;          //                                  (#5.x high)
  int children[8];    // mirror of NodeCold::children for the IEKF descent (-1 = none)
  double qk;          // normal^T plane_var(3:6, 3:6) normal   (#7.y)
  // ---- line 1
  double qA[6];       // upper triangle of plane_var(0:3, 0:3) (#8, #9, #10)
  double qb[3];       // plane_var(0:3, 3:6) * normal           (#11, #12.x)
  int pad[14];
};
static_assert(sizeof(NodeHot) == 256, "NodeHot is two 128-byte lines");

// PointCluster (types.hpp:115-175); P is symmetric by construction of every
// cluster the scope builds, stored as (0,0),(1,0),(2,0),(1,1),(2,1),(2,2)
struct Cluster
{
  double P[6];
  double v[3];
  int N;
  int pad;
};

// pointVar (types.hpp:177-182) with the covariance stored symmetric:
// (0,0),(0,1),(0,2),(1,1),(1,2),(2,2) = the upper triangle of the reference's matrix
struct PointRec
{
  double p[3];
  double v[6];
};

// OctoTree::point_fix is a chain of segments of the fixed-point pool (one per marginalised frame). The chain is
// stored in 128-byte blocks of up to 15 segments: walking a 40-frame chain costs 3 dependent loads, not 40.
#define VN_FIXSEG_PER_BLOCK 15
struct __align__(128) FixSeg
{
  int off[VN_FIXSEG_PER_BLOCK];
  int n;
  int cnt[VN_FIXSEG_PER_BLOCK];
  int next;
};
static_assert(sizeof(FixSeg) == 128, "one line per hop of the chain walk");

struct NodeCold
{
  Cluster pcr_add, pcr_fix;
  Cluster pcrs_local[VINA_MAX_WIN];
  double cov_add[45];  // 9x9 symmetric, upper triangle packed by rows
  double plane_var[21];  // Plane::plane_var, upper triangle packed by rows (the IEKF reads NodeHot::qA/qb/qk)
  double eig_value[3];
  double eig_vector[9];  // column-major
  int win_off[VINA_MAX_WIN];
  int win_cnt[VINA_MAX_WIN];
  int fix_head, fix_tail, fix_count;  // chain of FixSeg
  int last_num, opt_state, isexist, has_sw;
  int path;
  int root;  // node id of the root voxel this node belongs to
  unsigned long long rootkey;
  int pend_cnt, pend_off;  // per-insert scratch: points of this scan landing in the leaf
  int touch_stamp, in_slide;
  int children[8];
  double jour;  // OctoTree::jour of a root: distance travelled at its last multi_margi (local_mapping.cpp:36)
};

// One LiDAR BA factor = one entry of the reference's LidarFactor container (factors.hpp:10-40), filled by
// tras_opt (octree.cpp:498-521) for every plane leaf of the slide map with lambda_0 / lambda_1 <= 0.12
struct BaFactor
{
  Cluster local[VINA_MAX_WIN];  // sw->pcrs_local[mp[i]], body frame of window frame i
  Cluster fix;                  // pcr_fix (world frame)
  Cluster add;                  // pcr_add; overwritten by evaluate_only_residual
  double eig_value[3];
  double eig_vector[9];  // column-major
  double coe;
  int node, pad;
};

// control block of the P2P record exchange (shard_kernels.cu); lives in the owner's memory, written by peers
struct ShardChan
{
  int counts[VINA_MAX_WORLD][VINA_MAX_WORLD];  // row r: records rank r sends to every destination
  unsigned long long ready[VINA_MAX_WORLD];    // epoch at which row r is valid
  unsigned long long done[VINA_MAX_WORLD];     // epoch at which rank r's records have landed in this rank's inbox
};
#define VN_CHAN_BUILD 0  // map-build records (13 doubles), first region of the inbox
#define VN_CHAN_QUERY 1  // association queries (10 doubles), second region of the inbox
struct ShardCtrl
{
  ShardChan ch[2];
  double sums[VINA_MAX_WORLD][40];             // row r: the 34 IEKF sums of rank r's shard (this iteration)
  unsigned long long sready[VINA_MAX_WORLD];   // epoch at which row r of `sums` is valid
};
struct ShardPeers
{
  int rank, world;
  double* inbox[VINA_MAX_WORLD];    // peer pointers (own entry = local buffer)
  ShardCtrl* ctrl[VINA_MAX_WORLD];
};

struct PoseD
{
  double R[9];  // column-major
  double p[3];
};

struct DeskewPoses
{
  int m;
  int pad;
  vina_imu_pose pose[VINA_MAX_POSES];
  double R_end[9], p_end[3];
  double ext_R[9], ext_t[3];
};

// everything the map kernels need, passed by value
struct MapView
{
  HashSlot* slots;
  unsigned int hmask;
  NodeHot* hot;
  NodeCold* cold;
  int max_nodes;
  int* node_count;
  int* root_count;
  PointRec* win_pool[VINA_MAX_WIN];
  int* win_cursor;  // [VINA_MAX_WIN]
  long long win_cap;
  PointRec* fix_pool;
  FixSeg* fix_segs;
  int* fix_cursor;      // points
  int* fixseg_cursor;   // segments
  long long fix_cap;
  int fixseg_cap;
  // ids / chain blocks given back by the map pruning (local_mapping.cpp:317-341); the allocators pop these
  // stacks before they bump node_count / fixseg_cursor. Only the pruning pushes, and never while an insert runs.
  int* free_nodes;
  int* free_count;
  int* free_segs;
  int* free_seg_count;
  double jour;  // the value multi_margi stamps into the roots of surf_map_slide (local_mapping.cpp:36, 507)
  int* slide_list[2];
  int* slide_count;  // [2]
  int slide_cur;     // which of the two lists is surf_map_slide right now
  int slide_others;  // roots in the surf_map_slide shards of the OTHER ranks (0 on one GPU): the
                     // "fewer roots than threads" early-outs are rules about the whole map
  int* status;
  // config
  double voxel_size;
  double min_eigen_value;
  double thre[4];  // already inverted
  double min_point[4];
  int max_layer, max_points, win_size, thread_num;
  int mp[VINA_MAX_WIN];
};

struct ScanView
{
  // pointVar SoA (body frame): p[3][cap], v[6][cap]
  double* p[3];
  double* v[6];
  int n;
};

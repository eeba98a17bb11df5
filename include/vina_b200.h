/* vina_b200.h — C ABI of the B200-native per-scan hot path of VINA-SLAM.
 *
 * This is the drop-in boundary. The reference has no FFI
 * layer: its "interface" is a handful of C++ functions with Eigen/PCL/STL
 * arguments called from the single odometry thread. Each entry point below
 * names the reference function it replaces (file:line under the reference
 * tree); INTEGRATION.md shows the thin C++ adapters a maintainer adds on the
 * reference side. POD only: no Eigen/STL/torch types cross the boundary.
 * All 3x3 matrices are COLUMN-MAJOR (Eigen's default storage), so
 * `Eigen::Matrix3d::data()` can be passed as is.
 *
 * Conventions
 *  - every call returns an int: 0 = ok, negative = error (VINA_E_*); nothing
 *    here ever calls exit() (the reference does, imu_ekf.cpp:19-24);
 *    vina_last_error(ctx) gives the text.
 *  - a ctx belongs to ONE caller thread (the odometry thread, node.cpp:437);
 *    calls are ordered on the ctx's CUDA stream. Different ctxs are
 *    independent (replicas / other GPUs).
 *  - there is NO CPU fallback: without a CUDA device vina_ctx_create fails
 *    with VINA_E_CUDA. (The only entry points that run without a device are the
 *    stateless host pieces of the BA, vina_ba_imu_evaluate / vina_ba_solve, the
 *    scan / IMU pairing vina_sync_*, the message unpacking vina_decode_* and
 *    vina_shard_owner / vina_config_default:
 *    they are host work in the product too.)
 *
 * Groups: lifetime | deskew, down-sampling, var_init | IEKF accumulate / device loop |
 * map insert, recut, margi | map sharded over GPUs (records and IEKF queries through
 * NCCL or through peer memory) | the per-scan loop body (vina_odom_*) | batch replay |
 * sliding-window BA (vina_ba_*, vina_odom_set_ba) | scan front end (vina_scan_prepare) and scan / IMU pairing
 * (vina_sync_*) | map pruning behind the vehicle (vina_map_prune, vina_odom_idle).
 */
#ifndef VINA_B200_H
#define VINA_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VINA_OK 0
#define VINA_E_ARG -1      /* bad argument */
#define VINA_E_CUDA -2     /* CUDA runtime error / no device */
#define VINA_E_CAPACITY -3 /* a device pool (points, voxels, nodes, hash) is full */
#define VINA_E_ORDER -4    /* scan not time-sorted (lidar_decoder.cpp:30 contract) */
#define VINA_E_TIME -5     /* "LiDAR time regress" (imu_ekf.cpp:19-24) */
#define VINA_E_STATE -6    /* call sequence error (e.g. accumulate before begin) */

#define VINA_MAX_WIN 10   /* LocalBA.win_size of every reference yaml */
#define VINA_MAX_POSES 96 /* IMU poses per scan (reference: ~20 @200 Hz, ~40 @400 Hz) */
#define VINA_MAX_WORLD 16 /* ranks a hash-range-sharded map can span */
#define VINA_SHARD_RECORD_DOUBLES 13 /* one routed point: body p[3], world var (upper) [6], world p[3], int64 scan index */
#define VINA_SHARD_QUERY_DOUBLES 10  /* one routed association query: body p[3], body var (upper) [6], int64 scan index */

typedef struct vina_ctx vina_ctx;

/* The reference's mutable globals (src/mapping/octree.cpp:67-75, node.cpp:38,
 * 210-219, 256-259) plus device capacities. */
typedef struct vina_config
{
  double voxel_size;                /* Odometry.voxel_size */
  double min_eigen_value;           /* Odometry.min_eigen_value */
  double plane_eigen_value_thre[4]; /* LocalBA.plane_eigen_value_thre AS IN THE YAML; inverted on load like node.cpp:256-259 */
  double min_point[4];              /* {20,20,15,10}, node.cpp:219 */
  double dept_err, beam_err;        /* Odometry.dept_err / beam_err (deg) */
  double down_size;                 /* Odometry.down_size */
  double ext_R[9];                  /* Lid_rot_to_IMU (column-major), ekf_imu.hpp:24 */
  double ext_t[3];                  /* Lid_offset_to_IMU, ekf_imu.hpp:25 */
  double cov_gyr, cov_acc, rdw_gyr, rdw_acc; /* Odometry.* noise, node.cpp:211-214 */
  int32_t max_layer;                /* LocalBA.max_layer (<= 3) */
  int32_t max_points;               /* 100, octree.cpp:70 */
  int32_t win_size;                 /* LocalBA.win_size (<= VINA_MAX_WIN) */
  int32_t thread_num;               /* LocalBA.thread_num: only the "fewer roots than threads" early-outs
                                       (voxel_map.cpp:96-97, local_mapping.cpp:26-28,150-154) depend on it */
  /* device capacities (0 = default) */
  int32_t max_scan_points;          /* points per scan */
  int32_t max_nodes;                /* octree nodes (roots + children) */
  int32_t hash_capacity_log2;       /* open-addressing table slots = 2^this (>= 2x root voxels) */
  int32_t device;                   /* CUDA device ordinal */
  int64_t fix_pool_points;          /* capacity of the fixed-point pool (point_fix lists) */
  int64_t win_pool_points;          /* capacity of EACH per-frame window arena (SlideWindow::points) */
} vina_config;

void vina_config_default(vina_config* cfg);

/* IMUST as POD (include/vina_slam/core/types.hpp:43-54); R and cov column-major. */
typedef struct vina_state
{
  double t;
  double R[9];
  double p[3], v[3], bg[3], ba[3], g[3];
  double cov[225];
} vina_state;

/* one entry of IMUEKF::imu_poses (imu_ekf.cpp:62-63): offset time, pose,
 * velocity, mean angular velocity and world acceleration of the interval. */
typedef struct vina_imu_pose
{
  double t;
  double R[9];
  double p[3], v[3], w[3], a[3];
} vina_imu_pose;

typedef struct vina_pose
{
  double R[9];
  double p[3];
} vina_pose;

/* one sensor_msgs::Imu: stamp (s), gyro, accel */
typedef struct vina_imu
{
  double t;
  double gyr[3];
  double acc[3];
} vina_imu;

/* one octree node as exported for parity checks (same layout as the oracle's vo_node_record) */
typedef struct vina_node_record
{
  int64_t key[3];
  int32_t code; /* layer | path<<2, path = child index per level, 3 bits each */
  int32_t layer, octo_state, isexist, has_sw, is_plane, last_num, opt_state;
  int32_t N_add, N_fix, n_point_fix, n_win_points;
  int32_t N_local[16];
  double P_add[9], v_add[3], P_fix[9], v_fix[3];
  double eig_value[3], eig_vector[9];
  double center[3], normal[3], plane_var[36], radius;
  double cov_add[81];
  double voxel_center[3], quater_length;
} vina_node_record;

/* per-stage device time of the last vina_odom_step / vina_map_update (ms, CUDA events) */
typedef struct vina_timings
{
  float deskew_ms, downsample_ms, var_init_ms, iekf_ms, insert_ms, recut_ms, margi_ms, total_ms;
  float iekf_kernel_ms; /* sum over iterations of the accumulate kernel alone */
  int32_t iekf_iters;
  int32_t kernel_launches;
} vina_timings;

/* ---- lifetime ---------------------------------------------------------- */
int vina_ctx_create(const vina_config* cfg, vina_ctx** out);
void vina_ctx_destroy(vina_ctx* ctx);
const char* vina_last_error(vina_ctx* ctx);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the ctx's own */
int vina_ctx_set_stream(vina_ctx* ctx, void* cuda_stream);
int vina_ctx_sync(vina_ctx* ctx);

/* ---- a2: deskew — replaces the per-point loop of IMUEKF::motion_blur
 * (src/estimation/imu_ekf.cpp:114-144). xyzt = n x (x,y,z,curvature) float32,
 * time-sorted (lidar_decoder.cpp:30). The scan stays resident on the device. */
int vina_scan_upload(vina_ctx* ctx, const float* xyzt, int n);
int vina_scan_upload_device(vina_ctx* ctx, const void* d_xyzt, int n); /* same, from a DEVICE pointer */
int vina_deskew(vina_ctx* ctx, const vina_imu_pose* poses, int m, const double R_end[9], const double p_end[3]);
int vina_scan_download(vina_ctx* ctx, float* xyzt, int cap); /* returns n or <0 */
/* ---- scan front end, between the sensor driver and IMUEKF::process: the keep rule every decoder handler applies
 * (src/sensor/lidar_pointcloud_decoder.cpp:70, 96, 161, 190: point i stays iff i % point_filter_num == 0 and
 * x*x + y*y + z*z > blind2; blind2 = General.blind SQUARED as node.cpp:210 leaves it) and pcl_handler
 * (src/sensor/lidar_decoder.cpp:16-34: two dummy points for an empty cloud, sort by time offset, cut at 0.11 s) on
 * the device. xyzt = raw points in arrival order, any time order. The result (time-sorted; equal stamps keep their
 * arrival order, which the reference's std::sort leaves unspecified) becomes the context's scan, like
 * vina_scan_upload; n_out / t_last = its size and last time offset (one synchronisation). vina_odom_step_prepared
 * is vina_odom_step on that scan. VINA_E_ARG if no kept point lies within 0.11 s. */
int vina_scan_prepare(vina_ctx* ctx, const float* xyzt, int n, int point_filter_num, double blind2, int* n_out,
                      float* t_last);
int vina_scan_prepare_device(vina_ctx* ctx, const void* d_xyzt, int n, int point_filter_num, double blind2, int* n_out,
                             float* t_last);

/* ---- f1: down_sampling_voxel (include/vina_slam/core/point_utils.hpp:7-44) on
 * the device-resident scan, incl. the "<2000 points -> down_size/2" retry of
 * local_mapping.cpp:396-403. vina_down_upload instead installs a caller-made
 * down-sampled cloud (used by stage-wise parity tests). */
int vina_downsample(vina_ctx* ctx);
int vina_down_count(vina_ctx* ctx); /* points after down-sampling incl. the retry (one sync); <0 = error */
int vina_down_upload(vina_ctx* ctx, const float* xyzt, int n);
int vina_down_download(vina_ctx* ctx, float* xyzt, int cap); /* returns n_d or <0 */

/* ---- a3: var_init (src/core/point_utils.cpp:36-52). which: 0 = full scan, 1 = down-sampled */
int vina_var_init(vina_ctx* ctx, int which);
/* caller-provided pointVar arrays instead (pnt n x 3; var n x 9 column-major) */
int vina_pvec_upload(vina_ctx* ctx, int which, const double* pnt, const double* var, int n);
/* pnt n x 3, var n x 9 (symmetric storage expanded) */
int vina_pvec_download(vina_ctx* ctx, int which, double* pnt, double* var, int cap);

/* ---- a4-a6: one IEKF iteration of VINA_SLAM::LioStateEstimation's point loop
 * (src/pipeline/odometry.cpp:98-148) incl. match() (src/mapping/voxel_map.cpp:241-266),
 * OctoTree::match/inside (src/mapping/octree.cpp:551-595, 732-737).
 * begin: fixes the prior covariance blocks (odometry.cpp:105-106) and resets the
 * per-point leaf cache (odometry.cpp:79). accumulate: sums for state (R,p);
 * HTH 6x6 column-major, HTz 6, nnt 3x3 column-major. */
int vina_iekf_begin(vina_ctx* ctx, int which, const double rot_var[9], const double tsl_var[9]);
int vina_iekf_accumulate(vina_ctx* ctx, const double R[9], const double p[3], double HTH[36], double HTz[6],
                         double nnt[9], int32_t* match_num);
/* same sums, and additionally records the per-point association for vina_iekf_debug_assoc */
int vina_iekf_accumulate_debug(vina_ctx* ctx, const double R[9], const double p[3], double HTH[36], double HTz[6],
                               double nnt[9], int32_t* match_num);
/* association of the last accumulate_debug: keys n x 3, leaf code (-1 = none), flag, sigma_d */
int vina_iekf_debug_assoc(vina_ctx* ctx, int64_t* keys, int32_t* codes, uint8_t* flags, double* sigma, int cap);

/* ---- a8-a12: pvec_update + cut_voxel_multi + multi_recut + multi_margi
 * (src/core/point_utils.cpp:54-65; src/mapping/voxel_map.cpp:47-135;
 * src/mapping/octree.cpp:151-177, 203-228, 335-495; src/pipeline/local_mapping.cpp:17-84, 144-201).
 * insert works on the down-sampled device pointVar set; win_ord = win_count-1 is the frame ordinal. */
int vina_map_insert(vina_ctx* ctx, int win_ord, const double R[9], const double p[3], const double cov_rot[9],
                    const double cov_tsl[9]);
int vina_map_recut(vina_ctx* ctx, int win_count, const vina_pose* x_buf);
int vina_map_margi(vina_ctx* ctx, int win_count, const vina_pose* x_buf);
/* rotate the ring map mp[] after a marginalisation (local_mapping.cpp:521-526) */
/* sizes of the last map update, for the measurement (bench.py): [0] points inserted, [1] leaves touched, [2..5]
 * nodes under surf_map_slide per layer (as multi_recut listed them), [6] leaves subdivided. Synchronises. */
int vina_map_last_counts(vina_ctx* ctx, int32_t out[8]);
int vina_map_shift_window(vina_ctx* ctx);
int64_t vina_map_count(vina_ctx* ctx, int64_t* n_roots, int64_t* n_slide);
int64_t vina_map_export(vina_ctx* ctx, vina_node_record* out, int64_t cap);
/* ---- map pruning behind the vehicle: the `else if (release_flag)` branch of the idle path of
 * thd_odometry_localmapping (src/pipeline/local_mapping.cpp:317-341) with OctoTree::tras_ptr
 * (src/mapping/octree.cpp:597-608). vina_map_set_journey sets the `jour` argument of multi_margi (:36, :507): the
 * value the next vina_map_margi stamps into every root of surf_map_slide. vina_map_prune erases every root voxel
 * (with its subtree) for which (int)(jour - root.jour) >= horizon; horizon <= 0 = the reference's 700 (metres of
 * travel). Roots still in surf_map_slide are kept. The node ids and fixed-point chain blocks go back to the
 * allocators, the hash table is rebuilt from the surviving roots and the fixed-point pool is compacted. */
int vina_map_set_journey(vina_ctx* ctx, double jour);
int vina_map_prune(vina_ctx* ctx, double jour, int horizon, int64_t* roots_erased, int64_t* nodes_freed);

/* ---- map sharded by voxel-hash range over `world` ranks (one ctx per rank / GPU). A root voxel - and every
 * leaf below it - lives on rank owner(key) = floor(hash(key) * world / 2^32): cut_voxel_multi only ever
 * combines points of the same root voxel (voxel_map.cpp:86, 108-134), so insert / recut / margi stay local
 * once each point has reached its owner. Per scan:
 *   vina_shard_route          pvec_update (point_utils.cpp:54-65) + key + owner for points [first, first+count)
 *                             of the ctx's down-sampled pointVar set; writes `count` records to d_send, grouped
 *                             by owner rank 0..world-1 in scan order (stable); counts_out[r] = records for rank r
 *   <exchange>                all-to-all of the records (NCCL: ncclSend/ncclRecv or torch.distributed
 *                             all_to_all_single - see INTEGRATION.md); receive in source-rank order
 *   vina_shard_insert_begin   received records -> key -> root find/create (voxel_map.cpp:53-87);
 *                             reports this rank's distinct-root and slide-map counts
 *   <all-reduce of the two counts: the reference's "fewer roots than threads" early-outs
 *    (voxel_map.cpp:96-97, local_mapping.cpp:26-28, 150-154) are rules about the WHOLE map>
 *   vina_shard_insert_finish  OctoTree::allocate/push on the local roots (octree.cpp:151-228)
 *   vina_map_recut / vina_map_margi / vina_map_shift_window as on one GPU.
 * The union of the shards equals the single-GPU map bit for bit when ranks route ascending slices of the scan. */
int vina_shard_owner(int64_t kx, int64_t ky, int64_t kz, int world); /* pure function; -1 = key out of range */
int vina_shard_route(vina_ctx* ctx, int world, int first, int count, int64_t scan_index_base, const double R[9],
                     const double p[3], const double cov_rot[9], const double cov_tsl[9], void* d_send,
                     int32_t* counts_out);
int vina_shard_insert_begin(vina_ctx* ctx, const void* d_recv, int n, int win_ord, int32_t* local_roots,
                            int32_t* local_slide);
int vina_shard_insert_finish(vina_ctx* ctx, int win_ord, int global_roots, int global_slide);
/* The same exchange FUSED into the routing kernel over peer memory (NVLink): every record is stored straight into
 * its owner's inbox at its final, scan-ordered position; the counts table and the completion flags travel the
 * same way. No staging buffer, no collective call, no host synchronisation between routing and arrival.
 *   vina_shard_p2p_create    allocates this rank's inbox (inbox_records x 13 doubles) and control block; returns
 *                            their CUDA IPC handles (2 x 64 bytes) to be all-gathered by the caller
 *   vina_shard_p2p_connect   opens the peers' handles (world x 128 bytes, rank order)
 *   vina_shard_p2p_connect_local   peers living in the same process: raw pointers from vina_shard_p2p_pointers
 *   vina_shard_route_p2p     as vina_shard_route, but asynchronous and with the exchange inside. phase 0 = all of
 *                            it; phase 1 = only the part that never waits (owners, counts, my counts row to the
 *                            peers), phase 2 = only the part that waits for the peers' rows, stores the records
 *                            and signals - for several ranks driven from ONE host thread on ONE GPU (tests),
 *                            where a kernel must never wait for work that is enqueued after it
 *   vina_shard_insert_begin_p2p    waits (on the device) for every peer's records, then as vina_shard_insert_begin;
 *                            also returns the number of records received
 * A 2-int all-reduce per scan (see above) must separate two scans: it is what guarantees that every rank has
 * consumed its inbox before the next scan's records arrive. */
int vina_shard_p2p_create(vina_ctx* ctx, int rank, int world, int64_t inbox_records, void* ipc_handles_out);
int vina_shard_p2p_connect(vina_ctx* ctx, const void* all_handles);
int vina_shard_p2p_pointers(vina_ctx* ctx, void** inbox, void** ctrl);
int vina_shard_p2p_connect_local(vina_ctx* ctx, void* const* inbox_ptrs, void* const* ctrl_ptrs);
int vina_shard_route_p2p(vina_ctx* ctx, int first, int count, int64_t scan_index_base, const double R[9],
                         const double p[3], const double cov_rot[9], const double cov_tsl[9], int phase);
int vina_shard_insert_begin_p2p(vina_ctx* ctx, int win_ord, int32_t* n_recv, int32_t* local_roots, int32_t* local_slide);

/* Association against the sharded map (one IEKF iteration, odometry.cpp:98-148): every rank routes points
 * [first, first+count) of its FULL-scan pointVar set to the owners of the voxels they fall into under (R, p)
 * (vina_shard_query_route, 10-double records), the records are exchanged (all-to-all), each owner evaluates gate,
 * residual and Jacobian of what it received against its shard (vina_shard_query_accumulate -> the 34 packed
 * sums: 21 HTH upper by rows, 6 HTz, 6 nnt upper, match count, written to DEVICE memory d_sums34), the sums
 * are all-reduced and every rank applies the same update (vina_odom_iekf_host_begin / _update, the reference's
 * 15x15 route). No per-point leaf cache survives the exchange: every iteration looks its voxel up again, which
 * differs from the cached path (odometry.cpp:124-127) only for points that sit on a voxel face to within
 * float rounding of the key. */
int vina_shard_query_route(vina_ctx* ctx, int world, int first, int count, int64_t scan_index_base, const double R[9],
                           const double p[3], void* d_send, int32_t* counts_out);
int vina_shard_query_accumulate(vina_ctx* ctx, const void* d_recv, int n, const double R[9], const double p[3],
                                const double rot_var[9], const double tsl_var[9], double* d_sums34);
/* the reference's IEKF update on the host, one iteration at a time, for externally reduced sums
 * (odometry.cpp:82, 192-230). begin: x_prop = x_curr, P^-1, loop counters. update: returns 1 when the loop is
 * finished (converged twice or out of iterations; the covariance has then been updated), 0 otherwise. */
int vina_odom_iekf_host_begin(vina_ctx* ctx, int max_iter);
int vina_odom_iekf_host_update(vina_ctx* ctx, const double sums34[34]);
/* The same loop (VINA_SLAM::LioStateEstimation, odometry.cpp:64-255, against the sharded map) with the exchange and
 * the update fused into the kernels - no collective call and no host synchronisation inside the loop. Per
 * iteration every rank's routing kernel stores its queries straight into the owners' inboxes over peer memory
 * (vina_shard_p2p_create / _connect; query region behind the map-build region), the owners run the accumulate
 * kernel on what arrived, store their 34 sums into every peer's control block, and every rank adds the rows in
 * rank order (bitwise the same total everywhere) and applies the update of odometry.cpp:192-230 to its copy of
 * the device iterate; iterations after convergence return at once. x_curr (vina_odom_set_state / the previous
 * scan) must be the same on every rank; points [first, first+count) of the FULL-scan pointVar set are this rank's
 * share of the scan. phase VINA_SHARD_IEKF_ALL: stage, enqueue max_iter iterations, wait, install the result as
 * x_curr (one process per GPU). The other phases split that for several ranks driven from one host thread (a kernel
 * must never wait for work enqueued behind it): STAGE once on every rank; then per iteration ROUTE on every rank
 * (never waits), SEND on every rank, EVAL on every rank, SOLVE on every rank; FINISH on every rank. Extra
 * iterations after convergence are no-ops, like in the single-GPU loop. */
#define VINA_SHARD_IEKF_ALL 0
#define VINA_SHARD_IEKF_STAGE 1
#define VINA_SHARD_IEKF_ROUTE 2
#define VINA_SHARD_IEKF_SEND 3
#define VINA_SHARD_IEKF_EVAL 4
#define VINA_SHARD_IEKF_SOLVE 5
#define VINA_SHARD_IEKF_FINISH 6
int vina_odom_iekf_sharded_p2p(vina_ctx* ctx, int first, int count, int max_iter, int phase, int* iters_out,
                               int* not_degenerate);

/* ---- the per-scan loop body (src/pipeline/local_mapping.cpp:389-546), host
 * orchestration in C++ inside the library: a1 IMU propagation on the host
 * (imu_ekf.cpp:33-94), kernels for a2-a6/a8-a12, a7 solve on the host
 * (odometry.cpp:192-230). */
int vina_odom_set_state(vina_ctx* ctx, const vina_state* s);
int vina_odom_get_state(vina_ctx* ctx, vina_state* s);
int vina_odom_set_imu_anchor(vina_ctx* ctx, double last_pcl_end_time, const vina_imu* last_imu, double scale_gravity);
/* harness bootstrap (instead of the start-up phase below, when the first states are known): an already
 * deskewed scan at a known state -> downsample, var_init, map update */
int vina_odom_bootstrap(vina_ctx* ctx, const float* xyzt, int n, const vina_state* x_known);
/* ---- the start-up phase: VINA_SLAM::initialization (src/platform/ros2/node.cpp:293-366) as the loop calls it
 * (src/pipeline/local_mapping.cpp:362-388) - IMU initialisation (IMUEKF::IMU_init, imu_ekf.cpp:147-172), then for
 * win_size scans the kd-tree IEKF against a local map (lio_state_estimation_kdtree, odometry.cpp:267-439), then
 * Initialization::motion_init (initialization.cpp:158-367: map rebuilt from the re-deskewed frames, recut, gravity
 * BA LI_BA_OptimizerGravity, optimizers.cpp:746-826; gravity alignment), the window's first marginalisation.
 * vina_odom_cold_start puts a context into that phase (empty map, zero state); every scan then goes to
 * vina_odom_init_scan (raw scan: n x (x, y, z, time offset), time-sorted; its IMU batch as sync_packages hands it
 * over) until *status == 1: 0 = still collecting, 1 = initialised - the next scan goes to vina_odom_step -,
 * -1 = motion_init failed and the system was reset (node.cpp:368-408): keep feeding scans. x_out: x_curr. */
int vina_odom_cold_start(vina_ctx* ctx);
int vina_odom_init_scan(vina_ctx* ctx, const float* xyzt, int n, double pcl_beg_time, const vina_imu* imus, int m,
                        vina_state* x_out, int32_t* status);
/* one scan from HOST buffers. iekf_on_full: IEKF on the un-downsampled scan
 * (production, local_mapping.cpp:413) or the down-sampled one (:412).
 * max_iter <= 0: the reference's 20 (plain variant); 4 = the VNC_lio budget. */
int vina_odom_step(vina_ctx* ctx, const float* xyzt, int n, double pcl_beg_time, const vina_imu* imus, int m,
                   int iekf_on_full, int max_iter, vina_state* x_out);
int vina_odom_step_prepared(vina_ctx* ctx, double pcl_beg_time, const vina_imu* imus, int m, int iekf_on_full,
                            int max_iter, vina_state* x_out);
/* same scan, but the raw points are already in HBM: d_xyzt is a DEVICE pointer
 * (n x 4 float32) that is copied device-to-device into the ctx's scan buffer;
 * pcl_end_time = pcl_beg_time + curvature of the last point (sync.cpp:40),
 * which the caller knows. The timed "inputs resident" leg of bench.py. */
int vina_odom_step_resident(vina_ctx* ctx, const void* d_xyzt, int n, double pcl_beg_time, double pcl_end_time,
                            const vina_imu* imus, int m, int iekf_on_full, int max_iter, vina_state* x_out);
/* ---- batch replay: B independent sequences (contexts on the same GPU) advance one scan each per call.
 * Per-sequence stages run concurrently on the contexts' own streams; the IEKF iterations of all sequences are
 * ONE k_iekf launch per iteration (grid blocks x B). Results are identical to B separate vina_odom_step_resident
 * calls (same kernels, same per-sequence reduction order up to the number of blocks per sequence). */
typedef struct vina_batch vina_batch;
int vina_batch_create(vina_ctx** ctxs, int n, vina_batch** out); /* n <= 16 */
void vina_batch_destroy(vina_batch* b);
int vina_batch_step_resident(vina_batch* b, const void* const* d_xyzt, const int32_t* n, const double* pcl_beg_time,
                             const double* pcl_end_time, const vina_imu* const* imus, const int32_t* m,
                             int iekf_on_full, int max_iter, vina_state* x_out /* [B] or NULL */);
/* device time of each batched k_iekf launch of the last step (CUDA events on the batch stream) */
int vina_batch_iekf_time(vina_batch* b, float* ms_per_launch, int cap, int32_t* launches);
int vina_batch_sync(vina_batch* b);

/* stage-wise pieces of the step, for parity tests */
int vina_odom_propagate(vina_ctx* ctx, double pcl_beg_time, double pcl_end_time, const vina_imu* imus, int m,
                        vina_imu_pose* poses_out, int cap); /* returns #poses */
/* VINA_SLAM::LioStateEstimation (odometry.cpp:64-255) on the current state. vina_odom_iekf keeps the whole
 * iteration loop on the device (k_iekf accumulates, its last block solves and updates the device-resident
 * iterate; one host synchronisation per call). vina_odom_iekf_host is the same loop with the 15x15 update
 * (odometry.cpp:192-230) on the host, one 34-double readback per iteration - the cross-check of the former. */
int vina_odom_iekf(vina_ctx* ctx, int which, int max_iter, int* iters_out, int* not_degenerate);
int vina_odom_iekf_host(vina_ctx* ctx, int which, int max_iter, int* iters_out, int* not_degenerate);
int vina_odom_map_update(vina_ctx* ctx); /* pvec_update + insert + recut + (margi + shift) with x_curr */
int vina_odom_window(vina_ctx* ctx, int* win_count, int* mp, int cap);
/* distance travelled (`jour`, `last_pos`, `release_flag`: local_mapping.cpp:262-263, 272, 509-519), kept by
 * vina_odom_step / _step_resident / the batch step; and the idle path of the loop (:303-341), to be called when no
 * scan is waiting: prunes the map (vina_map_prune with the current journey) if release_flag is set. */
int vina_odom_journey(vina_ctx* ctx, double* jour, int* release_flag);
int vina_odom_idle(vina_ctx* ctx, int horizon, int64_t* roots_erased, int64_t* nodes_freed);
int vina_get_timings(vina_ctx* ctx, vina_timings* t);
/* ---- sliding-window BA, the LiDAR factor (the data-parallel part of LI_BA_Optimizer::damping_iter,
 * src/mapping/optimizers.cpp:430-517; the IMU pre-integration factors and the LM loop are host work and not
 * part of this library yet). The factor store is the device form of the reference's `voxhess` container.
 *   vina_ba_collect          tras_opt (octree.cpp:498-521, local_mapping.cpp:196-200): copy every plane leaf of the
 *                            slide map that k_recut marked as a factor (lambda_0 / lambda_1 <= 0.12) into the store.
 *                            Call between vina_map_recut and vina_map_margi. vina_ba_set_capture(ctx, 1) makes
 *                            vina_odom_map_update / vina_odom_step do it after every recut with a full window.
 *   vina_ba_lidar_hessian    LidarFactor::acc_evaluate2 (factors.cpp:22-126) over all factors for the window poses
 *                            xs: Hess (6 win x 6 win, column-major, lower blocks mirrored), JacT (6 win), residual.
 *   vina_ba_lidar_residual   LidarFactor::evaluate_only_residual (factors.cpp:128-158): the residual at candidate
 *                            poses; overwrites the stored factors' eigen-decomposition and pcr_add like the
 *                            reference's container. lam0 (nullable, cap entries) receives lambda_0 per factor.
 * Sums over factors run in the store's order (arrival order of an atomic cursor): results agree with the
 * reference to rounding (tests: 1e-9 of the largest entry), the per-factor eigenvalues bit for bit. */
/* The whole sliding-window BA inside vina_odom_step (LocalBA.if_BA: 1 of mid360.yaml / velodyne.yaml,
 * local_mapping.cpp:437-441, 492-497, 541-546): IMU pre-integration factors (imu_preintegration.cpp) and the
 * Levenberg-Marquardt loop of LI_BA_Optimizer::damping_iter (optimizers.cpp:430-517) on the host, the LiDAR factor on
 * the device, OctoTree::margi taking the re-evaluated factors back. Call before the first frame enters the window;
 * imu_coef <= 0 keeps LocalBA.imu_coef = 1e-4. BA runs once every pair of consecutive window frames has an IMU
 * factor (frames from vina_odom_bootstrap have none). */
int vina_odom_set_ba(vina_ctx* ctx, int on, double imu_coef);
int vina_odom_ba_stats(vina_ctx* ctx, int32_t* runs, int32_t* last_iters);
/* host-side pieces of the BA, stateless (no context, no device): one IMU pre-integration factor built from `m`
 * IMU samples (IMU_PRE::push_imu, imu_preintegration.cpp:32-100) and evaluated between two states
 * (give_evaluate, :102-163): residual r^T cov^-1 r, and - when jtj / gg are not null - J^T cov^-1 J (30 x 30,
 * column-major) and J^T cov^-1 r (30); the symmetric solve of the LM step (lower triangle of A is read). */
int vina_ba_imu_evaluate(const vina_config* cfg, const double bg[3], const double ba[3], const vina_imu* imus, int m,
                         double scale_gravity, const vina_state* s1, const vina_state* s2, double* residual, double* jtj,
                         double* gg);
int vina_ba_solve(const double* A, int n, const double* b, double* x);
int vina_ba_set_capture(vina_ctx* ctx, int on);
int vina_ba_collect(vina_ctx* ctx, int32_t* n_factors);
int vina_ba_count(vina_ctx* ctx, int32_t* n_factors);
int vina_ba_lidar_hessian(vina_ctx* ctx, const vina_pose* xs, int win, double* Hess, double* JacT, double* residual);
int vina_ba_lidar_residual(vina_ctx* ctx, const vina_pose* xs, int win, double* residual, double* lam0, int cap);

/* ---- pairing of scans and IMU samples in front of the step: the buffers of src/sensor/sync.cpp:5-16 and
 * sync_packages (:18-96) as a host object (no device, no context; thread-safe like the reference's mBuf).
 * push_imu = imu_handler (src/platform/ros2/subscribers.cpp:11-20); push_scan = the tail of pcl_handler
 * (src/sensor/lidar_decoder.cpp:36-43): t_start = the message stamp, t_last = back().curvature of the prepared
 * scan (vina_scan_prepare's t_last), tag = whatever identifies the scan for the caller. vina_sync_next =
 * sync_packages: 1 = a package is ready (tag, pcl_beg_time, pcl_end_time, the m IMU samples up to pcl_end_time);
 * 0 = the reference returns false and nothing was consumed (no scan, or the IMU stream has not passed the scan's
 * end yet); 2 = the reference returns false and scan `tag` is gone (<= 4 IMU samples, or the first scan with
 * point_notime); VINA_E_STATE = the IMU buffer ran dry (the reference exit(0)s, sync.cpp:79-82); VINA_E_CAPACITY =
 * more samples than `cap`: nothing was consumed, the scan stays held, *m = the number of samples it waits for - call
 * again with a larger buffer. point_notime != 0: scans without per-point time, frame interval as in sync.cpp:43-56. */
typedef struct vina_sync vina_sync;
int vina_sync_create(int point_notime, vina_sync** out);
void vina_sync_destroy(vina_sync* s);
int vina_sync_push_imu(vina_sync* s, const vina_imu* imu);
int vina_sync_push_scan(vina_sync* s, double t_start, double t_last, int64_t tag);
int vina_sync_pending(vina_sync* s, int32_t* scans, int32_t* imus);
int vina_sync_next(vina_sync* s, int64_t* tag, double* pcl_beg_time, double* pcl_end_time, vina_imu* imus, int cap,
                   int32_t* m);

/* ---- unpacking of the driver messages: LidarPointCloudDecoder::process and its handlers
 * (src/sensor/lidar_pointcloud_decoder.cpp:21-240), host functions (no device, no context). A PointCloud2 is its
 * data pointer, the point count (width * height) and the byte offsets of the fields pcl::fromROSMsg maps by name
 * into the handler's point struct (include/vina_slam/lidar_pointcloud_decoder.hpp:44-109): x, y, z FLOAT32 and the
 * time field (`time` FLOAT32 for Velodyne, `t` UINT32 ns for Ouster, `timestamp` FLOAT64 for Hesai / RoboSense;
 * t_datatype = the sensor_msgs::PointField datatype 6 / 7 / 8). Output: n_kept x (x, y, z, curvature) float32 in
 * arrival order, each handler's own time rule and keep rule applied (RoboSense: planar blind test; Velodyne without
 * usable stamps: time from the azimuth at omega_l deg/s) - to be handed to vina_scan_prepare with
 * point_filter_num = 1 and blind2 < 0 (sort + 0.11 s cut). blind2 = General.blind squared (node.cpp:210).
 * Returns the number of points written, VINA_E_CAPACITY if cap is too small. */
enum
{
  VINA_LIDAR_LIVOX = 0,
  VINA_LIDAR_VELODYNE = 1,
  VINA_LIDAR_OUSTER = 2,
  VINA_LIDAR_HESAI = 3,
  VINA_LIDAR_ROBOSENSE = 4,
  VINA_LIDAR_TARTANAIR = 5 /* LID_TYPE, lidar_pointcloud_decoder.hpp:20-28 */
};
typedef struct vina_pc2_layout
{
  int32_t point_step;
  int32_t off_x, off_y, off_z;
  int32_t off_t;      /* -1: none (TartanAir) */
  int32_t t_datatype; /* 6 = UINT32, 7 = FLOAT32, 8 = FLOAT64 */
  int32_t is_bigendian;
} vina_pc2_layout;
/* livox_ros_driver2::msg::CustomPoint */
typedef struct vina_livox_point
{
  uint32_t offset_time; /* ns from the message stamp */
  float x, y, z;
  uint8_t reflectivity, tag, line, pad;
} vina_livox_point;
int64_t vina_decode_pointcloud2(int lidar_type, const uint8_t* data, int64_t n_points, const vina_pc2_layout* layout,
                                double header_stamp, double omega_l, double blind2, int point_filter_num, float* xyzt,
                                int64_t cap);
int64_t vina_decode_livox(const vina_livox_point* pts, int64_t n_points, double blind2, int point_filter_num,
                          float* xyzt, int64_t cap);
/* back().curvature of the scan pcl_handler would queue for these decoded points (lidar_decoder.cpp:16-34): the largest
 * time offset not beyond 0.11 s, 0.09 for an empty cloud - the t_last of vina_sync_push_scan when the scan is sorted
 * and cut later, on the device. VINA_E_ARG if every stamp is beyond 0.11 s. */
int vina_scan_last_stamp(const float* xyzt, int64_t n, float* t_last);

/* record per-stage CUDA-event timings (adds event records + one sync per step) */
int vina_set_profiling(vina_ctx* ctx, int on);
/* vina_odom_step / _step_resident schedule (default on): down-sampling and the var_init of the map's point set run
 * on a side stream concurrently with the IEKF loop, and the map update (pvec_update, cut_voxel_multi, multi_recut,
 * multi_margi - local_mapping.cpp:425-507) is enqueued behind the loop with the new pose and the posterior
 * covariance blocks read from the device iterate, before the result has reached the host. off: everything in
 * stream order with the host in between (the schedule the per-stage timers of vina_set_profiling see). Both give
 * bitwise the same states and maps. */
int vina_set_overlap(vina_ctx* ctx, int on);
/* on: the host-to-device copy of vina_odom_step starts only behind the work that is already on the context's stream
 * when the call is made (default off: the copy runs on its own stream as early as the scan buffer is free and
 * overlaps the previous scan's map update). For timing a step end to end with events on that stream: with the
 * option on the copy lies inside the bracketed region. */
int vina_set_upload_ordered(vina_ctx* ctx, int on);
/* the IEKF iteration loop of LioStateEstimation (odometry.cpp:98-231) inside vina_odom_step as ONE persistent
 * cooperative launch (k_iekf_loop: scan resident in shared memory over the iterations, grid barrier and the a7
 * update between them; the front of the step then goes out as two fused launches in stream order). Default off =
 * one k_iekf launch per iteration with the update in its last block and the down-sampling on a side stream next to
 * them: on B200 the persistent kernel shortens the loop by ~30 us per scan but fills every SM, so the side stream no
 * longer overlaps and the whole step is ~5 % slower. Same associations either way; the sums
 * agree to rounding (different, but fixed, summation order). Environment: VINA_IEKF_LOOP=1. */
int vina_set_iekf_loop(vina_ctx* ctx, int on);

#ifdef __cplusplus
}
#endif
#endif

// The opaque context behind the C ABI (include/vina_b200.h): device pools,
// per-scan buffers, the CUDA stream and the host-side odometry state.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "vn_kernels.cuh"

struct OdomHost;  // host/vina_pipeline.cpp

struct vina_ctx
{
  vina_config cfg;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  std::string err;

  // scan buffers
  int cap_points = 0;
  float4* d_scan = nullptr;  // raw -> deskewed in place (x,y,z,curvature)
  float4* d_down = nullptr;  // down-sampled cloud
  int n_scan = 0;
  int n_down = 0;            // host copy (valid after a sync)
  bool n_down_pending = false;
  int* d_n_down = nullptr;
  int* h_n_down = nullptr;   // pinned
  // fused front (vn_front_fused): the count arrives through mapped memory, [0] sequence number, [1] count
  bool down_fuse_var_init = false;      // run_downsample: k_down_emit also does var_init(1) (set around the per-scan step)
  bool front_fused = true;
  bool n_down_mapped = false;           // the pending count is the one k_down_emit_all publishes
  bool split_overlap = false;           // k_split on the side stream next to the marginalisation (launch_map_recut_margi)
  bool batch_member = false;            // the context belongs to a vina_batch
  bool front_was_fused = false;         // the serial step's front went out as the two fused launches
  bool down_retried = false;            // vn_finish_downsample re-ran the down-sampling (< 2000 points rule)
  unsigned long long* h_down_pub = nullptr;
  unsigned long long* d_down_pub = nullptr;
  unsigned long long down_seq = 0;
  int* d_emit_counts = nullptr;
  unsigned long long* d_emit_bar = nullptr;
  ScanView pv[2];            // [0] full scan, [1] down-sampled
  int n_pv[2] = { 0, 0 };
  int* d_cache = nullptr;
  bool cache_is_reset = false;  // the fused deskew + var_init kernel has just written -1 everywhere
  DeskewPoses* d_poses = nullptr;
  DeskewPoses* h_poses = nullptr;  // pinned
  // scan front end (vina_scan_prepare; allocated on first use)
  float4* d_raw = nullptr;
  unsigned int* d_fkey[2] = { nullptr, nullptr };
  int* d_fidx[2] = { nullptr, nullptr };
  int* d_fhist = nullptr;
  unsigned long long* h_front_pub = nullptr;  // mapped: [0] sequence number, [1..4] the counters of k_front_sort
  unsigned long long* d_front_pub = nullptr;
  unsigned long long front_seq = 0;
  int* d_fbkt = nullptr;                    // bucket counts + cursors of the front end's fast path (2 x 1024)
  unsigned long long* d_fpairs = nullptr;   // (time key, arrival index) pairs grouped by bucket
  int* d_fcnt = nullptr;     // 2 counters + the last point's time offset
  int* h_fcnt = nullptr;     // pinned copy
  float front_t_last = 0.f;  // time offset of the last point of the prepared scan (pcl_end_time - pcl_beg_time)
  bool front_valid = false;  // d_scan holds a scan made by vina_scan_prepare
  // down-sampling scratch
  DownSlot* d_dtab = nullptr;
  unsigned int dmask = 0;
  int *d_slot_of = nullptr, *d_flag = nullptr, *d_scanbuf = nullptr, *d_block_sums = nullptr;
  // IEKF
  int iekf_which = -1;
  int iekf_blocks = 0;
  int iekf_variant = 0;
  unsigned long long iekf_seq = 0;
  double rot_var[9], tsl_var[9];
  double* d_partials = nullptr;
  unsigned int* d_ticket = nullptr;
  bool iekf_loop = false;                    // vina_set_iekf_loop: the iteration loop as one persistent launch (k_iekf_loop)
  unsigned long long* d_loop_bar = nullptr;  // grid-barrier words of k_iekf_loop
  double* d_loop_partials = nullptr;         // [2][34][sm_count]
  bool iterate_uploaded = false;             // iekf_upload_iterate ran ahead of iekf_stage (the overlapped step)
  bool iekf_looped = false;                  // the last enqueued loop went out as k_iekf_loop
  int loop_launches = 0;                     // profiling: launches of k_iekf_loop timed so far
  double* h_result = nullptr;  // pinned + mapped; the kernel's last block writes it
  double* d_result = nullptr;  // device alias of h_result
  IekfDev* d_iekf = nullptr;   // device-resident iterate (state, covariance, flags)
  IekfDev* h_iekf = nullptr;   // pinned staging of the upload
  IekfDev* h_pub = nullptr;    // mapped pinned: the converged iterate as published by k_publish_iterate
  IekfDev* d_pub = nullptr;    // device alias of h_pub
  unsigned long long* h_pub_flag = nullptr;  // mapped pinned sequence number, written after the data
  unsigned long long* d_pub_flag = nullptr;
  unsigned long long pub_seq = 0;
  cudaStream_t copy_stream = nullptr;  // host -> device scan uploads: overlap the previous scan's map update
  cudaEvent_t ev_scan_up = nullptr;    // the upload into d_scan has landed
  cudaEvent_t ev_scan_rd = nullptr;    // the last enqueued reader of d_scan is done
  bool scan_rd_valid = false;
  // vina_odom_step: the scan is uploaded in chunks, and the fused deskew kernel follows chunk by chunk (the copy of a
  // 240 000-point scan takes ~3x as long as its deskew: only the last chunk's kernel is left behind the copy)
  cudaEvent_t ev_chunk[4] = { nullptr, nullptr, nullptr, nullptr };
  int upload_chunks = 0;        // > 0: chunks of the scan in d_scan whose events the deskew still has to wait for
  int chunk_end[4] = { 0, 0, 0, 0 };
  bool upload_ordered = false;  // vina_set_upload_ordered: the upload starts behind the work already on the stream
  cudaEvent_t ev_step_begin = nullptr;
  bool overlap = true;                 // vina_set_overlap: the per-scan step forks / hands the pose over on the device
  cudaStream_t side_stream = nullptr;  // down-sampling + var_init of the map's point set, concurrent with the IEKF
  cudaEvent_t ev_collect_fork = nullptr, ev_collect_done = nullptr;  // k_recut_collect next to the insert's accumulation
  bool early_collect = true, collected_early = false;
  cudaEvent_t ev_fork = nullptr;       // the deskewed scan is ready (compute stream -> side stream)
  cudaEvent_t ev_join = nullptr;       // the down-sampled pointVar set is ready (side stream -> compute stream)
  cudaEvent_t ev_poses = nullptr;  // the pose-table staging buffer has been consumed
  bool poses_in_flight = false;
  IekfDebug dbg = { nullptr, nullptr, nullptr, nullptr };
  bool dbg_valid = false;
  // map
  MapView map;
  InsertScratch ins;
  LayerLists layers;
  unsigned int hash_slots = 0;
  int* d_prune = nullptr;  // counters of vina_map_prune
  int* d_status = nullptr;
  int* h_status = nullptr;  // pinned
  // hash-range sharding scratch (allocated on first use)
  unsigned char* d_sh_owner = nullptr;
  int* d_sh_hist = nullptr;
  int* d_sh_counts = nullptr;  // [VINA_MAX_WORLD] counts, then [VINA_MAX_WORLD + 1] segment starts
  int* h_sh_counts = nullptr;  // pinned
  // P2P record exchange (vina_shard_*_p2p)
  ShardPeers peers;
  double* p2p_inbox = nullptr;
  ShardCtrl* p2p_ctrl = nullptr;
  long long p2p_cap = 0;
  unsigned long long p2p_epoch = 0;
  unsigned long long p2p_qepoch = 0;  // query channel: one epoch per IEKF iteration
  int* d_n_recv = nullptr;
  long long* d_p2p_base = nullptr;  // my base offset in every owner's inbox (this scan)
  void* p2p_opened[2 * VINA_MAX_WORLD] = { nullptr };  // IPC mappings to close
  bool p2p_connected = false;
  int win_count_last = 0;

  // VINA_TRACE=1: host / device timeline of the overlapped step, printed to stderr when the ctx is destroyed
  bool trace = false;
  cudaEvent_t tr_ev[8] = { nullptr };
  double tr_host_us[8] = { 0 };
  double tr_dev_us[8] = { 0 };
  int tr_n = 0;
  // BA LiDAR factor store (allocated on first use): the device form of the reference's `voxhess`
  BaFactor* d_ba = nullptr;
  int ba_cap = 0;
  int* d_ba_n = nullptr;
  int ba_n = -1;             // host copy (valid after vina_ba_collect)
  double* d_ba_partial = nullptr;
  double* d_ba_out = nullptr;   // Hess (6 win)^2, JacT (6 win), residual
  double* d_ba_lam = nullptr;
  double* h_ba_out = nullptr;   // mapped pinned: the kernels write Hess / JacT / residual (or block sums) here
  double* d_ba_map = nullptr;   // device alias of h_ba_out
  unsigned long long* h_ba_flag = nullptr;  // mapped pinned completion sequence number
  unsigned long long* d_ba_flag = nullptr;
  unsigned int* d_ba_ticket = nullptr;
  unsigned long long ba_seq = 0;
  bool ba_capture = false;   // vina_ba_set_capture: collect after every recut with a full window
  // profiling
  bool profiling = false;
  cudaEvent_t ev[16];
  std::vector<cudaEvent_t> iekf_ev;  // per-iteration (begin, end) pairs of the device IEKF loop
  vina_timings tm;
  int launches = 0;

  // start-up phase (vina_odom_init_scan; allocated on first use): the local map of the kd-tree IEKF (two buffers:
  // appended to, then down-sampled into the other), per-point plane (distance, normal), block partials, the 28 sums
  float4* d_tree[2] = { nullptr, nullptr };
  int tree_cur = 0, n_tree = 0;
  double* d_init_ds = nullptr;
  double* d_init_dir = nullptr;
  double* d_init_part = nullptr;
  double* h_init_sums = nullptr;  // mapped pinned
  double* d_init_sums = nullptr;  // device alias

  OdomHost* odom = nullptr;
};

int vn_fail(vina_ctx* c, int code, const char* fmt, ...);
int vn_init_ensure(vina_ctx* c);        // buffers of the start-up phase
int vn_map_clear(vina_ctx* c);          // the map back to its state after vina_ctx_create (motion_init rebuilds it every round)
// down_sampling_voxel of an arbitrary device cloud (no "< 2000 points" retry); synchronises, returns the count
int vn_map_recut_margi_live(vina_ctx* c, int win_count, const vina_pose* x_buf);
int vn_scan_upload_chunked(vina_ctx* c, const float* xyzt, int n);
int vn_settle_upload(vina_ctx* c);
int vn_front_fused(vina_ctx* c, const vina_imu_pose* poses, int m, const double R_end[9], const double p_end[3]);
int vn_downsample_cloud(vina_ctx* c, const float4* in, int n, double size, float4* out, int* n_out);
// sum of n n^T over the normals (eigenvector of the smallest eigenvalue) of the collected BA factors, 3x3 column-major
int vn_ba_normal_scatter(vina_ctx* c, double nnt[9]);
int vn_init_assoc(vina_ctx* c, const double R[9], const double p[3], int refind, double sums28[28]);
int vn_init_tree_push(vina_ctx* c, const double R[9], const double p[3]);
// one frame of motion_init: re-deskew `n` retained raw points (host, time-sorted) with the backward pose table, world
// points / covariances, then cut_voxel into window slot `frame`
int vn_init_insert_frame(vina_ctx* c, const float* xyzt, int n, int n_skip, const vina_imu_pose* poses, int m,
                         const vina_state* x, int converged, int frame);
int vn_check_cuda(vina_ctx* c, cudaError_t e, const char* what);
// copy the device status word back (synchronises) and translate it
int vn_check_status(vina_ctx* c);
// wait for the sums of the last k_iekf launch (polls the sequence number in mapped memory)
int vn_iekf_wait(vina_ctx* c);
// enqueue the hand-over of the device iterate to the host / wait for it (polls the mapped sequence number)
int vn_iterate_publish(vina_ctx* c, cudaStream_t st);
int vn_iterate_wait(vina_ctx* c, cudaStream_t st);
// fill the launch descriptor of this context's sequence (R/p come from c->d_iekf)
void vn_iekf_fill_seq(vina_ctx* c, IekfSeq* q, bool debug);
// enqueue max_iter iterations of the IEKF against the sharded map, exchange and update on the device (vn_ctx.cu)
int vn_shard_iekf_enqueue(vina_ctx* c, int first, int count, int max_iter, int part);
int vn_mark_scan_read(vina_ctx* c);
int vn_deskew_var_init(vina_ctx* c, const vina_imu_pose* poses, int m, const double R_end[9], const double p_end[3]);
int vn_ba_collect_enqueue(vina_ctx* c);
int vn_ba_writeback_enqueue(vina_ctx* c);
int vn_ba_hess_enqueue(vina_ctx* c, const vina_pose* xs, int win);  // vina_ba_lidar_hessian in two halves
int vn_ba_hess_finish(vina_ctx* c, int win, double* Hess, double* JacT, double* residual);  // factor store -> leaves (octree.cpp:410-416), before margi  // tras_opt into the factor store (after recut, before margi)
// map update with the newest pose read from the device iterate (vn_ctx.cu)
int vn_map_insert_live(vina_ctx* c, int win_ord);
int vn_map_recut_live(vina_ctx* c, int win_count, const vina_pose* x_buf);
int vn_map_margi_live(vina_ctx* c, int win_count, const vina_pose* x_buf);
void odom_host_destroy(OdomHost* o);

/* ORACLE — TEST INFRASTRUCTURE ONLY (pinned against oracle/_ref, see vina_oracle.hpp).
 * Plain-C view of the CPU restatement so tests/ and bench.py's cpu_baseline
 * leg can drive it through ctypes. Matrices are column-major (Eigen default).
 */
#ifndef VINA_ORACLE_CAPI_H
#define VINA_ORACLE_CAPI_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct vo_config
{
  double voxel_size;
  double min_eigen_value;
  double plane_eigen_value_thre[4]; /* as in the yaml (NOT inverted); inverted on load like node.cpp:256-259 */
  double min_point[4];
  double dept_err, beam_err;
  double down_size;
  double ext_R[9]; /* Lid_rot_to_IMU, column-major */
  double ext_t[3];
  double cov_gyr, cov_acc, rdw_gyr, rdw_acc; /* Odometry.* (node.cpp:211-214) */
  int32_t max_layer;
  int32_t max_points;
  int32_t win_size;
  int32_t thread_num;
} vo_config;

typedef struct vo_state
{
  double t;
  double R[9];
  double p[3], v[3], bg[3], ba[3], g[3];
  double cov[225];
} vo_state;

/* one octree node as exported for parity (interior nodes included) */
typedef struct vo_node_record
{
  int64_t key[3];
  int32_t code; /* layer | path<<2, path = child index per level, 3 bits each */
  int32_t layer, octo_state, isexist, has_sw, is_plane, last_num, opt_state;
  int32_t N_add, N_fix, n_point_fix, n_win_points;
  int32_t N_local[16]; /* pcrs_local[mp[i]].N for frame ordinal i */
  double P_add[9], v_add[3], P_fix[9], v_fix[3];
  double eig_value[3], eig_vector[9];
  double center[3], normal[3], plane_var[36], radius;
  double cov_add[81];
  double voxel_center[3], quater_length;
} vo_node_record;

/* ---- stateless pieces ---- */
void vo_eig3(const double A[9], double vals[3], double vecs[9]);
void vo_inverse15(const double A[225], double out[225]);
void vo_exp(const double w[3], double R[9]);
void vo_exp_dt(const double w[3], double dt, double R[9]);
void vo_log(const double R[9], double w[3]);
void vo_var_init(int n, const float* xyz4, const double ext_R[9], const double ext_t[3], double dept_err,
                 double beam_err, double* pnt, double* var);
void vo_pvec_update(int n, const double* pnt, double* var, const double R[9], const double p[3],
                    const double cov[225], double* pwld);
void vo_voxel_keys(int n, const double* pw, double voxel_size, int64_t* keys);
int vo_down_sampling_voxel(int n, const float* xyz4_in, double voxel_size, float* xyz4_out);
/* decoder keep rule + pcl_handler (lidar_pointcloud_decoder.cpp:70; lidar_decoder.cpp:16-34): filter, stable sort by
 * time offset, cut at 0.11 s; blind2 = General.blind squared (node.cpp:210). Returns the point count, -1 if none. */
int vo_scan_prepare(int n, const float* xyz4_in, int point_filter_num, double blind2, float* xyz4_out);

/* ---- sync_packages and its buffers (src/sensor/sync.cpp:5-96; imu_handler subscribers.cpp:11-20; the tail of
 * pcl_handler lidar_decoder.cpp:36-43). vo_sync_next: 1 = true, 0 = false / nothing consumed, 2 = false / scan gone,
 * -6 = the reference exit(0)s, -3 = more IMU samples than cap. */
void* vo_sync_create(int point_notime);
void vo_sync_destroy(void* h);
void vo_sync_push_imu(void* h, const double imu7[7]);
void vo_sync_push_scan(void* h, double t_start, double t_last, int64_t tag);
int vo_sync_next(void* h, int64_t* tag, double* beg, double* end, double* imu7, int cap, int* m);

/* ---- per-sequence odometry (VINA_SLAM members of the per-scan loop) ---- */
void* vo_odom_create(const vo_config* cfg);
void vo_odom_destroy(void* h);
void vo_odom_set_state(void* h, const vo_state* s);
void vo_odom_get_state(void* h, vo_state* s);
void vo_odom_set_imu_anchor(void* h, double last_pcl_end_time, const double last_imu7[7], double scale_gravity);
void vo_odom_bootstrap(void* h, const float* xyz4, int n, const vo_state* x_known);
/* full scan: xyz4 = (x,y,z,curvature) in/out (deskewed on return); imu7 = m x (t,gx,gy,gz,ax,ay,az) */
int vo_odom_step(void* h, float* xyz4, int n, double pcl_beg_time, const double* imu7, int m, int iekf_on_full,
                 int max_iter);
/* start-up phase (VINA_SLAM::initialization, node.cpp:293-366 + local_mapping.cpp:362-388): cold_start switches a
 * freshly created odometry to it; init_scan: 0 = collecting, 1 = initialised (go on with vo_odom_step), -1 = failed */
void vo_odom_cold_start(void* h);
int vo_odom_init_scan(void* h, const float* xyz4, int n, double pcl_beg_time, const double* imu7, int m);
void vo_odom_stage_times(void* h, double t[4]); /* odom(deskew+var+iekf+pvec_update), insert, recut, margi */
int vo_odom_last_iters(void* h);
int vo_odom_last_down(void* h, float* xyz4, int cap); /* the down-sampled cloud of the last step */

/* stage-wise entries */
/* IMUEKF::motion_blur in one piece: propagation + deskew (xyz4 in/out) */
int vo_odom_motion_blur(void* h, float* xyz4, int n, double pcl_beg_time, double pcl_end_time, const double* imu7, int m);
/* match() (voxel_map.cpp:241-266) for n world points with their world covariances (n x 9 column-major) */
int vo_odom_match(void* h, int n, const double* wld, const double* var, uint8_t* flags, double* sigma, double* centers);
int vo_odom_propagate(void* h, double pcl_beg_time, double pcl_end_time, const double* imu7, int m);
int vo_odom_imu_poses(void* h, double* poses22, int cap); /* t,R(9),p,v,w,a per pose */
void vo_odom_deskew(void* h, float* xyz4, int n);
void vo_odom_set_dump(void* h, int on);
/* IEKF on caller-provided pointVar arrays (pnt n x 3, var n x 9 column-major); returns "not degenerate" */
int vo_odom_iekf(void* h, int n, const double* pnt, const double* var, int max_iter);
int vo_odom_iter_dump(void* h, int it, double HTH[36], double HTz[6], double nnt[9], int32_t* match_num,
                      int64_t* keys, int32_t* codes, uint8_t* flags, double* sigma, double R[9], double p[3]);
/* map update on caller-provided world-var pointVar + pose already in x_curr */
void vo_odom_map_update(void* h, int n, const double* pnt, const double* var);
int64_t vo_odom_map_count(void* h, int64_t* n_roots, int64_t* n_slide);
int64_t vo_odom_map_export(void* h, vo_node_record* out, int64_t cap);
int vo_odom_window(void* h, int* win_count, int* mp, int cap);
/* distance travelled / pruning of the map behind the vehicle (local_mapping.cpp:317-341, 509-519): the journey
 * bookkeeping runs inside vo_odom_step; vo_odom_idle is the `release_flag` branch of the idle path with the
 * reference's 700 m as a parameter. Returns the number of root voxels erased. */
void vo_odom_journey(void* h, double* jour, int* release_flag);
int vo_odom_idle(void* h, int horizon, int* nodes_freed);

/* ---- BA probe (the data-parallel part): LidarFactor::acc_evaluate2 and
 * evaluate_only_residual (factors.cpp:22-158) on a copy of the LiDAR factors and window poses taken between
 * multi_recut and multi_margi of the last map update with a full window (where damping_iter consumes them,
 * local_mapping.cpp:492-497). poses12 = win x (R 9 column-major, p 3). Hess is (6 win)^2 column-major. */
/* the whole sliding-window BA (LI_BA_Optimizer::damping_iter, LiDAR + IMU pre-integration factors) inside
 * vo_odom_step, like local_mapping.cpp:492-497 with if_BA: 1. It runs once every pair of consecutive window frames
 * has an IMU factor (frames inserted by vo_odom_bootstrap have none). imu_coef <= 0 keeps LocalBA.imu_coef = 1e-4. */
double vo_ba_imu_evaluate(const vo_config* cfg, const double bg[3], const double ba[3], const double* imu7, int m,
                          double scale_gravity, const vo_state* s1, const vo_state* s2, double* jtj, double* gg);
void vo_odom_set_ba(void* h, int on, double imu_coef);
void vo_odom_ba_stats(void* h, int* runs, int* last_iters);
void vo_odom_ba_probe(void* h, int on);
int vo_odom_ba_count(void* h);
int vo_odom_ba_poses(void* h, double* poses12, int cap);
int vo_odom_ba_hess(void* h, const double* poses12, int win, double* Hess, double* JacT, double* residual);
int vo_odom_ba_residual(void* h, const double* poses12, int win, double* residual, double* lam0, int cap);

#ifdef __cplusplus
}
#endif
#endif

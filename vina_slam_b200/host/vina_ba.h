// Sliding-window BA, host side (vina_ba.cpp): IMU pre-integration factors and the LM loop of
// LI_BA_Optimizer::damping_iter around the device LiDAR factor (csrc/ba_kernels.cu).
#pragma once
#include <deque>
#include <vector>
#include "vina_b200.h"

struct vina_ctx;
struct ImuPre;  // IMU_PRE (src/estimation/imu_preintegration.cpp)

// new IMU_PRE(bg, ba) + push_imu(imus): imus = the scan's IMU batch with its ends re-stamped to the scan boundaries
ImuPre* ba_imu_factor_new(const double* bg, const double* ba, const std::deque<vina_imu>& imus, double scale_gravity,
                          const vina_config& cfg);
void ba_imu_factor_delete(ImuPre* f);
int ba_damping_iter(vina_ctx* ctx, std::vector<vina_state>& xs, std::deque<ImuPre*>& imus_factor, double imu_coef,
                    int* iters_out);
// the same loop with the window's gravity vector as an additional unknown (LI_BA_OptimizerGravity::damping_iter,
// optimizers.cpp:746-826; used by the start-up phase): see vina_ba.cpp
int ba_damping_iter_ex(vina_ctx* ctx, std::vector<vina_state>& xs, std::deque<ImuPre*>& imus_factor, double imu_coef,
                       int* iters_out, bool gravity, int max_iter, double* resis);

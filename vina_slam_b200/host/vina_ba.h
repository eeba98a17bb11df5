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

// carrot.cuh — batched carrot selection / request construction (the plugin's front half).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "neompc.h"

namespace neompc {

struct CarrotConst {
  const double* plan;        // device [L][3]: x, y, yaw
  unsigned L;
  double max_transform_dist; // max(costmap size) * resolution / 2      (cpp:78-79)
  double la_min, la_max, la_close;
  float control_interval;    // 1 / controller_frequency                (cpp:246)
  const uint8_t* cells;      // device costmap or nullptr
  int W, H;
  double origin_x, origin_y, resolution;
  const uint8_t* raw_table;  // device [256]: costmap byte -> nav2 raw cost
  int fp_n;
  float fp_x[NEOMPC_MAX_FOOTPRINT_VERTICES], fp_y[NEOMPC_MAX_FOOTPRINT_VERTICES];
};

cudaError_t launch_build_requests(const CarrotConst& c, const neompc_robot_tick* d_ticks, unsigned n, uint32_t first_id,
                                  neompc_request* d_reqs, neompc_carrot_info* d_info, cudaStream_t stream);

}  // namespace neompc

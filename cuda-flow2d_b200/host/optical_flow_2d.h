// optical_flow_2d.h -- source-compatible replacement of the reference's solver entry
// (src/optical_flow/optical_flow_2d.h:43-71, optical_flow_base_2d.h:29-52): the same three methods
// and the `silent` member, implemented on top of the C ABI in include/flow2d.h.
#pragma once
#include "data2d.h"
#include "data_structs.h"
#include "operation_parameters.h"

struct flow2d_handle;

class OpticalFlow2D {
 public:
  OpticalFlow2D() = default;
  OpticalFlow2D(const OpticalFlow2D&) = delete;
  OpticalFlow2D& operator=(const OpticalFlow2D&) = delete;
  ~OpticalFlow2D();

  const char* GetName() const { return "Optical Flow 2D"; }
  // optical_flow_2d.cpp:48-57: allocates the device side for frames of data_size.width x height.
  bool Initialize(const DataSize3& data_size, DataConstancy data_constancy = DataConstancy::Grey);
  // optical_flow_2d.cpp:142-569.  Required keys (all by pointer, non-owning): warp_levels_count:size_t,
  // warp_scale_factor:float, outer_iterations_count:size_t, inner_iterations_count:size_t,
  // equation_alpha:float, equation_smoothness:float, equation_data:float, median_radius:size_t,
  // gaussian_sigma:float.  A missing key prints a message and returns, like upstream.
  void ComputeFlow(Data2D& frame_0, Data2D& frame_1, Data2D& flow_u, Data2D& flow_v, OperationParameters& params);
  void Destroy();
  size_t GetMaxWarpLevel(size_t width, size_t height, float scale_factor) const;

  bool silent = false;
  int device = 0;              // CUDA device of the handle (the reference always uses device 0)
  float last_gpu_time_ms = 0;  // what upstream prints as "Total GPU computation time"
  // 0 after a successful ComputeFlow, else the flow2d_status of the failure (upstream's ComputeFlow is void and
  // swallows every error: cuda_utils.h:33-51); lets a caller skip its writers instead of saving garbage
  int last_status() const { return last_status_; }

 private:
  flow2d_handle* handle_ = nullptr;
  DataSize3 size_{0, 0, 0};
  int last_status_ = 0;
};

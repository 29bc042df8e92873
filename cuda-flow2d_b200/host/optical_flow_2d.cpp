#include "optical_flow_2d.h"

#include <cstdio>

#include "flow2d.h"

OpticalFlow2D::~OpticalFlow2D() { Destroy(); }

bool OpticalFlow2D::Initialize(const DataSize3& data_size, DataConstancy data_constancy) {
  Destroy();
  if (data_constancy == DataConstancy::LogDerivatives) {
    std::printf("Error: '%s': the LogDerivatives data term is not supported (its reference kernel reads wrong halos).\n", GetName());
    return false;
  }
  const int constancy = data_constancy == DataConstancy::Gradient ? FLOW2D_GRADIENT : FLOW2D_GREY;
  const int rc = flow2d_create(&handle_, device, data_size.width, data_size.height, constancy);
  if (rc != FLOW2D_OK) {
    std::printf("Error: '%s' initialisation failed (flow2d_create: %d).\n", GetName(), rc);
    handle_ = nullptr;
    return false;
  }
  size_ = data_size;
  size_.pitch = flow2d_pitch_elems(handle_) * sizeof(float);
  return true;
}

void OpticalFlow2D::Destroy() {
  if (handle_) flow2d_destroy(handle_);
  handle_ = nullptr;
}

size_t OpticalFlow2D::GetMaxWarpLevel(size_t width, size_t height, float scale_factor) const {
  return flow2d_max_warp_level(width, height, scale_factor);
}

namespace {
template <typename T>
bool get(const OperationParameters& params, const char* owner, const char* key, T* out) {
  void* p = params.GetValuePtr(key);
  if (!p) {
    std::printf("Operation: '%s'. Missing parameter '%s'.\n", owner, key);
    return false;
  }
  *out = *static_cast<T*>(p);
  return true;
}
}  // namespace

void OpticalFlow2D::ComputeFlow(Data2D& frame_0, Data2D& frame_1, Data2D& flow_u, Data2D& flow_v,
                                OperationParameters& params) {
  last_status_ = FLOW2D_ERR_INVALID_ARGUMENT;  // until the solve has succeeded
  if (!handle_) {
    std::printf("Error: '%s' was not initialized.\n", GetName());
    return;
  }
  flow2d_params p;
  flow2d_default_params(&p);
  if (!get(params, GetName(), "warp_levels_count", &p.warp_levels_count) ||
      !get(params, GetName(), "warp_scale_factor", &p.warp_scale_factor) ||
      !get(params, GetName(), "outer_iterations_count", &p.outer_iterations_count) ||
      !get(params, GetName(), "inner_iterations_count", &p.inner_iterations_count) ||
      !get(params, GetName(), "equation_alpha", &p.equation_alpha) ||
      !get(params, GetName(), "equation_smoothness", &p.equation_smoothness) ||
      !get(params, GetName(), "equation_data", &p.equation_data) ||
      !get(params, GetName(), "median_radius", &p.median_radius) ||
      !get(params, GetName(), "gaussian_sigma", &p.gaussian_sigma))
    return;
  // optional, beyond the reference's nine keys: "report_residuals" (bool*) prints the per-level residual norms
  if (void* opt = params.GetValuePtr("report_residuals")) p.report_residuals = *static_cast<bool*>(opt) ? 1 : 0;
  // ... and the opt-in solver extensions of include/flow2d.h (absent keys = the reference's behaviour)
  if (void* opt = params.GetValuePtr("solver_scheme")) p.scheme = *static_cast<int*>(opt);
  if (void* opt = params.GetValuePtr("solver_omega")) p.omega = *static_cast<float*>(opt);
  if (void* opt = params.GetValuePtr("data_term")) p.data_term = *static_cast<int*>(opt);
  if (void* opt = params.GetValuePtr("data_gamma")) p.gamma = *static_cast<float*>(opt);
  if (void* opt = params.GetValuePtr("residual_tolerance")) p.residual_tolerance = *static_cast<float*>(opt);
  if (void* opt = params.GetValuePtr("residual_check_every")) p.residual_check_every = *static_cast<int*>(opt);
  if (void* opt = params.GetValuePtr("cascaded_restriction")) p.cascaded_restriction = *static_cast<bool*>(opt) ? 1 : 0;
  // the reference times every level's solve and prints it unless `silent` (cuda_operation_solve_2d.cpp:302-311)
  if (void* opt = params.GetValuePtr("report_level_times")) p.report_level_times = *static_cast<bool*>(opt) ? 1 : 0;
  for (Data2D* d : {&frame_0, &frame_1, &flow_u, &flow_v})
    if (d->Width() != size_.width || d->Height() != size_.height || !d->DataPtr()) {
      std::printf("Error: '%s': frames and flow fields must be %zux%zu.\n", GetName(), size_.width, size_.height);
      return;
    }
  std::printf("\nStarting optical flow computation...\n");
  const int rc = flow2d_compute(handle_, frame_0.DataPtr(), frame_1.DataPtr(), flow_u.DataPtr(), flow_v.DataPtr(), &p);
  last_status_ = rc;
  if (rc != FLOW2D_OK) {
    std::fprintf(stderr, "Error: '%s': %s (%d)\n", GetName(), flow2d_last_error(handle_), rc);
    return;
  }
  long long launches = 0;
  int levels = 0;
  flow2d_last_stats(handle_, &launches, &levels, &last_gpu_time_ms);
  if (!silent) std::printf("Levels: %d, kernel launches: %lld\n", levels, launches);
  if (p.report_level_times) {
    float lv[FLOW2D_MAX_LEVELS], sv[FLOW2D_MAX_LEVELS];
    int n = 0;
    if (flow2d_level_times(handle_, lv, sv, FLOW2D_MAX_LEVELS, &n) == FLOW2D_OK)
      for (int i = 0; i < n; i++) std::printf("Level %3d (coarsest first): %8.4fs, solve %8.4fs\n", i, lv[i] / 1000., sv[i] / 1000.);
  }
  if (p.residual_tolerance > 0.f) {
    int it[FLOW2D_MAX_LEVELS], n = 0;
    if (flow2d_level_outer_iterations(handle_, it, FLOW2D_MAX_LEVELS, &n) == FLOW2D_OK) {
      long long total = 0;
      for (int i = 0; i < n; i++) total += it[i];
      std::printf("Outer iterations run: %lld of %lld (convergence test at %g)\n", total, (long long)n * (long long)p.outer_iterations_count,
                  (double)p.residual_tolerance);
    }
  }
  if (p.report_residuals) {
    double ru[FLOW2D_MAX_LEVELS], rv[FLOW2D_MAX_LEVELS];
    int n = 0;
    if (flow2d_level_residuals(handle_, ru, rv, FLOW2D_MAX_LEVELS, &n) == FLOW2D_OK)
      for (int i = 0; i < n; i++) std::printf("Level %3d (coarsest first): residual rms u %.6e  v %.6e\n", i, ru[i], rv[i]);
  }
  std::printf("Total GPU computation time: % 4.4fs\n", last_gpu_time_ms / 1000.);
}

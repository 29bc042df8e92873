// cuda_operations.cpp -- see cuda_operations.h.  Every Execute forwards to one flow2d_stage_* call of the C ABI.
#include "cuda_operations.h"

#include <cstdio>

#include "flow2d.h"

namespace {
// GET_PARAM_OR_RETURN of the reference (src/utils/common_utils.h): by-name lookup, message and early return when absent
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
float* ptr(CUdeviceptr p) { return reinterpret_cast<float*>(static_cast<size_t>(p)); }
}  // namespace

#define GET_OR_RETURN(type, var, key) \
  type var;                           \
  if (!get(params, GetName(), key, &var)) return

size_t CudaOperationBase::ContainerPitchBytes(size_t width) { return (width + 127) / 128 * 128 * sizeof(float); }

bool CudaOperationBase::Initialize(const OperationParameters* params) {
  if (handle_) return true;
  DataSize3 container_size;
  if (!params || !get(*params, GetName(), "container_size", &container_size)) return false;
  last_status_ = flow2d_create(&handle_, device, container_size.width, container_size.height, constancy_);
  if (last_status_ != FLOW2D_OK) {
    std::printf("Operation: '%s'. Initialization failed (%d).\n", GetName(), last_status_);
    handle_ = nullptr;
  }
  return handle_ != nullptr;
}

void CudaOperationBase::Destroy() {
  if (handle_) flow2d_destroy(handle_);
  handle_ = nullptr;
}

CudaOperationBase::~CudaOperationBase() { Destroy(); }

bool CudaOperationBase::CheckSize(const DataSize3& s) {
  if (s.pitch != 0 && s.pitch != flow2d_pitch_elems(handle_) * sizeof(float)) {
    std::printf("Operation '%s': Error. data_size.pitch %zu differs from the container pitch %zu.\n", GetName(), s.pitch,
                flow2d_pitch_elems(handle_) * sizeof(float));
    last_status_ = FLOW2D_ERR_INVALID_ARGUMENT;
    return false;
  }
  return true;
}

#define FINISH(call)                                                                                   \
  do {                                                                                                 \
    last_status_ = (call);                                                                             \
    if (last_status_ == FLOW2D_OK) last_status_ = flow2d_synchronize(handle_);                         \
    if (last_status_ != FLOW2D_OK)                                                                     \
      std::printf("Operation '%s': Error. %s (%d)\n", GetName(), flow2d_last_error(handle_), last_status_); \
  } while (0)

void CudaOperationConvolution2D::Execute(OperationParameters& params) {
  if (!IsInitialized()) return;
  GET_OR_RETURN(CUdeviceptr, dev_input, "dev_input");
  GET_OR_RETURN(CUdeviceptr, dev_output, "dev_output");
  GET_OR_RETURN(CUdeviceptr, dev_temp, "dev_temp");
  GET_OR_RETURN(DataSize3, data_size, "data_size");
  GET_OR_RETURN(float, gaussian_sigma, "gaussian_sigma");
  (void)dev_temp;
  if (dev_input == dev_output) {
    std::printf("Operation '%s': Error. Input buffer cannot serve as output buffer.", GetName());
    return;
  }
  if (!CheckSize(data_size)) return;
  FINISH(flow2d_stage_blur(handle_, ptr(dev_input), ptr(dev_output), data_size.width, data_size.height, gaussian_sigma));
}

void CudaOperationResample2D::Execute(OperationParameters& params) {
  if (!IsInitialized()) return;
  GET_OR_RETURN(CUdeviceptr, dev_input, "dev_input");
  GET_OR_RETURN(CUdeviceptr, dev_output, "dev_output");
  GET_OR_RETURN(CUdeviceptr, dev_temp, "dev_temp");
  GET_OR_RETURN(DataSize3, data_size, "data_size");
  GET_OR_RETURN(DataSize3, resample_size, "resample_size");
  (void)dev_temp;
  if (dev_input == dev_output) {
    std::printf("Operation '%s': Error. Input buffer cannot serve as output buffer.", GetName());
    return;
  }
  if (!CheckSize(data_size)) return;
  FINISH(flow2d_stage_resample(handle_, ptr(dev_input), data_size.width, data_size.height, ptr(dev_output), resample_size.width,
                               resample_size.height));
}

void CudaOperationRegistration2D::Execute(OperationParameters& params) {
  if (!IsInitialized()) return;
  GET_OR_RETURN(CUdeviceptr, dev_frame_0, "dev_frame_0");
  GET_OR_RETURN(CUdeviceptr, dev_frame_1, "dev_frame_1");
  GET_OR_RETURN(CUdeviceptr, dev_flow_u, "dev_flow_u");
  GET_OR_RETURN(CUdeviceptr, dev_flow_v, "dev_flow_v");
  GET_OR_RETURN(CUdeviceptr, dev_output, "dev_output");
  GET_OR_RETURN(float, hx, "hx");
  GET_OR_RETURN(float, hy, "hy");
  GET_OR_RETURN(DataSize3, data_size, "data_size");
  if (dev_frame_1 == dev_output) {
    std::printf("Operation '%s': Error. Input buffer cannot serve as output buffer.", GetName());
    return;
  }
  if (!CheckSize(data_size)) return;
  FINISH(flow2d_stage_warp(handle_, ptr(dev_frame_0), ptr(dev_frame_1), ptr(dev_flow_u), ptr(dev_flow_v), ptr(dev_output),
                           data_size.width, data_size.height, hx, hy));
}

bool CudaOperationSolve2D::Initialize(const OperationParameters* params) {
  if (params)
    if (void* c = params->GetValuePtr("data_constancy"))
      constancy_ = *static_cast<DataConstancy*>(c) == DataConstancy::Gradient ? FLOW2D_GRADIENT : FLOW2D_GREY;
  return CudaOperationBase::Initialize(params);
}

void CudaOperationSolve2D::Execute(OperationParameters& params) {
  if (!IsInitialized()) return;
  GET_OR_RETURN(CUdeviceptr, dev_frame_0, "dev_frame_0");
  GET_OR_RETURN(CUdeviceptr, dev_frame_1, "dev_frame_1");
  GET_OR_RETURN(CUdeviceptr, dev_flow_u, "dev_flow_u");
  GET_OR_RETURN(CUdeviceptr, dev_flow_v, "dev_flow_v");
  GET_OR_RETURN(CUdeviceptr, dev_phi, "dev_phi");
  GET_OR_RETURN(CUdeviceptr, dev_ksi, "dev_ksi");
  // by pointer upstream (swapped in place there); here the result is written to *dev_flow_du / *dev_flow_dv directly
  GET_OR_RETURN(CUdeviceptr, dev_flow_du, "dev_flow_du");
  GET_OR_RETURN(CUdeviceptr, dev_flow_dv, "dev_flow_dv");
  GET_OR_RETURN(CUdeviceptr, dev_temp_du, "dev_temp_du");
  GET_OR_RETURN(CUdeviceptr, dev_temp_dv, "dev_temp_dv");
  (void)dev_temp_du; (void)dev_temp_dv;
  flow2d_params p;
  flow2d_default_params(&p);
  if (!get(params, GetName(), "outer_iterations_count", &p.outer_iterations_count) ||
      !get(params, GetName(), "inner_iterations_count", &p.inner_iterations_count) ||
      !get(params, GetName(), "equation_alpha", &p.equation_alpha) ||
      !get(params, GetName(), "equation_smoothness", &p.equation_smoothness) ||
      !get(params, GetName(), "equation_data", &p.equation_data))
    return;
  GET_OR_RETURN(float, hx, "hx");
  GET_OR_RETURN(float, hy, "hy");
  GET_OR_RETURN(DataSize3, data_size, "data_size");
  if (!CheckSize(data_size)) return;
  FINISH(flow2d_stage_solve(handle_, ptr(dev_frame_0), ptr(dev_frame_1), ptr(dev_flow_u), ptr(dev_flow_v), ptr(dev_flow_du),
                            ptr(dev_flow_dv), ptr(dev_phi), ptr(dev_ksi), data_size.width, data_size.height, hx, hy, &p));
}

void CudaOperationAdd2D::Execute(OperationParameters& params) {
  if (!IsInitialized()) return;
  GET_OR_RETURN(CUdeviceptr, operand_0, "operand_0");
  GET_OR_RETURN(CUdeviceptr, operand_1, "operand_1");
  GET_OR_RETURN(DataSize3, data_size, "data_size");
  if (!CheckSize(data_size)) return;
  FINISH(flow2d_stage_add(handle_, ptr(operand_0), ptr(operand_1), data_size.width, data_size.height));
}

void CudaOperationMedian2D::Execute(OperationParameters& params) {
  if (!IsInitialized()) return;
  GET_OR_RETURN(CUdeviceptr, dev_input, "dev_input");
  GET_OR_RETURN(CUdeviceptr, dev_output, "dev_output");
  GET_OR_RETURN(DataSize3, data_size, "data_size");
  GET_OR_RETURN(size_t, radius, "radius");
  if (dev_input == dev_output) {
    std::printf("Operation '%s': Error. Input buffer cannot serve as output buffer.", GetName());
    return;
  }
  if (!CheckSize(data_size)) return;
  FINISH(flow2d_stage_median(handle_, ptr(dev_input), ptr(dev_output), data_size.width, data_size.height, radius));
}

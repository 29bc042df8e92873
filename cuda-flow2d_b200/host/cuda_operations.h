// cuda_operations.h -- source-compatible operator classes of the reference's cuda_operations/ layer on top of the C ABI
// (include/flow2d.h: flow2d_stage_*).  Same class names, the same three methods
//     bool Initialize(const OperationParameters* params);  void Execute(OperationParameters& params);  void Destroy();
// (src/cuda_operations/cuda_operation_base.h:44-51) and the same by-name parameters, so that code written against the
// reference's operators -- and per-operator parity tests -- compile unchanged:
//   Initialize   "container_size": DataSize3 (src/optical_flow/optical_flow_2d.cpp:67-69), "data_constancy": DataConstancy
//                (solve only, optional).  DataSize3.pitch is in BYTES like upstream; 0 = take the handle's pitch.
//   Convolution  dev_input, dev_output, dev_temp: CUdeviceptr, data_size, gaussian_sigma   (cuda_operation_convolution_2d.cpp:144-148)
//   Resample     dev_input, dev_output, dev_temp, data_size, resample_size                 (cuda_operation_resample_2d.cpp:88-92)
//   Registration dev_frame_0, dev_frame_1, dev_flow_u, dev_flow_v, dev_output, hx, hy, data_size (cuda_operation_registration_2d.cpp:86-98)
//   Solve        dev_frame_0, dev_frame_1, dev_flow_u, dev_flow_v, dev_phi, dev_ksi by value; dev_flow_du, dev_flow_dv,
//                dev_temp_du, dev_temp_dv by pointer (the result is in *dev_flow_du / *dev_flow_dv, as upstream after its
//                swaps); outer_iterations_count, inner_iterations_count: size_t; equation_alpha, equation_smoothness,
//                equation_data, hx, hy: float; data_size; data_constancy                    (cuda_operation_solve_2d.cpp:119-179)
//   Add          operand_0, operand_1, data_size                                           (cuda_operation_add_2d.cpp:85-87)
//   Median       dev_input, dev_output, data_size, radius: size_t                          (cuda_operation_median_2d.cpp:86-92)
// Behaviour kept: a missing key prints "Operation: '<name>'. Missing parameter '<key>'." and returns; in-place use is
// refused with upstream's message for convolution / resample / registration / median; Execute before Initialize is a
// no-op.  Differences: device buffers are CONTAINERS of the handle's pitch (ContainerPitchBytes(); the reference's
// cuMemAllocPitch gives the same 512-byte granularity), "dev_temp" is accepted and not needed, Execute returns after
// the operator has finished (the reference leaves the kernel on the NULL stream), and last_status() reports the C-ABI
// status that upstream's void Execute swallows.
#pragma once
#include <cstddef>

#include "data_structs.h"
#include "operation_parameters.h"

struct flow2d_handle;
#if !defined(__cuda_cuda_h__) && !defined(CUDA_VERSION)
typedef unsigned long long CUdeviceptr;
#endif

class CudaOperationBase {
 public:
  const char* GetName() const { return name_; }
  virtual bool Initialize(const OperationParameters* params = nullptr);
  virtual void Execute(OperationParameters& params) = 0;
  virtual void Destroy();
  virtual ~CudaOperationBase();
  int last_status() const { return last_status_; }
  // pitch in bytes of the containers an operator initialised for `width` works on
  static size_t ContainerPitchBytes(size_t width);
  int device = 0;

 protected:
  explicit CudaOperationBase(const char* name) : name_(name) {}
  bool IsInitialized() const { return handle_ != nullptr; }
  bool CheckSize(const DataSize3& data_size);
  const char* name_;
  flow2d_handle* handle_ = nullptr;
  int constancy_ = 0;
  int last_status_ = 0;
};

class CudaOperationConvolution2D : public CudaOperationBase {
 public:
  CudaOperationConvolution2D() : CudaOperationBase("Convolution 2D") {}
  void Execute(OperationParameters& params) override;
};
class CudaOperationResample2D : public CudaOperationBase {
 public:
  CudaOperationResample2D() : CudaOperationBase("Resampling 2D") {}
  void Execute(OperationParameters& params) override;
};
class CudaOperationRegistration2D : public CudaOperationBase {
 public:
  CudaOperationRegistration2D() : CudaOperationBase("Registration 2D") {}
  void Execute(OperationParameters& params) override;
};
class CudaOperationSolve2D : public CudaOperationBase {
 public:
  CudaOperationSolve2D() : CudaOperationBase("Solve 2D") {}
  bool Initialize(const OperationParameters* params = nullptr) override;
  void Execute(OperationParameters& params) override;
  bool silent = true;
};
class CudaOperationAdd2D : public CudaOperationBase {
 public:
  CudaOperationAdd2D() : CudaOperationBase("Add 2D") {}
  void Execute(OperationParameters& params) override;
};
class CudaOperationMedian2D : public CudaOperationBase {
 public:
  CudaOperationMedian2D() : CudaOperationBase("Median 2D") {}
  void Execute(OperationParameters& params) override;
};

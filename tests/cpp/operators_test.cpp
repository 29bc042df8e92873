// operators_test.cpp -- drives the source-compatible operator classes (cuda-flow2d_b200/host/cuda_operations.h) exactly
// the way OpticalFlow2D::ComputeFlow drives the reference's (src/optical_flow/optical_flow_2d.cpp:218-449): named
// parameters, CUdeviceptr containers, Initialize / Execute / Destroy.  Reads dense float32 images, writes dense results;
// tests/test_operators_gpu.py compares them with the CPU oracle.
//   operators_test <W> <H> <dir>     inputs <dir>/f0.raw f1.raw u.raw v.raw; outputs blur, resample, warp, du, dv, phi, ksi, add, median
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "cuda_operations.h"

static size_t W, H, pitch;  // pitch in bytes

static CUdeviceptr container() {
  void* p = nullptr;
  if (cudaMalloc(&p, pitch * H) != cudaSuccess) { std::printf("cudaMalloc failed\n"); std::exit(2); }
  cudaMemset(p, 0, pitch * H);
  return (CUdeviceptr)(size_t)p;
}
static CUdeviceptr upload(const std::string& path) {
  std::vector<float> host(W * H);
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f || std::fread(host.data(), sizeof(float), W * H, f) != W * H) { std::printf("cannot read %s\n", path.c_str()); std::exit(2); }
  std::fclose(f);
  CUdeviceptr d = container();
  cudaMemcpy2D((void*)(size_t)d, pitch, host.data(), W * sizeof(float), W * sizeof(float), H, cudaMemcpyHostToDevice);
  return d;
}
static void download(CUdeviceptr d, size_t w, size_t h, const std::string& path) {
  std::vector<float> host(w * h);
  cudaMemcpy2D(host.data(), w * sizeof(float), (void*)(size_t)d, pitch, w * sizeof(float), h, cudaMemcpyDeviceToHost);
  FILE* f = std::fopen(path.c_str(), "wb");
  std::fwrite(host.data(), sizeof(float), w * h, f);
  std::fclose(f);
}

int main(int argc, char** argv) {
  if (argc != 4) return 1;
  W = std::atoi(argv[1]); H = std::atoi(argv[2]);
  const std::string dir = std::string(argv[3]) + "/";
  pitch = CudaOperationBase::ContainerPitchBytes(W);
  DataSize3 container_size = {W, H, pitch};
  DataConstancy data_constancy = DataConstancy::Grey;
  OperationParameters init;
  init.PushValuePtr("container_size", &container_size);
  init.PushValuePtr("data_constancy", &data_constancy);

  CudaOperationConvolution2D conv;
  CudaOperationResample2D resample;
  CudaOperationRegistration2D registration;
  CudaOperationSolve2D solve;
  CudaOperationAdd2D add;
  CudaOperationMedian2D median;
  CudaOperationBase* ops[6] = {&conv, &resample, &registration, &solve, &add, &median};
  {  // Execute before Initialize is a no-op, a missing key is a message
    OperationParameters none;
    conv.Execute(none);
  }
  for (CudaOperationBase* op : ops)
    if (!op->Initialize(&init)) { std::printf("Initialize of '%s' failed\n", op->GetName()); return 3; }

  CUdeviceptr f0 = upload(dir + "f0.raw"), f1 = upload(dir + "f1.raw"), u = upload(dir + "u.raw"), v = upload(dir + "v.raw");
  CUdeviceptr out = container(), temp = container(), warped = container();
  CUdeviceptr du = container(), dv = container(), tdu = container(), tdv = container(), phi = container(), ksi = container();
  DataSize3 data_size = {W, H, pitch};
  float hx = 1.25f, hy = 1.5f;

  {  // optical_flow_2d.cpp:225-233
    float gaussian_sigma = 1.5f;
    OperationParameters p;
    p.PushValuePtr("dev_input", &f0); p.PushValuePtr("dev_output", &out); p.PushValuePtr("dev_temp", &temp);
    p.PushValuePtr("data_size", &data_size); p.PushValuePtr("gaussian_sigma", &gaussian_sigma);
    conv.Execute(p);
    download(out, W, H, dir + "blur.raw");
    OperationParameters bad;  // in-place is refused, nothing is written
    bad.PushValuePtr("dev_input", &f0); bad.PushValuePtr("dev_output", &f0); bad.PushValuePtr("dev_temp", &temp);
    bad.PushValuePtr("data_size", &data_size); bad.PushValuePtr("gaussian_sigma", &gaussian_sigma);
    conv.Execute(bad);
    OperationParameters missing;
    missing.PushValuePtr("dev_input", &f0);
    conv.Execute(missing);
  }
  DataSize3 resample_size = {(W * 7 + 9) / 10, (H * 7 + 9) / 10, pitch};
  {  // optical_flow_2d.cpp:284-291
    OperationParameters p;
    p.PushValuePtr("dev_input", &f0); p.PushValuePtr("dev_output", &out); p.PushValuePtr("dev_temp", &temp);
    p.PushValuePtr("data_size", &data_size); p.PushValuePtr("resample_size", &resample_size);
    resample.Execute(p);
    download(out, resample_size.width, resample_size.height, dir + "resample.raw");
  }
  {  // optical_flow_2d.cpp:347-357
    OperationParameters p;
    p.PushValuePtr("dev_frame_0", &f0); p.PushValuePtr("dev_frame_1", &f1); p.PushValuePtr("dev_flow_u", &u);
    p.PushValuePtr("dev_flow_v", &v); p.PushValuePtr("dev_output", &warped); p.PushValuePtr("data_size", &data_size);
    p.PushValuePtr("hx", &hx); p.PushValuePtr("hy", &hy);
    registration.Execute(p);
    download(warped, W, H, dir + "warp.raw");
  }
  {  // optical_flow_2d.cpp:369-392
    size_t outer_iterations_count = 3, inner_iterations_count = 5;
    float equation_alpha = 20.f, equation_smoothness = 0.001f, equation_data = 0.001f;
    OperationParameters p;
    p.PushValuePtr("dev_frame_0", &f0); p.PushValuePtr("dev_frame_1", &warped); p.PushValuePtr("dev_flow_u", &u);
    p.PushValuePtr("dev_flow_v", &v); p.PushValuePtr("dev_flow_du", &du); p.PushValuePtr("dev_flow_dv", &dv);
    p.PushValuePtr("dev_phi", &phi); p.PushValuePtr("dev_ksi", &ksi); p.PushValuePtr("dev_temp_du", &tdu);
    p.PushValuePtr("dev_temp_dv", &tdv); p.PushValuePtr("outer_iterations_count", &outer_iterations_count);
    p.PushValuePtr("inner_iterations_count", &inner_iterations_count); p.PushValuePtr("equation_alpha", &equation_alpha);
    p.PushValuePtr("equation_smoothness", &equation_smoothness); p.PushValuePtr("equation_data", &equation_data);
    p.PushValuePtr("hx", &hx); p.PushValuePtr("hy", &hy); p.PushValuePtr("data_size", &data_size);
    p.PushValuePtr("data_constancy", &data_constancy);
    solve.Execute(p);
    download(du, W, H, dir + "du.raw"); download(dv, W, H, dir + "dv.raw");
    download(phi, W, H, dir + "phi.raw"); download(ksi, W, H, dir + "ksi.raw");
  }
  {  // optical_flow_2d.cpp:411-421: u += du
    OperationParameters p;
    p.PushValuePtr("operand_0", &u); p.PushValuePtr("operand_1", &du); p.PushValuePtr("data_size", &data_size);
    add.Execute(p);
    download(u, W, H, dir + "add.raw");
  }
  {  // optical_flow_2d.cpp:431-441
    size_t radius = 5;
    OperationParameters p;
    p.PushValuePtr("dev_input", &u); p.PushValuePtr("dev_output", &out); p.PushValuePtr("data_size", &data_size);
    p.PushValuePtr("radius", &radius);
    median.Execute(p);
    download(out, W, H, dir + "median.raw");
  }
  int rc = 0;
  for (CudaOperationBase* op : ops) {
    if (op->last_status() != 0) { std::printf("'%s' ended with status %d\n", op->GetName(), op->last_status()); rc = 4; }
    op->Destroy();
  }
  std::printf(rc == 0 ? "\nOPERATORS_DONE\n" : "\nOPERATORS_FAILED\n");
  return rc;
}

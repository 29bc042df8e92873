/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A small driver around the UNMODIFIED reference classes (compiled from /root/reference by
 * oracle/Makefile `make ref`; no reference source is copied into this repo).  It replaces the
 * reference's main.cpp so that the reference's own CUDA build can be
 *   (a) run end to end with every solver parameter and the data-term enum chosen on the command
 *       line, looped and timed (`flow`), and
 *   (b) driven one operator at a time on caller-supplied buffers (`conv`, `resample`, `warp`,
 *       `solve`, `add`, `median`), which is the per-stage oracle of the GPU parity tests.
 * The operators are the reference's own CudaOperation*2D::Execute (its launch geometry, its
 * kernels, its named-parameter interface: SURVEY.md section 8b).
 *
 * All image files are headerless row-major float32, dense (pitch = width of the CONTAINER, W).
 * Kernels are looked up by the reference itself at <dir of this executable>/kernels/<name>.ptx.
 */
#include <cuda.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "src/cuda_operations/2d/cuda_operation_add_2d.h"
#include "src/cuda_operations/2d/cuda_operation_convolution_2d.h"
#include "src/cuda_operations/2d/cuda_operation_median_2d.h"
#include "src/cuda_operations/2d/cuda_operation_registration_2d.h"
#include "src/cuda_operations/2d/cuda_operation_resample_2d.h"
#include "src/cuda_operations/2d/cuda_operation_solve_2d.h"
#include "src/data_types/data2d.h"
#include "src/data_types/data_structs.h"
#include "src/data_types/operation_parameters.h"
#include "src/optical_flow/optical_flow_2d.h"
#include "src/utils/cuda_utils.h"



static bool read_raw(const char* name, std::vector<float>& v, size_t n) {
  v.resize(n);
  FILE* f = std::fopen(name, "rb");
  if (!f) { std::fprintf(stderr, "cannot open %s\n", name); return false; }
  size_t got = std::fread(v.data(), sizeof(float), n, f);
  std::fclose(f);
  if (got != n) { std::fprintf(stderr, "short read %s (%zu of %zu)\n", name, got, n); return false; }
  return true;
}

static bool write_raw(const char* name, const float* p, size_t n) {
  FILE* f = std::fopen(name, "wb");
  if (!f) { std::fprintf(stderr, "cannot open %s for writing\n", name); return false; }
  size_t put = std::fwrite(p, sizeof(float), n, f);
  std::fclose(f);
  return put == n;
}

/* Pitched containers allocated the way the reference does (optical_flow_2d.cpp:117-121),
 * pre-filled with NaN so that a read outside the written area shows up in the output. */
struct Containers {
  DataSize3 size{0, 0, 0};
  std::vector<CUdeviceptr> ptrs;
  bool init(size_t W, size_t H, int count) {
    size.width = W; size.height = H;
    for (int i = 0; i < count; i++) {
      CUdeviceptr p; size_t pitch;
      if (CheckCudaError(cuMemAllocPitch(&p, &pitch, W * sizeof(float), H, sizeof(float)))) return false;
      size.pitch = pitch;
      CheckCudaError(cuMemsetD32(p, 0x7fc00000u, pitch / 4 * H));
      ptrs.push_back(p);
    }
    return true;
  }
  void up(int i, const std::vector<float>& h) {
    CUDA_MEMCPY2D c; std::memset(&c, 0, sizeof(c));
    c.srcMemoryType = CU_MEMORYTYPE_HOST; c.srcHost = h.data(); c.srcPitch = size.width * 4;
    c.dstMemoryType = CU_MEMORYTYPE_DEVICE; c.dstDevice = ptrs[i]; c.dstPitch = size.pitch;
    c.WidthInBytes = size.width * 4; c.Height = size.height;
    CheckCudaError(cuMemcpy2D(&c));
  }
  void down(CUdeviceptr p, std::vector<float>& h) {
    h.resize(size.width * size.height);
    CUDA_MEMCPY2D c; std::memset(&c, 0, sizeof(c));
    c.srcMemoryType = CU_MEMORYTYPE_DEVICE; c.srcDevice = p; c.srcPitch = size.pitch;
    c.dstMemoryType = CU_MEMORYTYPE_HOST; c.dstHost = h.data(); c.dstPitch = size.width * 4;
    c.WidthInBytes = size.width * 4; c.Height = size.height;
    CheckCudaError(cuMemcpy2D(&c));
  }
};

static int usage() {
  std::fprintf(stderr,
    "ref_harness flow f0 f1 W H out_prefix levels scale outer inner alpha e_smooth e_data median sigma constancy warmup reps\n"
    "ref_harness conv W H sigma in out\n"
    "ref_harness resample W H iw ih ow oh in out\n"
    "ref_harness warp W H cw ch hx hy f0 f1 u v out\n"
    "ref_harness solve W H cw ch hx hy alpha e_smooth e_data outer inner constancy f0 f1 u v du_out dv_out phi_out ksi_out\n"
    "ref_harness add W H cw ch a b out\n"
    "ref_harness median W H cw ch radius in out\n");
  return 64;
}

static int run_flow(int argc, char** argv) {
  if (argc != 19) return usage();
  const char* f0n = argv[2]; const char* f1n = argv[3];
  size_t W = std::strtoull(argv[4], 0, 10), H = std::strtoull(argv[5], 0, 10);
  std::string out = argv[6];
  size_t levels = std::strtoull(argv[7], 0, 10);
  float scale = std::strtof(argv[8], 0);
  size_t outer = std::strtoull(argv[9], 0, 10), inner = std::strtoull(argv[10], 0, 10);
  float alpha = std::strtof(argv[11], 0), es = std::strtof(argv[12], 0), ed = std::strtof(argv[13], 0);
  size_t median = std::strtoull(argv[14], 0, 10);
  float sigma = std::strtof(argv[15], 0);
  int constancy = std::atoi(argv[16]);
  int warmup = std::atoi(argv[17]), reps = std::atoi(argv[18]);

  Data2D frame_0, frame_1;
  if (!frame_0.ReadRAWFromFileF32(f0n, W, H) || !frame_1.ReadRAWFromFileF32(f1n, W, H)) return 2;

  OpticalFlow2D optical_flow;
  DataSize3 data_size = {W, H, 1};
  DataConstancy dc = constancy == 1 ? DataConstancy::Gradient
                   : constancy == 2 ? DataConstancy::LogDerivatives : DataConstancy::Grey;
  if (!optical_flow.Initialize(data_size, dc)) return 4;
  optical_flow.silent = true;

  Data2D flow_u(W, H), flow_v(W, H);
  OperationParameters params;
  params.PushValuePtr("warp_levels_count", &levels);
  params.PushValuePtr("warp_scale_factor", &scale);
  params.PushValuePtr("outer_iterations_count", &outer);
  params.PushValuePtr("inner_iterations_count", &inner);
  params.PushValuePtr("equation_alpha", &alpha);
  params.PushValuePtr("equation_smoothness", &es);
  params.PushValuePtr("equation_data", &ed);
  params.PushValuePtr("median_radius", &median);
  params.PushValuePtr("gaussian_sigma", &sigma);

  for (int i = 0; i < warmup + reps; i++) {
    CheckCudaError(cuCtxSynchronize());
    auto t0 = std::chrono::steady_clock::now();
    optical_flow.ComputeFlow(frame_0, frame_1, flow_u, flow_v, params);
    CheckCudaError(cuCtxSynchronize());
    auto t1 = std::chrono::steady_clock::now();
    double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    std::printf("\nREF_MS %s %.6f\n", i < warmup ? "warmup" : "timed", ms);
  }
  if (out != "-") {
    write_raw((out + "u.raw").c_str(), flow_u.DataPtr(), W * H);
    write_raw((out + "v.raw").c_str(), flow_v.DataPtr(), W * H);
  }
  optical_flow.Destroy();
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) return usage();
  CUcontext ctx;
  if (!InitCudaContextWithFirstAvailableDevice(&ctx)) return 1;
  std::string cmd = argv[1];
  int rc = 0;

  if (cmd == "flow") {
    rc = run_flow(argc, argv);
    cuCtxDestroy(ctx);
    return rc;
  }

  if (argc < 4) return usage();
  size_t W = std::strtoull(argv[2], 0, 10), H = std::strtoull(argv[3], 0, 10);
  const size_t N = W * H;
  Containers c;
  if (!c.init(W, H, 12)) return 4;
  DataConstancy grey = DataConstancy::Grey;
  OperationParameters init;
  init.PushValuePtr("container_size", &c.size);
  init.PushValuePtr("data_constancy", &grey);
  std::vector<float> h0, h1, h2, h3, o0, o1;

  if (cmd == "conv" && argc == 7) {
    float sigma = std::strtof(argv[4], 0);
    if (!read_raw(argv[5], h0, N)) return 2;
    c.up(0, h0);
    CudaOperationConvolution2D op;
    if (!op.Initialize(&init)) return 4;
    DataSize3 ds = {W, H, 0};
    OperationParameters p;
    p.PushValuePtr("dev_input", &c.ptrs[0]);
    p.PushValuePtr("dev_output", &c.ptrs[1]);
    p.PushValuePtr("dev_temp", &c.ptrs[2]);
    p.PushValuePtr("data_size", &ds);
    p.PushValuePtr("gaussian_sigma", &sigma);
    op.Execute(p);
    CheckCudaError(cuCtxSynchronize());
    c.down(c.ptrs[1], o0);
    write_raw(argv[6], o0.data(), N);
  } else if (cmd == "resample" && argc == 10) {
    DataSize3 in_size = {std::strtoull(argv[4], 0, 10), std::strtoull(argv[5], 0, 10), 0};
    DataSize3 out_size = {std::strtoull(argv[6], 0, 10), std::strtoull(argv[7], 0, 10), 0};
    if (!read_raw(argv[8], h0, N)) return 2;
    c.up(0, h0);
    CudaOperationResample2D op;
    if (!op.Initialize(&init)) return 4;
    OperationParameters p;
    p.PushValuePtr("dev_input", &c.ptrs[0]);
    p.PushValuePtr("dev_output", &c.ptrs[1]);
    p.PushValuePtr("dev_temp", &c.ptrs[2]);
    p.PushValuePtr("data_size", &in_size);
    p.PushValuePtr("resample_size", &out_size);
    op.Execute(p);
    CheckCudaError(cuCtxSynchronize());
    c.down(c.ptrs[1], o0);
    write_raw(argv[9], o0.data(), N);
  } else if (cmd == "warp" && argc == 13) {
    DataSize3 ds = {std::strtoull(argv[4], 0, 10), std::strtoull(argv[5], 0, 10), 0};
    float hx = std::strtof(argv[6], 0), hy = std::strtof(argv[7], 0);
    if (!read_raw(argv[8], h0, N) || !read_raw(argv[9], h1, N) || !read_raw(argv[10], h2, N) ||
        !read_raw(argv[11], h3, N)) return 2;
    c.up(0, h0); c.up(1, h1); c.up(2, h2); c.up(3, h3);
    CudaOperationRegistration2D op;
    if (!op.Initialize(&init)) return 4;
    OperationParameters p;
    p.PushValuePtr("dev_frame_0", &c.ptrs[0]);
    p.PushValuePtr("dev_frame_1", &c.ptrs[1]);
    p.PushValuePtr("dev_flow_u", &c.ptrs[2]);
    p.PushValuePtr("dev_flow_v", &c.ptrs[3]);
    p.PushValuePtr("dev_output", &c.ptrs[4]);
    p.PushValuePtr("data_size", &ds);
    p.PushValuePtr("hx", &hx);
    p.PushValuePtr("hy", &hy);
    op.Execute(p);
    CheckCudaError(cuCtxSynchronize());
    c.down(c.ptrs[4], o0);
    write_raw(argv[12], o0.data(), N);
  } else if (cmd == "solve" && argc == 22) {
    DataSize3 ds = {std::strtoull(argv[4], 0, 10), std::strtoull(argv[5], 0, 10), 0};
    float hx = std::strtof(argv[6], 0), hy = std::strtof(argv[7], 0);
    float alpha = std::strtof(argv[8], 0), es = std::strtof(argv[9], 0), ed = std::strtof(argv[10], 0);
    size_t outer = std::strtoull(argv[11], 0, 10), inner = std::strtoull(argv[12], 0, 10);
    int constancy = std::atoi(argv[13]);
    DataConstancy dc = constancy == 1 ? DataConstancy::Gradient
                     : constancy == 2 ? DataConstancy::LogDerivatives : DataConstancy::Grey;
    if (!read_raw(argv[14], h0, N) || !read_raw(argv[15], h1, N) || !read_raw(argv[16], h2, N) ||
        !read_raw(argv[17], h3, N)) return 2;
    c.up(0, h0); c.up(1, h1); c.up(2, h2); c.up(3, h3);
    OperationParameters init2;
    init2.PushValuePtr("container_size", &c.size);
    init2.PushValuePtr("data_constancy", &dc);
    CudaOperationSolve2D op;
    if (!op.Initialize(&init2)) return 4;
    op.silent = true;
    CUdeviceptr du = c.ptrs[4], dv = c.ptrs[5], phi = c.ptrs[6], ksi = c.ptrs[7], tdu = c.ptrs[8], tdv = c.ptrs[9];
    OperationParameters p;
    p.PushValuePtr("dev_frame_0", &c.ptrs[0]);
    p.PushValuePtr("dev_frame_1", &c.ptrs[1]);
    p.PushValuePtr("dev_flow_u", &c.ptrs[2]);
    p.PushValuePtr("dev_flow_v", &c.ptrs[3]);
    p.PushValuePtr("dev_flow_du", &du);
    p.PushValuePtr("dev_flow_dv", &dv);
    p.PushValuePtr("dev_phi", &phi);
    p.PushValuePtr("dev_ksi", &ksi);
    p.PushValuePtr("dev_temp_du", &tdu);
    p.PushValuePtr("dev_temp_dv", &tdv);
    p.PushValuePtr("data_constancy", &dc);
    p.PushValuePtr("outer_iterations_count", &outer);
    p.PushValuePtr("inner_iterations_count", &inner);
    p.PushValuePtr("equation_alpha", &alpha);
    p.PushValuePtr("equation_smoothness", &es);
    p.PushValuePtr("equation_data", &ed);
    p.PushValuePtr("data_size", &ds);
    p.PushValuePtr("hx", &hx);
    p.PushValuePtr("hy", &hy);
    op.Execute(p);
    CheckCudaError(cuCtxSynchronize());
    c.down(du, o0); write_raw(argv[18], o0.data(), N);
    c.down(dv, o0); write_raw(argv[19], o0.data(), N);
    c.down(phi, o0); write_raw(argv[20], o0.data(), N);
    c.down(ksi, o0); write_raw(argv[21], o0.data(), N);
  } else if (cmd == "add" && argc == 9) {
    DataSize3 ds = {std::strtoull(argv[4], 0, 10), std::strtoull(argv[5], 0, 10), 0};
    if (!read_raw(argv[6], h0, N) || !read_raw(argv[7], h1, N)) return 2;
    c.up(0, h0); c.up(1, h1);
    CudaOperationAdd2D op;
    if (!op.Initialize(&init)) return 4;
    OperationParameters p;
    p.PushValuePtr("operand_0", &c.ptrs[0]);
    p.PushValuePtr("operand_1", &c.ptrs[1]);
    p.PushValuePtr("data_size", &ds);
    op.Execute(p);
    CheckCudaError(cuCtxSynchronize());
    c.down(c.ptrs[0], o0);
    write_raw(argv[8], o0.data(), N);
  } else if (cmd == "median" && argc == 9) {
    DataSize3 ds = {std::strtoull(argv[4], 0, 10), std::strtoull(argv[5], 0, 10), 0};
    size_t radius = std::strtoull(argv[6], 0, 10);
    if (!read_raw(argv[7], h0, N)) return 2;
    c.up(0, h0);
    CudaOperationMedian2D op;
    if (!op.Initialize(&init)) return 4;
    OperationParameters p;
    p.PushValuePtr("dev_input", &c.ptrs[0]);
    p.PushValuePtr("dev_output", &c.ptrs[1]);
    p.PushValuePtr("data_size", &ds);
    p.PushValuePtr("radius", &radius);
    op.Execute(p);
    CheckCudaError(cuCtxSynchronize());
    c.down(c.ptrs[1], o0);
    write_raw(argv[8], o0.data(), N);
  } else {
    rc = usage();
  }
  cuCtxDestroy(ctx);
  return rc;
}

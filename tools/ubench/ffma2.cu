// Micro-benchmark: scalar FFMA against packed FFMA2 / FADD2 (fma.rn.f32x2) throughput on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void k_ffma(float* out, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = __ffma2_rn(x[i], aa, bb);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// three distinct register operands per instruction (what a stencil's fma looks like)
__global__ void k_ffma_3reg(float* out, float a, float b) {
  float x[16], y[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { x[i] = threadIdx.x * 0.001f + i; y[i] = a + i; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], y[i], y[(i + 5) & 15]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2_3reg(float* out, float a, float b) {
  float2 x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i); y[i] = make_float2(a + i, b + i); }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = __ffma2_rn(x[i], y[i], y[(i + 3) & 7]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
static void run(const char* name, K kern, float* d, int threads, double flops_per_thread) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int blocks = 148 * 8;
  kern<<<blocks, threads>>>(d, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  kern<<<blocks, threads>>>(d, 1.0001f, 0.5f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double tf = flops_per_thread * threads * blocks / (ms * 1e-3) / 1e12;
  std::printf("%-14s threads %4d: %.3f ms  %.1f TFLOP/s (fma = 2 flop)\n", name, threads, ms, tf);
}

int main() {
  float* d;
  cudaMalloc(&d, 148 * 8 * 1024 * sizeof(float));
  for (int threads : {256, 384, 768, 1024}) {
    run("FFMA", k_ffma, d, threads, 2.0 * 16 * ITERS);
    run("FFMA2", k_ffma2, d, threads, 2.0 * 16 * ITERS);
    run("FFMA 3reg", k_ffma_3reg, d, threads, 2.0 * 16 * ITERS);
    run("FFMA2 3reg", k_ffma2_3reg, d, threads, 2.0 * 16 * ITERS);
  }
  std::printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// common.cuh -- shared device/host helpers of the flow2d-b200 kernels (sm_100a only).
//
// Arithmetic contract: every translation unit is compiled with -fmad=false, so
// `a*b+c` is two separately rounded fp32 operations and the ONLY fused multiply-adds are explicit
// fmaf() calls, placed where the reference kernels (nvcc -ptx + ptxas sm_100) have an FFMA.
// `/`, sqrtf and 1.f/x are IEEE round-to-nearest (nvcc defaults -prec-div=true -prec-sqrt=true),
// i.e. div.rn / sqrt.rn / rcp.rn like the reference.  No fast-math anywhere.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace flow2d {

// Mirror without edge repeat: -1 -> 1, n -> n-2 (reference: solve_2d.cu:75-76,88-89,101-102 and
// median_2d.cu:110-146).  The result is additionally clamped into [0, n-1] so that cells further
// outside than the reference ever reads (they are never used) still map to valid memory.
__host__ __device__ __forceinline__ int mirror_clamp(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - i - 2;
  i = i < 0 ? 0 : i;
  return i >= n ? n - 1 : i;
}

struct LevelGeom {
  int w, h;      // level size in pixels
  int pitch;     // container pitch in floats (multiple of 4; every row start is 16-byte aligned)
  float hx, hy;  // grid spacing of this level in finest-level pixels
};

constexpr int kMaxGaussRadius = 16;  // reference: 16-px halo of convolutionRowsKernel (convolution_2d.cu:70-73)
struct GaussTaps {
  int radius;
  float c[2 * kMaxGaussRadius + 1];
};

}  // namespace flow2d

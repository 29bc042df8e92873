// median.cu -- flow update + intermediate median filtering.  sm_100a.
//
// Replaces 2 x add_2d (src/kernels/add_2d.cu:33-46) and 2 x median_2d
// (src/kernels/median_2d.cu:87-299) of one pyramid level by ONE launch that handles both flow
// components: out = median_R(a + b) with the reference's mirrored border (no edge repeat).
//
// The reference gathers the R x R window and insertion-sorts it in local memory, then takes
// element R*R/2.  For NaN-free data that is the true median, a pure selection (compare/select, no
// arithmetic), so any exact selection is bit identical.  Here: "forgetful selection" entirely in
// registers -- keep R*R/2+2 candidates, repeatedly drop the current min and max and take in the
// next window element; the last survivor is the median.
#include "kernels.h"

namespace flow2d {

__device__ __forceinline__ void cswap(float& a, float& b) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo;
  b = hi;
}

// Moves the minimum of v[0..S) to v[0] and the maximum to v[S-1] (S >= 2).
template <int S>
__device__ __forceinline__ void min_max_to_ends(float* v) {
#pragma unroll
  for (int i = 0; i + 1 < S; i += 2) cswap(v[i], v[i + 1]);  // pairs: min at even, max at odd index
  // minimum among even slots (and the unpaired last one) -> v[0]
#pragma unroll
  for (int i = 2; i < S; i += 2) cswap(v[0], v[i]);
  // maximum among odd slots (and everything that might still be large) -> v[S-1]
  if ((S & 1) == 0) {
#pragma unroll
    for (int i = 1; i + 2 < S; i += 2) cswap(v[i], v[S - 1]);
  } else {
    // S odd: the last element is unpaired; compare all odd slots against it
#pragma unroll
    for (int i = 1; i < S - 1; i += 2) cswap(v[i], v[S - 1]);
    // v[S-1] took part in the min pass above (it is an even index), so it may have received a
    // larger value from v[0]'s chain; nothing else to do: the max of all elements is either an
    // odd slot or the unpaired one, and both were merged into v[S-1].
  }
}

template <int S, int N_LEFT>
struct Forget {
  // v holds S live candidates, `next` points at N_LEFT window elements still to be taken in.
  __device__ static __forceinline__ float run(float* v, const float* next, int stride_unused) {
    min_max_to_ends<S>(v);
    if constexpr (N_LEFT == 0) {
      static_assert(S == 3, "selection must end with three candidates");
      return v[1];
    } else {
      v[0] = next[0];  // drop min (slot 0) and max (slot S-1); the new element takes slot 0
      return Forget<S - 1, N_LEFT - 1>::run(v, next + 1, stride_unused);
    }
  }
};

template <int R>
__device__ __forceinline__ float median_window(const float* win) {
  constexpr int N = R * R;
  constexpr int S0 = N / 2 + 2;
  float v[S0];
#pragma unroll
  for (int i = 0; i < S0; i++) v[i] = win[i];
  return Forget<S0, N - S0>::run(v, win + S0, 0);
}

constexpr int kMedTW = 32, kMedTH = 8;

struct MedianPair {
  const float* a[2];
  const float* b[2];  // may be null: out = median(a)
  float* out[2];
};

template <int R>
__global__ void __launch_bounds__(kMedTW * kMedTH)
add_median_kernel(MedianPair io, int w, int h, int pitch, int row0, int row1) {
  constexpr int R2 = R / 2;
  constexpr int SW = kMedTW + 2 * R2, SH = kMedTH + 2 * R2;
  __shared__ float tile[SH * SW];
  const float* __restrict__ a = blockIdx.z ? io.a[1] : io.a[0];
  const float* __restrict__ b = blockIdx.z ? io.b[1] : io.b[0];
  float* __restrict__ out = blockIdx.z ? io.out[1] : io.out[0];
  const int x0 = blockIdx.x * kMedTW, y0 = row0 + blockIdx.y * kMedTH;  // rows [row0, row1) of the level (all unless slabbed)
  const int tid = threadIdx.y * kMedTW + threadIdx.x;
  for (int i = tid; i < SH * SW; i += kMedTW * kMedTH) {
    const int ly = i / SW, lx = i - ly * SW;
    const int gx = mirror_clamp(x0 - R2 + lx, w), gy = mirror_clamp(y0 - R2 + ly, h);
    const size_t g = (size_t)gy * pitch + gx;
    tile[i] = b ? a[g] + b[g] : a[g];
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= w || y >= row1) return;
  if constexpr (R == 1) {
    out[(size_t)y * pitch + x] = tile[threadIdx.y * SW + threadIdx.x];
  } else {
    float win[R * R];
#pragma unroll
    for (int iy = 0; iy < R; iy++)
#pragma unroll
      for (int ix = 0; ix < R; ix++) win[iy * R + ix] = tile[(threadIdx.y + iy) * SW + threadIdx.x + ix];
    out[(size_t)y * pitch + x] = median_window<R>(win);
  }
}

// radius: already normalised to 1, 3, 5 or 7.  count = 1 or 2 images.
void launch_add_median(cudaStream_t st, const float* const* a, const float* const* b, float* const* out, int count,
                       int w, int h, int pitch, int radius, int row0, int row1) {
  if (row1 <= row0) { row0 = 0; row1 = h; }
  MedianPair io;
  for (int i = 0; i < 2; i++) {
    int k = i < count ? i : 0;
    io.a[i] = a[k];
    io.b[i] = b ? b[k] : nullptr;
    io.out[i] = out[k];
  }
  dim3 block(kMedTW, kMedTH), grid((w + kMedTW - 1) / kMedTW, (row1 - row0 + kMedTH - 1) / kMedTH, count);
  switch (radius) {
    case 1: add_median_kernel<1><<<grid, block, 0, st>>>(io, w, h, pitch, row0, row1); break;
    case 3: add_median_kernel<3><<<grid, block, 0, st>>>(io, w, h, pitch, row0, row1); break;
    case 5: add_median_kernel<5><<<grid, block, 0, st>>>(io, w, h, pitch, row0, row1); break;
    default: add_median_kernel<7><<<grid, block, 0, st>>>(io, w, h, pitch, row0, row1); break;
  }
}

// a += b (add_2d.cu:43-44), kept for the per-stage API.
__global__ void __launch_bounds__(256) add_kernel(float* __restrict__ a, const float* __restrict__ b, int w, int h, int pitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < w && y < h) a[(size_t)y * pitch + x] += b[(size_t)y * pitch + x];
}

void launch_add(cudaStream_t st, float* a, const float* b, int w, int h, int pitch) {
  dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
  add_kernel<<<grid, block, 0, st>>>(a, b, w, h, pitch);
}

void preload_median_kernels() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, add_median_kernel<1>);
  cudaFuncGetAttributes(&a, add_median_kernel<3>);
  cudaFuncGetAttributes(&a, add_median_kernel<5>);
  cudaFuncGetAttributes(&a, add_median_kernel<7>);
  cudaFuncGetAttributes(&a, add_kernel);
}

}  // namespace flow2d

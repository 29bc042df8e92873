// pyramid.cu -- presmoothing, area resampling (restriction + prolongation), backward warping and
// derivative / motion-tensor assembly.  sm_100a, compiled with -fmad=false (see common.cuh).
//
// Replaces, with identical results:
//   convolutionRowsKernel + convolutionColumnsKernel  src/kernels/convolution_2d.cu:74-261
//   resample_x + resample_y                           src/kernels/resample_2d.cu:34-118
//   registration_2d                                   src/kernels/registration_2d.cu:34-74
//   the fx/fy/ft (+ gradient-constancy tensor) part of solve_2d / solve_2d_grad
//                                                     src/kernels/solve_2d.cu:311-329, 798-884
#include "kernels.h"

namespace flow2d {

// ---------------------------------------------------------------------------------------------
// Gaussian presmoothing: one fused kernel (rows pass into shared memory, then columns pass).
// Zero padding in both directions, ascending-j fma chain from sum = 0 (convolution_2d.cu:150-166).
// Algorithmic traffic: 1 read + 1 write per pixel (the reference's two kernels: 2R + 2W).
// ---------------------------------------------------------------------------------------------
constexpr int kBlurTW = 64;  // output tile width  (threads in x)
constexpr int kBlurTH = 32;  // output tile height (each of the 8 thread rows produces 4 outputs)

__global__ void __launch_bounds__(512)
blur_kernel(const float* __restrict__ in, float* __restrict__ out, int w, int h, int pitch, GaussTaps taps) {
  extern __shared__ float smem[];
  const int r = taps.radius;
  const int IW = kBlurTW + 2 * r;        // staged input width
  const int IH = kBlurTH + 2 * r;        // staged input / row-pass height
  float* s_in = smem;                    // IH x IW   input tile, zero outside the image
  float* s_row = smem + IH * IW;         // IH x TW   row-pass result, zero for rows outside the image
  const int x0 = blockIdx.x * kBlurTW, y0 = blockIdx.y * kBlurTH;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthreads = blockDim.x * blockDim.y;

  for (int i = tid; i < IH * IW; i += nthreads) {
    int ly = i / IW, lx = i - ly * IW;
    int gx = x0 - r + lx, gy = y0 - r + ly;
    s_in[i] = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? in[(size_t)gy * pitch + gx] : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < IH * kBlurTW; i += nthreads) {
    int ly = i / kBlurTW, lx = i - ly * kBlurTW;
    int gy = y0 - r + ly;
    float sum = 0.f;
    if (gy >= 0 && gy < h) {
      const float* row = s_in + ly * IW + lx + r;
      for (int j = -r; j <= r; j++) sum = fmaf(taps.c[r - j], row[j], sum);
    }
    s_row[i] = sum;
  }
  __syncthreads();
  for (int i = tid; i < kBlurTH * kBlurTW; i += nthreads) {
    int ly = i / kBlurTW, lx = i - ly * kBlurTW;
    int gx = x0 + lx, gy = y0 + ly;
    if (gx < w && gy < h) {
      const float* col = s_row + (ly + r) * kBlurTW + lx;
      float sum = 0.f;
      for (int j = -r; j <= r; j++) sum = fmaf(taps.c[r - j], col[j * kBlurTW], sum);
      out[(size_t)gy * pitch + gx] = sum;
    }
  }
}

void launch_blur(cudaStream_t st, const float* in, float* out, int w, int h, int pitch, const GaussTaps& taps) {
  const int r = taps.radius;
  size_t smem = sizeof(float) * ((kBlurTH + 2 * r) * (kBlurTW + 2 * r) + (kBlurTH + 2 * r) * kBlurTW);
  dim3 grid((w + kBlurTW - 1) / kBlurTW, (h + kBlurTH - 1) / kBlurTH), block(64, 8);
  blur_kernel<<<grid, block, smem, st>>>(in, out, w, h, pitch, taps);
}

// ---------------------------------------------------------------------------------------------
// Area resampling, one axis per kernel, up to four images per launch (blockIdx.z): the two frames that a
// level restricts and the two flow components it prolongates go through the same two launches.
// The index path is the reference's fp32 arithmetic verbatim (resample_2d.cu:44-51):
//   delta = in/(float)out (div.rn), left_f = x*delta, right_f = (x+1)*delta (plain mul),
//   left_i = floor, right_i = min(in, ceil); accumulation value = fma(frac, in[..], value).
// ---------------------------------------------------------------------------------------------
struct ResampleBatch {  // up to four images per launch (blockIdx.z), each with its own geometry
  ResampleJob job[4];
};

// One thread = four consecutive outputs along x (one float4 store; the y pass also reads float4 rows of the x-pass
// result): four independent fma chains in flight per thread instead of one.  Every output is still the reference's
// sequential chain value = fma(frac_j, in[left_i + j], value), j ascending, times the normalisation (its order is
// part of the result).
struct ResampleCell {
  int left_i, n;
  float left_f, right_f;
};
__device__ __forceinline__ ResampleCell resample_cell(int o, int in_n, float delta) {
  ResampleCell c;
  c.left_f = (float)o * delta;
  c.right_f = (float)(o + 1) * delta;
  c.left_i = (int)floorf(c.left_f);
  c.n = min(in_n, (int)ceilf(c.right_f)) - c.left_i;
  return c;
}
__device__ __forceinline__ float resample_frac(const ResampleCell& c, int j, float delta) {
  float frac = 1.f;
  if (j == 0) frac = (float)(c.left_i + 1) - c.left_f;
  if (j == c.n - 1) frac = c.right_f - (float)(c.left_i + j);
  if (c.n == 1) frac = delta;
  return frac;
}

template <bool ALONG_X>
__global__ void __launch_bounds__(256)
resample_kernel(ResampleBatch b, int pitch) {
  const ResampleJob& jb = b.job[blockIdx.z];
  const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  // rows this pass writes: the x pass the input rows [iy0, iy1) the y pass will read, the y pass the output rows [oy0, oy1)
  const int y = (ALONG_X ? jb.iy0 : jb.oy0) + blockIdx.y * blockDim.y + threadIdx.y;
  const int out_w = jb.ow, y_end = ALONG_X ? jb.iy1 : jb.oy1;
  const int in_n = ALONG_X ? jb.iw : jb.ih, out_n = ALONG_X ? jb.ow : jb.oh;
  if (x >= out_w || y >= y_end) return;
  const float* __restrict__ in = ALONG_X ? jb.in : jb.tmp;
  float* __restrict__ out = ALONG_X ? jb.tmp : jb.out;
  const float delta = (float)in_n / (float)out_n;
  const float normalization = (float)out_n / (float)in_n;
  float value[4] = {0.f, 0.f, 0.f, 0.f};
  if (ALONG_X) {
    const float* row = in + (size_t)y * pitch;
    ResampleCell c[4];
    int nmax = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      c[i] = resample_cell(min(x + i, out_w - 1), in_n, delta);
      nmax = max(nmax, c[i].n);
    }
    for (int j = 0; j < nmax; j++) {
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (j < c[i].n) value[i] = fmaf(resample_frac(c[i], j, delta), row[c[i].left_i + j], value[i]);
    }
  } else {
    const ResampleCell c = resample_cell(y, in_n, delta);
    const float* src = in + (size_t)c.left_i * pitch + x;
    if (x + 3 < out_w) {
      // The chain of an output is sequential (its order is part of the result) but its loads are not: eight rows are
      // requested before the first of them is consumed.  A coarse level has few outputs with chains of 50 - 200 rows;
      // one load per chain step made the pass a string of dependent memory round trips (23 us per level at 1024^2).
      constexpr int kAhead = 8;
      for (int j0 = 0; j0 < c.n; j0 += kAhead) {
        float4 v[kAhead];
#pragma unroll
        for (int k = 0; k < kAhead; k++)
          v[k] = *reinterpret_cast<const float4*>(src + (size_t)min(j0 + k, c.n - 1) * pitch);
#pragma unroll
        for (int k = 0; k < kAhead; k++) {
          if (j0 + k < c.n) {
            const float frac = resample_frac(c, j0 + k, delta);
            value[0] = fmaf(frac, v[k].x, value[0]);
            value[1] = fmaf(frac, v[k].y, value[1]);
            value[2] = fmaf(frac, v[k].z, value[2]);
            value[3] = fmaf(frac, v[k].w, value[3]);
          }
        }
      }
    } else {
      for (int j = 0; j < c.n; j++) {
        const float frac = resample_frac(c, j, delta);
        for (int i = 0; x + i < out_w; i++) value[i] = fmaf(frac, src[(size_t)j * pitch + i], value[i]);
      }
    }
  }
  float* o = out + (size_t)y * pitch + x;
  if (x + 3 < out_w) {
    *reinterpret_cast<float4*>(o) = make_float4(value[0] * normalization, value[1] * normalization, value[2] * normalization,
                                                value[3] * normalization);
  } else {
    for (int i = 0; x + i < out_w; i++) o[i] = value[i] * normalization;
  }
}

// The x pass with the input staged through shared memory.  The plain kernel above reads its taps straight from global
// memory: lanes of a warp are delta floats apart per load instruction, which ran the full-resolution frames at 22 % of
// the HBM copy bandwidth (ncu, 8192x8192: 370 us per level for 512 MB).  Here a CTA takes 8 input rows and a chunk of
// at most 128 consecutive outputs, loads the input span of the chunk with coalesced float4 rows into shared memory
// and runs the same sequential chains from there.  The chunk shrinks with delta so that the span fits (coarse levels:
// few outputs, long chains).  Same arithmetic, same order.
constexpr int kRxRows = 8, kRxSpan = 1032;  // staged floats per row: chunk * delta + alignment slack

__global__ void __launch_bounds__(256)
resample_x_staged_kernel(ResampleBatch b, int pitch) {
  __shared__ __align__(16) float tile[kRxRows][kRxSpan];
  const ResampleJob& jb = b.job[blockIdx.z];
  const int chunk = jb.chunk;
  const int x0 = blockIdx.x * chunk;
  const int row0 = jb.iy0 + blockIdx.y * kRxRows;
  if (x0 >= jb.ow || row0 >= jb.iy1) return;  // uniform per CTA
  const int in_n = jb.iw, out_n = jb.ow;
  const float delta = (float)in_n / (float)out_n;
  const float normalization = (float)out_n / (float)in_n;
  const int xe = min(x0 + chunk, out_n);                                   // outputs [x0, xe)
  const int s0 = (int)floorf((float)x0 * delta) & ~3;                      // first staged input index (16-byte aligned)
  const int e1 = min(in_n, (int)ceilf((float)xe * delta));                 // one past the last input index read
  const int span4 = (e1 - s0 + 3) >> 2;                                    // float4s per row (<= kRxSpan / 4 by construction)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int row = row0 + warp;
    if (row < jb.iy1) {
      const float4* src = reinterpret_cast<const float4*>(jb.in + (size_t)row * pitch + s0);
      float4* dst = reinterpret_cast<float4*>(tile[warp]);
      for (int i = lane; i < span4; i += 32) dst[i] = src[i];  // (reads at most up to the row's pitch: containers are padded)
    }
  }
  __syncthreads();
  const int row = row0 + warp;
  const int x = x0 + 4 * lane;
  if (row >= jb.iy1 || x >= xe) return;
  const float* trow = tile[warp] - s0;  // trow[input index]
  float value[4] = {0.f, 0.f, 0.f, 0.f};
  ResampleCell c[4];
  int nmax = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    c[i] = resample_cell(min(x + i, out_n - 1), in_n, delta);
    nmax = max(nmax, c[i].n);
  }
  for (int j = 0; j < nmax; j++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (j < c[i].n) value[i] = fmaf(resample_frac(c[i], j, delta), trow[c[i].left_i + j], value[i]);
  }
  float* o = jb.tmp + (size_t)row * pitch + x;
  if (x + 3 < out_n) {
    *reinterpret_cast<float4*>(o) = make_float4(value[0] * normalization, value[1] * normalization, value[2] * normalization,
                                                value[3] * normalization);
  } else {
    for (int i = 0; x + i < out_n; i++) o[i] = value[i] * normalization;
  }
}

// One x pass and one y pass for `count` (1..4) images of possibly different sizes.  A job with oy0 < oy1 < oh produces
// only those output rows (row-slab mode): the x pass then covers just the input rows they are made of.
void launch_resample_batch(cudaStream_t st, const ResampleJob* jobs, int count, int pitch) {
  ResampleBatch b;
  int gw = 0, gh_x = 0, gh_y = 0;
  for (int i = 0; i < 4; i++) {
    ResampleJob& j = b.job[i];
    j = jobs[i < count ? i : 0];
    if (j.oy1 <= j.oy0) { j.oy0 = 0; j.oy1 = j.oh; }  // default: all rows
    // input rows [iy0, iy1) that the output rows [oy0, oy1) read: same fp32 expressions as the kernel
    const float delta = (float)j.ih / (float)j.oh;
    j.iy0 = (int)floorf((float)j.oy0 * delta);
    const int hi = (int)ceilf((float)j.oy1 * delta);
    j.iy1 = hi < j.ih ? hi : j.ih;
    if (i < count) {
      gw = j.ow > gw ? j.ow : gw;
      gh_x = j.iy1 - j.iy0 > gh_x ? j.iy1 - j.iy0 : gh_x;
      gh_y = j.oy1 - j.oy0 > gh_y ? j.oy1 - j.oy0 : gh_y;
    }
  }
  dim3 block(32, 8);
  // staged x pass: every job gets the largest chunk (multiple of 4, <= 128) whose input span fits the shared tile
  bool staged = true;
  int gx = 0;
  for (int i = 0; i < 4; i++) {
    ResampleJob& j = b.job[i];
    const float delta = (float)j.iw / (float)j.ow;
    int chunk = (int)((float)(kRxSpan - 12) / delta) & ~3;
    chunk = chunk > 128 ? 128 : chunk;
    if (chunk < 4) staged = false;
    j.chunk = chunk;
    if (i < count && chunk >= 4) gx = (j.ow + chunk - 1) / chunk > gx ? (j.ow + chunk - 1) / chunk : gx;
  }
  if (staged) resample_x_staged_kernel<<<dim3(gx, (gh_x + kRxRows - 1) / kRxRows, count), 256, 0, st>>>(b, pitch);
  else resample_kernel<true><<<dim3((gw + 127) / 128, (gh_x + 7) / 8, count), block, 0, st>>>(b, pitch);
  resample_kernel<false><<<dim3((gw + 127) / 128, (gh_y + 7) / 8, count), block, 0, st>>>(b, pitch);
}

// (iw x ih) -> (ow x oh) for `count` (1 or 2) images; tmp[] are scratch containers.
void launch_resample(cudaStream_t st, const float* const* in, float* const* tmp, float* const* out, int count,
                     int iw, int ih, int ow, int oh, int pitch) {
  ResampleJob jobs[2];
  for (int i = 0; i < count && i < 2; i++) jobs[i] = ResampleJob{in[i], tmp[i], out[i], iw, ih, ow, oh, 0, 0, 0, 0, 0};
  launch_resample_batch(st, jobs, count < 2 ? count : 2, pitch);
}

// ---------------------------------------------------------------------------------------------
// Backward bilinear registration of frame 1 by the current flow (registration_2d.cu:48-72).
// As compiled upstream: x_f = fma(rcp.rn(hx), u, (float)x); bounds (float)(w-1); the blend is
// fma(w11,f11, fma(w01,f01, fma(f00,w00, w10*f10))).  Index work (floor, clamp, OOB/NaN test)
// is bit exact by construction.
// ---------------------------------------------------------------------------------------------
// Branch-free so that the 16 gathers of a thread's four pixels are all in flight at once: the taps of an
// out-of-range pixel are clamped into the image (loaded, never used) and frame 0 is read under a predicate.
struct WarpTap {
  int i00, i01, i10, i11;  // element offsets of the four taps in frame 1
  float w00, w10, w01, w11;
  bool oob;
};
__device__ __forceinline__ WarpTap warp_tap(float u, float v, int xx, int yy, int w, int h, int pitch, float rhx, float rhy,
                                            float bx, float by) {
  WarpTap t;
  const float x_f = fmaf(rhx, u, (float)xx);
  const float y_f = fmaf(rhy, v, (float)yy);
  // (x_f < 0 || x_f > bx || y_f < 0 || y_f > by || isnan(x_f) || isnan(y_f)), registration_2d.cu:57
  t.oob = !(x_f >= 0.f && x_f <= bx && y_f >= 0.f && y_f <= by);
  const float xs = t.oob ? 0.f : x_f, ys = t.oob ? 0.f : y_f;
  const int x = (int)xs, y = (int)ys;  // == floorf for the non-negative in-range values
  const float dx = xs - (float)x, dy = ys - (float)y;
  const int x1 = min(w - 1, x + 1), y1 = min(h - 1, y + 1);
  const float ox = 1.f - dx, oy = 1.f - dy;
  t.w00 = ox * oy; t.w10 = dx * oy; t.w01 = ox * dy; t.w11 = dx * dy;
  const int r0 = y * pitch, r1 = y1 * pitch;
  t.i00 = r0 + x; t.i10 = r0 + x1; t.i01 = r1 + x; t.i11 = r1 + x1;
  return t;
}
__device__ __forceinline__ float warp_blend(const WarpTap& t, float f00, float f10, float f01, float f11, float own) {
  float val = t.w10 * f10;
  val = fmaf(f00, t.w00, val);
  val = fmaf(t.w01, f01, val);
  val = fmaf(t.w11, f11, val);
  return t.oob ? own : val;
}

// One thread = four consecutive pixels: float4 loads of u, v, one float4 store, 16 gathers in flight.
__global__ void __launch_bounds__(256)
warp_kernel(const float* __restrict__ f0, const float* __restrict__ f1, const float* __restrict__ u,
            const float* __restrict__ v, float* __restrict__ out, int w, int h, int pitch, float rhx, float rhy, int y0, int y1) {
  const int xx = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int yy = y0 + blockIdx.y * blockDim.y + threadIdx.y;  // rows [y0, y1) of the level (all of them unless slabbed)
  if (xx >= w || yy >= y1) return;
  const int c = yy * pitch + xx;
  const float bx = (float)(w - 1), by = (float)(h - 1);
  if (xx + 3 < w) {
    const float4 uu = *reinterpret_cast<const float4*>(u + c), vv = *reinterpret_cast<const float4*>(v + c);
    const float us[4] = {uu.x, uu.y, uu.z, uu.w}, vs[4] = {vv.x, vv.y, vv.z, vv.w};
    WarpTap t[4];
    float f00[4], f10[4], f01[4], f11[4], own[4], o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) t[i] = warp_tap(us[i], vs[i], xx + i, yy, w, h, pitch, rhx, rhy, bx, by);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      f00[i] = f1[t[i].i00]; f10[i] = f1[t[i].i10]; f01[i] = f1[t[i].i01]; f11[i] = f1[t[i].i11];
      own[i] = t[i].oob ? f0[c + i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) o[i] = warp_blend(t[i], f00[i], f10[i], f01[i], f11[i], own[i]);
    *reinterpret_cast<float4*>(out + c) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
    for (int i = 0; xx + i < w; i++) {
      const WarpTap t = warp_tap(u[c + i], v[c + i], xx + i, yy, w, h, pitch, rhx, rhy, bx, by);
      out[c + i] = warp_blend(t, f1[t.i00], f1[t.i10], f1[t.i01], f1[t.i11], t.oob ? f0[c + i] : 0.f);
    }
  }
}

void launch_warp(cudaStream_t st, const float* f0, const float* f1, const float* u, const float* v, float* out,
                 const LevelGeom& g, int y0, int y1) {
  if (y1 <= y0) { y0 = 0; y1 = g.h; }
  dim3 block(32, 8), grid((g.w + 127) / 128, (y1 - y0 + 7) / 8);
  warp_kernel<<<grid, block, 0, st>>>(f0, f1, u, v, out, g.w, g.h, g.pitch, 1.f / g.hx, 1.f / g.hy, y0, y1);
}

// ---------------------------------------------------------------------------------------------
// Derivative assembly.  fx, fy, ft depend only on frame 0 and the warped frame 1, i.e. they are
// constant for a whole pyramid level, while the reference recomputes them in every one of the
// outer*(1+inner) kernel launches of the level (solve_2d.cu:164-174, 311-321).  They are computed
// here once per level with the reference's expression tree:
//   fx = (((f0[x+1]-f0[x-1]) + f1[x+1]) - f1[x-1]) / (4*hx)   (mirrored neighbours, div.rn)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
derivatives_kernel(const float* __restrict__ f0, const float* __restrict__ f1, float* __restrict__ fx,
                   float* __restrict__ fy, float* __restrict__ ft, int w, int h, int pitch, float hx4, float hy4, int y0, int y1) {
  // one thread = four consecutive pixels (float4 rows; the two x neighbours outside the strip are scalar loads)
  const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= y1) return;
  const int ym = mirror_clamp(y - 1, h), yp = mirror_clamp(y + 1, h);
  const size_t row = (size_t)y * pitch, up = (size_t)ym * pitch, dn = (size_t)yp * pitch;
  if (x + 3 < w) {
    const float4 a0 = *reinterpret_cast<const float4*>(f0 + row + x), a1 = *reinterpret_cast<const float4*>(f1 + row + x);
    const float4 u0 = *reinterpret_cast<const float4*>(f0 + up + x), u1 = *reinterpret_cast<const float4*>(f1 + up + x);
    const float4 d0 = *reinterpret_cast<const float4*>(f0 + dn + x), d1 = *reinterpret_cast<const float4*>(f1 + dn + x);
    const int xl = mirror_clamp(x - 1, w), xr = mirror_clamp(x + 4, w);
    const float l0 = f0[row + xl], l1 = f1[row + xl], r0 = f0[row + xr], r1 = f1[row + xr];
    const float c0[6] = {l0, a0.x, a0.y, a0.z, a0.w, r0}, c1[6] = {l1, a1.x, a1.y, a1.z, a1.w, r1};
    const float p0[4] = {u0.x, u0.y, u0.z, u0.w}, p1[4] = {u1.x, u1.y, u1.z, u1.w};
    const float q0[4] = {d0.x, d0.y, d0.z, d0.w}, q1[4] = {d1.x, d1.y, d1.z, d1.w};
    float ox[4], oy[4], ot[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      ox[i] = (((c0[i + 2] - c0[i]) + c1[i + 2]) - c1[i]) / hx4;
      oy[i] = (((q0[i] - p0[i]) + q1[i]) - p1[i]) / hy4;
      ot[i] = c1[i + 1] - c0[i + 1];
    }
    *reinterpret_cast<float4*>(fx + row + x) = make_float4(ox[0], ox[1], ox[2], ox[3]);
    *reinterpret_cast<float4*>(fy + row + x) = make_float4(oy[0], oy[1], oy[2], oy[3]);
    *reinterpret_cast<float4*>(ft + row + x) = make_float4(ot[0], ot[1], ot[2], ot[3]);
  } else {
    for (int i = 0; x + i < w; i++) {
      const int xc = x + i, xm = mirror_clamp(xc - 1, w), xp = mirror_clamp(xc + 1, w);
      fx[row + xc] = (((f0[row + xp] - f0[row + xm]) + f1[row + xp]) - f1[row + xm]) / hx4;
      fy[row + xc] = (((f0[dn + xc] - f0[up + xc]) + f1[dn + xc]) - f1[up + xc]) / hy4;
      ft[row + xc] = f1[row + xc] - f0[row + xc];
    }
  }
}

// Gradient-constancy motion tensor (solve_2d.cu:868-884) from the fx/fy/ft planes.  The reference
// takes the central differences inside a 16x8 CUDA block whose 1-px halo of fx/fy/ft is the
// block-edge thread's OWN value (solve_2d.cu:813-841), so its result depends on that tiling; it is
// reproduced here (tx = x%16, ty = y%8).  The cell just outside the image in a partial block is
// uninitialised shared memory upstream; here it is defined as the own value.
// hx_1 = (float)(1.0/(2.0*hx)) is computed in double upstream and passed in.
__global__ void __launch_bounds__(256)
grad_tensor_kernel(const float* __restrict__ fx, const float* __restrict__ fy, const float* __restrict__ ft,
                   float* __restrict__ J11, float* __restrict__ J22, float* __restrict__ J12,
                   float* __restrict__ J13, float* __restrict__ J23, int w, int h, int pitch, float hx_1, float hy_1, int y0, int y1) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= y1) return;
  const int tx = x & 15, ty = y & 7;
  const int xl = (tx == 0) ? x : x - 1;
  const int xr = (tx == 15 || x + 1 >= w) ? x : x + 1;
  const int yu = (ty == 0) ? y : y - 1;
  const int yb = (ty == 7 || y + 1 >= h) ? y : y + 1;
  const size_t row = (size_t)y * pitch, up = (size_t)yu * pitch, dn = (size_t)yb * pitch;
  const float fxx = (fx[row + xr] - fx[row + xl]) * hx_1;
  const float fxy = (fx[dn + x] - fx[up + x]) * hy_1;
  const float fyy = (fy[dn + x] - fy[up + x]) * hy_1;
  const float fxt = (ft[row + xr] - ft[row + xl]) * hx_1;
  const float fyt = (ft[dn + x] - ft[up + x]) * hy_1;
  const size_t c = row + x;
  J11[c] = fmaf(fxx, fxx, fxy * fxy);
  J22[c] = fmaf(fxy, fxy, fyy * fyy);
  J12[c] = fmaf(fxx, fxy, fxy * fyy);
  J13[c] = fmaf(fxx, fxt, fxy * fyt);
  J23[c] = fmaf(fxy, fxt, fyy * fyt);
}

void launch_derivatives(cudaStream_t st, const float* f0, const float* f1w, float* fx, float* fy, float* ft,
                        const LevelGeom& g, int y0, int y1) {
  if (y1 <= y0) { y0 = 0; y1 = g.h; }
  dim3 block(32, 8), grid((g.w + 127) / 128, (y1 - y0 + 7) / 8);
  derivatives_kernel<<<grid, block, 0, st>>>(f0, f1w, fx, fy, ft, g.w, g.h, g.pitch, g.hx * 4.f, g.hy * 4.f, y0, y1);
}

void launch_grad_tensor(cudaStream_t st, const float* fx, const float* fy, const float* ft, float* const* J,
                        const LevelGeom& g, int y0, int y1) {
  if (y1 <= y0) { y0 = 0; y1 = g.h; }
  dim3 block(32, 8), grid((g.w + 31) / 32, (y1 - y0 + 7) / 8);
  const float hx_1 = (float)(1.0 / (2.0 * (double)g.hx)), hy_1 = (float)(1.0 / (2.0 * (double)g.hy));
  grad_tensor_kernel<<<grid, block, 0, st>>>(fx, fy, ft, J[0], J[1], J[2], J[3], J[4], g.w, g.h, g.pitch, hx_1, hy_1, y0, y1);
}

// Lazy module loading (the CUDA 12 default) loads a kernel at its first launch and may synchronise the context to do so:
// a rank whose stream spins on a neighbour's flag would then block the very launch it is waiting for when two ranks share
// a device.  flow2d_slab_connect therefore loads every kernel up front.
void preload_pyramid_kernels() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, blur_kernel);
  cudaFuncGetAttributes(&a, resample_kernel<true>);
  cudaFuncGetAttributes(&a, resample_kernel<false>);
  cudaFuncGetAttributes(&a, resample_x_staged_kernel);
  cudaFuncGetAttributes(&a, warp_kernel);
  cudaFuncGetAttributes(&a, derivatives_kernel);
  cudaFuncGetAttributes(&a, grad_tensor_kernel);
}

}  // namespace flow2d

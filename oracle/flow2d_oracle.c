/*
 * flow2d_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see flow2d_oracle.h).
 *
 * Plain C, fp32, literal restatement of axruff/cuda-flow2d's ComputeFlow path.
 * Build with -ffp-contract=off: every fused multiply-add below is an explicit fmaf()
 * placed exactly where the reference's SASS (nvcc 12.9 -ptx, ptxas -arch=sm_100) has an FFMA;
 * everything else is a separately rounded fp32 operation.  IEEE `/`, sqrtf and 1.f/x are
 * correctly rounded and equal div.rn / sqrt.rn / rcp.rn.
 */
#include "flow2d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX(x, y) ((size_t)(y) * pitch + (size_t)(x))

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* Mirror without edge repeat: -1 -> 1, n -> n-2 (solve_2d.cu:75-76,88-89,101-102;
 * median_2d.cu:110-146). */
static inline long mirror(long i, long n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - i - 2;
  return i;
}

/* ---------------------------------------------------------------------------------------------
 * Level table (integer results from fp32 math: must be bit exact)
 * optical_flow_base_2d.cpp:36-59
 * ------------------------------------------------------------------------------------------- */
size_t oracle_max_warp_level(size_t width, size_t height, float scale_factor) {
  size_t r_width = 1, r_height = 1, level_counter = 1;
  while (scale_factor < 1.f) {
    float scale = powf(scale_factor, (float)level_counter);
    r_width = (size_t)ceilf((float)width * scale);
    r_height = (size_t)ceilf((float)height * scale);
    if (r_width < 4 || r_height < 4) break;
    ++level_counter;
  }
  if (r_width == 1 || r_height == 1) --level_counter;
  return level_counter;
}

/* optical_flow_2d.cpp:268-272 */
void oracle_level_geometry(size_t W, size_t H, float scale_factor, int level,
                           size_t* cw, size_t* ch, float* hx, float* hy) {
  float scale = powf(scale_factor, (float)level);
  *cw = (size_t)ceilf((float)W * scale);
  *ch = (size_t)ceilf((float)H * scale);
  *hx = (float)W / (float)(*cw);
  *hy = (float)H / (float)(*ch);
}

/* ---------------------------------------------------------------------------------------------
 * Gaussian presmoothing.  Taps: cuda_operation_convolution_2d.cpp:83-112 (double math -> float,
 * float running-sum normalisation).  Convolution: convolution_2d.cu:150-166 / 244-258: zero
 * padding, `sum += c[r-j]*s[x+j]` for j=-r..r ascending, contracted to an fma chain from sum=0.
 * ------------------------------------------------------------------------------------------- */
int oracle_gauss_taps(float sigma, float* taps) {
  const float pixel_size = 1.0f;
  const size_t precision = 3;
  size_t radius = (size_t)((float)precision * sigma / pixel_size);
  int r = (int)radius;
  for (int i = -r; i <= r; i++) {
    float arg = -((float)(i * i) * pixel_size * pixel_size); /* int*int, then float products */
    float val = (float)(1.0 / ((double)sigma * sqrt(2.0 * 3.1415926)) *
                        exp((double)arg / (2.0 * (double)sigma * (double)sigma)));
    taps[i + r] = val;
  }
  float sum = 0.0f;
  for (int i = 0; i < 2 * r + 1; i++) sum = sum + taps[i];
  for (int i = 0; i < 2 * r + 1; i++) taps[i] = taps[i] / sum;
  return r;
}

void oracle_blur(const float* in, float* out, size_t w, size_t h, size_t pitch, float sigma) {
  float taps[64];
  int r = oracle_gauss_taps(sigma, taps);
  float* tmp = (float*)malloc(sizeof(float) * pitch * h);
  long W = (long)w, Hh = (long)h;
  /* rows: convolution_2d.cu:74-168 */
#pragma omp parallel for schedule(static)
  for (long y = 0; y < Hh; y++)
    for (long x = 0; x < W; x++) {
      float sum = 0.f;
      for (int j = -r; j <= r; j++) {
        long xx = x + j;
        float s = (xx >= 0 && xx < W) ? in[IDX(xx, y)] : 0.f;
        sum = fmaf(taps[r - j], s, sum);
      }
      tmp[IDX(x, y)] = sum;
    }
  /* columns: convolution_2d.cu:181-261 */
#pragma omp parallel for schedule(static)
  for (long y = 0; y < Hh; y++)
    for (long x = 0; x < W; x++) {
      float sum = 0.f;
      for (int j = -r; j <= r; j++) {
        long yy = y + j;
        float s = (yy >= 0 && yy < Hh) ? tmp[IDX(x, yy)] : 0.f;
        sum = fmaf(taps[r - j], s, sum);
      }
      out[IDX(x, y)] = sum;
    }
  free(tmp);
}

/* ---------------------------------------------------------------------------------------------
 * Area resampling.  resample_2d.cu:34-118.  As compiled: delta = div.rn((float)in,(float)out),
 * left_f = delta*(float)x, right_f = delta*(float)(x+1) (plain mul), floor/ceil -> int,
 * value = fma(frac, in[..], value), out = value * norm.
 * ------------------------------------------------------------------------------------------- */
void oracle_resample_cells(size_t in_n, size_t out_n, size_t x, int* left_i, int* right_i) {
  float delta = (float)in_n / (float)out_n;
  float left_f = (float)x * delta;
  float right_f = (float)(x + 1) * delta;
  *left_i = (int)floorf(left_f);
  int ri = (int)ceilf(right_f);
  *right_i = ri < (int)in_n ? ri : (int)in_n;
}

static inline float resample_1d(const float* in, size_t stride, size_t in_n, size_t out_n, size_t x) {
  float delta = (float)in_n / (float)out_n;
  float normalization = (float)out_n / (float)in_n;
  float left_f = (float)x * delta;
  float right_f = (float)(x + 1) * delta;
  int left_i = (int)floorf(left_f);
  int right_i = (int)ceilf(right_f);
  if (right_i > (int)in_n) right_i = (int)in_n;
  float value = 0.f;
  int n = right_i - left_i;
  for (int j = 0; j < n; j++) {
    float frac = 1.f;
    if (j == 0) frac = (float)(left_i + 1) - left_f;
    if (j == n - 1) frac = right_f - (float)(left_i + j);
    if (n == 1) frac = delta;
    value = fmaf(frac, in[(size_t)(left_i + j) * stride], value);
  }
  return value * normalization;
}

void oracle_resample(const float* in, size_t iw, size_t ih, float* out, size_t ow, size_t oh, size_t pitch) {
  /* x pass: out_w x in_h into a temp (cuda_operation_resample_2d.cpp:99-101) */
  float* tmp = (float*)malloc(sizeof(float) * pitch * (ih > oh ? ih : oh));
#pragma omp parallel for schedule(static)
  for (long y = 0; y < (long)ih; y++)
    for (size_t x = 0; x < ow; x++) tmp[IDX(x, y)] = resample_1d(in + IDX(0, y), 1, iw, ow, x);
  /* y pass (cuda_operation_resample_2d.cpp:103-105) */
#pragma omp parallel for schedule(static)
  for (long y = 0; y < (long)oh; y++)
    for (size_t x = 0; x < ow; x++) out[IDX(x, y)] = resample_1d(tmp + x, pitch, ih, oh, (size_t)y);
  free(tmp);
}

/* ---------------------------------------------------------------------------------------------
 * Backward bilinear registration.  registration_2d.cu:48-72.  As compiled:
 * x_f = fma(rcp.rn(hx), u, (float)x); bound = (float)(w-1); blend = fma(w11,f11, fma(w01,f01,
 * fma(f00,w00, w10*f10))).
 * ------------------------------------------------------------------------------------------- */
void oracle_warp(const float* f0, const float* f1, const float* u, const float* v,
                 size_t w, size_t h, size_t pitch, float hx, float hy, float* out) {
  const float rhx = 1.f / hx, rhy = 1.f / hy;
  const float bx = (float)(w - 1), by = (float)(h - 1);
#pragma omp parallel for schedule(static)
  for (long yy = 0; yy < (long)h; yy++)
    for (size_t xx = 0; xx < w; xx++) {
      float x_f = fmaf(rhx, u[IDX(xx, yy)], (float)xx);
      float y_f = fmaf(rhy, v[IDX(xx, yy)], (float)yy);
      if ((x_f < 0.f) || (x_f > bx) || (y_f < 0.f) || (y_f > by) || isnan(x_f) || isnan(y_f)) {
        out[IDX(xx, yy)] = f0[IDX(xx, yy)];
      } else {
        int x = (int)floorf(x_f);
        int y = (int)floorf(y_f);
        float dx = x_f - (float)x;
        float dy = y_f - (float)y;
        int x1 = (int)(w - 1) < x + 1 ? (int)(w - 1) : x + 1;
        int y1 = (int)(h - 1) < y + 1 ? (int)(h - 1) : y + 1;
        float ox = 1.f - dx, oy = 1.f - dy;
        float w00 = ox * oy, w10 = dx * oy, w01 = ox * dy, w11 = dx * dy;
        float val = w10 * f1[IDX(x1, y)];
        val = fmaf(f1[IDX(x, y)], w00, val);
        val = fmaf(w01, f1[IDX(x, y1)], val);
        val = fmaf(w11, f1[IDX(x1, y1)], val);
        out[IDX(xx, yy)] = val;
      }
    }
}

/* ---------------------------------------------------------------------------------------------
 * Robust weights.  solve_2d.cu:139-197.
 * ------------------------------------------------------------------------------------------- */
void oracle_phi_ksi(const float* f0, const float* f1, const float* u, const float* v,
                    const float* du, const float* dv, size_t w, size_t h, size_t pitch,
                    float hx, float hy, float e_smooth, float e_data, float* phi, float* ksi) {
  const float hx2 = hx + hx, hy2 = hy + hy; /* 2.f*hx compiles to add.f32 hx,hx */
  const float hx4 = hx * 4.f, hy4 = hy * 4.f;
  long W = (long)w, Hh = (long)h;
#pragma omp parallel for schedule(static)
  for (long y = 0; y < Hh; y++) {
    long ym = mirror(y - 1, Hh), yp = mirror(y + 1, Hh);
    for (long x = 0; x < W; x++) {
      long xm = mirror(x - 1, W), xp = mirror(x + 1, W);
      float dux = (((u[IDX(xp, y)] - u[IDX(xm, y)]) + du[IDX(xp, y)]) - du[IDX(xm, y)]) / hx2;
      float duy = (((u[IDX(x, yp)] - u[IDX(x, ym)]) + du[IDX(x, yp)]) - du[IDX(x, ym)]) / hy2;
      float dvx = (((v[IDX(xp, y)] - v[IDX(xm, y)]) + dv[IDX(xp, y)]) - dv[IDX(xm, y)]) / hx2;
      float dvy = (((v[IDX(x, yp)] - v[IDX(x, ym)]) + dv[IDX(x, yp)]) - dv[IDX(x, ym)]) / hy2;
      float t = duy * duy;
      t = fmaf(dux, dux, t);
      t = fmaf(dvx, dvx, t);
      t = fmaf(dvy, dvy, t);
      t = fmaf(e_smooth, e_smooth, t);
      float r = sqrtf(t);
      phi[IDX(x, y)] = 1.f / (r + r);

      float fx = (((f0[IDX(xp, y)] - f0[IDX(xm, y)]) + f1[IDX(xp, y)]) - f1[IDX(xm, y)]) / hx4;
      float fy = (((f0[IDX(x, yp)] - f0[IDX(x, ym)]) + f1[IDX(x, yp)]) - f1[IDX(x, ym)]) / hy4;
      float ft = f1[IDX(x, y)] - f0[IDX(x, y)];
      float J11 = fx * fx, J22 = fy * fy, J12 = fx * fy, J13 = fx * ft, J23 = fy * ft;
      float d_u = du[IDX(x, y)], d_v = dv[IDX(x, y)];
      float a = J13 + fmaf(J11, d_u, J12 * d_v);
      float b = J23 + fmaf(J12, d_u, J22 * d_v);
      float c = fmaf(ft, ft, fmaf(J13, d_u, J23 * d_v));
      float s = fmaf(d_u, a, d_v * b) + c;
      s = s * ((s > 0.f) ? 1.f : 0.f);
      float q = sqrtf(fmaf(e_data, e_data, s));
      ksi[IDX(x, y)] = 1.f / (q + q);
    }
  }
}

/* One pixel of the lagged linear system, given the motion tensor entries.
 * solve_2d.cu:332-374 (identical text in solve_2d_grad, 889-931). */
static inline void jacobi_update(const float* u, const float* v, const float* du, const float* dv,
                                 const float* phi, const float* ksi, long x, long y, long W, long Hh,
                                 size_t pitch, float hx_2, float hy_2,
                                 float J11, float J22, float J12, float J13, float J23,
                                 float* out_du, float* out_dv) {
  long xm = mirror(x - 1, W), xp = mirror(x + 1, W);
  long ym = mirror(y - 1, Hh), yp = mirror(y + 1, Hh);
  float wxp = hx_2 * ((x < W - 1) ? 1.f : 0.f);
  float wxm = hx_2 * ((x > 0) ? 1.f : 0.f);
  float wyp = hy_2 * ((y < Hh - 1) ? 1.f : 0.f);
  float wym = hy_2 * ((y > 0) ? 1.f : 0.f);
  float pc = phi[IDX(x, y)];
  float phi_xp = (phi[IDX(xp, y)] + pc) * 0.5f;
  float phi_xm = (phi[IDX(xm, y)] + pc) * 0.5f;
  float phi_yp = (phi[IDX(x, yp)] + pc) * 0.5f;
  float phi_ym = (phi[IDX(x, ym)] + pc) * 0.5f;
  float axp = wxp * phi_xp, axm = wxm * phi_xm, ayp = wyp * phi_yp, aym = wym * phi_ym;
  float sumH = ((axp + axm) + ayp) + aym;

  float uc = u[IDX(x, y)], vc = v[IDX(x, y)];
  float sumU = axm * ((u[IDX(xm, y)] + du[IDX(xm, y)]) - uc);
  sumU = fmaf(axp, (u[IDX(xp, y)] + du[IDX(xp, y)]) - uc, sumU);
  sumU = fmaf(ayp, (u[IDX(x, yp)] + du[IDX(x, yp)]) - uc, sumU);
  sumU = fmaf(aym, (u[IDX(x, ym)] + du[IDX(x, ym)]) - uc, sumU);
  float sumV = axm * ((v[IDX(xm, y)] + dv[IDX(xm, y)]) - vc);
  sumV = fmaf(axp, (v[IDX(xp, y)] + dv[IDX(xp, y)]) - vc, sumV);
  sumV = fmaf(ayp, (v[IDX(x, yp)] + dv[IDX(x, yp)]) - vc, sumV);
  sumV = fmaf(aym, (v[IDX(x, ym)] + dv[IDX(x, ym)]) - vc, sumV);

  float k = ksi[IDX(x, y)];
  /* ptxas fuses (-J13) - J12*dv into one FFMA(-J12, dv, -J13) */
  float r_du = fmaf(k, fmaf(-J12, dv[IDX(x, y)], -J13), sumU) / fmaf(J11, k, sumH);
  float r_dv = fmaf(k, fmaf(-J12, r_du, -J23), sumV) / fmaf(J22, k, sumH);
  *out_du = r_du;
  *out_dv = r_dv;
}

/* solve_2d.cu:308-375 */
void oracle_sweep_grey(const float* f0, const float* f1, const float* u, const float* v,
                       const float* du, const float* dv, const float* phi, const float* ksi,
                       size_t w, size_t h, size_t pitch, float hx, float hy, float alpha,
                       float* du_out, float* dv_out) {
  const float hx4 = hx * 4.f, hy4 = hy * 4.f;
  const float hx_2 = alpha / (hx * hx), hy_2 = alpha / (hy * hy);
  long W = (long)w, Hh = (long)h;
#pragma omp parallel for schedule(static)
  for (long y = 0; y < Hh; y++) {
    long ym = mirror(y - 1, Hh), yp = mirror(y + 1, Hh);
    for (long x = 0; x < W; x++) {
      long xm = mirror(x - 1, W), xp = mirror(x + 1, W);
      float fx = (((f0[IDX(xp, y)] - f0[IDX(xm, y)]) + f1[IDX(xp, y)]) - f1[IDX(xm, y)]) / hx4;
      float fy = (((f0[IDX(x, yp)] - f0[IDX(x, ym)]) + f1[IDX(x, yp)]) - f1[IDX(x, ym)]) / hy4;
      float ft = f1[IDX(x, y)] - f0[IDX(x, y)];
      jacobi_update(u, v, du, dv, phi, ksi, x, y, W, Hh, pitch, hx_2, hy_2,
                    fx * fx, fy * fy, fx * fy, fx * ft, fy * ft,
                    &du_out[IDX(x, y)], &dv_out[IDX(x, y)]);
    }
  }
}

/* solve_2d.cu:795-884.  fx/fy/ft planes per 16x8 CUDA block; the 1-px halo of those planes is the
 * block-edge thread's OWN value (813-841), not the true neighbour.  In a partial block the cell
 * just outside the image is never written by the reference (uninitialised shared memory); the
 * oracle defines it as the own value too and parity there is reported separately (F5). */
void oracle_sweep_grad(const float* f0, const float* f1, const float* u, const float* v,
                       const float* du, const float* dv, const float* phi, const float* ksi,
                       size_t w, size_t h, size_t pitch, float hx, float hy, float alpha,
                       float* du_out, float* dv_out) {
  const float hx4 = hx * 4.f, hy4 = hy * 4.f;
  const float hx_2 = alpha / (hx * hx), hy_2 = alpha / (hy * hy);
  const float hx_1 = (float)(1.0 / (2.0 * (double)hx)), hy_1 = (float)(1.0 / (2.0 * (double)hy));
  long W = (long)w, Hh = (long)h;
  float* FX = (float*)malloc(sizeof(float) * pitch * h);
  float* FY = (float*)malloc(sizeof(float) * pitch * h);
  float* FT = (float*)malloc(sizeof(float) * pitch * h);
#pragma omp parallel for schedule(static)
  for (long y = 0; y < Hh; y++) {
    long ym = mirror(y - 1, Hh), yp = mirror(y + 1, Hh);
    for (long x = 0; x < W; x++) {
      long xm = mirror(x - 1, W), xp = mirror(x + 1, W);
      FX[IDX(x, y)] = (((f0[IDX(xp, y)] - f0[IDX(xm, y)]) + f1[IDX(xp, y)]) - f1[IDX(xm, y)]) / hx4;
      FY[IDX(x, y)] = (((f0[IDX(x, yp)] - f0[IDX(x, ym)]) + f1[IDX(x, yp)]) - f1[IDX(x, ym)]) / hy4;
      FT[IDX(x, y)] = f1[IDX(x, y)] - f0[IDX(x, y)];
    }
  }
#pragma omp parallel for schedule(static)
  for (long y = 0; y < Hh; y++)
    for (long x = 0; x < W; x++) {
      long tx = x % 16, ty = y % 8;
      long xl = (tx == 0) ? x : x - 1;
      long xr = (tx == 15 || x + 1 >= W) ? x : x + 1;
      long yu = (ty == 0) ? y : y - 1;
      long yb = (ty == 7 || y + 1 >= Hh) ? y : y + 1;
      float fxx = (FX[IDX(xr, y)] - FX[IDX(xl, y)]) * hx_1;
      float fxy = (FX[IDX(x, yb)] - FX[IDX(x, yu)]) * hy_1;
      float fyy = (FY[IDX(x, yb)] - FY[IDX(x, yu)]) * hy_1;
      float fxt = (FT[IDX(xr, y)] - FT[IDX(xl, y)]) * hx_1;
      float fyt = (FT[IDX(x, yb)] - FT[IDX(x, yu)]) * hy_1;
      float J11 = fmaf(fxx, fxx, fxy * fxy);
      float J22 = fmaf(fxy, fxy, fyy * fyy);
      float J12 = fmaf(fxx, fxy, fxy * fyy);
      float J13 = fmaf(fxx, fxt, fxy * fyt);
      float J23 = fmaf(fxy, fxt, fyy * fyt);
      jacobi_update(u, v, du, dv, phi, ksi, x, y, W, Hh, pitch, hx_2, hy_2, J11, J22, J12, J13, J23,
                    &du_out[IDX(x, y)], &dv_out[IDX(x, y)]);
    }
  free(FX);
  free(FY);
  free(FT);
}

/* EXTENSION (see the header): residual of the system jacobi_update relaxes. */
void oracle_residual(const float* f0, const float* f1, const float* u, const float* v, const float* du,
                     const float* dv, const float* phi, const float* ksi, size_t w, size_t h, size_t pitch,
                     float hx, float hy, float alpha, int constancy, double* rms_u, double* rms_v) {
  const float hx4 = hx * 4.f, hy4 = hy * 4.f;
  const float hx_2 = alpha / (hx * hx), hy_2 = alpha / (hy * hy);
  const float hx_1 = (float)(1.0 / (2.0 * (double)hx)), hy_1 = (float)(1.0 / (2.0 * (double)hy));
  long W = (long)w, Hh = (long)h;
  float* FX = (float*)malloc(sizeof(float) * pitch * h);
  float* FY = (float*)malloc(sizeof(float) * pitch * h);
  float* FT = (float*)malloc(sizeof(float) * pitch * h);
  for (long y = 0; y < Hh; y++) {
    long ym = mirror(y - 1, Hh), yp = mirror(y + 1, Hh);
    for (long x = 0; x < W; x++) {
      long xm = mirror(x - 1, W), xp = mirror(x + 1, W);
      FX[IDX(x, y)] = (((f0[IDX(xp, y)] - f0[IDX(xm, y)]) + f1[IDX(xp, y)]) - f1[IDX(xm, y)]) / hx4;
      FY[IDX(x, y)] = (((f0[IDX(x, yp)] - f0[IDX(x, ym)]) + f1[IDX(x, yp)]) - f1[IDX(x, ym)]) / hy4;
      FT[IDX(x, y)] = f1[IDX(x, y)] - f0[IDX(x, y)];
    }
  }
  double su2 = 0.0, sv2 = 0.0;
  for (long y = 0; y < Hh; y++)
    for (long x = 0; x < W; x++) {
      float J11, J22, J12, J13, J23;
      if (constancy == ORACLE_GRADIENT) {
        long tx = x % 16, ty = y % 8;
        long xl = (tx == 0) ? x : x - 1, xr = (tx == 15 || x + 1 >= W) ? x : x + 1;
        long yu = (ty == 0) ? y : y - 1, yb = (ty == 7 || y + 1 >= Hh) ? y : y + 1;
        float fxx = (FX[IDX(xr, y)] - FX[IDX(xl, y)]) * hx_1, fxy = (FX[IDX(x, yb)] - FX[IDX(x, yu)]) * hy_1;
        float fyy = (FY[IDX(x, yb)] - FY[IDX(x, yu)]) * hy_1;
        float fxt = (FT[IDX(xr, y)] - FT[IDX(xl, y)]) * hx_1, fyt = (FT[IDX(x, yb)] - FT[IDX(x, yu)]) * hy_1;
        J11 = fmaf(fxx, fxx, fxy * fxy); J22 = fmaf(fxy, fxy, fyy * fyy); J12 = fmaf(fxx, fxy, fxy * fyy);
        J13 = fmaf(fxx, fxt, fxy * fyt); J23 = fmaf(fxy, fxt, fyy * fyt);
      } else {
        float fx = FX[IDX(x, y)], fy = FY[IDX(x, y)], ft = FT[IDX(x, y)];
        J11 = fx * fx; J22 = fy * fy; J12 = fx * fy; J13 = fx * ft; J23 = fy * ft;
      }
      long xm = mirror(x - 1, W), xp = mirror(x + 1, W), ym = mirror(y - 1, Hh), yp = mirror(y + 1, Hh);
      float wxp = hx_2 * ((x < W - 1) ? 1.f : 0.f), wxm = hx_2 * ((x > 0) ? 1.f : 0.f);
      float wyp = hy_2 * ((y < Hh - 1) ? 1.f : 0.f), wym = hy_2 * ((y > 0) ? 1.f : 0.f);
      float pc = phi[IDX(x, y)];
      double axp = wxp * ((phi[IDX(xp, y)] + pc) * 0.5f), axm = wxm * ((phi[IDX(xm, y)] + pc) * 0.5f);
      double ayp = wyp * ((phi[IDX(x, yp)] + pc) * 0.5f), aym = wym * ((phi[IDX(x, ym)] + pc) * 0.5f);
      double k = ksi[IDX(x, y)], d_u = du[IDX(x, y)], d_v = dv[IDX(x, y)];
      double sU = (double)u[IDX(x, y)] + d_u, sV = (double)v[IDX(x, y)] + d_v;
#define SU(X, Y) ((double)u[IDX(X, Y)] + (double)du[IDX(X, Y)])
#define SV(X, Y) ((double)v[IDX(X, Y)] + (double)dv[IDX(X, Y)])
      double lap_u = axm * (SU(xm, y) - sU) + axp * (SU(xp, y) - sU) + ayp * (SU(x, yp) - sU) + aym * (SU(x, ym) - sU);
      double lap_v = axm * (SV(xm, y) - sV) + axp * (SV(xp, y) - sV) + ayp * (SV(x, yp) - sV) + aym * (SV(x, ym) - sV);
#undef SU
#undef SV
      double ru = k * (-(double)J13 - (double)J12 * d_v - (double)J11 * d_u) + lap_u;
      double rv = k * (-(double)J23 - (double)J12 * d_u - (double)J22 * d_v) + lap_v;
      su2 += ru * ru;
      sv2 += rv * rv;
    }
  free(FX);
  free(FY);
  free(FT);
  *rms_u = sqrt(su2 / ((double)w * (double)h));
  *rms_v = sqrt(sv2 / ((double)w * (double)h));
}

/* cuda_operation_solve_2d.cpp:229-299 */
void oracle_solve_level(const float* f0, const float* f1, const float* u, const float* v,
                        float* du, float* dv, float* phi, float* ksi, float* tmp_du, float* tmp_dv,
                        size_t w, size_t h, size_t pitch, float hx, float hy,
                        const oracle_params* p) {
  float *a_du = du, *a_dv = dv, *b_du = tmp_du, *b_dv = tmp_dv;
  for (size_t y = 0; y < h; y++) {
    memset(a_du + IDX(0, y), 0, w * sizeof(float));
    memset(a_dv + IDX(0, y), 0, w * sizeof(float));
  }
  for (size_t i = 0; i < p->outer_iterations_count; ++i) {
    /* the reference always builds ksi from the brightness tensor (solve_2d.cu:164-196, F5) */
    oracle_phi_ksi(f0, f1, u, v, a_du, a_dv, w, h, pitch, hx, hy, p->equation_smoothness,
                   p->equation_data, phi, ksi);
    for (size_t j = 0; j < p->inner_iterations_count; ++j) {
      if (p->constancy == ORACLE_GRADIENT)
        oracle_sweep_grad(f0, f1, u, v, a_du, a_dv, phi, ksi, w, h, pitch, hx, hy, p->equation_alpha, b_du, b_dv);
      else
        oracle_sweep_grey(f0, f1, u, v, a_du, a_dv, phi, ksi, w, h, pitch, hx, hy, p->equation_alpha, b_du, b_dv);
      float* t;
      t = a_du; a_du = b_du; b_du = t;
      t = a_dv; a_dv = b_dv; b_dv = t;
    }
  }
  if (a_du != du) {
    for (size_t y = 0; y < h; y++) {
      memcpy(du + IDX(0, y), a_du + IDX(0, y), w * sizeof(float));
      memcpy(dv + IDX(0, y), a_dv + IDX(0, y), w * sizeof(float));
    }
  }
}

/* add_2d.cu:43-44 */
void oracle_add(float* a, const float* b, size_t w, size_t h, size_t pitch) {
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++) a[IDX(x, y)] += b[IDX(x, y)];
}

/* median_2d.cu:281-297: gather the radius x radius window (mirror boundary), insertion sort,
 * take element len/2.  cuda_operation_median_2d.cpp:100-111: radius 1 = copy, even radius -= 1,
 * 3..7 supported. */
int oracle_median(const float* in, float* out, size_t w, size_t h, size_t pitch, size_t radius) {
  if (radius == 1) {
    for (size_t y = 0; y < h; y++) memcpy(out + IDX(0, y), in + IDX(0, y), w * sizeof(float));
    return 0;
  }
  if (radius % 2 == 0) radius -= 1;
  if (radius < 3 || radius > 7) return 1;
  long r2 = (long)(radius / 2), W = (long)w, Hh = (long)h, R = (long)radius;
#pragma omp parallel for schedule(static)
  for (long y = 0; y < Hh; y++)
    for (long x = 0; x < W; x++) {
      float buffer[49];
      for (long iy = 0; iy < R; ++iy)
        for (long ix = 0; ix < R; ++ix)
          buffer[iy * R + ix] = in[IDX(mirror(x - ix + r2, W), mirror(y - iy + r2, Hh))];
      long len = R * R;
      for (long i = 0; i < len; i++) { /* insertionSort, median_2d.cu:54-65 */
        float temp = buffer[i];
        long j;
        for (j = i - 1; j >= 0 && temp < buffer[j]; j--) buffer[j + 1] = buffer[j];
        buffer[j + 1] = temp;
      }
      out[IDX(x, y)] = buffer[len / 2];
    }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * The whole path.  optical_flow_2d.cpp:142-569.
 * ------------------------------------------------------------------------------------------- */
int oracle_compute_flow(const float* f0_in, const float* f1_in, size_t W, size_t H,
                        const oracle_params* p, float* u_out, float* v_out) {
  const size_t pitch = W, N = W * H;
  float* buf = (float*)calloc(12 * N, sizeof(float));
  if (!buf) return -1;
  float *frame0 = buf, *frame1 = buf + N, *frame0_res = buf + 2 * N, *frame1_res = buf + 3 * N;
  float *u = buf + 4 * N, *v = buf + 5 * N, *du = buf + 6 * N, *dv = buf + 7 * N;
  float *t0 = buf + 8 * N, *t1 = buf + 9 * N, *t2 = buf + 10 * N, *t3 = buf + 11 * N;
  float* sw;
#define SWAP(a, b) do { sw = a; a = b; b = sw; } while (0)
  memcpy(frame0, f0_in, N * sizeof(float));
  memcpy(frame1, f1_in, N * sizeof(float));

  /* presmoothing, optical_flow_2d.cpp:218-246 */
  if (p->gaussian_sigma > 0.0f) {
    oracle_blur(frame0, u, W, H, pitch, p->gaussian_sigma);
    oracle_blur(frame1, v, W, H, pitch, p->gaussian_sigma);
    SWAP(frame0, u);
    SWAP(frame1, v);
  }

  size_t max_level = oracle_max_warp_level(W, H, p->warp_scale_factor);
  int level = (int)(p->warp_levels_count < max_level ? p->warp_levels_count : max_level) - 1;
  size_t pw = 0, ph = 0;

  while (level >= 0) {
    size_t cw, ch;
    float hx, hy;
    oracle_level_geometry(W, H, p->warp_scale_factor, level, &cw, &ch, &hx, &hy);

    /* frames: always resampled from the ORIGINAL size (279-305) */
    if (level == 0) {
      SWAP(frame0, frame0_res);
      SWAP(frame1, frame1_res);
    } else {
      oracle_resample(frame0, W, H, frame0_res, cw, ch, pitch);
      oracle_resample(frame1, W, H, frame1_res, cw, ch, pitch);
    }
    /* flow: zero at the first level, else prolongated from the previous level (308-341) */
    if (pw == 0) {
      memset(u, 0, N * sizeof(float));
      memset(v, 0, N * sizeof(float));
    } else {
      oracle_resample(u, pw, ph, du, cw, ch, pitch);
      oracle_resample(v, pw, ph, dv, cw, ch, pitch);
      SWAP(u, du);
      SWAP(v, dv);
    }
    /* backward registration (344-363) */
    oracle_warp(frame0_res, frame1_res, u, v, cw, ch, pitch, hx, hy, t0);
    SWAP(frame1_res, t0);
    /* solve (366-406): result in du, dv */
    oracle_solve_level(frame0_res, frame1_res, u, v, du, dv, t0, t1, t2, t3, cw, ch, pitch, hx, hy, p);
    /* add (409-422) */
    oracle_add(u, du, cw, ch, pitch);
    oracle_add(v, dv, cw, ch, pitch);
    pw = cw;
    ph = ch;
    --level;
    /* median (428-449); an unsupported radius still swaps the stale temp in (F9) */
    oracle_median(u, t0, cw, ch, pitch, p->median_radius);
    SWAP(u, t0);
    oracle_median(v, t0, cw, ch, pitch, p->median_radius);
    SWAP(v, t0);
  }
  if (pw != 0) { /* F9: with no level run the reference returns stale memory; the oracle returns 0 */
    memcpy(u_out, u, N * sizeof(float));
    memcpy(v_out, v, N * sizeof(float));
  } else {
    memset(u_out, 0, N * sizeof(float));
    memset(v_out, 0, N * sizeof(float));
  }
#undef SWAP
  free(buf);
  return 0;
}

/*
 * flow2d_oracle_ext.c -- CPU oracle of the OPT-IN EXTENSIONS (TEST INFRASTRUCTURE ONLY, see flow2d_oracle.h).
 *
 * Nothing in this file restates the reference: SURVEY.md 8(f) ranks 3 and 4 name solver features that
 * axruff/cuda-flow2d does not have (a relaxation factor, red-black ordering, a convergence test, level-to-level
 * restriction: the reference runs fixed Jacobi counts, cuda_operation_solve_2d.cpp:229-299, solve_2d.cu:361-374, and
 * restricts every level from the original frames, optical_flow_2d.cpp:279-305) or has only in a broken form
 * (solve_2d_log, solve_2d.cu:391-669: derivative halos taken from the 16x8 CUDA block instead of the neighbours; the
 * README's brightness + gradient model, README.md:32-34, is never combined in code).  This file DEFINES those
 * extensions operation by operation so that the sm_100a kernels of csrc/solve_ext.cu can be checked bit for bit
 * (the log term to a tolerance: libm's log and CUDA's differ in the last place).  It is a specification with a
 * checker's role, never reference parity.
 *
 * Built with -ffp-contract=off like flow2d_oracle.c: every fused multiply-add is an explicit fmaf().
 */
#include "flow2d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define IDX(x, y) ((size_t)(y) * pitch + (size_t)(x))

static inline long mirror(long i, long n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - i - 2;
  return i;
}

/* log(1 + f) of a frame (the transform of solve_2d.cu:509-524), evaluated in double and rounded once; intensities below
 * zero (the reference would produce NaN for f <= -1) count as zero */
static void log_frame(const float* in, float* out, size_t w, size_t h, size_t pitch) {
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++) out[IDX(x, y)] = (float)log((double)fmaxf(in[IDX(x, y)], 0.f) + 1.0);
}

/* Motion tensor of a level, six planes J11 J22 J12 J13 J23 J33.  The first derivatives are the reference's
 * (solve_2d.cu:308-331); the second derivatives are central differences of those planes over the TRUE neighbours,
 * mirrored at the image border (the reference takes the block-edge thread's own value, F5). */
void oracle_ext_tensor(const float* f0, const float* f1w, size_t w, size_t h, size_t pitch, float hx, float hy,
                       int data_term, float gamma, float* const J[6]) {
  const float hx4 = hx * 4.f, hy4 = hy * 4.f;
  const float hx_1 = (float)(1.0 / (2.0 * (double)hx)), hy_1 = (float)(1.0 / (2.0 * (double)hy));
  long W = (long)w, Hh = (long)h;
  float* FX = (float*)malloc(sizeof(float) * pitch * h);
  float* FY = (float*)malloc(sizeof(float) * pitch * h);
  float* FT = (float*)malloc(sizeof(float) * pitch * h);
  float *g0 = NULL, *g1 = NULL;
  if (data_term == ORACLE_TERM_LOG_GRADIENT) {
    g0 = (float*)malloc(sizeof(float) * pitch * h);
    g1 = (float*)malloc(sizeof(float) * pitch * h);
    log_frame(f0, g0, w, h, pitch);
    log_frame(f1w, g1, w, h, pitch);
    f0 = g0;
    f1w = g1;
  }
  for (long y = 0; y < Hh; y++) {
    long ym = mirror(y - 1, Hh), yp = mirror(y + 1, Hh);
    for (long x = 0; x < W; x++) {
      long xm = mirror(x - 1, W), xp = mirror(x + 1, W);
      FX[IDX(x, y)] = (((f0[IDX(xp, y)] - f0[IDX(xm, y)]) + f1w[IDX(xp, y)]) - f1w[IDX(xm, y)]) / hx4;
      FY[IDX(x, y)] = (((f0[IDX(x, yp)] - f0[IDX(x, ym)]) + f1w[IDX(x, yp)]) - f1w[IDX(x, ym)]) / hy4;
      FT[IDX(x, y)] = f1w[IDX(x, y)] - f0[IDX(x, y)];
    }
  }
  for (long y = 0; y < Hh; y++) {
    long ym = mirror(y - 1, Hh), yp = mirror(y + 1, Hh);
    for (long x = 0; x < W; x++) {
      long xm = mirror(x - 1, W), xp = mirror(x + 1, W);
      const float fx = FX[IDX(x, y)], fy = FY[IDX(x, y)], ft = FT[IDX(x, y)];
      const float B[6] = {fx * fx, fy * fy, fx * fy, fx * ft, fy * ft, ft * ft};
      float G[6] = {0, 0, 0, 0, 0, 0};
      if (data_term != ORACLE_TERM_DEFAULT) {
        const float fxx = (FX[IDX(xp, y)] - FX[IDX(xm, y)]) * hx_1, fxy = (FX[IDX(x, yp)] - FX[IDX(x, ym)]) * hy_1;
        const float fyy = (FY[IDX(x, yp)] - FY[IDX(x, ym)]) * hy_1;
        const float fxt = (FT[IDX(xp, y)] - FT[IDX(xm, y)]) * hx_1, fyt = (FT[IDX(x, yp)] - FT[IDX(x, ym)]) * hy_1;
        G[0] = fmaf(fxx, fxx, fxy * fxy);
        G[1] = fmaf(fxy, fxy, fyy * fyy);
        G[2] = fmaf(fxx, fxy, fxy * fyy);
        G[3] = fmaf(fxx, fxt, fxy * fyt);
        G[4] = fmaf(fxy, fxt, fyy * fyt);
        G[5] = fmaf(fxt, fxt, fyt * fyt);
      }
      for (int k = 0; k < 6; k++) {
        float v;
        if (data_term == ORACLE_TERM_DEFAULT) v = B[k];
        else if (data_term == ORACLE_TERM_COMBINED) v = fmaf(gamma, G[k], B[k]);
        else v = G[k];
        J[k][IDX(x, y)] = v;
      }
    }
  }
  free(FX); free(FY); free(FT);
  free(g0); free(g1);
}

/* Robust weights from the tensor planes: phi as solve_2d.cu:141-162, ksi = psi'(w^T J w) with w = (du, dv, 1) and the
 * FULL tensor (the reference always uses the brightness tensor here, solve_2d.cu:164-196). */
void oracle_ext_phi_ksi(const float* const J[6], const float* u, const float* v, const float* du, const float* dv,
                        size_t w, size_t h, size_t pitch, float hx, float hy, float e_smooth, float e_data, float* phi,
                        float* ksi) {
  const float hx2 = hx + hx, hy2 = hy + hy;
  long W = (long)w, Hh = (long)h;
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

      const size_t c = IDX(x, y);
      const float J11 = J[0][c], J22 = J[1][c], J12 = J[2][c], J13 = J[3][c], J23 = J[4][c], J33 = J[5][c];
      const float d_u = du[c], d_v = dv[c];
      float a = J13 + fmaf(J11, d_u, J12 * d_v);
      float b = J23 + fmaf(J12, d_u, J22 * d_v);
      float cc = J33 + fmaf(J13, d_u, J23 * d_v);
      float s = fmaf(d_u, a, d_v * b) + cc;
      s = s * ((s > 0.f) ? 1.f : 0.f);
      float q = sqrtf(fmaf(e_data, e_data, s));
      ksi[c] = 1.f / (q + q);
    }
  }
}

/* One relaxation of one pixel of the lagged system: the Jacobi / Gauss-Seidel value of solve_2d.cu:332-374 blended with
 * the old value by omega.  Neighbours are read from (du, dv): the previous iterate (Jacobi, out of place) or whatever is
 * there (red-black, in place). */
static inline void relax_px(const float* const J[6], const float* u, const float* v, const float* du, const float* dv,
                            const float* phi, const float* ksi, long x, long y, long W, long Hh, size_t pitch,
                            float hx_2, float hy_2, float omega, float* out_du, float* out_dv) {
  long xm = mirror(x - 1, W), xp = mirror(x + 1, W);
  long ym = mirror(y - 1, Hh), yp = mirror(y + 1, Hh);
  float wxp = hx_2 * ((x < W - 1) ? 1.f : 0.f), wxm = hx_2 * ((x > 0) ? 1.f : 0.f);
  float wyp = hy_2 * ((y < Hh - 1) ? 1.f : 0.f), wym = hy_2 * ((y > 0) ? 1.f : 0.f);
  const size_t c = IDX(x, y);
  float pc = phi[c];
  float axp = wxp * ((phi[IDX(xp, y)] + pc) * 0.5f), axm = wxm * ((phi[IDX(xm, y)] + pc) * 0.5f);
  float ayp = wyp * ((phi[IDX(x, yp)] + pc) * 0.5f), aym = wym * ((phi[IDX(x, ym)] + pc) * 0.5f);
  float sumH = ((axp + axm) + ayp) + aym;
  float uc = u[c], vc = v[c];
  float sumU = axm * ((u[IDX(xm, y)] + du[IDX(xm, y)]) - uc);
  sumU = fmaf(axp, (u[IDX(xp, y)] + du[IDX(xp, y)]) - uc, sumU);
  sumU = fmaf(ayp, (u[IDX(x, yp)] + du[IDX(x, yp)]) - uc, sumU);
  sumU = fmaf(aym, (u[IDX(x, ym)] + du[IDX(x, ym)]) - uc, sumU);
  float sumV = axm * ((v[IDX(xm, y)] + dv[IDX(xm, y)]) - vc);
  sumV = fmaf(axp, (v[IDX(xp, y)] + dv[IDX(xp, y)]) - vc, sumV);
  sumV = fmaf(ayp, (v[IDX(x, yp)] + dv[IDX(x, yp)]) - vc, sumV);
  sumV = fmaf(aym, (v[IDX(x, ym)] + dv[IDX(x, ym)]) - vc, sumV);
  const float k = ksi[c], J11 = J[0][c], J22 = J[1][c], J12 = J[2][c], J13 = J[3][c], J23 = J[4][c];
  const float du_old = du[c], dv_old = dv[c];
  float r_du = fmaf(k, fmaf(-J12, dv_old, -J13), sumU) / fmaf(J11, k, sumH);
  if (omega != 1.f) r_du = fmaf(omega, r_du - du_old, du_old);
  float r_dv = fmaf(k, fmaf(-J12, r_du, -J23), sumV) / fmaf(J22, k, sumH);
  if (omega != 1.f) r_dv = fmaf(omega, r_dv - dv_old, dv_old);
  *out_du = r_du;
  *out_dv = r_dv;
}

/* RMS residual of the lagged system with tensor planes (double arithmetic on the fp32 fields, as oracle_residual) */
void oracle_ext_residual(const float* const J[6], const float* u, const float* v, const float* du, const float* dv,
                         const float* phi, const float* ksi, size_t w, size_t h, size_t pitch, float hx, float hy,
                         float alpha, double* rms_u, double* rms_v) {
  const float hx_2 = alpha / (hx * hx), hy_2 = alpha / (hy * hy);
  long W = (long)w, Hh = (long)h;
  double su2 = 0.0, sv2 = 0.0;
  for (long y = 0; y < Hh; y++)
    for (long x = 0; x < W; x++) {
      long xm = mirror(x - 1, W), xp = mirror(x + 1, W), ym = mirror(y - 1, Hh), yp = mirror(y + 1, Hh);
      float wxp = hx_2 * ((x < W - 1) ? 1.f : 0.f), wxm = hx_2 * ((x > 0) ? 1.f : 0.f);
      float wyp = hy_2 * ((y < Hh - 1) ? 1.f : 0.f), wym = hy_2 * ((y > 0) ? 1.f : 0.f);
      const size_t c = IDX(x, y);
      float pc = phi[c];
      double axp = wxp * ((phi[IDX(xp, y)] + pc) * 0.5f), axm = wxm * ((phi[IDX(xm, y)] + pc) * 0.5f);
      double ayp = wyp * ((phi[IDX(x, yp)] + pc) * 0.5f), aym = wym * ((phi[IDX(x, ym)] + pc) * 0.5f);
      double k = ksi[c], d_u = du[c], d_v = dv[c];
      double sU = (double)u[c] + d_u, sV = (double)v[c] + d_v;
#define SU(X, Y) ((double)u[IDX(X, Y)] + (double)du[IDX(X, Y)])
#define SV(X, Y) ((double)v[IDX(X, Y)] + (double)dv[IDX(X, Y)])
      double lap_u = axm * (SU(xm, y) - sU) + axp * (SU(xp, y) - sU) + ayp * (SU(x, yp) - sU) + aym * (SU(x, ym) - sU);
      double lap_v = axm * (SV(xm, y) - sV) + axp * (SV(xp, y) - sV) + ayp * (SV(x, yp) - sV) + aym * (SV(x, ym) - sV);
#undef SU
#undef SV
      double ru = k * (-(double)J[3][c] - (double)J[2][c] * d_v - (double)J[0][c] * d_u) + lap_u;
      double rv = k * (-(double)J[4][c] - (double)J[2][c] * d_u - (double)J[1][c] * d_v) + lap_v;
      su2 += ru * ru;
      sv2 += rv * rv;
    }
  *rms_u = sqrt(su2 / ((double)w * (double)h));
  *rms_v = sqrt(sv2 / ((double)w * (double)h));
}

static int ext_solver_needed(const oracle_ext* e) {
  return e->scheme != ORACLE_SCHEME_JACOBI || (e->omega != 0.f && e->omega != 1.f) || e->data_term != ORACLE_TERM_DEFAULT;
}

/* The solve of one level with the extensions.  Returns the number of outer iterations that ran.
 *   - no extension that changes the relaxation (Jacobi, omega 1, default term): the reference's own iteration
 *     (oracle_phi_ksi + oracle_sweep_grey / _grad), optionally ended early by the residual test;
 *   - otherwise: tensor planes + oracle_ext_phi_ksi + relax_px in Jacobi (double-buffered) or red-black (in place, the
 *     cells with even x+y first) order.
 * Early exit: after every `residual_check_every`-th outer iteration the RMS residual of that iteration's lagged system
 * is formed; the level ends when both components are <= residual_tolerance. */
int oracle_ext_solve_level(const float* f0, const float* f1w, const float* u, const float* v, float* du, float* dv,
                           float* phi, float* ksi, size_t w, size_t h, size_t pitch, float hx, float hy,
                           const oracle_params* p, const oracle_ext* e) {
  const size_t n = pitch * h;
  const float omega = (e->omega == 0.f) ? 1.f : e->omega;
  const int every = e->residual_check_every > 0 ? e->residual_check_every : 1;
  const int ext = ext_solver_needed(e);
  float* Jbuf = NULL;
  float* J[6] = {0, 0, 0, 0, 0, 0};
  float* tdu = (float*)calloc(n, sizeof(float));
  float* tdv = (float*)calloc(n, sizeof(float));
  float *a_du = du, *a_dv = dv, *b_du = tdu, *b_dv = tdv;
  const float hx_2 = p->equation_alpha / (hx * hx), hy_2 = p->equation_alpha / (hy * hy);
  long W = (long)w, Hh = (long)h;
  int used = 0;
  if (ext) {
    Jbuf = (float*)calloc(6 * n, sizeof(float));
    for (int k = 0; k < 6; k++) J[k] = Jbuf + (size_t)k * n;
    oracle_ext_tensor(f0, f1w, w, h, pitch, hx, hy, e->data_term, e->gamma, J);
  }
  for (size_t y = 0; y < h; y++) {
    memset(a_du + IDX(0, y), 0, w * sizeof(float));
    memset(a_dv + IDX(0, y), 0, w * sizeof(float));
  }
  for (size_t o = 0; o < p->outer_iterations_count; ++o) {
    if (ext) oracle_ext_phi_ksi((const float* const*)J, u, v, a_du, a_dv, w, h, pitch, hx, hy, p->equation_smoothness, p->equation_data, phi, ksi);
    else oracle_phi_ksi(f0, f1w, u, v, a_du, a_dv, w, h, pitch, hx, hy, p->equation_smoothness, p->equation_data, phi, ksi);
    for (size_t j = 0; j < p->inner_iterations_count; ++j) {
      if (!ext) {
        if (p->constancy == ORACLE_GRADIENT)
          oracle_sweep_grad(f0, f1w, u, v, a_du, a_dv, phi, ksi, w, h, pitch, hx, hy, p->equation_alpha, b_du, b_dv);
        else
          oracle_sweep_grey(f0, f1w, u, v, a_du, a_dv, phi, ksi, w, h, pitch, hx, hy, p->equation_alpha, b_du, b_dv);
      } else if (e->scheme == ORACLE_SCHEME_JACOBI) {
        for (long y = 0; y < Hh; y++)
          for (long x = 0; x < W; x++)
            relax_px((const float* const*)J, u, v, a_du, a_dv, phi, ksi, x, y, W, Hh, pitch, hx_2, hy_2, omega,
                     &b_du[IDX(x, y)], &b_dv[IDX(x, y)]);
      } else {
        for (int colour = 0; colour < 2; colour++)
          for (long y = 0; y < Hh; y++)
            for (long x = 0; x < W; x++)
              if (((x + y) & 1) == colour)
                relax_px((const float* const*)J, u, v, a_du, a_dv, phi, ksi, x, y, W, Hh, pitch, hx_2, hy_2, omega,
                         &a_du[IDX(x, y)], &a_dv[IDX(x, y)]);
      }
      if (!ext || e->scheme == ORACLE_SCHEME_JACOBI) {
        float* t;
        t = a_du; a_du = b_du; b_du = t;
        t = a_dv; a_dv = b_dv; b_dv = t;
      }
    }
    used = (int)o + 1;
    if (e->residual_tolerance > 0.f && p->inner_iterations_count > 0 && used % every == 0) {
      double ru, rv;
      if (ext) oracle_ext_residual((const float* const*)J, u, v, a_du, a_dv, phi, ksi, w, h, pitch, hx, hy, p->equation_alpha, &ru, &rv);
      else oracle_residual(f0, f1w, u, v, a_du, a_dv, phi, ksi, w, h, pitch, hx, hy, p->equation_alpha, p->constancy, &ru, &rv);
      if (ru <= (double)e->residual_tolerance && rv <= (double)e->residual_tolerance) break;
    }
  }
  if (a_du != du)
    for (size_t y = 0; y < h; y++) {
      memcpy(du + IDX(0, y), a_du + IDX(0, y), w * sizeof(float));
      memcpy(dv + IDX(0, y), a_dv + IDX(0, y), w * sizeof(float));
    }
  free(tdu); free(tdv); free(Jbuf);
  return used;
}

/* The whole path (oracle_compute_flow's stage order) with the extensions.  outer_used (optional, `cap` entries):
 * outer iterations that ran per level, coarsest level first. */
int oracle_ext_compute_flow(const float* f0_in, const float* f1_in, size_t W, size_t H, const oracle_params* p,
                            const oracle_ext* e, float* u_out, float* v_out, int* outer_used, int cap) {
  const size_t pitch = W, N = W * H;
  size_t max_level = oracle_max_warp_level(W, H, p->warp_scale_factor);
  const int levels = (int)(p->warp_levels_count < max_level ? p->warp_levels_count : max_level);
  float* buf = (float*)calloc(12 * N, sizeof(float));
  if (!buf) return -1;
  float *frame0 = buf, *frame1 = buf + N, *r0 = buf + 2 * N, *r1 = buf + 3 * N;
  float *u = buf + 4 * N, *v = buf + 5 * N, *du = buf + 6 * N, *dv = buf + 7 * N;
  float *t0 = buf + 8 * N, *t1 = buf + 9 * N, *t2 = buf + 10 * N, *t3 = buf + 11 * N;
  float* sw;
#define SWAP(a, b) do { sw = a; a = b; b = sw; } while (0)
  memcpy(frame0, f0_in, N * sizeof(float));
  memcpy(frame1, f1_in, N * sizeof(float));
  if (p->gaussian_sigma > 0.0f) {
    oracle_blur(frame0, u, W, H, pitch, p->gaussian_sigma);
    oracle_blur(frame1, v, W, H, pitch, p->gaussian_sigma);
    SWAP(frame0, u);
    SWAP(frame1, v);
  }
  /* cascaded restriction: level l from level l-1 (the reference restricts every level from the original) */
  float** pyr0 = NULL; float** pyr1 = NULL;
  if (e->cascaded_restriction && levels > 1) {
    pyr0 = (float**)calloc((size_t)levels, sizeof(float*));
    pyr1 = (float**)calloc((size_t)levels, sizeof(float*));
    size_t lw = W, lh = H;
    const float *s0 = frame0, *s1 = frame1;
    for (int l = 1; l < levels; l++) {
      size_t cw, ch; float hx, hy;
      oracle_level_geometry(W, H, p->warp_scale_factor, l, &cw, &ch, &hx, &hy);
      pyr0[l] = (float*)calloc(N, sizeof(float));
      pyr1[l] = (float*)calloc(N, sizeof(float));
      oracle_resample(s0, lw, lh, pyr0[l], cw, ch, pitch);
      oracle_resample(s1, lw, lh, pyr1[l], cw, ch, pitch);
      s0 = pyr0[l]; s1 = pyr1[l]; lw = cw; lh = ch;
    }
  }
  int level = levels - 1, li = 0;
  size_t pw = 0, ph = 0;
  while (level >= 0) {
    size_t cw, ch;
    float hx, hy;
    oracle_level_geometry(W, H, p->warp_scale_factor, level, &cw, &ch, &hx, &hy);
    const float *l0, *l1;
    if (level == 0) { l0 = frame0; l1 = frame1; }
    else if (pyr0) { l0 = pyr0[level]; l1 = pyr1[level]; }
    else {
      oracle_resample(frame0, W, H, r0, cw, ch, pitch);
      oracle_resample(frame1, W, H, r1, cw, ch, pitch);
      l0 = r0; l1 = r1;
    }
    if (pw == 0) {
      memset(u, 0, N * sizeof(float));
      memset(v, 0, N * sizeof(float));
    } else {
      oracle_resample(u, pw, ph, du, cw, ch, pitch);
      oracle_resample(v, pw, ph, dv, cw, ch, pitch);
      SWAP(u, du);
      SWAP(v, dv);
    }
    oracle_warp(l0, l1, u, v, cw, ch, pitch, hx, hy, t0);
    const int used = oracle_ext_solve_level(l0, t0, u, v, du, dv, t1, t2, cw, ch, pitch, hx, hy, p, e);
    if (outer_used && li < cap) outer_used[li] = used;
    ++li;
    oracle_add(u, du, cw, ch, pitch);
    oracle_add(v, dv, cw, ch, pitch);
    pw = cw; ph = ch;
    --level;
    oracle_median(u, t3, cw, ch, pitch, p->median_radius);
    SWAP(u, t3);
    oracle_median(v, t3, cw, ch, pitch, p->median_radius);
    SWAP(v, t3);
  }
  memcpy(u_out, u, N * sizeof(float));
  memcpy(v_out, v, N * sizeof(float));
#undef SWAP
  if (pyr0) {
    for (int l = 1; l < levels; l++) { free(pyr0[l]); free(pyr1[l]); }
    free(pyr0); free(pyr1);
  }
  free(buf);
  return li;
}

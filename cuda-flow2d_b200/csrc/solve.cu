// solve.cu -- the inner solver: robust weights (phi, ksi) + temporally blocked Jacobi sweeps of the
// lagged nonlinear system, fused into one kernel.  sm_100a, compiled with -fmad=false.
//
// Replaces compute_phi_ksi (src/kernels/solve_2d.cu:43-198) and solve_2d / solve_2d_grad
// (src/kernels/solve_2d.cu:200-377, 683-953) as driven by CudaOperationSolve2D::Execute
// (src/cuda_operations/2d/cuda_operation_solve_2d.cpp:229-299), with identical results:
// the reference's scheme is a double-buffered JACOBI sweep (all neighbour reads from the previous
// iterate) with one Gauss-Seidel coupling inside the pixel (dv' uses du'), no relaxation factor.
// Jacobi is tiling independent, so S sweeps can be run on a shared-memory tile with an S-pixel
// halo and give bit-identical values to S separate global sweeps.
//
// One CTA = 512 threads = one 64x64 region L of the level (output tile O plus halo).  A thread owns
// two 1x4 pixel strips for the whole pass; everything that is constant per pixel during the sweeps
// lives in registers or in thread-private shared-memory planes, and only s_u = u+du, s_v = v+dv are
// exchanged: x neighbours by warp shuffle, y neighbours through double-buffered shared planes.
//
//   phase A  load u, v, du, dv (strip-wise float4) -> registers + shared planes; fx, fy, ft -> regs
//   phase B  phi on L minus apron, ksi; phi -> shared plane         (or load phi/ksi of an earlier pass)
//   phase C  edge weights, sumH, the two denominators, -J12, -J13, -J23
//   phase D  S sweeps; sweep k updates O grown by S-k pixels (clipped to the image)
//   phase E  store du, dv (and phi, ksi if a later pass of the same outer iteration needs them)
//
// "Resident" mode: if the whole level fits one region, a single CTA runs ALL outer iterations and
// all inner sweeps of the level without leaving the SM (grid = 1).
#include "kernels.h"

namespace flow2d {

constexpr int LW = kSolveLW, LH = kSolveLH;  // 64 x 64
constexpr int NT = 512;
constexpr int PL = LW * LH;  // floats per shared plane

// shared planes
enum {
  P_U = 0, P_V, P_DU, P_DV,        // phase A/B: neighbour access for phi
  P_PHI,                           // phase B/C
  P_SU0, P_SV0, P_SU1, P_SV1,      // phase D: double-buffered s_u, s_v
  P_NJ13, P_NJ23,                  // thread-private: -J13, -J23
  kNumPlanes,
  // thread-private planes aliased onto the phase A/B planes (dead after the barrier that ends phase B)
  P_EYP = P_U, P_EYM = P_V, P_DENU = P_DU, P_DENV = P_DV,
  // phase B -> C hand-over of J11, J22 through planes that are first written by sweep 1
  P_J11 = P_SU1, P_J22 = P_SV1
};

size_t solve_pass_smem_bytes() { return sizeof(float) * PL * kNumPlanes; }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void unpack(const float4& q, float (&v)[4]) {
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}

// Own strip (gx..gx+3, gy) of a global plane.  Cells outside the image are clamped to valid
// memory; their values are never used in arithmetic (see nb_* below).
__device__ __forceinline__ void load_strip(const float* __restrict__ p, int gx, int gy, int w, int h, int pitch,
                                           float (&v)[4]) {
  const int my = min(max(gy, 0), h - 1);
  const float* row = p + (size_t)my * pitch;
  if (gx >= 0 && gx + 3 < w) {
    unpack(ld4(row + gx), v);
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = row[min(max(gx + i, 0), w - 1)];
  }
}

// Neighbour selection with the reference's mirrored border (index -1 -> 1, n -> n-2): at the image
// border the mirrored neighbour IS the opposite neighbour, so no out-of-image value is ever used.
__device__ __forceinline__ float nb_lo(bool at_lo_border, float lo, float hi) { return at_lo_border ? hi : lo; }
__device__ __forceinline__ float nb_hi(bool at_hi_border, float lo, float hi) { return at_hi_border ? lo : hi; }

template <bool GRAD>
__global__ void __launch_bounds__(NT, 1) solve_pass_kernel(const SolveArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int lane_c = tid & 15;  // strip column within the region
  const int lx = 4 * lane_c;
  const int w = a.w, h = a.h, pitch = a.pitch;

  const int ox0 = blockIdx.x * a.ow, oy0 = blockIdx.y * a.oh;
  const int ox1 = min(w, ox0 + a.ow), oy1 = min(h, oy0 + a.oh);
  const int lx0 = ox0 - a.halo_x, ly0 = oy0 - a.halo_y;
  const int gx = lx0 + lx;  // multiple of 4

  // per-strip constants
  int gy[2], soff[2];
  gy[0] = ly0 + (tid >> 4);
  gy[1] = gy[0] + 32;
  soff[0] = (tid >> 4) * LW + lx;
  soff[1] = soff[0] + 32 * LW;

  // per-pixel flags: inside the image and inside the region minus its 1-cell apron
  bool live[2][4];
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int x = gx + i, y = gy[s], lxx = lx + i, lyy = y - ly0;
      live[s][i] = x >= 0 && x < w && y >= 0 && y < h && lxx >= 1 && lxx <= LW - 2 && lyy >= 1 && lyy <= LH - 2;
    }

  const float hx2 = a.hx + a.hx, hy2 = a.hy + a.hy;
  const float hx_2 = a.alpha / (a.hx * a.hx), hy_2 = a.alpha / (a.hy * a.hy);

  float uc[2][4], vc[2][4], du[2][4], dv[2][4];
#pragma unroll
  for (int s = 0; s < 2; s++) {
    load_strip(a.u, gx, gy[s], w, h, pitch, uc[s]);
    load_strip(a.v, gx, gy[s], w, h, pitch, vc[s]);
    if (a.du_in) {
      load_strip(a.du_in, gx, gy[s], w, h, pitch, du[s]);
      load_strip(a.dv_in, gx, gy[s], w, h, pitch, dv[s]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) du[s][i] = dv[s][i] = 0.f;
    }
  }

  float ksi[2][4], nJ12[2][4], su[2][4], sv[2][4], ex[2][5];

  for (int outer = 0; outer < a.outer; ++outer) {
    // ---------------- phase A: publish u, v, du, dv ----------------
    if (outer > 0) __syncthreads();  // previous iteration's readers of the aliased planes are done
#pragma unroll
    for (int s = 0; s < 2; s++) {
      st4(sm + P_U * PL + soff[s], uc[s]);
      st4(sm + P_V * PL + soff[s], vc[s]);
      st4(sm + P_DU * PL + soff[s], du[s]);
      st4(sm + P_DV * PL + soff[s], dv[s]);
    }
    __syncthreads();

    // ---------------- phase B: phi, ksi ----------------
    float phi[2][4];
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const int y = gy[s];
      const bool y_lo = (y == 0), y_hi = (y == h - 1);
      float fx[4], fy[4], ft[4];
      load_strip(a.fx, gx, y, w, h, pitch, fx);
      load_strip(a.fy, gx, y, w, h, pitch, fy);
      load_strip(a.ft, gx, y, w, h, pitch, ft);

      if (a.phi_in) {
        load_strip(a.phi_in, gx, y, w, h, pitch, phi[s]);
        load_strip(a.ksi_in, gx, y, w, h, pitch, ksi[s]);
      } else {
        // x neighbours of the strip ends come from the adjacent lanes
        const float uL = __shfl_up_sync(0xffffffffu, uc[s][3], 1), uR = __shfl_down_sync(0xffffffffu, uc[s][0], 1);
        const float vL = __shfl_up_sync(0xffffffffu, vc[s][3], 1), vR = __shfl_down_sync(0xffffffffu, vc[s][0], 1);
        const float duL = __shfl_up_sync(0xffffffffu, du[s][3], 1), duR = __shfl_down_sync(0xffffffffu, du[s][0], 1);
        const float dvL = __shfl_up_sync(0xffffffffu, dv[s][3], 1), dvR = __shfl_down_sync(0xffffffffu, dv[s][0], 1);
        float uU[4], uD[4], vU[4], vD[4], duU[4], duD[4], dvU[4], dvD[4];
        // rows 0 and LH-1 of the region are apron (never live): clamp their neighbour row
        const int up = soff[s] - ((tid >> 4) + 32 * s > 0 ? LW : 0);
        const int dn = soff[s] + ((tid >> 4) + 32 * s < LH - 1 ? LW : 0);
        unpack(ld4(sm + P_U * PL + up), uU); unpack(ld4(sm + P_V * PL + up), vU);
        unpack(ld4(sm + P_DU * PL + up), duU); unpack(ld4(sm + P_DV * PL + up), dvU);
        unpack(ld4(sm + P_U * PL + dn), uD); unpack(ld4(sm + P_V * PL + dn), vD);
        unpack(ld4(sm + P_DU * PL + dn), duD); unpack(ld4(sm + P_DV * PL + dn), dvD);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          phi[s][i] = 0.f;
          ksi[s][i] = 0.f;
          if (live[s][i]) {
            const int x = gx + i;
            const bool x_lo = (x == 0), x_hi = (x == w - 1);
            const float ul_raw = i > 0 ? uc[s][i - 1] : uL, ur_raw = i < 3 ? uc[s][i + 1] : uR;
            const float vl_raw = i > 0 ? vc[s][i - 1] : vL, vr_raw = i < 3 ? vc[s][i + 1] : vR;
            const float dul_raw = i > 0 ? du[s][i - 1] : duL, dur_raw = i < 3 ? du[s][i + 1] : duR;
            const float dvl_raw = i > 0 ? dv[s][i - 1] : dvL, dvr_raw = i < 3 ? dv[s][i + 1] : dvR;
            const float u_l = nb_lo(x_lo, ul_raw, ur_raw), u_r = nb_hi(x_hi, ul_raw, ur_raw);
            const float v_l = nb_lo(x_lo, vl_raw, vr_raw), v_r = nb_hi(x_hi, vl_raw, vr_raw);
            const float du_l = nb_lo(x_lo, dul_raw, dur_raw), du_r = nb_hi(x_hi, dul_raw, dur_raw);
            const float dv_l = nb_lo(x_lo, dvl_raw, dvr_raw), dv_r = nb_hi(x_hi, dvl_raw, dvr_raw);
            const float u_u = nb_lo(y_lo, uU[i], uD[i]), u_d = nb_hi(y_hi, uU[i], uD[i]);
            const float v_u = nb_lo(y_lo, vU[i], vD[i]), v_d = nb_hi(y_hi, vU[i], vD[i]);
            const float du_u = nb_lo(y_lo, duU[i], duD[i]), du_d = nb_hi(y_hi, duU[i], duD[i]);
            const float dv_u = nb_lo(y_lo, dvU[i], dvD[i]), dv_d = nb_hi(y_hi, dvU[i], dvD[i]);
            // solve_2d.cu:141-162
            const float dux = (((u_r - u_l) + du_r) - du_l) / hx2;
            const float duy = (((u_d - u_u) + du_d) - du_u) / hy2;
            const float dvx = (((v_r - v_l) + dv_r) - dv_l) / hx2;
            const float dvy = (((v_d - v_u) + dv_d) - dv_u) / hy2;
            float t = duy * duy;
            t = fmaf(dux, dux, t);
            t = fmaf(dvx, dvx, t);
            t = fmaf(dvy, dvy, t);
            t = fmaf(a.e_smooth, a.e_smooth, t);
            const float r = sqrtf(t);
            phi[s][i] = 1.f / (r + r);
            // solve_2d.cu:176-196 (always the brightness tensor, also in gradient mode)
            const float j11 = fx[i] * fx[i], j22 = fy[i] * fy[i], j12 = fx[i] * fy[i];
            const float j13 = fx[i] * ft[i], j23 = fy[i] * ft[i];
            const float d_u = du[s][i], d_v = dv[s][i];
            const float ta = j13 + fmaf(j11, d_u, j12 * d_v);
            const float tb = j23 + fmaf(j12, d_u, j22 * d_v);
            const float tc = fmaf(ft[i], ft[i], fmaf(j13, d_u, j23 * d_v));
            float sq = fmaf(d_u, ta, d_v * tb) + tc;
            sq = sq * ((sq > 0.f) ? 1.f : 0.f);
            const float q = sqrtf(fmaf(a.e_data, a.e_data, sq));
            ksi[s][i] = 1.f / (q + q);
          }
        }
      }
      // motion tensor of the sweep (solve_2d.cu:324-329 / 879-884)
      {
        float J11[4], J22[4], nJ13[4], nJ23[4];
        if (GRAD) {
          float j12[4];
          load_strip(a.J[0], gx, y, w, h, pitch, J11);
          load_strip(a.J[1], gx, y, w, h, pitch, J22);
          load_strip(a.J[2], gx, y, w, h, pitch, j12);
          load_strip(a.J[3], gx, y, w, h, pitch, nJ13);
          load_strip(a.J[4], gx, y, w, h, pitch, nJ23);
#pragma unroll
          for (int i = 0; i < 4; i++) { nJ12[s][i] = -j12[i]; nJ13[i] = -nJ13[i]; nJ23[i] = -nJ23[i]; }
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            J11[i] = fx[i] * fx[i];
            J22[i] = fy[i] * fy[i];
            nJ12[s][i] = -(fx[i] * fy[i]);
            nJ13[i] = -(fx[i] * ft[i]);
            nJ23[i] = -(fy[i] * ft[i]);
          }
        }
        st4(sm + P_J11 * PL + soff[s], J11);
        st4(sm + P_J22 * PL + soff[s], J22);
        st4(sm + P_NJ13 * PL + soff[s], nJ13);
        st4(sm + P_NJ23 * PL + soff[s], nJ23);
      }
      st4(sm + P_PHI * PL + soff[s], phi[s]);
    }
    __syncthreads();  // phi published; P_U..P_DV are dead from here on

    // ---------------- phase C: weights and denominators ----------------
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const int y = gy[s];
      const bool y_lo = (y == 0), y_hi = (y == h - 1);
      const float pL = __shfl_up_sync(0xffffffffu, phi[s][3], 1), pR = __shfl_down_sync(0xffffffffu, phi[s][0], 1);
      float pU[4], pD[4];
      unpack(ld4(sm + P_PHI * PL + soff[s] - ((tid >> 4) + 32 * s > 0 ? LW : 0)), pU);
      unpack(ld4(sm + P_PHI * PL + soff[s] + ((tid >> 4) + 32 * s < LH - 1 ? LW : 0)), pD);
      const float wyp = hy_2 * ((y < h - 1) ? 1.f : 0.f), wym = hy_2 * ((y > 0) ? 1.f : 0.f);
      float eyp[4], eym[4], denU[4], denV[4], J11[4], J22[4];
      unpack(ld4(sm + P_J11 * PL + soff[s]), J11);
      unpack(ld4(sm + P_J22 * PL + soff[s]), J22);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int x = gx + i;
        const bool x_lo = (x == 0), x_hi = (x == w - 1);
        const float pl_raw = i > 0 ? phi[s][i - 1] : pL, pr_raw = i < 3 ? phi[s][i + 1] : pR;
        const float p_l = nb_lo(x_lo, pl_raw, pr_raw), p_r = nb_hi(x_hi, pl_raw, pr_raw);
        const float p_u = nb_lo(y_lo, pU[i], pD[i]), p_d = nb_hi(y_hi, pU[i], pD[i]);
        const float pc = phi[s][i];
        // solve_2d.cu:333-349
        const float wxp = hx_2 * ((x < w - 1) ? 1.f : 0.f), wxm = hx_2 * ((x > 0) ? 1.f : 0.f);
        const float axp = wxp * ((p_r + pc) * 0.5f);
        const float axm = wxm * ((p_l + pc) * 0.5f);
        eyp[i] = wyp * ((p_d + pc) * 0.5f);
        eym[i] = wym * ((p_u + pc) * 0.5f);
        // axm(x) == axp(x-1) bit for bit for x >= 1 (same products, commuted add), so only the
        // strip's first axm is kept separately
        if (i == 0) ex[s][0] = axm;
        ex[s][i + 1] = axp;
        const float sumH = ((axp + axm) + eyp[i]) + eym[i];
        denU[i] = fmaf(J11[i], ksi[s][i], sumH);
        denV[i] = fmaf(J22[i], ksi[s][i], sumH);
        su[s][i] = uc[s][i] + du[s][i];
        sv[s][i] = vc[s][i] + dv[s][i];
      }
      st4(sm + P_EYP * PL + soff[s], eyp);
      st4(sm + P_EYM * PL + soff[s], eym);
      st4(sm + P_DENU * PL + soff[s], denU);
      st4(sm + P_DENV * PL + soff[s], denV);
      st4(sm + P_SU0 * PL + soff[s], su[s]);
      st4(sm + P_SV0 * PL + soff[s], sv[s]);
      if (a.phi_out && !a.phi_in) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int x = gx + i;
          if (x >= ox0 && x < ox1 && y >= oy0 && y < oy1) {
            a.phi_out[(size_t)y * pitch + x] = phi[s][i];
            a.ksi_out[(size_t)y * pitch + x] = ksi[s][i];
          }
        }
      }
    }
    __syncthreads();

    // ---------------- phase D: Jacobi sweeps ----------------
    for (int k = 1; k <= a.sweeps; ++k) {
      const int grow = a.sweeps - k;
      const int wx0 = max(0, ox0 - grow), wx1 = min(w, ox1 + grow);
      const int wy0 = max(0, oy0 - grow), wy1 = min(h, oy1 + grow);
      const float* cur_u = sm + ((k & 1) ? P_SU0 : P_SU1) * PL;
      const float* cur_v = sm + ((k & 1) ? P_SV0 : P_SV1) * PL;
      float* nxt_u = sm + ((k & 1) ? P_SU1 : P_SU0) * PL;
      float* nxt_v = sm + ((k & 1) ? P_SV1 : P_SV0) * PL;
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const int y = gy[s];
        const bool y_lo = (y == 0), y_hi = (y == h - 1);
        const float suL = __shfl_up_sync(0xffffffffu, su[s][3], 1), suR = __shfl_down_sync(0xffffffffu, su[s][0], 1);
        const float svL = __shfl_up_sync(0xffffffffu, sv[s][3], 1), svR = __shfl_down_sync(0xffffffffu, sv[s][0], 1);
        const bool row_active = y >= wy0 && y < wy1;
        float nu[4], nv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { nu[i] = su[s][i]; nv[i] = sv[s][i]; }
        if (row_active && gx + 3 >= wx0 && gx < wx1) {
          float suU[4], suD[4], svU[4], svD[4], eyp[4], eym[4], denU[4], denV[4], nJ13v[4], nJ23v[4];
          unpack(ld4(cur_u + soff[s] - LW), suU); unpack(ld4(cur_u + soff[s] + LW), suD);
          unpack(ld4(cur_v + soff[s] - LW), svU); unpack(ld4(cur_v + soff[s] + LW), svD);
          unpack(ld4(sm + P_EYP * PL + soff[s]), eyp); unpack(ld4(sm + P_EYM * PL + soff[s]), eym);
          unpack(ld4(sm + P_DENU * PL + soff[s]), denU); unpack(ld4(sm + P_DENV * PL + soff[s]), denV);
          unpack(ld4(sm + P_NJ13 * PL + soff[s]), nJ13v); unpack(ld4(sm + P_NJ23 * PL + soff[s]), nJ23v);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int x = gx + i;
            if (x >= wx0 && x < wx1) {
              const bool x_lo = (x == 0), x_hi = (x == w - 1);
              const float sul_raw = i > 0 ? su[s][i - 1] : suL, sur_raw = i < 3 ? su[s][i + 1] : suR;
              const float svl_raw = i > 0 ? sv[s][i - 1] : svL, svr_raw = i < 3 ? sv[s][i + 1] : svR;
              const float axm = ex[s][i], axp = ex[s][i + 1];
              const float u0 = uc[s][i], v0 = vc[s][i];
              // solve_2d.cu:350-367 as compiled (fma chain; (-J13) - J12*dv fused by ptxas)
              float sumU = axm * (nb_lo(x_lo, sul_raw, sur_raw) - u0);
              sumU = fmaf(axp, nb_hi(x_hi, sul_raw, sur_raw) - u0, sumU);
              sumU = fmaf(eyp[i], nb_hi(y_hi, suU[i], suD[i]) - u0, sumU);
              sumU = fmaf(eym[i], nb_lo(y_lo, suU[i], suD[i]) - u0, sumU);
              float sumV = axm * (nb_lo(x_lo, svl_raw, svr_raw) - v0);
              sumV = fmaf(axp, nb_hi(x_hi, svl_raw, svr_raw) - v0, sumV);
              sumV = fmaf(eyp[i], nb_hi(y_hi, svU[i], svD[i]) - v0, sumV);
              sumV = fmaf(eym[i], nb_lo(y_lo, svU[i], svD[i]) - v0, sumV);
              const float kk = ksi[s][i];
              const float r_du = fmaf(kk, fmaf(nJ12[s][i], dv[s][i], nJ13v[i]), sumU) / denU[i];
              const float r_dv = fmaf(kk, fmaf(nJ12[s][i], r_du, nJ23v[i]), sumV) / denV[i];
              du[s][i] = r_du;
              dv[s][i] = r_dv;
              nu[i] = u0 + r_du;
              nv[i] = v0 + r_dv;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) { su[s][i] = nu[i]; sv[s][i] = nv[i]; }
        st4(nxt_u + soff[s], nu);
        st4(nxt_v + soff[s], nv);
      }
      __syncthreads();
    }
  }

  // ---------------- phase E: store du, dv of the output tile ----------------
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int y = gy[s];
    if (y < oy0 || y >= oy1) continue;
    float* rdu = a.du_out + (size_t)y * pitch;
    float* rdv = a.dv_out + (size_t)y * pitch;
    if (gx >= ox0 && gx + 3 < ox1) {
      st4(rdu + gx, du[s]);
      st4(rdv + gx, dv[s]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int x = gx + i;
        if (x >= ox0 && x < ox1) { rdu[x] = du[s][i]; rdv[x] = dv[s][i]; }
      }
    }
  }
}

cudaError_t solve_pass_configure() {
  cudaError_t e = cudaFuncSetAttribute(solve_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)solve_pass_smem_bytes());
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(solve_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)solve_pass_smem_bytes());
}

void launch_solve_pass(cudaStream_t st, const SolveArgs& a, bool grad, int grid_x, int grid_y) {
  dim3 grid(grid_x, grid_y), block(NT);
  if (grad) solve_pass_kernel<true><<<grid, block, solve_pass_smem_bytes(), st>>>(a);
  else solve_pass_kernel<false><<<grid, block, solve_pass_smem_bytes(), st>>>(a);
}

}  // namespace flow2d

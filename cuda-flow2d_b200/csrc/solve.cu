// solve.cu -- the inner solver: robust weights (phi, ksi) + temporally blocked Jacobi sweeps of the
// lagged nonlinear system, fused into one kernel.  sm_100a, compiled with -fmad=false.
//
// Replaces compute_phi_ksi (src/kernels/solve_2d.cu:43-198) and solve_2d / solve_2d_grad
// (src/kernels/solve_2d.cu:200-377, 683-953) as driven by CudaOperationSolve2D::Execute
// (src/cuda_operations/2d/cuda_operation_solve_2d.cpp:229-299), with identical results:
// the reference's scheme is a double-buffered JACOBI sweep (all neighbour reads from the previous
// iterate) with one Gauss-Seidel coupling inside the pixel (dv' uses du'), no relaxation factor.
//
// Temporal blocking.  One CTA = 768 threads = one 64x48 region of the level; a thread owns one 1x4
// pixel strip for the whole pass.  EVERY cell of the region is updated in EVERY sweep, without any
// activity mask: a cell's value after sweep k is exact iff its inputs were exact, so under Jacobi
// the set of exact cells simply shrinks by one ring per sweep from the region's edge (phi needs one
// ring, the edge weights another, then one per sweep).  With a halo of S+1 cells the output tile is
// exact after S sweeps and bit-identical to S separate global sweeps; what happens in the rings
// outside it is never stored.  At the image border nothing shrinks: the reference's mirrored
// border (index -1 -> 1, n -> n-2) means "the mirrored neighbour IS the opposite neighbour", which
// is a row-offset choice in y and a register select in x (only in CTAs that touch the border).
// Cells of a border region that lie outside the image are made inert (zero weights, ksi = 0,
// denominator 1) so that they stay finite and cheap.
//
// Everything that is constant per pixel during the sweeps lives in registers or in thread-private
// shared-memory planes (conflict-free LDS.128); only s_u = u+du, s_v = v+dv are exchanged:
// x neighbours by warp shuffle, y neighbours through double-buffered shared planes, one
// __syncthreads per sweep.
//
//   phase A  the loads of the pass: planes that are only handed over or published go global -> shared by
//            cp.async; a first pass of an outer iteration loads fx, fy, ft, du, dv into registers and
//            computes ksi
//   phase B  u/v/du/dv published for the neighbour rows; phi
//            (skipped when phi/ksi of an earlier pass of the same outer iteration were staged by phase A)
//   phase C  motion tensor, edge weights, sumH, the two denominators and their fast-path reciprocals
//   phase D  S Jacobi sweeps (unrolled by buffer parity; PTX ld/st.shared with immediate plane offsets)
//   phase E  store du, dv of the output tile
//
// Mid-size and tiny levels use the one-pixel-per-thread kernels further down (solve_small_pass, solve_tiny).
//
// "Resident" mode: if the whole level fits one region, a single CTA runs ALL outer iterations and
// all inner sweeps of the level without leaving the SM (grid = 1).
#include <type_traits>

#include "kernels.h"
#include "solve_common.cuh"
#include "solve_onepx.cuh"

namespace flow2d {

constexpr int NT = LH * (LW / 4);            // one thread per 1x4 strip: 768 threads, <= 80 registers each

// shared planes (14 x 12 KiB = 168 KiB; the rest of the 228 KiB stays L1)
enum {
  P_SU0 = 0, P_SV0, P_SU1, P_SV1,   // phase D: double-buffered s_u, s_v (read by the rows above/below)
  P_EYP, P_EYM,                     // thread-private from here on: y edge weights
  P_DENU, P_DENV, P_RU, P_RV,       // the two denominators and their fast-path reciprocals
  P_NJ13, P_NJ23, P_KSI, P_NJ12,
  kNumPlanes,
  // phase B planes, aliased onto planes that are first written after the barrier ending phase B
  P_U = P_EYP, P_V = P_EYM, P_DU = P_DENU, P_DV = P_DENV,
  P_PHI = P_SU1,                    // read in phase C; sweep 1 is the first writer of SU1
  P_J11 = P_RU, P_J22 = P_RV,       // thread-private hand-over A -> C (gradient constancy: the loaded tensor)
  P_FX = P_RU, P_FY = P_RV, P_FT = P_NJ12  // same for brightness constancy: fx, fy, ft
};

size_t solve_pass_smem_bytes() { return sizeof(float) * PL * kNumPlanes; }

struct Strip {  // per-thread state that lives in registers during the sweeps
  float uc[4], vc[4], dv[4];
  float ksi[4], nJ12[4];
  float su[4], sv[4];
  float ex[5];  // x edge weights: ex[i] lies between pixels x-1+i and x+i
  bool den_ok;  // all eight denominators of the strip are safe for the fast division path
};

// x neighbours of strip element i: inside the strip from registers, at its ends from the adjacent
// lanes (L, R); at the image border (BORDER CTAs only) the mirrored neighbour is the opposite one.
template <bool BORDER>
__device__ __forceinline__ void x_nb(const float (&c)[4], float L, float R, int i, bool x_lo, int i_hi, float& l,
                                     float& r) {
  l = i > 0 ? c[i - 1] : L;
  r = i < 3 ? c[i + 1] : R;
  if (BORDER) {
    const float l0 = l;
    l = (x_lo && i == 0) ? r : l;
    r = (i == i_hi) ? l0 : r;
  }
}

template <bool GRAD, bool BORDER, bool TIMING>
__device__ __forceinline__ void pass_body(const SolveArgs& a, float* sm) {
  const int tid = threadIdx.x;
  if (TIMING) stamp(a, 0);
  const int row = tid >> 4;       // row of the region
  const int lx = 4 * (tid & 15);  // first column of the strip within the region
  const int w = a.w, h = a.h, pitch = a.pitch;

  const int ox0 = blockIdx.x * a.ow, oy0 = a.y0 + blockIdx.y * a.oh;
  const int ox1 = min(w, ox0 + a.ow), oy1 = min(a.y1, oy0 + a.oh);
  const int gx = ox0 - a.halo_x + lx;  // multiple of 4
  const int gy = oy0 - a.halo_y + row;
  const unsigned sb = (unsigned)__cvta_generic_to_shared(sm) + 4u * (unsigned)(row * LW + lx);  // this strip, plane 0
  const int last_row = (int)(blockDim.x >> 4) - 1;  // LH-1, or fewer rows for a small resident level
  // Neighbour rows.  Image border: the mirrored neighbour is the opposite neighbour.  The first and
  // last row of the region have no neighbour row in shared memory (they are never exact anyway).
  int up_off = row > 0 ? -LW : LW, dn_off = row < last_row ? LW : -LW;
  if (BORDER) {
    if (gy == 0 && row < last_row) up_off = LW;
    if (gy == h - 1 && row > 0) dn_off = -LW;
  }
  const unsigned sb_up = sb + 4 * up_off, sb_dn = sb + 4 * dn_off;
  const bool x_lo = BORDER && gx == 0;        // only element 0 of a strip can be x == 0 (gx % 4 == 0)
  const int i_hi = BORDER ? w - 1 - gx : -1;  // element index of x == w-1 in this strip, if 0..3
  bool inside[4];                             // cells outside the image exist only in BORDER CTAs
#pragma unroll
  for (int i = 0; i < 4; i++) inside[i] = !BORDER || (gx + i >= 0 && gx + i < w && gy >= 0 && gy < h);

  StripAddr sa;
  {
    const int my = min(max(gy, 0), h - 1);
    sa.off = my * pitch + gx;
    sa.interior = !BORDER || (gx >= 0 && gx + 3 < w);  // an interior region has no partial strip
  }

  Strip t;
  float du[4];  // increment in u: live in phases A-C and as the sweeps' result, not across the sweeps
  const bool later = a.phi_in != nullptr;  // a later pass of an outer iteration: phi, ksi come from its first pass
  // Programmatic dependent launch: this grid may have been started while the previous pass was still
  // draining.  Everything the previous pass does not write (u, v, the derivative planes) is requested
  // first; du, dv, phi, ksi only after griddepcontrol.wait (= previous grid complete and flushed).
  if (a.pdl) asm volatile("griddepcontrol.launch_dependents;");

  for (int outer = 0; outer < a.outer; ++outer) {
    // (the sweep loop of the previous outer iteration ended with a barrier: every plane is free)
    // ------ phase A: loads; ksi (solve_2d.cu:176-196); hand-over of fx, fy, ft or the gradient tensor ------
    float phi[4];
    stage_strip<P_U>(a.u, sa, gx, w, sb);
    stage_strip<P_V>(a.v, sa, gx, w, sb);
    if (later) {
      // nothing is computed from these planes before phase C: all of them go straight to shared memory
      if (!GRAD) {
        stage_strip<P_FX>(a.fx, sa, gx, w, sb);
        stage_strip<P_FY>(a.fy, sa, gx, w, sb);
        stage_strip<P_FT>(a.ft, sa, gx, w, sb);
      }
      if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
      stage_strip<P_DU>(a.du_in, sa, gx, w, sb);
      stage_strip<P_DV>(a.dv_in, sa, gx, w, sb);
      stage_strip<P_PHI>(a.phi_in, sa, gx, w, sb);
      stage_strip<P_KSI>(a.ksi_in, sa, gx, w, sb);
    } else {
      float fx[4], fy[4], ft[4], ksi[4];
      load_strip(a.fx, sa, gx, w, fx);
      load_strip(a.fy, sa, gx, w, fy);
      load_strip(a.ft, sa, gx, w, ft);
      if (outer == 0) {
        if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
        if (a.du_in) {
          load_strip(a.du_in, sa, gx, w, du);
          load_strip(a.dv_in, sa, gx, w, t.dv);
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) du[i] = t.dv[i] = 0.f;
        }
      } else {
        // resident mode: du of the previous outer iteration comes back from this thread's own store
        // (dv is still in its registers)
        load_strip(a.du_out, sa, gx, w, du);
      }
      // own pixel only; always the brightness tensor, also in gradient mode
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float J11 = fx[i] * fx[i], J22 = fy[i] * fy[i], J12 = fx[i] * fy[i], J13 = fx[i] * ft[i], J23 = fy[i] * ft[i];
        const float d_u = du[i], d_v = t.dv[i];
        const float ta = J13 + fmaf(J11, d_u, J12 * d_v);
        const float tb = J23 + fmaf(J12, d_u, J22 * d_v);
        const float tc = fmaf(ft[i], ft[i], fmaf(J13, d_u, J23 * d_v));
        float sq = fmaf(d_u, ta, d_v * tb) + tc;
        sq = sq * ((sq > 0.f) ? 1.f : 0.f);
        const float q = sqrtf(fmaf(a.e_data, a.e_data, sq));
        ksi[i] = 1.f / (q + q);
      }
      sts4<P_KSI>(sb, ksi);
      if (!GRAD) {
        // brightness constancy: the tensor is five products of fx, fy, ft; three planes are handed to
        // phase C instead of five
        sts4<P_FX>(sb, fx);
        sts4<P_FY>(sb, fy);
        sts4<P_FT>(sb, ft);
      }
    }
    if (GRAD) {
      float J11[4], J22[4], J12[4], J13[4], J23[4];
      load_strip(a.J[0], sa, gx, w, J11);
      load_strip(a.J[1], sa, gx, w, J22);
      load_strip(a.J[2], sa, gx, w, J12);
      load_strip(a.J[3], sa, gx, w, J13);
      load_strip(a.J[4], sa, gx, w, J23);
#pragma unroll
      for (int i = 0; i < 4; i++) { J12[i] = -J12[i]; J13[i] = -J13[i]; J23[i] = -J23[i]; }
      sts4<P_J11>(sb, J11);
      sts4<P_J22>(sb, J22);
      sts4<P_NJ12>(sb, J12);
      sts4<P_NJ13>(sb, J13);
      sts4<P_NJ23>(sb, J23);
    }

    if (TIMING) stamp(a, 1);
    if (!later) {
      // ---------------- phase B: phi (solve_2d.cu:141-162) ----------------
      sts4<P_DU>(sb, du);
      sts4<P_DV>(sb, t.dv);
      cp_async_wait_all();  // u, v of this strip have landed ...
      __syncthreads();      // ... and everybody's are visible
      lds4<P_U>(sb, t.uc);
      lds4<P_V>(sb, t.vc);
      float dux[4], duy[4], dvx[4], dvy[4], num[4];
      const float hx2 = a.hx + a.hx, hy2 = a.hy + a.hy;
      const float rhx2 = fast_path_rcp(hx2), rhy2 = fast_path_rcp(hy2);
      {
        float nU[4], nD[4];
        lds4<P_U>(sb_up, nU);
        lds4<P_U>(sb_dn, nD);
#pragma unroll
        for (int i = 0; i < 4; i++) num[i] = nD[i] - nU[i];
        lds4<P_DU>(sb_up, nU);
        lds4<P_DU>(sb_dn, nD);
#pragma unroll
        for (int i = 0; i < 4; i++) num[i] = (num[i] + nD[i]) - nU[i];
        div_rn4(num, hy2, rhy2, duy);
        lds4<P_V>(sb_up, nU);
        lds4<P_V>(sb_dn, nD);
#pragma unroll
        for (int i = 0; i < 4; i++) num[i] = nD[i] - nU[i];
        lds4<P_DV>(sb_up, nU);
        lds4<P_DV>(sb_dn, nD);
#pragma unroll
        for (int i = 0; i < 4; i++) num[i] = (num[i] + nD[i]) - nU[i];
        div_rn4(num, hy2, rhy2, dvy);
      }
      {
        const float uL = __shfl_up_sync(0xffffffffu, t.uc[3], 1), uR = __shfl_down_sync(0xffffffffu, t.uc[0], 1);
        const float dL = __shfl_up_sync(0xffffffffu, du[3], 1), dR = __shfl_down_sync(0xffffffffu, du[0], 1);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float l, r, dl, dr;
          x_nb<BORDER>(t.uc, uL, uR, i, x_lo, i_hi, l, r);
          x_nb<BORDER>(du, dL, dR, i, x_lo, i_hi, dl, dr);
          num[i] = ((r - l) + dr) - dl;
        }
        div_rn4(num, hx2, rhx2, dux);
      }
      {
        const float vL = __shfl_up_sync(0xffffffffu, t.vc[3], 1), vR = __shfl_down_sync(0xffffffffu, t.vc[0], 1);
        const float dL = __shfl_up_sync(0xffffffffu, t.dv[3], 1), dR = __shfl_down_sync(0xffffffffu, t.dv[0], 1);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float l, r, dl, dr;
          x_nb<BORDER>(t.vc, vL, vR, i, x_lo, i_hi, l, r);
          x_nb<BORDER>(t.dv, dL, dR, i, x_lo, i_hi, dl, dr);
          num[i] = ((r - l) + dr) - dl;
        }
        div_rn4(num, hx2, rhx2, dvx);
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float s = duy[i] * duy[i];
        s = fmaf(dux[i], dux[i], s);
        s = fmaf(dvx[i], dvx[i], s);
        s = fmaf(dvy[i], dvy[i], s);
        s = fmaf(a.e_smooth, a.e_smooth, s);
        const float r = sqrtf(s);
        phi[i] = 1.f / (r + r);
      }
      sts4<P_PHI>(sb, phi);
    } else {
      cp_async_wait_all();
    }
    __syncthreads();  // phi published; every reader of the neighbours' P_U..P_DV is done
    if (TIMING) stamp(a, 2);

    // ---------------- phase C: weights and denominators (solve_2d.cu:333-349, 363, 367) ----------------
    {
      if (later) {  // this strip's own values, staged by phase A (their planes are recycled below)
        lds4<P_U>(sb, t.uc);
        lds4<P_V>(sb, t.vc);
        lds4<P_DU>(sb, du);
        lds4<P_DV>(sb, t.dv);
        lds4<P_PHI>(sb, phi);
      }
      const float pL = __shfl_up_sync(0xffffffffu, phi[3], 1), pR = __shfl_down_sync(0xffffffffu, phi[0], 1);
      float pU[4], pD[4], J11[4], J22[4], ksi[4];
      lds4<P_PHI>(sb_up, pU);
      lds4<P_PHI>(sb_dn, pD);
      lds4<P_KSI>(sb, ksi);
      if (BORDER) {  // cells outside the image are inert
#pragma unroll
        for (int i = 0; i < 4; i++) ksi[i] = inside[i] ? ksi[i] : 0.f;
      }
      if (GRAD) {
        lds4<P_J11>(sb, J11);
        lds4<P_J22>(sb, J22);
        lds4<P_NJ12>(sb, t.nJ12);
      } else {
        float gfx[4], gfy[4], gft[4], nJ13[4], nJ23[4];
        lds4<P_FX>(sb, gfx);
        lds4<P_FY>(sb, gfy);
        lds4<P_FT>(sb, gft);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          J11[i] = gfx[i] * gfx[i];
          J22[i] = gfy[i] * gfy[i];
          t.nJ12[i] = -(gfx[i] * gfy[i]);
          nJ13[i] = -(gfx[i] * gft[i]);
          nJ23[i] = -(gfy[i] * gft[i]);
        }
        sts4<P_NJ13>(sb, nJ13);
        sts4<P_NJ23>(sb, nJ23);
      }
      if (a.phi_out && !later) {  // a later pass of this outer iteration reloads the robust weights
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int x = gx + i;
          if (x >= ox0 && x < ox1 && gy >= oy0 && gy < oy1) {
            a.phi_out[(size_t)gy * pitch + x] = phi[i];
            a.ksi_out[(size_t)gy * pitch + x] = ksi[i];
          }
        }
      }
      const float hx_2 = a.hx_2, hy_2 = a.hy_2;  // alpha / h^2, divided once on the host (IEEE, same bits)
      // Neumann boundary through zero weights (solve_2d.cu:337-340)
      float wyp = hy_2, wym = hy_2;
      if (BORDER) {
        wyp = hy_2 * ((gy < h - 1) ? 1.f : 0.f);
        wym = hy_2 * ((gy > 0) ? 1.f : 0.f);
      }
      float eyp[4], eym[4], denU[4], denV[4], rU[4], rV[4];
      t.den_ok = true;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float p_l, p_r;
        x_nb<BORDER>(phi, pL, pR, i, x_lo, i_hi, p_l, p_r);
        const float pc = phi[i];
        float wxp = hx_2, wxm = hx_2;
        if (BORDER) {
          wxp = hx_2 * ((gx + i < w - 1) ? 1.f : 0.f);
          wxm = hx_2 * ((gx + i > 0) ? 1.f : 0.f);
        }
        float axp = wxp * ((p_r + pc) * 0.5f);
        float axm = wxm * ((p_l + pc) * 0.5f);
        eyp[i] = wyp * ((pD[i] + pc) * 0.5f);
        eym[i] = wym * ((pU[i] + pc) * 0.5f);
        const float sumH = ((axp + axm) + eyp[i]) + eym[i];
        denU[i] = fmaf(J11[i], ksi[i], sumH);
        denV[i] = fmaf(J22[i], ksi[i], sumH);
        if (!inside[i]) {  // inert cell: stays at zero increment, on the fast division path
          axp = axm = eyp[i] = eym[i] = 0.f;
          denU[i] = denV[i] = 1.f;
        }
        // axm(x) == axp(x-1) bit for bit for x >= 1 (same products, commuted add): only the strip's
        // first axm is kept separately
        if (i == 0) t.ex[0] = axm;
        t.ex[i + 1] = axp;
        rU[i] = fast_path_rcp(denU[i]);
        rV[i] = fast_path_rcp(denV[i]);
        t.den_ok = t.den_ok && rU[i] != 0.f && rV[i] != 0.f;
        t.ksi[i] = ksi[i];
        t.su[i] = t.uc[i] + du[i];
        t.sv[i] = t.vc[i] + t.dv[i];
      }
      // (the compiler would otherwise keep phi and rebuild the three inner x weights in every sweep)
#pragma unroll
      for (int i = 1; i < 4; i++) t.ex[i] = keep(t.ex[i]);
      sts4<P_EYP>(sb, eyp);
      sts4<P_EYM>(sb, eym);
      sts4<P_DENU>(sb, denU);
      sts4<P_DENV>(sb, denV);
      sts4<P_RU>(sb, rU);
      sts4<P_RV>(sb, rV);
      sts4<P_SU0>(sb, t.su);
      sts4<P_SV0>(sb, t.sv);
    }
    __syncthreads();
    if (TIMING) stamp(a, 3);

    // ---------------- phase D: Jacobi sweeps (solve_2d.cu:350-367 as compiled) ----------------
    // Rows further than sweeps-k from the output tile can no longer influence it: whole warps (two
    // rows) outside that window skip the sweep (nothing reads what they would write).
    const int need = keep(max(oy0 - gy, gy - oy1 + 1));  // this row matters as long as sweeps-k >= need
    // one sweep; ODD = the sweep number is odd: reads s_u/s_v buffer 0, writes buffer 1
    const unsigned kb = keep(sb), kb_up = keep(sb_up), kb_dn = keep(sb_dn);  // address registers of the sweeps
    auto sweep = [&](auto odd, int k) {
      constexpr bool ODD = decltype(odd)::value;
      constexpr int CU = ODD ? P_SU0 : P_SU1, CV = ODD ? P_SV0 : P_SV1, NU = ODD ? P_SU1 : P_SU0, NV = ODD ? P_SV1 : P_SV0;
      if (!__any_sync(0xffffffffu, a.sweeps - k >= need)) {
        __syncthreads();
        return;
      }
      const float suL = __shfl_up_sync(0xffffffffu, t.su[3], 1), suR = __shfl_down_sync(0xffffffffu, t.su[0], 1);
      const float svL = __shfl_up_sync(0xffffffffu, t.sv[3], 1), svR = __shfl_down_sync(0xffffffffu, t.sv[0], 1);
      float nU[4], nD[4], eyp[4], eym[4], sumU[4], sumV[4];
      lds4<P_EYP>(kb, eyp);
      lds4<P_EYM>(kb, eym);
      lds4<CU>(kb_up, nU);
      lds4<CU>(kb_dn, nD);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float l, r;
        x_nb<BORDER>(t.su, suL, suR, i, x_lo, i_hi, l, r);
        const float u0 = t.uc[i];
        // a mul, then an fma chain in the order xm, xp, yp, ym
        float s = t.ex[i] * (l - u0);
        s = fmaf(t.ex[i + 1], r - u0, s);
        s = fmaf(eyp[i], nD[i] - u0, s);
        sumU[i] = fmaf(eym[i], nU[i] - u0, s);
      }
      lds4<CV>(kb_up, nU);
      lds4<CV>(kb_dn, nD);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float l, r;
        x_nb<BORDER>(t.sv, svL, svR, i, x_lo, i_hi, l, r);
        const float v0 = t.vc[i];
        float s = t.ex[i] * (l - v0);
        s = fmaf(t.ex[i + 1], r - v0, s);
        s = fmaf(eyp[i], nD[i] - v0, s);
        sumV[i] = fmaf(eym[i], nU[i] - v0, s);
      }
      // (-J13) - J12*dv is one FFMA in the reference SASS
      float nj[4], den[4], rcp[4], num[4], rdv[4];
      const float (&ksi)[4] = t.ksi;
      const float (&nJ12)[4] = t.nJ12;
      lds4<P_NJ13>(kb, nj);
      lds4<P_DENU>(kb, den);
      lds4<P_RU>(kb, rcp);
#pragma unroll
      for (int i = 0; i < 4; i++) num[i] = fmaf(ksi[i], fmaf(nJ12[i], t.dv[i], nj[i]), sumU[i]);
      div_rn4(num, den, rcp, t.den_ok, du);
      lds4<P_NJ23>(kb, nj);
      lds4<P_DENV>(kb, den);
      lds4<P_RV>(kb, rcp);
#pragma unroll
      for (int i = 0; i < 4; i++) num[i] = fmaf(ksi[i], fmaf(nJ12[i], du[i], nj[i]), sumV[i]);
      div_rn4(num, den, rcp, t.den_ok, rdv);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        t.dv[i] = rdv[i];
        t.su[i] = t.uc[i] + du[i];
        t.sv[i] = t.vc[i] + rdv[i];
      }
      sts4<NU>(kb, t.su);
      sts4<NV>(kb, t.sv);
      __syncthreads();
    };
    for (int k = 1; k <= a.sweeps; k += 2) {
      sweep(std::true_type{}, k);
      if (k + 1 <= a.sweeps) sweep(std::false_type{}, k + 1);
    }

    if (TIMING) stamp(a, 4);
    // ---------------- phase E: store du, dv of the output tile ----------------
    if (gy >= oy0 && gy < oy1) {
      float* rdu = a.du_out + (size_t)gy * pitch;
      float* rdvp = a.dv_out + (size_t)gy * pitch;
      if (gx >= ox0 && gx + 3 < ox1) {
        st4(rdu + gx, du);
        st4(rdvp + gx, t.dv);
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int x = gx + i;
          if (x >= ox0 && x < ox1) { rdu[x] = du[i]; rdvp[x] = t.dv[i]; }
        }
      }
    }
    if (TIMING) stamp(a, 5);
  }
}

// TIMING: the variant with the %globaltimer stamps of flow2d_debug_timing (tools/phase_timing.py); the
// production kernel carries neither the stamps nor the registers of their address arithmetic.
template <bool GRAD, bool TIMING>
__global__ void __launch_bounds__(NT, 1) solve_pass_kernel(const SolveArgs a) {
  extern __shared__ __align__(16) float sm[];
  if (a.stop && *a.stop) return;  // the level has converged (flow2d_params.residual_tolerance)
  // does this CTA's region reach the image border (or beyond)?
  const int lx0 = blockIdx.x * a.ow - a.halo_x, ly0 = a.y0 + blockIdx.y * a.oh - a.halo_y;
  const bool border = lx0 <= 0 || lx0 + LW >= a.w || ly0 <= 0 || ly0 + (int)(blockDim.x >> 4) >= a.h;
  if (border) pass_body<GRAD, true, TIMING>(a, sm);
  else pass_body<GRAD, false, TIMING>(a, sm);
}

cudaError_t solve_pass_configure() {
  const int bytes = (int)solve_pass_smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(solve_pass_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(solve_pass_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(solve_pass_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(solve_pass_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  return e;
}

void launch_solve_pass(cudaStream_t st, const SolveArgs& a, bool grad, int grid_x, int grid_y, int rows) {
  // rows: region rows actually needed (resident mode of a small level); whole warps = pairs of rows
  const int nrows = rows <= 0 || rows > LH ? LH : (rows + 1) / 2 * 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid_x, grid_y);
  cfg.blockDim = dim3(nrows * (LW / 4));
  cfg.dynamicSmemBytes = solve_pass_smem_bytes();
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = a.pdl ? 1 : 0;  // only between consecutive passes of one solve (see SolveArgs::pdl)
  if (a.timing) {
    if (grad) cudaLaunchKernelEx(&cfg, solve_pass_kernel<true, true>, a);
    else cudaLaunchKernelEx(&cfg, solve_pass_kernel<false, true>, a);
  } else {
    if (grad) cudaLaunchKernelEx(&cfg, solve_pass_kernel<true, false>, a);
    else cudaLaunchKernelEx(&cfg, solve_pass_kernel<false, false>, a);
  }
}

// ---------------------------------------------------------------------------------------------
// solve_tiny: the whole solve of a level of at most 1024 pixels (the coarsest ~20 levels of every
// pyramid) in one CTA with ONE THREAD PER PIXEL.  Those levels are pure latency: with one pixel per
// thread every per-pixel constant stays in registers, a sweep is ~50 dependent instructions and
// four shared-memory reads, and the mirrored border is folded into four precomputed neighbour
// indices.  Same arithmetic, operation for operation, as solve_pass.
// ---------------------------------------------------------------------------------------------
constexpr int kTinyMax = 1024;

// (div1 / sqrt1 / rcp1, the nine shared planes and one_px_outer live in solve_onepx.cuh: solve_cluster.cu shares them)

template <bool GRAD>
__global__ void __launch_bounds__(kTinyMax, 1) solve_tiny_kernel(const SolveArgs a) {
  __shared__ __align__(16) float sq[kOnePxPlanes * kTinyMax];
  const int w = a.w, h = a.h, n = w * h;
  const int t = threadIdx.x;
  const bool on = t < n;  // (threads beyond the level only take part in the barriers; their cells are inert)
  const int x = on ? t % w : 0, y = on ? t / w : 0;
  // mirrored neighbours (index -1 -> 1, n -> n-2)
  const int il = y * w + (x == 0 ? 1 : x - 1), ir = y * w + (x == w - 1 ? w - 2 : x + 1);
  const int iu = (y == 0 ? 1 : y - 1) * w + x, id = (y == h - 1 ? h - 2 : y + 1) * w + x;
  const size_t g = (size_t)y * a.pitch + x;
  asm volatile("griddepcontrol.launch_dependents;");

  OnePx c;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sq);
  c.ac = base + 4u * t; c.al = base + 4u * (on ? il : t); c.ar = base + 4u * (on ? ir : t);
  c.au = base + 4u * (on ? iu : t); c.ad = base + 4u * (on ? id : t);
  c.uc = c.vc = c.fx = c.fy = c.ft = 0.f;
  c.J11 = c.J22 = c.nJ12 = c.nJ13 = c.nJ23 = 0.f;
  float du = 0.f, dv = 0.f;
  if (on) {
    c.uc = a.u[g]; c.vc = a.v[g]; c.fx = a.fx[g]; c.fy = a.fy[g]; c.ft = a.ft[g];
    if (a.du_in) { du = a.du_in[g]; dv = a.dv_in[g]; }
    if (GRAD) {
      c.J11 = a.J[0][g]; c.J22 = a.J[1][g]; c.nJ12 = -a.J[2][g]; c.nJ13 = -a.J[3][g]; c.nJ23 = -a.J[4][g];
    } else {
      c.J11 = c.fx * c.fx; c.J22 = c.fy * c.fy; c.nJ12 = -(c.fx * c.fy); c.nJ13 = -(c.fx * c.ft); c.nJ23 = -(c.fy * c.ft);
    }
  }
  stq<kTinyMax, Q_U>(c.ac, c.uc);
  stq<kTinyMax, Q_V>(c.ac, c.vc);
  c.hx2 = a.hx + a.hx; c.hy2 = a.hy + a.hy;
  c.rhx2 = fast_path_rcp(c.hx2); c.rhy2 = fast_path_rcp(c.hy2);
  c.wxp = a.hx_2 * ((x < w - 1) ? 1.f : 0.f); c.wxm = a.hx_2 * ((x > 0) ? 1.f : 0.f);
  c.wyp = a.hy_2 * ((y < h - 1) ? 1.f : 0.f); c.wym = a.hy_2 * ((y > 0) ? 1.f : 0.f);
  c.e_smooth = a.e_smooth; c.e_data = a.e_data;
  c.live = on;

  float phi = 0.f, ksi = 0.f;
  for (int outer = 0; outer < a.outer; ++outer) one_px_outer<kTinyMax>(c, a.sweeps, a.exact != 0, du, dv, phi, ksi);
  if (on) {
    a.du_out[g] = du; a.dv_out[g] = dv;
    if (a.phi_out) { a.phi_out[g] = phi; a.ksi_out[g] = ksi; }
  }
}

bool solve_tiny_fits(int w, int h) { return w >= 2 && h >= 2 && w * h <= kTinyMax; }

void launch_solve_tiny(cudaStream_t st, const SolveArgs& a, bool grad) {
  const int threads = (a.w * a.h + 31) / 32 * 32;
  if (grad) solve_tiny_kernel<true><<<1, threads, 0, st>>>(a);
  else solve_tiny_kernel<false><<<1, threads, 0, st>>>(a);
}

// ---------------------------------------------------------------------------------------------
// solve_small_pass: one pass (phi, ksi, weights + S sweeps) of a MID-SIZE level with one thread per
// pixel: 1024 threads = one 32x32 region, output tile (32 - 2(S+1))^2.  The halo redundancy is large
// (S = 5: 400 of 1024 cells are output), but levels between ~3 000 and ~200 000 pixels cannot fill
// the GPU with 64x48 regions anyway (a 584x388 level is 143 CTAs): what counts there is the latency
// of one CTA, and one pixel per thread cuts the dependent chain of a pass by four.  Mask-free like
// solve_pass: every cell is updated in every sweep, exactness shrinks by one ring per sweep.
// ---------------------------------------------------------------------------------------------
// The region edge TS is 32, 24 or 16: a smaller region is a shorter pass as long as the level's regions
// still fit the SMs one to one; the scheduler picks the size per level (flow2d_api.cu: run_solve).
template <bool GRAD, int TS>
__global__ void __launch_bounds__(TS * TS, 1) solve_small_pass_kernel(const SolveArgs a) {
  constexpr int N = TS * TS;
  __shared__ __align__(16) float sq[kOnePxPlanes * N];
  if (a.stop && *a.stop) return;  // the level has converged (flow2d_params.residual_tolerance)
  const int w = a.w, h = a.h;
  const int t = threadIdx.x, ly = t / TS, lx = t - ly * TS;
  const int ox0 = blockIdx.x * a.ow, oy0 = a.y0 + blockIdx.y * a.oh;
  const int ox1 = min(w, ox0 + a.ow), oy1 = min(a.y1, oy0 + a.oh);
  const int gx = ox0 - a.halo_x + lx, gy = oy0 - a.halo_y + ly;
  // neighbours inside the region; at the image border the mirrored neighbour is the opposite one; the
  // region's own edge cells have no outer neighbour (they are never exact): any in-range cell will do
  const int il = (gx == 0 || lx == 0) ? t + 1 : t - 1, ir = (gx == w - 1 || lx == TS - 1) ? t - 1 : t + 1;
  const int iu = (gy == 0 || ly == 0) ? t + TS : t - TS, id = (gy == h - 1 || ly == TS - 1) ? t - TS : t + TS;
  const size_t g = (size_t)min(max(gy, 0), h - 1) * a.pitch + min(max(gx, 0), w - 1);

  OnePx c;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sq);
  c.ac = base + 4u * t; c.al = base + 4u * il; c.ar = base + 4u * ir; c.au = base + 4u * iu; c.ad = base + 4u * id;
  asm volatile("griddepcontrol.launch_dependents;");  // the next kernel of the stream may be scheduled (it waits for us)
  c.uc = a.u[g]; c.vc = a.v[g]; c.fx = a.fx[g]; c.fy = a.fy[g]; c.ft = a.ft[g];
  if (GRAD) {
    c.J11 = a.J[0][g]; c.J22 = a.J[1][g]; c.nJ12 = -a.J[2][g]; c.nJ13 = -a.J[3][g]; c.nJ23 = -a.J[4][g];
  } else {
    c.J11 = c.fx * c.fx; c.J22 = c.fy * c.fy; c.nJ12 = -(c.fx * c.fy); c.nJ13 = -(c.fx * c.ft); c.nJ23 = -(c.fy * c.ft);
  }
  if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  float du = 0.f, dv = 0.f;
  if (a.du_in) { du = a.du_in[g]; dv = a.dv_in[g]; }
  stq<N, Q_U>(c.ac, c.uc);
  stq<N, Q_V>(c.ac, c.vc);
  c.hx2 = a.hx + a.hx; c.hy2 = a.hy + a.hy;
  c.rhx2 = fast_path_rcp(c.hx2); c.rhy2 = fast_path_rcp(c.hy2);
  c.wxp = a.hx_2 * ((gx < w - 1) ? 1.f : 0.f); c.wxm = a.hx_2 * ((gx > 0) ? 1.f : 0.f);
  c.wyp = a.hy_2 * ((gy < h - 1) ? 1.f : 0.f); c.wym = a.hy_2 * ((gy > 0) ? 1.f : 0.f);
  c.e_smooth = a.e_smooth; c.e_data = a.e_data;
  c.live = gx >= 0 && gx < w && gy >= 0 && gy < h;

  float phi, ksi;
  one_px_outer<N>(c, a.sweeps, a.exact != 0, du, dv, phi, ksi);
  if (gx >= ox0 && gx < ox1 && gy >= oy0 && gy < oy1) {
    const size_t o = (size_t)gy * a.pitch + gx;
    a.du_out[o] = du;
    a.dv_out[o] = dv;
    if (a.phi_out) { a.phi_out[o] = phi; a.ksi_out[o] = ksi; }
  }
}

template <int TS>
static void launch_small_ts(const cudaLaunchConfig_t& cfg, const SolveArgs& a, bool grad) {
  if (grad) cudaLaunchKernelEx(&cfg, solve_small_pass_kernel<true, TS>, a);
  else cudaLaunchKernelEx(&cfg, solve_small_pass_kernel<false, TS>, a);
}

void launch_solve_small_pass(cudaStream_t st, const SolveArgs& a, bool grad, int grid_x, int grid_y, int ts) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid_x, grid_y);
  cfg.blockDim = dim3(ts * ts);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = a.pdl ? 1 : 0;
  if (ts == 16) launch_small_ts<16>(cfg, a, grad);
  else if (ts == 24) launch_small_ts<24>(cfg, a, grad);
  else launch_small_ts<32>(cfg, a, grad);
}

void preload_solve_kernels() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, solve_pass_kernel<false, false>);
  cudaFuncGetAttributes(&a, solve_pass_kernel<true, false>);
  cudaFuncGetAttributes(&a, solve_tiny_kernel<false>);
  cudaFuncGetAttributes(&a, solve_tiny_kernel<true>);
  cudaFuncGetAttributes(&a, solve_small_pass_kernel<false, 16>);
  cudaFuncGetAttributes(&a, solve_small_pass_kernel<false, 24>);
  cudaFuncGetAttributes(&a, solve_small_pass_kernel<false, 32>);
  cudaFuncGetAttributes(&a, solve_small_pass_kernel<true, 16>);
  cudaFuncGetAttributes(&a, solve_small_pass_kernel<true, 24>);
  cudaFuncGetAttributes(&a, solve_small_pass_kernel<true, 32>);
}

}  // namespace flow2d

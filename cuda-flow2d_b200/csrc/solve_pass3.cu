// solve_pass3.cu -- the tiled solve pass of the large pyramid levels, third generation: PERSISTENT CTAs whose input
// tiles arrive by TMA (cp.async.bulk.tensor.2d + mbarrier) while the previous tile is still being swept.
// sm_100a, compiled with -fmad=false.  Brightness constancy; gradient constancy stays on solve_pass2.cu.
//
// Replaces compute_phi_ksi (src/kernels/solve_2d.cu:43-198) and solve_2d (src/kernels/solve_2d.cu:200-377) as driven by
// CudaOperationSolve2D::Execute (src/cuda_operations/2d/cuda_operation_solve_2d.cpp:229-299), with identical results.
// The arithmetic, the 64x48 region with its shrinking exactness, the register-resident strips (two 1x4 strips per
// thread, packed fp32) and the mirrored / inert border handling are those of solve_pass2.cu; what changed is how the
// inputs get there:
//
//   * one CTA per SM walks over the tiles of the level (tile = blockIdx.x, += gridDim.x).  The 5 - 9 input planes of a
//     tile (u, v, fx, fy, ft [, du, dv [, phi, ksi]]) are fetched as 64x48 boxes of tensor maps by ONE thread -- one
//     cp.async.bulk.tensor per plane, completion counted on an mbarrier -- into staging planes of their own.  Cells of a
//     box outside the level are filled with zeros by the hardware (they are inert: zero weights, ksi = 0, denominator 1,
//     and no cell inside the image ever reads them), so border tiles need no clamped scalar loads.  A box may start at
//     any column, so the x halo is S+1 rounded up to 2 instead of to the 4 columns of an aligned float4: with five
//     sweeps per pass the output tile is 52 x 36 instead of 48 x 36 of the 64 x 48 region;
//   * as soon as phase C has consumed the staging planes the loads of the NEXT tile are issued; they land while the
//     sweeps of the current tile run.  solve_pass2 had one CTA per SM as well (168 registers x 384 threads), so nothing
//     hid its load phase: ncu showed 18 % of its samples there, half of them waiting for memory, plus spills of
//     freshly loaded float4s (profiles/r02/solve_2048_ncu_summary.txt);
//   * phase B reads the neighbour rows of u, v, du, dv straight from the staging planes (the whole region is there), so
//     the publish step of solve_pass2 (8 STS.128 per thread and one CTA barrier per pass) is gone; a later pass of an
//     outer iteration reads the neighbours' phi from its staging plane as well.
//
// Shared memory: 6 work planes (72 KiB, as solve_pass2) + 9 staging planes of 68 x 48 (115 KiB) + the mbarrier.
#include <cuda.h>

#include <type_traits>

#include "kernels.h"
#include "solve_common.cuh"
#include "solve_packed.cuh"

namespace flow2d {

constexpr int NT3 = (LH / 2) * (LW / 4);  // 384 threads

enum {
  W_SU0 = 0, W_SV0, W_SU1, W_SV1,  // sweeps: double-buffered s_u = u+du, s_v = v+dv (read by the rows above / below)
  W_RU, W_RV,                      // thread-private: fast-path reciprocals of the two denominators
  kWorkPlanes,
  W_PHI = W_SU1                    // phase C (neighbour rows of phi, first pass of an outer iteration); sweep 1 is the first writer of SU1
};
enum { T_U = 0, T_V, T_FX, T_FY, T_FT, T_DU, T_DV, T_PHI, T_KSI, kStagePlanes };  // staging planes, written by TMA only
constexpr unsigned kPlaneBytes = PL * 4;
// A TMA box must start on a 16-byte boundary of its row, while the region of a tile starts at an EVEN column (x halo =
// S+1 rounded up to 2).  When that column is not a multiple of 4 (PAD2: halo 2 or 6, i.e. 1, 4 or 5 sweeps per pass) a
// staging plane is a 68 x 48 box that starts two columns left of the region: its rows are 272 bytes apart and the strips
// in it only 8-byte aligned, so it is read with 64-bit shared loads (twice the instructions and, with lanes 16 bytes
// apart, twice the wavefronts of an aligned LDS.128 -- measured 7 % of a pass, against 17 % fewer tiles).  Otherwise
// the box is the region itself and is read like a work plane.
template <bool PAD2> struct Stage {
  static constexpr int TW = PAD2 ? LW + 4 : LW, TPL = TW * LH;
  static constexpr unsigned kBytes = TPL * 4;  // 13056 = 102 * 128 | 12288
};
constexpr unsigned kMaxStageBytes = Stage<true>::kBytes;

size_t solve_pass3_smem_bytes() { return (size_t)kPlaneBytes * kWorkPlanes + (size_t)kMaxStageBytes * kStagePlanes + 16; }

// four pixels of a staging plane: plane and row offset are immediates of ONE per-thread address (strip A, staging plane 0)
template <bool PAD2, int PLANE, int DROW>
__device__ __forceinline__ Q ldt(unsigned addr) {
  Q q;
  constexpr int OFF = (PLANE * Stage<PAD2>::TPL + DROW * Stage<PAD2>::TW) * 4;
  if (PAD2) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2 + %3];" : "=f"(q.lo.x), "=f"(q.lo.y) : "r"(addr), "n"(OFF) : "memory");
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2 + %3];" : "=f"(q.hi.x), "=f"(q.hi.y) : "r"(addr), "n"(OFF + 8) : "memory");
  } else {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4 + %5];"
                 : "=f"(q.lo.x), "=f"(q.lo.y), "=f"(q.hi.x), "=f"(q.hi.y)
                 : "r"(addr), "n"(OFF)
                 : "memory");
  }
  return q;
}

struct Pass3Maps {
  CUtensorMap m[9];  // u v fx fy ft du dv phi ksi (the staging planes in order); unused entries are copies of m[0]
};

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// one 68x48 box of a plane: global (tensor map, element coordinates x, y; may lie partly or wholly outside) -> shared
__device__ __forceinline__ void tma_load_box(unsigned dst, const CUtensorMap* map, int x, int y, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(bar)
               : "memory");
}

struct Tile3 {
  int ox0, oy0, ox1, oy1;  // output tile
  int gx0, gy0;            // region origin (may be negative; gx0 is even)
  int bx0;                 // first column of the staged box: gx0 rounded down to a multiple of 4
  bool border;
};
__device__ __forceinline__ Tile3 tile_of(const SolveArgs& a, int tile, int tiles_x) {
  Tile3 t;
  const int by = tile / tiles_x, bx = tile - by * tiles_x;
  t.ox0 = bx * a.ow; t.oy0 = a.y0 + by * a.oh;
  t.ox1 = min(a.w, t.ox0 + a.ow); t.oy1 = min(a.y1, t.oy0 + a.oh);
  t.gx0 = t.ox0 - a.halo_x; t.gy0 = t.oy0 - a.halo_y;
  t.bx0 = t.gx0 - (a.halo_x & 2);  // ox0 is a multiple of 4 (ow = 64 - 2 * halo_x, halo_x even)
  t.border = t.gx0 <= 0 || t.gx0 + LW >= a.w || t.gy0 <= 0 || t.gy0 + LH >= a.h;
  return t;
}

// phases A - E of one tile; the staging planes hold its inputs.  Returns after the stores of the tile.
// `issue_next` is called by thread 0 once the staging planes are free again (after the barrier that ends phase C).
template <bool BORDER, bool PAD2, typename IssueNext>
__device__ __forceinline__ void pass3_tile(const SolveArgs& a, const Tile3& tl, unsigned smem0, unsigned stage0, IssueNext issue_next) {
  const int tid = threadIdx.x;
  const int trow = tid >> 4;      // thread row: region rows 2*trow (strip A) and 2*trow + 1 (strip B)
  const int lx = 4 * (tid & 15);  // first column of the strips within the region
  const int w = a.w, h = a.h, pitch = a.pitch;
  const int ox0 = tl.ox0, oy0 = tl.oy0, ox1 = tl.ox1, oy1 = tl.oy1;
  const int gx = tl.gx0 + lx;  // even (a TMA box may start at any column: the halo is S+1 rounded up to 2, not to 4)
  const int gyA = tl.gy0 + 2 * trow, gyB = gyA + 1;
  const unsigned sb = keep(smem0 + 4u * (unsigned)(2 * trow * LW + lx));  // strip A, work plane 0
  constexpr int TW = Stage<PAD2>::TW;
  const unsigned tb = keep(stage0 + 4u * (unsigned)(2 * trow * TW + (PAD2 ? 2 : 0) + lx));  // strip A, staging plane 0
  constexpr int kLastT = LH / 2 - 1;
  // neighbour rows as in solve_pass2: the row above A and the row below B (the other two are the thread's own strips);
  // BORDER tiles take all four through run-time row offsets with the image-border mirror folded in
  int upA = trow > 0 ? -1 : 1, dnA = 1, upB = 0, dnB = trow < kLastT ? 2 : 0;
  if (BORDER) {
    if (gyA == 0) upA = 1;
    if (gyA == h - 1 && trow > 0) dnA = -1;
    if (gyB == 0) upB = trow < kLastT ? 2 : 0;
    if (gyB == h - 1) dnB = 0;
  }
  const unsigned a_upA = keep(sb + 4u * (unsigned)(upA * LW)), a_dnB = keep(sb + 4u * (unsigned)(dnB * LW));
  const unsigned a_dnA = BORDER ? keep(sb + 4u * (unsigned)(dnA * LW)) : sb, a_upB = BORDER ? keep(sb + 4u * (unsigned)(upB * LW)) : sb;
  const int x_lo = BORDER ? -gx : -1;         // element index of x == 0 in this strip, if 0..3
  const int i_hi = BORDER ? w - 1 - gx : -1;  // element index of x == w-1 in this strip, if 0..3
  bool insA[4], insB[4];                      // cells outside the image exist only in BORDER tiles
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const bool xin = gx + i >= 0 && gx + i < w;
    insA[i] = !BORDER || (xin && gyA >= 0 && gyA < h);
    insB[i] = !BORDER || (xin && gyB >= 0 && gyB < h);
  }
  const bool later = a.phi_in != nullptr;  // a later pass of an outer iteration: phi, ksi come from its first pass

  Strip2 A, B;
  Q duA, duB;  // increment in u: live in phases A-C and as the sweeps' result
  Q phiA, phiB;
  Q J11A, J22A, J11B, J22B;
  // ------ phase A: own cells from the staging planes; ksi (solve_2d.cu:176-196); motion tensor ------
  {
    A.uc = ldt<PAD2, T_U, 0>(tb); B.uc = ldt<PAD2, T_U, 1>(tb);
    A.vc = ldt<PAD2, T_V, 0>(tb); B.vc = ldt<PAD2, T_V, 1>(tb);
    const Q fxA = ldt<PAD2, T_FX, 0>(tb), fxB = ldt<PAD2, T_FX, 1>(tb);
    const Q fyA = ldt<PAD2, T_FY, 0>(tb), fyB = ldt<PAD2, T_FY, 1>(tb);
    const Q ftA = ldt<PAD2, T_FT, 0>(tb), ftB = ldt<PAD2, T_FT, 1>(tb);
    if (a.du_in) {
      duA = ldt<PAD2, T_DU, 0>(tb); duB = ldt<PAD2, T_DU, 1>(tb);
      A.dv = ldt<PAD2, T_DV, 0>(tb); B.dv = ldt<PAD2, T_DV, 1>(tb);
    } else {
      duA = duB = A.dv = B.dv = qsplat(0.f);
    }
    if (later) {
      phiA = ldt<PAD2, T_PHI, 0>(tb); phiB = ldt<PAD2, T_PHI, 1>(tb);
      A.ksi = ldt<PAD2, T_KSI, 0>(tb); B.ksi = ldt<PAD2, T_KSI, 1>(tb);
    }
    auto tensor_ksi = [&](const Q& fx, const Q& fy, const Q& ft, const Q& d_u, const Q& d_v, Strip2& t, Q& J11o, Q& J22o) {
      const Q J11 = qmul(fx, fx), J22 = qmul(fy, fy), J12 = qmul(fx, fy);
      if (!later) {
        // J13, J23 feed an ADD below: scalar products (a packed product feeding a packed add would be contracted)
        float fxv[4], fyv[4], ftv[4], J13[4], J23[4];
        qarr(fx, fxv); qarr(fy, fyv); qarr(ft, ftv);
#pragma unroll
        for (int i = 0; i < 4; i++) { J13[i] = fxv[i] * ftv[i]; J23[i] = fyv[i] * ftv[i]; }
        const Q J13q = qfrom(J13), J23q = qfrom(J23);
        const Q fa = qfma(J11, d_u, qmul(J12, d_v));
        const Q fb = qfma(J12, d_u, qmul(J22, d_v));
        const Q tc = qfma(ft, ft, qfma(J13q, d_u, qmul(J23q, d_v)));
        float fav[4], fbv[4], ta[4], tb[4];
        qarr(fa, fav); qarr(fb, fbv);
#pragma unroll
        for (int i = 0; i < 4; i++) { ta[i] = J13[i] + fav[i]; tb[i] = J23[i] + fbv[i]; }
        const Q sq0 = qadd(qfma(d_u, qfrom(ta), qmul(d_v, qfrom(tb))), tc);
        float sq[4];
        qarr(sq0, sq);
#pragma unroll
        for (int i = 0; i < 4; i++) sq[i] = sq[i] * ((sq[i] > 0.f) ? 1.f : 0.f);
        const Q arg = qfma(qsplat(a.e_data), qsplat(a.e_data), qfrom(sq));
        bool ok = true;
        t.ksi = half_rsqrt4<false>(arg, ok);
        if (__builtin_expect(!ok, 0)) t.ksi = half_rsqrt4<true>(arg, ok);
        t.nJ13 = qneg(J13q); t.nJ23 = qneg(J23q);
      } else {
        t.nJ13 = qneg(qmul(fx, ft));
        t.nJ23 = qneg(qmul(fy, ft));
      }
      J11o = J11; J22o = J22; t.nJ12 = qneg(J12);
    };
    tensor_ksi(fxA, fyA, ftA, duA, A.dv, A, J11A, J22A);
    tensor_ksi(fxB, fyB, ftB, duB, B.dv, B, J11B, J22B);
  }

  Q pU_A, pD_A, pU_B, pD_B;  // phi of the rows above / below the two strips
  // the same neighbour rows in the staging planes
  const unsigned t_upA = tb + 4u * (unsigned)(upA * TW), t_dnB = tb + 4u * (unsigned)(dnB * TW);
  const unsigned t_dnA = tb + 4u * (unsigned)(dnA * TW), t_upB = tb + 4u * (unsigned)(upB * TW);
  if (!later) {
    // ---------------- phase B: phi (solve_2d.cu:141-162); neighbour rows straight from the staging planes ----------------
    const float hx2 = a.hx + a.hx, hy2 = a.hy + a.hy;
    const float rhx2 = fast_path_rcp(hx2), rhy2 = fast_path_rcp(hy2);
    const bool have_d = a.du_in != nullptr;
    // numerators of the y differences: ((below - above) + d_below) - d_above
    auto ynum = [&](auto plane_f, auto plane_d, const Q& fA, const Q& fB, const Q& dA_, const Q& dB_, Q& nA, Q& nB) {
      constexpr int PF = decltype(plane_f)::value, PD = decltype(plane_d)::value;
      Q fU_A, fD_A, fU_B, fD_B, dU_A, dD_A, dU_B, dD_B;
      fU_A = ldt<PAD2, PF, 0>(t_upA); fD_B = ldt<PAD2, PF, 0>(t_dnB);
      if (have_d) { dU_A = ldt<PAD2, PD, 0>(t_upA); dD_B = ldt<PAD2, PD, 0>(t_dnB); }
      else dU_A = dD_B = qsplat(0.f);
      if (!BORDER) {
        fD_A = fB; dD_A = dB_; fU_B = fA; dU_B = dA_;  // the other strip of this thread
      } else {
        fD_A = ldt<PAD2, PF, 0>(t_dnA); fU_B = ldt<PAD2, PF, 0>(t_upB);
        if (have_d) { dD_A = ldt<PAD2, PD, 0>(t_dnA); dU_B = ldt<PAD2, PD, 0>(t_upB); }
        else dD_A = dU_B = qsplat(0.f);
      }
      nA = qsub(qadd(qsub(fD_A, fU_A), dD_A), dU_A);
      nB = qsub(qadd(qsub(fD_B, fU_B), dD_B), dU_B);
    };
    auto xnum = [&](const Q& f, const Q& d_) {
      Q l, r, dl, dr;
      x_shift<BORDER>(f, __shfl_up_sync(0xffffffffu, f.hi.y, 1), __shfl_down_sync(0xffffffffu, f.lo.x, 1), x_lo, i_hi, l, r);
      x_shift<BORDER>(d_, __shfl_up_sync(0xffffffffu, d_.hi.y, 1), __shfl_down_sync(0xffffffffu, d_.lo.x, 1), x_lo, i_hi, dl, dr);
      return qsub(qadd(qsub(r, l), dr), dl);
    };
    Q n[8];  // numerators: duy A,B; dvy A,B; dux A,B; dvx A,B
    ynum(std::integral_constant<int, T_U>{}, std::integral_constant<int, T_DU>{}, A.uc, B.uc, duA, duB, n[0], n[1]);
    ynum(std::integral_constant<int, T_V>{}, std::integral_constant<int, T_DV>{}, A.vc, B.vc, A.dv, B.dv, n[2], n[3]);
    n[4] = xnum(A.uc, duA); n[5] = xnum(B.uc, duB);
    n[6] = xnum(A.vc, A.dv); n[7] = xnum(B.vc, B.dv);
    const Q dy = qsplat(hy2), ry = qsplat(rhy2), dx = qsplat(hx2), rx = qsplat(rhx2);
    Q q[8];
    Range rg;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      q[i] = fastdiv(n[i], i < 4 ? dy : dx, i < 4 ? ry : rx);
      rg.add(n[i]);
    }
    if (__builtin_expect(!(rg.ok() && rhx2 != 0.f && rhy2 != 0.f), 0)) {
      const Q dd[8] = {dy, dy, dy, dy, dx, dx, dx, dx}, rr[8] = {ry, ry, ry, ry, rx, rx, rx, rx};
      fix_div_n<8>(n, dd, rr, q);
    }
    auto phi_of = [&](auto exact, const Q& dux, const Q& duy, const Q& dvx, const Q& dvy, bool& ok) {
      Q s_ = qmul(duy, duy);
      s_ = qfma(dux, dux, s_);
      s_ = qfma(dvx, dvx, s_);
      s_ = qfma(dvy, dvy, s_);
      s_ = qfma(qsplat(a.e_smooth), qsplat(a.e_smooth), s_);
      return half_rsqrt4<decltype(exact)::value>(s_, ok);
    };
    bool ok = true;
    phiA = phi_of(std::false_type{}, q[4], q[0], q[6], q[2], ok);
    phiB = phi_of(std::false_type{}, q[5], q[1], q[7], q[3], ok);
    if (__builtin_expect(!ok, 0)) {
      phiA = phi_of(std::true_type{}, q[4], q[0], q[6], q[2], ok);
      phiB = phi_of(std::true_type{}, q[5], q[1], q[7], q[3], ok);
    }
    stsq<W_PHI, 0>(sb, phiA);
    stsq<W_PHI, 1>(sb, phiB);
    __syncthreads();  // phi published
    pU_A = ldsq_at<W_PHI>(a_upA); pD_B = ldsq_at<W_PHI>(a_dnB);
    if (!BORDER) { pD_A = phiB; pU_B = phiA; }
    else { pD_A = ldsq_at<W_PHI>(a_dnA); pU_B = ldsq_at<W_PHI>(a_upB); }
  } else {
    pU_A = ldt<PAD2, T_PHI, 0>(t_upA); pD_B = ldt<PAD2, T_PHI, 0>(t_dnB);
    if (!BORDER) { pD_A = phiB; pU_B = phiA; }
    else { pD_A = ldt<PAD2, T_PHI, 0>(t_dnA); pU_B = ldt<PAD2, T_PHI, 0>(t_upB); }
  }

  // ---------------- phase C: weights and denominators (solve_2d.cu:333-349, 363, 367) ----------------
  bool den_ok = true;
  {
    const float hx_2 = a.hx_2, hy_2 = a.hy_2;  // alpha / h^2, divided once on the host (IEEE, same bits)
    auto weights = [&](Strip2& t, const Q& phi, const Q& pU, const Q& pD, const Q& J11, const Q& J22, const Q& du_, int gy,
                       const bool (&inside)[4], Q& rU, Q& rV) {
      if (BORDER) {  // cells outside the image are inert
        float k[4];
        qarr(t.ksi, k);
#pragma unroll
        for (int i = 0; i < 4; i++) k[i] = inside[i] ? k[i] : 0.f;
        t.ksi = qfrom(k);
      }
      if (a.phi_out && !later) {  // a later pass of this outer iteration reloads the robust weights
        float pv[4], kv[4];
        qarr(phi, pv); qarr(t.ksi, kv);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int x = gx + i;
          if (x >= ox0 && x < ox1 && gy >= oy0 && gy < oy1) {
            a.phi_out[(size_t)gy * pitch + x] = pv[i];
            a.ksi_out[(size_t)gy * pitch + x] = kv[i];
          }
        }
      }
      const float pL = __shfl_up_sync(0xffffffffu, phi.hi.y, 1), pR = __shfl_down_sync(0xffffffffu, phi.lo.x, 1);
      Q p_l, p_r;
      x_shift<BORDER>(phi, pL, pR, x_lo, i_hi, p_l, p_r);
      // Neumann boundary through zero weights (solve_2d.cu:337-340)
      Q wxp = qsplat(hx_2), wxm = qsplat(hx_2);
      float wyp = hy_2, wym = hy_2;
      if (BORDER) {
        float xp[4], xm[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          xp[i] = hx_2 * ((gx + i < w - 1) ? 1.f : 0.f);
          xm[i] = hx_2 * ((gx + i > 0) ? 1.f : 0.f);
        }
        wxp = qfrom(xp); wxm = qfrom(xm);
        wyp = hy_2 * ((gy < h - 1) ? 1.f : 0.f);
        wym = hy_2 * ((gy > 0) ? 1.f : 0.f);
      }
      const Q half = qsplat(0.5f);
      const Q axp = qmul(wxp, qmul(qadd(p_r, phi), half));
      const Q axm = qmul(wxm, qmul(qadd(p_l, phi), half));
      const Q eyp = qmul(qsplat(wyp), qmul(qadd(pD, phi), half));
      const Q eym = qmul(qsplat(wym), qmul(qadd(pU, phi), half));
      // sumH adds four PRODUCTS: scalar adds (see solve_pass2.cu)
      float xpv[4], xmv[4], ypv[4], ymv[4], sH[4];
      qarr(axp, xpv); qarr(axm, xmv); qarr(eyp, ypv); qarr(eym, ymv);
#pragma unroll
      for (int i = 0; i < 4; i++) sH[i] = ((xpv[i] + xmv[i]) + ypv[i]) + ymv[i];
      const Q sumH = qfrom(sH);
      Q denU = qfma(J11, t.ksi, sumH), denV = qfma(J22, t.ksi, sumH);
      if (BORDER) {  // inert cell: stays at zero increment, on the fast division path
        float dU[4], dV[4];
        qarr(denU, dU); qarr(denV, dV);
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (!inside[i]) { xpv[i] = xmv[i] = ypv[i] = ymv[i] = 0.f; dU[i] = dV[i] = 1.f; }
        denU = qfrom(dU); denV = qfrom(dV);
      }
      t.exm = qmake(xmv[0], keep(xpv[0]), keep(xpv[1]), keep(xpv[2]));
      t.exp = qfrom(xpv);
      t.eyp = qfrom(ypv);
      t.eym = qfrom(ymv);
      t.denU = denU; t.denV = denV;
      rU = fast_path_rcp4(denU, den_ok);
      rV = fast_path_rcp4(denV, den_ok);
      t.su = qadd(t.uc, du_);
      t.sv = qadd(t.vc, t.dv);
    };
    Q rUA, rVA, rUB, rVB;
    weights(A, phiA, pU_A, pD_A, J11A, J22A, duA, gyA, insA, rUA, rVA);
    weights(B, phiB, pU_B, pD_B, J11B, J22B, duB, gyB, insB, rUB, rVB);
    stsq<W_RU, 0>(sb, rUA); stsq<W_RU, 1>(sb, rUB);
    stsq<W_RV, 0>(sb, rVA); stsq<W_RV, 1>(sb, rVB);
    stsq<W_SU0, 0>(sb, A.su); stsq<W_SU0, 1>(sb, B.su);
    stsq<W_SV0, 0>(sb, A.sv); stsq<W_SV0, 1>(sb, B.sv);
  }
  __syncthreads();  // s_u, s_v published; nobody reads the staging planes any more
  if (tid == 0) issue_next();

  // ---------------- phase D: Jacobi sweeps (solve_2d.cu:350-367 as compiled); identical to solve_pass2 ----------------
  const int need = min(max(oy0 - gyA, gyA - oy1 + 1), max(oy0 - gyB, gyB - oy1 + 1));
  constexpr unsigned kBufBytes = 2u * kPlaneBytes;  // SU0,SV0 -> SU1,SV1
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform
  unsigned cur = 0, nxt = kBufBytes;
#pragma unroll 1
  for (int k = 1; k <= a.sweeps; ++k) {
    if (__any_sync(0xffffffffu, a.sweeps - k >= need)) {
      auto sums = [&](auto plane, const Q& cA, const Q& cB, const Q& u0A, const Q& u0B, Q& outA, Q& outB) {
        constexpr int PC = decltype(plane)::value;
        Q nU_A, nD_A, nU_B, nD_B;
        nU_A = ldsq_at<PC>(a_upA + cur); nD_B = ldsq_at<PC>(a_dnB + cur);
        if (!BORDER) {
          nD_A = cB; nU_B = cA;  // the other strip of this thread: registers
        } else {
          nD_A = ldsq_at<PC>(a_dnA + cur); nU_B = ldsq_at<PC>(a_upB + cur);
        }
        Q l, r;
        x_shift<BORDER>(cA, __shfl_up_sync(0xffffffffu, cA.hi.y, 1), __shfl_down_sync(0xffffffffu, cA.lo.x, 1), x_lo, i_hi, l, r);
        // a mul, then an fma chain in the order xm, xp, yp, ym
        Q s_ = qmul(A.exm, qsub(l, u0A));
        s_ = qfma(A.exp, qsub(r, u0A), s_);
        s_ = qfma(A.eyp, qsub(nD_A, u0A), s_);
        outA = qfma(A.eym, qsub(nU_A, u0A), s_);
        x_shift<BORDER>(cB, __shfl_up_sync(0xffffffffu, cB.hi.y, 1), __shfl_down_sync(0xffffffffu, cB.lo.x, 1), x_lo, i_hi, l, r);
        s_ = qmul(B.exm, qsub(l, u0B));
        s_ = qfma(B.exp, qsub(r, u0B), s_);
        s_ = qfma(B.eyp, qsub(nD_B, u0B), s_);
        outB = qfma(B.eym, qsub(nU_B, u0B), s_);
      };
      Q sumUA, sumUB, sumVA, sumVB;
      sums(std::integral_constant<int, W_SU0>{}, A.su, B.su, A.uc, B.uc, sumUA, sumUB);
      sums(std::integral_constant<int, W_SV0>{}, A.sv, B.sv, A.vc, B.vc, sumVA, sumVB);
      const Q rUA = ldsq<W_RU, 0>(sb), rUB = ldsq<W_RU, 1>(sb), rVA = ldsq<W_RV, 0>(sb), rVB = ldsq<W_RV, 1>(sb);
      // (-J13) - J12*dv is one FFMA in the reference SASS
      const Q numUA = qfma(A.ksi, qfma(A.nJ12, A.dv, A.nJ13), sumUA), numUB = qfma(B.ksi, qfma(B.nJ12, B.dv, B.nJ13), sumUB);
      duA = fastdiv(numUA, A.denU, rUA);
      duB = fastdiv(numUB, B.denU, rUB);
      Q numVA = qfma(A.ksi, qfma(A.nJ12, duA, A.nJ23), sumVA), numVB = qfma(B.ksi, qfma(B.nJ12, duB, B.nJ23), sumVB);
      Q dvA = fastdiv(numVA, A.denV, rVA), dvB = fastdiv(numVB, B.denV, rVB);
      Range rg;
      rg.add(numUA); rg.add(numUB); rg.add(numVA); rg.add(numVB);
      if (__builtin_expect(!(den_ok && rg.ok()), 0)) {
        // a zero, tiny, huge or non-finite dividend (or an unsafe denominator) somewhere in these 16 quotients
        {
          const Q nn[2] = {numUA, numUB}, dd[2] = {A.denU, B.denU}, rr[2] = {rUA, rUB};
          Q qq[2] = {duA, duB};
          fix_div_n<2>(nn, dd, rr, qq);
          duA = qq[0]; duB = qq[1];
        }
        numVA = qfma(A.ksi, qfma(A.nJ12, duA, A.nJ23), sumVA);
        numVB = qfma(B.ksi, qfma(B.nJ12, duB, B.nJ23), sumVB);
        {
          const Q nn[2] = {numVA, numVB}, dd[2] = {A.denV, B.denV}, rr[2] = {rVA, rVB};
          Q qq[2] = {fastdiv(numVA, A.denV, rVA), fastdiv(numVB, B.denV, rVB)};
          fix_div_n<2>(nn, dd, rr, qq);
          dvA = qq[0]; dvB = qq[1];
        }
      }
      A.dv = dvA; B.dv = dvB;
      A.su = qadd(A.uc, duA); B.su = qadd(B.uc, duB);
      A.sv = qadd(A.vc, dvA); B.sv = qadd(B.vc, dvB);
      stsq<W_SU0, 0>(sb + nxt, A.su); stsq<W_SU0, 1>(sb + nxt, B.su);
      stsq<W_SV0, 0>(sb + nxt, A.sv); stsq<W_SV0, 1>(sb + nxt, B.sv);
    }
    // pairwise named barriers between neighbouring warps (see solve_pass2.cu)
    if (warp > 0) asm volatile("bar.sync %0, 64;" ::"r"(warp) : "memory");
    if (warp < NT3 / 32 - 1) asm volatile("bar.sync %0, 64;" ::"r"(warp + 1) : "memory");
    const unsigned t_ = cur; cur = nxt; nxt = t_;
  }

  // ---------------- phase E: store du, dv of the output tile ----------------
  auto store = [&](int gy, const Q& du_, const Q& dv_) {
    if (gy >= oy0 && gy < oy1) {
      float* rdu = a.du_out + (size_t)gy * pitch;
      float* rdvp = a.dv_out + (size_t)gy * pitch;
      // gx is even and rows are 16-byte aligned: the pairs (gx, gx+1) and (gx+2, gx+3) are aligned float2s
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const int x = gx + 2 * k;
        const float2 pu = k ? du_.hi : du_.lo, pv = k ? dv_.hi : dv_.lo;
        if (x >= ox0 && x + 1 < ox1) {
          *reinterpret_cast<float2*>(rdu + x) = pu;
          *reinterpret_cast<float2*>(rdvp + x) = pv;
        } else {
          if (x >= ox0 && x < ox1) { rdu[x] = pu.x; rdvp[x] = pv.x; }
          if (x + 1 >= ox0 && x + 1 < ox1) { rdu[x + 1] = pu.y; rdvp[x + 1] = pv.y; }
        }
      }
    }
  };
  store(gyA, duA, A.dv);
  store(gyB, duB, B.dv);
}

template <bool PAD2>
__global__ void __launch_bounds__(NT3, 1) solve_pass3_kernel(const SolveArgs a, const __grid_constant__ Pass3Maps maps, int tiles_x, int tiles) {
  constexpr unsigned kStageBytes = Stage<PAD2>::kBytes;
  extern __shared__ __align__(128) float sm[];
  if (a.stop && *a.stop) return;  // the level has converged (flow2d_params.residual_tolerance)
  const unsigned smem0 = (unsigned)__cvta_generic_to_shared(sm);
  const unsigned stage0 = smem0 + kPlaneBytes * kWorkPlanes;
  const unsigned bar = stage0 + kMaxStageBytes * kStagePlanes;
  const int tid = threadIdx.x;
  const bool later = a.phi_in != nullptr, have_d = a.du_in != nullptr;
  const unsigned nplanes = later ? 9u : have_d ? 7u : 5u;
  if (a.pdl) asm volatile("griddepcontrol.launch_dependents;");
  int tile = blockIdx.x;
  if (tile >= tiles) return;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // the loads of a tile: planes the previous pass of the solve does not write first, du / dv / phi / ksi after
  // griddepcontrol.wait (programmatic dependent launch: this grid may have started while that pass was draining)
  auto issue = [&](int t, bool first) {
    const Tile3 tl = tile_of(a, t, tiles_x);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy reads of the staging planes
    mbar_expect_tx(bar, nplanes * kStageBytes);
#pragma unroll
    for (int p = 0; p < 5; p++) tma_load_box(stage0 + (T_U + p) * kStageBytes, &maps.m[p], tl.bx0, tl.gy0, bar);
    if (first && a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (have_d) {
      tma_load_box(stage0 + T_DU * kStageBytes, &maps.m[5], tl.bx0, tl.gy0, bar);
      tma_load_box(stage0 + T_DV * kStageBytes, &maps.m[6], tl.bx0, tl.gy0, bar);
    }
    if (later) {
      tma_load_box(stage0 + T_PHI * kStageBytes, &maps.m[7], tl.bx0, tl.gy0, bar);
      tma_load_box(stage0 + T_KSI * kStageBytes, &maps.m[8], tl.bx0, tl.gy0, bar);
    }
  };
  if (tid == 0) issue(tile, true);
  if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");  // before this CTA's first store over the previous pass's input
  unsigned parity = 0;
  for (; tile < tiles; tile += gridDim.x) {
    const Tile3 tl = tile_of(a, tile, tiles_x);
    const int next = tile + gridDim.x;
    auto issue_next = [&]() {
      if (next < tiles) issue(next, false);
    };
    mbar_wait(bar, parity);
    parity ^= 1u;
    if (tl.border) pass3_tile<true, PAD2>(a, tl, smem0, stage0, issue_next);
    else pass3_tile<false, PAD2>(a, tl, smem0, stage0, issue_next);
    __syncthreads();  // the work planes are free for the next tile
  }
}

cudaError_t solve_pass3_configure() {
  cudaError_t e = cudaFuncSetAttribute(solve_pass3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_pass3_smem_bytes());
  if (e == cudaSuccess) e = cudaFuncSetAttribute(solve_pass3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_pass3_smem_bytes());
  return e;
}

// ---- host side: tensor maps (one per input plane of the launch) ----
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
bool make_map(CUtensorMap* m, const float* plane, int w, int h, int pitch, int box_w) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)LH}, estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(plane), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

bool solve_pass3_available() { return encode_tiled() != nullptr; }

// grid_x, grid_y: tiles of the level as for solve_pass2; ctas: persistent CTAs to launch (<= tiles)
bool launch_solve_pass3(cudaStream_t st, const SolveArgs& a, int grid_x, int grid_y, int ctas) {
  Pass3Maps maps;
  const float* planes[9] = {a.u, a.v, a.fx, a.fy, a.ft, a.du_in, a.dv_in, a.phi_in, a.ksi_in};
  const bool pad2 = (a.halo_x & 2) != 0;  // the region starts two columns right of a 16-byte boundary
  for (int i = 0; i < 9; i++)
    if (!make_map(&maps.m[i], planes[i] ? planes[i] : a.u, a.w, a.h, a.pitch, pad2 ? Stage<true>::TW : Stage<false>::TW)) return false;
  const int tiles = grid_x * grid_y;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas < tiles ? ctas : tiles);
  cfg.blockDim = dim3(NT3);
  cfg.dynamicSmemBytes = solve_pass3_smem_bytes();
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = a.pdl ? 1 : 0;  // only between consecutive passes of one solve (see SolveArgs::pdl)
  if (pad2) return cudaLaunchKernelEx(&cfg, solve_pass3_kernel<true>, a, maps, grid_x, tiles) == cudaSuccess;
  return cudaLaunchKernelEx(&cfg, solve_pass3_kernel<false>, a, maps, grid_x, tiles) == cudaSuccess;
}

void preload_solve_pass3_kernels() {
  cudaFuncAttributes at;
  cudaFuncGetAttributes(&at, solve_pass3_kernel<false>);
  cudaFuncGetAttributes(&at, solve_pass3_kernel<true>);
}

}  // namespace flow2d

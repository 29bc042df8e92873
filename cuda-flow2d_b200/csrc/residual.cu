// residual.cu -- opt-in diagnostics beyond the reference: the RMS residual of
// the lagged linear system that the inner sweeps relax.  The reference runs a fixed number of sweeps
// and never looks at a norm (cuda_operation_solve_2d.cpp:229-299); this kernel is off the default
// path (flow2d_params.report_residuals / flow2d_stage_residual) and changes no result.
//
// For the (phi, ksi) of the last outer iteration and the increment (du, dv) left by its sweeps:
//   r_u = ksi*(-J13 - J12*dv - J11*du) + sum_n a_n*((u_n + du_n) - (u + du))
//   r_v = ksi*(-J23 - J12*du - J22*dv) + sum_n a_n*((v_n + dv_n) - (v + dv))
// with the edge weights a_n of the sweep (solve_2d.cu:333-349).  Both vanish at the fixed point of
// the Jacobi update (solve_2d.cu:361-374).  The per-pixel arithmetic is double precision on the fp32
// fields, so the number does not depend on operation order; the sum over the level is a warp-shuffle
// tree + one shared-memory round per CTA + one double atomicAdd per CTA and component.
#include "kernels.h"

namespace flow2d {

// r_u, r_v of one pixel (zero outside the level)
template <bool GRAD>
__device__ __forceinline__ void residual_px(const float* __restrict__ fx, const float* __restrict__ fy, const float* __restrict__ ft,
                                            const ResidualJ& J, const float* __restrict__ u, const float* __restrict__ v,
                                            const float* __restrict__ du, const float* __restrict__ dv, const float* __restrict__ phi,
                                            const float* __restrict__ ksi, int w, int h, int pitch, float hx_2, float hy_2,
                                            double& ru, double& rv) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  ru = 0.0; rv = 0.0;
  if (x < w && y < h) {
    const int xm = mirror_clamp(x - 1, w), xp = mirror_clamp(x + 1, w);
    const int ym = mirror_clamp(y - 1, h), yp = mirror_clamp(y + 1, h);
    const size_t c = (size_t)y * pitch + x, l = (size_t)y * pitch + xm, r = (size_t)y * pitch + xp;
    const size_t up = (size_t)ym * pitch + x, dn = (size_t)yp * pitch + x;
    // the weights exactly as the sweep forms them (fp32), everything after that in double
    const float wxp = hx_2 * ((x < w - 1) ? 1.f : 0.f), wxm = hx_2 * ((x > 0) ? 1.f : 0.f);
    const float wyp = hy_2 * ((y < h - 1) ? 1.f : 0.f), wym = hy_2 * ((y > 0) ? 1.f : 0.f);
    const float pc = phi[c];
    const double axp = wxp * ((phi[r] + pc) * 0.5f), axm = wxm * ((phi[l] + pc) * 0.5f);
    const double ayp = wyp * ((phi[dn] + pc) * 0.5f), aym = wym * ((phi[up] + pc) * 0.5f);
    double J11, J22, J12, J13, J23;
    if (GRAD) {
      J11 = J.p[0][c]; J22 = J.p[1][c]; J12 = J.p[2][c]; J13 = J.p[3][c]; J23 = J.p[4][c];
    } else {
      const float gx = fx[c], gy = fy[c], gt = ft[c];
      J11 = gx * gx; J22 = gy * gy; J12 = gx * gy; J13 = gx * gt; J23 = gy * gt;  // fp32 products, as in the sweep
    }
    const double k = ksi[c], d_u = du[c], d_v = dv[c];
    const double su = (double)u[c] + d_u, sv = (double)v[c] + d_v;
    auto s_u = [&](size_t i) { return (double)u[i] + (double)du[i]; };
    auto s_v = [&](size_t i) { return (double)v[i] + (double)dv[i]; };
    const double lap_u = axm * (s_u(l) - su) + axp * (s_u(r) - su) + ayp * (s_u(dn) - su) + aym * (s_u(up) - su);
    const double lap_v = axm * (s_v(l) - sv) + axp * (s_v(r) - sv) + ayp * (s_v(dn) - sv) + aym * (s_v(up) - sv);
    ru = k * (-J13 - J12 * d_v - J11 * d_u) + lap_u;
    rv = k * (-J23 - J12 * d_u - J22 * d_v) + lap_v;
  }
}

// sum of a, b over the CTA (32 x 8 threads); valid in thread (0, 0)
__device__ __forceinline__ void cta_sum2(double& a, double& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  __shared__ double sa[8], sb[8];
  const int warp = (threadIdx.y * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sa[warp] = a; sb[warp] = b; }
  __syncthreads();
  if (warp == 0) {
    a = lane < 8 ? sa[lane] : 0.0;
    b = lane < 8 ? sb[lane] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
  }
}

template <bool GRAD>
__global__ void __launch_bounds__(256)
residual_kernel(const float* __restrict__ fx, const float* __restrict__ fy, const float* __restrict__ ft, ResidualJ J,
                const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ du,
                const float* __restrict__ dv, const float* __restrict__ phi, const float* __restrict__ ksi, int w, int h,
                int pitch, float hx_2, float hy_2, double* __restrict__ sums) {
  double ru, rv;
  residual_px<GRAD>(fx, fy, ft, J, u, v, du, dv, phi, ksi, w, h, pitch, hx_2, hy_2, ru, rv);
  double a = ru * ru, b = rv * rv;
  cta_sum2(a, b);
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    atomicAdd(&sums[0], a);
    atomicAdd(&sums[1], b);
  }
}

// The convergence test of flow2d_params.residual_tolerance: the same residual, summed in a FIXED order (per-CTA partial
// sums, added up by the last CTA to finish) so that the decision does not depend on the order in which CTAs retire, then
// compared on the device: when both RMS values are <= tol the level's stop word is set to `which` (1: the increment is in
// the solver's result pair, 2: in its scratch pair) and every later kernel of the level returns at once.
template <bool GRAD>
__global__ void __launch_bounds__(256)
residual_decide_kernel(const float* __restrict__ fx, const float* __restrict__ fy, const float* __restrict__ ft, ResidualJ J,
                       const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ du,
                       const float* __restrict__ dv, const float* __restrict__ phi, const float* __restrict__ ksi, int w, int h,
                       int pitch, float hx_2, float hy_2, ResidualDecide d) {
  if (*d.stop) return;
  double ru, rv;
  residual_px<GRAD>(fx, fy, ft, J, u, v, du, dv, phi, ksi, w, h, pitch, hx_2, hy_2, ru, rv);
  double a = ru * ru, b = rv * rv;
  cta_sum2(a, b);
  const unsigned nblocks = gridDim.x * gridDim.y, me = blockIdx.y * gridDim.x + blockIdx.x;
  __shared__ bool last;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    d.partials[2 * me] = a;
    d.partials[2 * me + 1] = b;
    __threadfence();
    last = atomicAdd(d.counter, 1u) == nblocks - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // fixed order: thread t adds partials t, t + 256, ... ; then the fixed tree of cta_sum2
  const unsigned t = threadIdx.y * blockDim.x + threadIdx.x;
  a = 0.0; b = 0.0;
  for (unsigned i = t; i < nblocks; i += 256) {
    a += __ldcg(d.partials + 2 * i);
    b += __ldcg(d.partials + 2 * i + 1);
  }
  __syncthreads();
  cta_sum2(a, b);
  if (t == 0) {
    *d.counter = 0;
    const double n = (double)w * (double)h;
    const double rms_u = sqrt(a / n), rms_v = sqrt(b / n);
    if (d.sums) { d.sums[0] = a; d.sums[1] = b; }
    if (rms_u <= (double)d.tol && rms_v <= (double)d.tol) {
      *d.iterations = d.outer_done;
      __threadfence();
      *d.stop = d.which;
    }
  }
}

void launch_residual(cudaStream_t st, const float* fx, const float* fy, const float* ft, const float* const* J, bool grad,
                     const float* u, const float* v, const float* du, const float* dv, const float* phi, const float* ksi,
                     const LevelGeom& g, float alpha, double* sums) {
  ResidualJ j;
  for (int i = 0; i < 5; i++) j.p[i] = J ? J[i] : nullptr;
  const float hx_2 = alpha / (g.hx * g.hx), hy_2 = alpha / (g.hy * g.hy);  // solve_2d.cu:333-334
  dim3 block(32, 8), grid((g.w + 31) / 32, (g.h + 7) / 8);
  if (grad) residual_kernel<true><<<grid, block, 0, st>>>(fx, fy, ft, j, u, v, du, dv, phi, ksi, g.w, g.h, g.pitch, hx_2, hy_2, sums);
  else residual_kernel<false><<<grid, block, 0, st>>>(fx, fy, ft, j, u, v, du, dv, phi, ksi, g.w, g.h, g.pitch, hx_2, hy_2, sums);
}

void launch_residual_decide(cudaStream_t st, const float* fx, const float* fy, const float* ft, const float* const* J, bool grad,
                            const float* u, const float* v, const float* du, const float* dv, const float* phi, const float* ksi,
                            const LevelGeom& g, float alpha, const ResidualDecide& d) {
  ResidualJ j;
  for (int i = 0; i < 5; i++) j.p[i] = J ? J[i] : nullptr;
  const float hx_2 = alpha / (g.hx * g.hx), hy_2 = alpha / (g.hy * g.hy);
  dim3 block(32, 8), grid((g.w + 31) / 32, (g.h + 7) / 8);
  if (grad) residual_decide_kernel<true><<<grid, block, 0, st>>>(fx, fy, ft, j, u, v, du, dv, phi, ksi, g.w, g.h, g.pitch, hx_2, hy_2, d);
  else residual_decide_kernel<false><<<grid, block, 0, st>>>(fx, fy, ft, j, u, v, du, dv, phi, ksi, g.w, g.h, g.pitch, hx_2, hy_2, d);
}

}  // namespace flow2d

// solve_ext.cu -- the OPT-IN solver extensions of SURVEY.md 8(f) ranks 3 and 4.  sm_100a, compiled with -fmad=false.
//
// None of this exists upstream: the reference relaxes with a fixed number of Jacobi sweeps without a relaxation factor
// (solve_2d.cu:361-374, cuda_operation_solve_2d.cpp:229-299), builds the gradient-constancy tensor from derivative
// planes whose halo is the 16x8 CUDA block's own edge value (solve_2d.cu:813-841; same in solve_2d_log, 391-669) and
// never combines the brightness and gradient terms its README lists (README.md:32-34).  The kernels below implement
//   - a relaxation factor omega and red-black ordering (flow2d_params.scheme, .omega),
//   - data terms on tensor planes with TRUE neighbour halos: gradient constancy, gradient constancy of log(1+f), and
//     brightness + gamma * gradient (flow2d_params.data_term, .gamma), with the robust data weight psi'(w^T J w) on the
//     full tensor,
// operation for operation as the CPU specification of the extensions in the test tree defines them (include/flow2d.h names it; never
// reference parity).  They are off the default path: flow2d_params with all extension fields zero never gets here.
//
// Decomposition: one thread per pixel and one launch per sweep (red-black: per colour), i.e. the reference's own
// HBM-bound shape, 6 tensor planes + phi + ksi + u + v + du + dv per pixel and sweep.  The temporally blocked kernels of
// solve_pass2.cu stay reserved for the reference-exact iteration; red-black ordering shrinks the exact region of a tile
// by two rings per sweep and would halve their output tile.
//
// Early exit (flow2d_params.residual_tolerance): every kernel of a level reads the level's stop word first and returns
// when it is set (written by residual.cu: launch_residual_decide), so a converged level costs empty launches only --
// no host round trip, and the schedule stays a replayable CUDA graph.
#include "kernels.h"

namespace flow2d {

// log(1 + max(f, 0)), the transform of solve_2d.cu:509-524 (which yields NaN for f <= -1), in double and rounded once
// (so that CPU and GPU agree except where the two double logarithms differ in their last place)
__global__ void __launch_bounds__(256) ext_log_kernel(const float* __restrict__ in, float* __restrict__ out, int w, int h, int pitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const size_t c = (size_t)y * pitch + x;
  out[c] = (float)log((double)fmaxf(in[c], 0.f) + 1.0);
}

// Tensor planes J11 J22 J12 J13 J23 J33 from fx, fy, ft (launch_derivatives): brightness B = (fx, fy, ft)(fx, fy, ft)^T,
// gradient G from central differences of the derivative planes over the true (mirrored) neighbours.
__global__ void __launch_bounds__(256)
ext_tensor_kernel(const float* __restrict__ fx, const float* __restrict__ fy, const float* __restrict__ ft, ExtTensor J,
                  int w, int h, int pitch, float hx_1, float hy_1, int term, float gamma) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int xm = mirror_clamp(x - 1, w), xp = mirror_clamp(x + 1, w), ym = mirror_clamp(y - 1, h), yp = mirror_clamp(y + 1, h);
  const size_t row = (size_t)y * pitch, up = (size_t)ym * pitch, dn = (size_t)yp * pitch, c = row + x;
  const float gx = fx[c], gy = fy[c], gt = ft[c];
  const float B[6] = {gx * gx, gy * gy, gx * gy, gx * gt, gy * gt, gt * gt};
  float G[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (term != FLOW2D_TERM_DEFAULT) {
    const float fxx = (fx[row + xp] - fx[row + xm]) * hx_1, fxy = (fx[dn + x] - fx[up + x]) * hy_1;
    const float fyy = (fy[dn + x] - fy[up + x]) * hy_1;
    const float fxt = (ft[row + xp] - ft[row + xm]) * hx_1, fyt = (ft[dn + x] - ft[up + x]) * hy_1;
    G[0] = fmaf(fxx, fxx, fxy * fxy);
    G[1] = fmaf(fxy, fxy, fyy * fyy);
    G[2] = fmaf(fxx, fxy, fxy * fyy);
    G[3] = fmaf(fxx, fxt, fxy * fyt);
    G[4] = fmaf(fxy, fxt, fyy * fyt);
    G[5] = fmaf(fxt, fxt, fyt * fyt);
  }
#pragma unroll
  for (int k = 0; k < 6; k++) {
    float v;
    if (term == FLOW2D_TERM_DEFAULT) v = B[k];
    else if (term == FLOW2D_TERM_COMBINED) v = fmaf(gamma, G[k], B[k]);
    else v = G[k];
    J.p[k][c] = v;
  }
}

// phi as solve_2d.cu:141-162; ksi = psi'(w^T J w), w = (du, dv, 1), on the full tensor
__global__ void __launch_bounds__(256)
ext_phi_ksi_kernel(ExtTensor J, const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ du,
                   const float* __restrict__ dv, float* __restrict__ phi, float* __restrict__ ksi, int w, int h, int pitch,
                   float hx2, float hy2, float e_smooth, float e_data, const int* __restrict__ stop) {
  if (stop && *stop) return;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int xm = mirror_clamp(x - 1, w), xp = mirror_clamp(x + 1, w), ym = mirror_clamp(y - 1, h), yp = mirror_clamp(y + 1, h);
  const size_t row = (size_t)y * pitch, up = (size_t)ym * pitch, dn = (size_t)yp * pitch, c = row + x;
  float zl = 0.f, zr = 0.f, zu = 0.f, zd = 0.f, wl = 0.f, wr = 0.f, wu = 0.f, wd = 0.f, d_u = 0.f, d_v = 0.f;
  if (du) {  // null: the increment is still zero (first outer iteration)
    zl = du[row + xm]; zr = du[row + xp]; zu = du[up + x]; zd = du[dn + x];
    wl = dv[row + xm]; wr = dv[row + xp]; wu = dv[up + x]; wd = dv[dn + x];
    d_u = du[c]; d_v = dv[c];
  }
  const float dux = (((u[row + xp] - u[row + xm]) + zr) - zl) / hx2;
  const float duy = (((u[dn + x] - u[up + x]) + zd) - zu) / hy2;
  const float dvx = (((v[row + xp] - v[row + xm]) + wr) - wl) / hx2;
  const float dvy = (((v[dn + x] - v[up + x]) + wd) - wu) / hy2;
  float t = duy * duy;
  t = fmaf(dux, dux, t);
  t = fmaf(dvx, dvx, t);
  t = fmaf(dvy, dvy, t);
  t = fmaf(e_smooth, e_smooth, t);
  const float r = sqrtf(t);
  phi[c] = 1.f / (r + r);
  const float J11 = J.p[0][c], J22 = J.p[1][c], J12 = J.p[2][c], J13 = J.p[3][c], J23 = J.p[4][c], J33 = J.p[5][c];
  const float a = J13 + fmaf(J11, d_u, J12 * d_v);
  const float b = J23 + fmaf(J12, d_u, J22 * d_v);
  const float cc = J33 + fmaf(J13, d_u, J23 * d_v);
  float s = fmaf(d_u, a, d_v * b) + cc;
  s = s * ((s > 0.f) ? 1.f : 0.f);
  const float q = sqrtf(fmaf(e_data, e_data, s));
  ksi[c] = 1.f / (q + q);
}

// One relaxation of every pixel (JACOBI: du_in -> du_out) or of the pixels of one colour in place (red-black:
// du_out == du_in, (x + y) & 1 == colour).  The update is the reference's (solve_2d.cu:332-374) blended with the old
// value by omega.
template <bool RED_BLACK>
__global__ void __launch_bounds__(256)
ext_sweep_kernel(ExtTensor J, const float* __restrict__ u, const float* __restrict__ v, const float* du_in, const float* dv_in,
                 const float* __restrict__ phi, const float* __restrict__ ksi, float* du_out, float* dv_out, int w, int h,
                 int pitch, float hx_2, float hy_2, float omega, int colour, const int* __restrict__ stop) {
  if (stop && *stop) return;
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (RED_BLACK) x = 2 * x + ((y + colour) & 1);  // the cells of this colour in row y
  if (x >= w || y >= h) return;
  const int xm = mirror_clamp(x - 1, w), xp = mirror_clamp(x + 1, w), ym = mirror_clamp(y - 1, h), yp = mirror_clamp(y + 1, h);
  const size_t row = (size_t)y * pitch, up = (size_t)ym * pitch, dn = (size_t)yp * pitch, c = row + x;
  const float wxp = hx_2 * ((x < w - 1) ? 1.f : 0.f), wxm = hx_2 * ((x > 0) ? 1.f : 0.f);
  const float wyp = hy_2 * ((y < h - 1) ? 1.f : 0.f), wym = hy_2 * ((y > 0) ? 1.f : 0.f);
  const float pc = phi[c];
  const float axp = wxp * ((phi[row + xp] + pc) * 0.5f), axm = wxm * ((phi[row + xm] + pc) * 0.5f);
  const float ayp = wyp * ((phi[dn + x] + pc) * 0.5f), aym = wym * ((phi[up + x] + pc) * 0.5f);
  const float sumH = ((axp + axm) + ayp) + aym;
  const float uc = u[c], vc = v[c];
  float sumU = axm * ((u[row + xm] + du_in[row + xm]) - uc);
  sumU = fmaf(axp, (u[row + xp] + du_in[row + xp]) - uc, sumU);
  sumU = fmaf(ayp, (u[dn + x] + du_in[dn + x]) - uc, sumU);
  sumU = fmaf(aym, (u[up + x] + du_in[up + x]) - uc, sumU);
  float sumV = axm * ((v[row + xm] + dv_in[row + xm]) - vc);
  sumV = fmaf(axp, (v[row + xp] + dv_in[row + xp]) - vc, sumV);
  sumV = fmaf(ayp, (v[dn + x] + dv_in[dn + x]) - vc, sumV);
  sumV = fmaf(aym, (v[up + x] + dv_in[up + x]) - vc, sumV);
  const float k = ksi[c], J11 = J.p[0][c], J22 = J.p[1][c], J12 = J.p[2][c], J13 = J.p[3][c], J23 = J.p[4][c];
  const float du_old = du_in[c], dv_old = dv_in[c];
  float r_du = fmaf(k, fmaf(-J12, dv_old, -J13), sumU) / fmaf(J11, k, sumH);
  if (omega != 1.f) r_du = fmaf(omega, r_du - du_old, du_old);
  float r_dv = fmaf(k, fmaf(-J12, r_du, -J23), sumV) / fmaf(J22, k, sumH);
  if (omega != 1.f) r_dv = fmaf(omega, r_dv - dv_old, dv_old);
  du_out[c] = r_du;
  dv_out[c] = r_dv;
}

// After an early exit the increment may sit in the scratch pair: *stop == 2 means "result in (src_du, src_dv)".
__global__ void __launch_bounds__(256)
ext_pick_kernel(const int* __restrict__ stop, const float* __restrict__ src_du, const float* __restrict__ src_dv,
                float* __restrict__ dst_du, float* __restrict__ dst_dv, int w, int h, int pitch) {
  if (*stop != 2) return;
  const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const size_t c = (size_t)y * pitch + x;  // whole float4s: pitch % 4 == 0, the padding is never read as data
  *reinterpret_cast<float4*>(dst_du + c) = *reinterpret_cast<const float4*>(src_du + c);
  *reinterpret_cast<float4*>(dst_dv + c) = *reinterpret_cast<const float4*>(src_dv + c);
}

static dim3 px_grid(int w, int h) { return dim3((w + 31) / 32, (h + 7) / 8); }

void launch_ext_log(cudaStream_t st, const float* in, float* out, const LevelGeom& g) {
  ext_log_kernel<<<px_grid(g.w, g.h), dim3(32, 8), 0, st>>>(in, out, g.w, g.h, g.pitch);
}

void launch_ext_tensor(cudaStream_t st, const float* fx, const float* fy, const float* ft, const ExtTensor& J, const LevelGeom& g,
                       int term, float gamma) {
  const float hx_1 = (float)(1.0 / (2.0 * (double)g.hx)), hy_1 = (float)(1.0 / (2.0 * (double)g.hy));
  ext_tensor_kernel<<<px_grid(g.w, g.h), dim3(32, 8), 0, st>>>(fx, fy, ft, J, g.w, g.h, g.pitch, hx_1, hy_1, term, gamma);
}

void launch_ext_phi_ksi(cudaStream_t st, const ExtTensor& J, const float* u, const float* v, const float* du, const float* dv,
                        float* phi, float* ksi, const LevelGeom& g, float e_smooth, float e_data, const int* stop) {
  ext_phi_ksi_kernel<<<px_grid(g.w, g.h), dim3(32, 8), 0, st>>>(J, u, v, du, dv, phi, ksi, g.w, g.h, g.pitch, g.hx + g.hx,
                                                                g.hy + g.hy, e_smooth, e_data, stop);
}

void launch_ext_sweep(cudaStream_t st, const ExtTensor& J, const float* u, const float* v, const float* du_in, const float* dv_in,
                      const float* phi, const float* ksi, float* du_out, float* dv_out, const LevelGeom& g, float alpha,
                      float omega, int colour, const int* stop) {
  const float hx_2 = alpha / (g.hx * g.hx), hy_2 = alpha / (g.hy * g.hy);  // solve_2d.cu:333-334
  if (colour < 0)
    ext_sweep_kernel<false><<<px_grid(g.w, g.h), dim3(32, 8), 0, st>>>(J, u, v, du_in, dv_in, phi, ksi, du_out, dv_out, g.w, g.h,
                                                                       g.pitch, hx_2, hy_2, omega, 0, stop);
  else
    ext_sweep_kernel<true><<<px_grid((g.w + 1) / 2, g.h), dim3(32, 8), 0, st>>>(J, u, v, du_in, dv_in, phi, ksi, du_out, dv_out,
                                                                                g.w, g.h, g.pitch, hx_2, hy_2, omega, colour, stop);
}

void launch_ext_pick(cudaStream_t st, const int* stop, const float* src_du, const float* src_dv, float* dst_du, float* dst_dv,
                     const LevelGeom& g) {
  ext_pick_kernel<<<dim3((g.w + 127) / 128, (g.h + 7) / 8), dim3(32, 8), 0, st>>>(stop, src_du, src_dv, dst_du, dst_dv, g.w, g.h, g.pitch);
}

}  // namespace flow2d

// kernels.h -- launch interface of the sm_100a kernels (internal; the public boundary is include/flow2d.h).
#pragma once

#include "../../include/flow2d.h"
#include "common.cuh"

namespace flow2d {

// ---- pyramid.cu ----
void launch_blur(cudaStream_t st, const float* in, float* out, int w, int h, int pitch, const GaussTaps& taps);
void launch_resample(cudaStream_t st, const float* const* in, float* const* tmp, float* const* out, int count,
                     int iw, int ih, int ow, int oh, int pitch);
// the same for up to four images of different sizes in one x launch + one y launch (tmp holds ow x ih)
struct ResampleJob {
  const float* in;
  float* tmp;
  float* out;
  int iw, ih, ow, oh;
  int oy0, oy1;  // output rows to produce (oy1 <= oy0: all of them); row-slab mode
  int iy0, iy1;  // filled in by the launcher: the input rows those output rows are made of
  int chunk;     // filled in by the launcher: outputs per CTA of the staged x pass
};
void launch_resample_batch(cudaStream_t st, const ResampleJob* jobs, int count, int pitch);
// y0, y1: rows of the level to produce (y1 <= y0: all of them); row-slab mode
void launch_warp(cudaStream_t st, const float* f0, const float* f1, const float* u, const float* v, float* out,
                 const LevelGeom& g, int y0 = 0, int y1 = 0);
void launch_derivatives(cudaStream_t st, const float* f0, const float* f1w, float* fx, float* fy, float* ft,
                        const LevelGeom& g, int y0 = 0, int y1 = 0);
void launch_grad_tensor(cudaStream_t st, const float* fx, const float* fy, const float* ft, float* const* J,
                        const LevelGeom& g, int y0 = 0, int y1 = 0);

// ---- median.cu ----
void launch_add_median(cudaStream_t st, const float* const* a, const float* const* b, float* const* out, int count,
                       int w, int h, int pitch, int radius, int row0 = 0, int row1 = 0);
void launch_add(cudaStream_t st, float* a, const float* b, int w, int h, int pitch);

// ---- solve.cu ----
constexpr int kSolveLW = 64, kSolveLH = 48;  // shared-memory region of one CTA (output tile + halo)

struct SolveArgs {
  const float *fx, *fy, *ft;      // brightness derivatives of the level (launch_derivatives)
  const float* J[5];              // gradient-constancy tensor J11 J22 J12 J13 J23 (gradient mode only)
  const float *u, *v;             // flow of the level (constant during the solve)
  const float *du_in, *dv_in;     // current increment; null = all zero (first pass of a level)
  const float *phi_in, *ksi_in;   // null = compute the robust weights in this pass from du_in/dv_in
  float *du_out, *dv_out;
  float *phi_out, *ksi_out;       // null = do not store the robust weights
  int w, h, pitch;
  float hx, hy, alpha, e_smooth, e_data;
  float hx_2, hy_2;               // alpha / hx^2, alpha / hy^2 (solve_2d.cu:333-334), computed by the caller in fp32
  int sweeps;                     // Jacobi sweeps in this pass
  int outer;                      // 1, or (resident mode, grid 1x1) the number of outer iterations
  int ow, oh;                     // output tile of one CTA (ow % 4 == 0)
  int halo_x, halo_y;             // region origin = tile origin - halo; halo_x % 4 == 0
  int y0, y1;                     // rows of the level this launch produces (0, h unless the level is slabbed)
  unsigned long long* timing;     // debug: 8 globaltimer stamps per CTA (null = off)
  int exact;                      // one-pixel kernels: 1 = plain IEEE div / sqrt / rcp from the start (see one_px_outer)
  const int* stop;                // early exit (flow2d_params.residual_tolerance): return at once when *stop != 0; null = off
  int pdl;                        // 1 = launched with programmatic stream serialization: the previous kernel on the
                                  // stream is the previous pass of this solve (it writes only du/dv/phi/ksi)
};

size_t solve_pass_smem_bytes();
cudaError_t solve_pass_configure();
// rows > 0: launch only that many region rows (resident mode of a level with few rows)
void launch_solve_pass(cudaStream_t st, const SolveArgs& a, bool grad, int grid_x, int grid_y, int rows = 0);

// second generation of the tiled pass (solve_pass2.cu): eight pixels per thread, packed fp32; a.outer must be 1
size_t solve_pass2_smem_bytes();
cudaError_t solve_pass2_configure();
void launch_solve_pass2(cudaStream_t st, const SolveArgs& a, bool grad, int grid_x, int grid_y);

// third generation (solve_pass3.cu): persistent CTAs, input tiles by TMA (cp.async.bulk.tensor + mbarrier) prefetched during
// the sweeps of the previous tile; brightness constancy only.  launch_solve_pass3 returns false when the tensor maps or the
// launch could not be made (the caller then uses solve_pass2).  ctas: persistent CTAs to launch (the SM count).
size_t solve_pass3_smem_bytes();
cudaError_t solve_pass3_configure();
bool solve_pass3_available();
bool launch_solve_pass3(cudaStream_t st, const SolveArgs& a, int grid_x, int grid_y, int ctas);

// one pass of a mid-size level with one thread per pixel: ts x ts regions (ts = 32, 24 or 16), a.halo_x = a.halo_y =
// a.sweeps + 1, a.ow = a.oh = ts - 2 * halo; phi/ksi are always computed in the pass (a.phi_in must be null)
constexpr int kSmallTS = 32;
void launch_solve_small_pass(cudaStream_t st, const SolveArgs& a, bool grad, int grid_x, int grid_y, int ts = kSmallTS);
// whole solve of a level of <= 1024 pixels in one CTA, one thread per pixel (a.outer, a.sweeps = inner)
bool solve_tiny_fits(int w, int h);
void launch_solve_tiny(cudaStream_t st, const SolveArgs& a, bool grad);

// ---- solve_cluster.cu: a mid-size level on one thread-block cluster (one thread per pixel, halos through distributed
// shared memory, barrier.cluster per sweep) ----
constexpr int kClusterMaxCtas = 16;   // 8 is the portable cluster size; 16 needs cudaFuncAttributeNonPortableClusterSizeAllowed
constexpr int kClusterMinBlock = 8;   // a CTA's block is at least 8 x 8 cells (bounds its shared planes, see cluster_plane)
struct ClusterGeom {
  int cx, cy;   // CTAs of a cluster in x and y: launched as (cx * cy, 1, 1), rank r at (r % cx, r / cx)
  int tw, th;   // block of one CTA, tw * th <= threads, both >= kClusterMinBlock
  int ncx;      // regions per row of the level (pass mode); 1 = the region covers the level
};
// whole level: a.outer / a.sweeps = all iterations, a.ow = cx * tw, a.oh = cy * th, a.halo_x = a.halo_y = 0, clusters = 1
// pass: a.outer = 1, a.halo_x = a.halo_y = a.sweeps + 1, a.ow = cx * tw - 2 * halo, a.oh = cy * th - 2 * halo,
//       clusters = ncx * region rows.  threads: whole warps covering the block (tw * th <= threads <= 1024).
void launch_solve_cluster(cudaStream_t st, const SolveArgs& a, bool grad, const ClusterGeom& cg, int threads, int clusters);
// clusters of `csize` CTAs x `threads` threads the device holds at once; 0 = that shape cannot be launched here
int solve_cluster_max_active(int csize, int threads);
void preload_solve_cluster_kernels();

// ---- slab.cu: halo exchange between neighbour GPUs through peer-mapped mailboxes ----
struct SlabPushDir {
  float* dst[2];                // receive buffers of the two fields in the neighbour's mailbox (peer mapping)
  unsigned long long* flag;     // the neighbour's flag for messages from this side
  unsigned long long epoch;
  int row0, rows;               // rows [row0, row0 + rows) of the sender's containers; rows = 0: nothing to send
};
struct SlabPush {
  const float* field[2];
  SlabPushDir dir[2];           // 0: to the rank above, 1: to the rank below
  unsigned* counter;            // sender-local, zero between launches
  int w, pitch;
};
struct SlabUnpackDir {
  const float* src[2];          // receive buffers in the own mailbox
  const unsigned long long* flag;
  unsigned long long epoch;     // wait until *flag >= epoch
  int row0, rows;               // destination rows in the own containers
};
struct SlabUnpack {
  float* field[2];
  SlabUnpackDir dir[2];         // 0: from the rank above, 1: from the rank below
  unsigned* error;              // set to 1 when a wait timed out (flow2d_slab_status)
  int w, pitch;
};
void launch_slab_push(cudaStream_t st, const SlabPush& p);
void launch_slab_unpack(cudaStream_t st, const SlabUnpack& p);

// loads every kernel of a translation unit now instead of at its first launch (see pyramid.cu: preload_pyramid_kernels)
void preload_pyramid_kernels();
void preload_median_kernels();
void preload_solve_kernels();
void preload_solve_pass2_kernels();
void preload_solve_pass3_kernels();
void preload_slab_kernels();

// ---- residual.cu (opt-in diagnostics, not on the default path) ----
struct ResidualJ { const float* p[5]; };
// adds the sums of r_u^2 and r_v^2 over the level to sums[0], sums[1]
void launch_residual(cudaStream_t st, const float* fx, const float* fy, const float* ft, const float* const* J, bool grad,
                     const float* u, const float* v, const float* du, const float* dv, const float* phi, const float* ksi,
                     const LevelGeom& g, float alpha, double* sums);

// the convergence test of flow2d_params.residual_tolerance (see residual.cu)
struct ResidualDecide {
  double* partials;      // 2 doubles per CTA of the launch
  unsigned* counter;     // zero between launches
  double* sums;          // optional: receives the two sums of squares (overwritten)
  int* stop;             // the level's stop word
  int* iterations;       // receives outer_done when the level stops
  float tol;
  int which, outer_done;
};
void launch_residual_decide(cudaStream_t st, const float* fx, const float* fy, const float* ft, const float* const* J, bool grad,
                            const float* u, const float* v, const float* du, const float* dv, const float* phi, const float* ksi,
                            const LevelGeom& g, float alpha, const ResidualDecide& d);

// ---- solve_ext.cu (opt-in extensions: relaxation factor, red-black ordering, tensor data terms) ----
struct ExtTensor { float* p[6]; };  // J11 J22 J12 J13 J23 J33
void launch_ext_log(cudaStream_t st, const float* in, float* out, const LevelGeom& g);
void launch_ext_tensor(cudaStream_t st, const float* fx, const float* fy, const float* ft, const ExtTensor& J, const LevelGeom& g,
                       int term, float gamma);
void launch_ext_phi_ksi(cudaStream_t st, const ExtTensor& J, const float* u, const float* v, const float* du, const float* dv,
                        float* phi, float* ksi, const LevelGeom& g, float e_smooth, float e_data, const int* stop);
// colour < 0: Jacobi (du_in -> du_out); 0 / 1: the cells of that colour in place (du_out == du_in)
void launch_ext_sweep(cudaStream_t st, const ExtTensor& J, const float* u, const float* v, const float* du_in, const float* dv_in,
                      const float* phi, const float* ksi, float* du_out, float* dv_out, const LevelGeom& g, float alpha,
                      float omega, int colour, const int* stop);
void launch_ext_pick(cudaStream_t st, const int* stop, const float* src_du, const float* src_dv, float* dst_du, float* dst_dv,
                     const LevelGeom& g);

}  // namespace flow2d

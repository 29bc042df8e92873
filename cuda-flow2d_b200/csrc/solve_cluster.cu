// solve_cluster.cu -- the solve of a MID-SIZE level on one thread-block CLUSTER: up to 16 CTAs, one thread per pixel,
// every CTA holding a tw x th block of the level (plus a one-cell ring) in its shared memory; a cell on the edge of a
// block is pushed into the ring of the CTA next door (st.shared::cluster over the SM-to-SM network) whenever it is
// published, so every read is a local LDS, and the CTAs meet at one barrier.cluster per Jacobi sweep.
// sm_100a, compiled with -fmad=false.
//
// Replaces the launch loop of CudaOperationSolve2D::Execute (src/cuda_operations/2d/cuda_operation_solve_2d.cpp:238-300:
// `outer` x (compute_phi_ksi + `inner` x solve_2d), one cuStreamSynchronize each) for levels of up to 16 x 1024 pixels,
// with identical results (the arithmetic is one_px_outer of solve_onepx.cuh, shared with solve_tiny / solve_small_pass).
//
// Two uses of the same kernel:
//   whole level   the cluster's region (cx*tw) x (cy*th) covers the level: ALL outer iterations and inner sweeps run in
//                 ONE launch without leaving the SMs.  No halo at all: the only redundant cells are the padding of the
//                 region beyond the image (inert).  solve_small_pass needs one launch per outer iteration and computes
//                 1024 cells for 400 results (32x32 region, halo S+1 = 6); here every cell computed is a result.
//   pass          a grid of clusters, each solving one pass (phi, ksi, weights + S sweeps) of a region with an S+1 halo,
//                 like solve_small_pass with a 128x128 instead of a 32x32 region: 82 % of the cells are results.
//
// A cluster is launched as (C, 1, 1) CTAs; rank r sits at (r % cx, r / cx) of the region.  An edge cell's push addresses
// are mapped once (mapa) and every plane is an immediate offset from them.  The mirrored image border (-1 -> 1,
// n -> n-2) is folded into the neighbour addresses as in solve_tiny.  The "were the fast divisions safe" vote of
// one_px_outer is cluster-wide (ClusterLink): all CTAs repeat an outer iteration or none does, so every CTA executes the
// same number of cluster barriers.
//
// Measured on B200 (profiles/r02/cluster_ab/, solve_cluster_ncu_summary.txt): a barrier-to-barrier phase
// takes 0.31 us (the barrier: MEMBAR.ALL.GPU for the pushed stores + UCGABAR + CCTL.IVALL) + 0.0155 us per warp of the
// CTA, so a level takes about as long as with one solve_small_pass launch per outer iteration (<= 4 096 px) or longer --
// but on <= 16 CTAs instead of 40 launches of 16-81: the scheduler uses it where SM time counts (flow2d_api.cu:
// cluster_mode).  Next step: st.async + mbarrier::complete_tx between neighbours instead of the full barrier.
#include "kernels.h"
#include "solve_common.cuh"
#include "solve_onepx.cuh"
#include "solve_cluster_geom.h"

namespace flow2d {

template <bool GRAD, int N>
__global__ void __launch_bounds__(N, 1) solve_cluster_kernel(const SolveArgs a, const ClusterGeom cg) {
  constexpr int PS = cluster_plane(N);
  __shared__ __align__(16) float sq[kOnePxPlanes * PS];
  __shared__ unsigned redo_word;
  const int w = a.w, h = a.h;
  const int t = threadIdx.x;
  const unsigned rank = cluster_ctarank();
  ClusterLevel lv;
  lv.w = w; lv.h = h; lv.ow = a.ow; lv.oh = a.oh; lv.halo = a.halo_x; lv.y0 = a.y0; lv.y1 = a.y1;
  const ClusterCell cc = cluster_cell(cg, lv, (int)rank, (int)cluster_id_x(), t);  // (solve_cluster_geom.h)
  const int gx = cc.gx, gy = cc.gy;
  const size_t g = (size_t)min(max(gy, 0), h - 1) * a.pitch + min(max(gx, 0), w - 1);
  asm volatile("griddepcontrol.launch_dependents;");  // the next kernel of the stream may be scheduled (it waits for us)

  OnePx c;
  ClusterLink link;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sq);
  c.ac = base + 4u * (unsigned)cc.ac;
  c.al = base + 4u * (unsigned)cc.al; c.ar = base + 4u * (unsigned)cc.ar;
  c.au = base + 4u * (unsigned)cc.au; c.ad = base + 4u * (unsigned)cc.ad;
  link.push_h = cc.push_h_rank >= 0 ? cluster_map(base + 4u * (unsigned)cc.push_h, (unsigned)cc.push_h_rank) : 0u;
  link.push_v = cc.push_v_rank >= 0 ? cluster_map(base + 4u * (unsigned)cc.push_v, (unsigned)cc.push_v_rank) : 0u;
  link.redo_flag = (unsigned)__cvta_generic_to_shared(&redo_word);
  link.ranks = (unsigned)(cg.cx * cg.cy);
  link.epoch = 0u;
  c.uc = a.u[g]; c.vc = a.v[g]; c.fx = a.fx[g]; c.fy = a.fy[g]; c.ft = a.ft[g];
  if (GRAD) {
    c.J11 = a.J[0][g]; c.J22 = a.J[1][g]; c.nJ12 = -a.J[2][g]; c.nJ13 = -a.J[3][g]; c.nJ23 = -a.J[4][g];
  } else {
    c.J11 = c.fx * c.fx; c.J22 = c.fy * c.fy; c.nJ12 = -(c.fx * c.fy); c.nJ13 = -(c.fx * c.ft); c.nJ23 = -(c.fy * c.ft);
  }
  if (t == 0) redo_word = 0u;
  // every CTA of the cluster is running (its shared memory exists) before anyone stores into it
  cluster_barrier();
  pub<PS, Q_U, true>(c.ac, link, c.uc);
  pub<PS, Q_V, true>(c.ac, link, c.vc);
  if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  float du = 0.f, dv = 0.f;
  if (a.du_in) { du = a.du_in[g]; dv = a.dv_in[g]; }
  c.hx2 = a.hx + a.hx; c.hy2 = a.hy + a.hy;
  c.rhx2 = fast_path_rcp(c.hx2); c.rhy2 = fast_path_rcp(c.hy2);
  c.wxp = a.hx_2 * ((gx < w - 1) ? 1.f : 0.f); c.wxm = a.hx_2 * ((gx > 0) ? 1.f : 0.f);
  c.wyp = a.hy_2 * ((gy < h - 1) ? 1.f : 0.f); c.wym = a.hy_2 * ((gy > 0) ? 1.f : 0.f);
  c.e_smooth = a.e_smooth; c.e_data = a.e_data;
  c.live = cc.live != 0;

  float phi = 0.f, ksi = 0.f;
  // (the first cluster barrier inside one_px_outer orders the stores above against the neighbours' reads; the last one
  // follows the last store into another CTA's shared memory, so a CTA may exit as soon as it is through)
  for (int outer = 0; outer < a.outer; ++outer) {
    link.epoch = (unsigned)outer + 1u;
    one_px_outer<PS, true>(c, a.sweeps, a.exact != 0, du, dv, phi, ksi, link);
  }
  if (cc.out) {
    const size_t o = (size_t)gy * a.pitch + gx;
    a.du_out[o] = du;
    a.dv_out[o] = dv;
    if (a.phi_out) { a.phi_out[o] = phi; a.ksi_out[o] = ksi; }
  }
}

namespace {

template <int N>
const void* cluster_kernel(bool grad) {
  return grad ? (const void*)solve_cluster_kernel<true, N> : (const void*)solve_cluster_kernel<false, N>;
}
const void* cluster_kernel(bool grad, int threads) {
  return threads <= 256 ? cluster_kernel<256>(grad) : threads <= 512 ? cluster_kernel<512>(grad) : cluster_kernel<1024>(grad);
}

}  // namespace

// How many clusters of `csize` CTAs x `threads` threads the device can hold at once (0 = that shape cannot be launched).
int solve_cluster_max_active(int csize, int threads) {
  if (csize < 1 || csize > kClusterMaxCtas) return 0;
  int total = 0;
  for (int grad = 0; grad < 2; ++grad) {
    const void* fn = cluster_kernel(grad != 0, threads);
    if (csize > 8 && cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      (void)cudaGetLastError();
      return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize);
    cfg.blockDim = dim3(threads <= 256 ? 256 : threads <= 512 ? 512 : 1024);
    cfg.dynamicSmemBytes = 0;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess) {
      (void)cudaGetLastError();
      return 0;
    }
    if (grad == 0 || n < total) total = n;
  }
  return total;
}

// clusters: regions of the level (1 in whole-level mode); cg.cx * cg.cy CTAs each; `threads`: whole warps covering a block
// (the instantiation, i.e. the size of the shared planes, follows from it: up to 256, 512 or 1024 cells)
void launch_solve_cluster(cudaStream_t st, const SolveArgs& a, bool grad, const ClusterGeom& cg, int threads, int clusters) {
  const int csize = cg.cx * cg.cy;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize * clusters);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = a.pdl ? 2 : 1;
  void* args[2] = {(void*)&a, (void*)&cg};
  cudaLaunchKernelExC(&cfg, cluster_kernel(grad, threads), args);
}

void preload_solve_cluster_kernels() {
  cudaFuncAttributes fa;
  for (int grad = 0; grad < 2; ++grad) {
    cudaFuncGetAttributes(&fa, cluster_kernel<256>(grad != 0));
    cudaFuncGetAttributes(&fa, cluster_kernel<512>(grad != 0));
    cudaFuncGetAttributes(&fa, cluster_kernel<1024>(grad != 0));
  }
}

}  // namespace flow2d

// solve_cluster_geom.h -- where a thread of solve_cluster_kernel sits: its cell of the level, the shared-memory offsets of
// the cell and of its four neighbours, and where an edge cell is pushed to in the CTA next door.  Host + device: the
// kernel (solve_cluster.cu) and flow2d_debug_cluster_cell (flow2d_api.cu; tests/test_cluster_cpu.py runs a stencil over
// these very indices on the CPU) share this one function, so the index algebra is tested without a GPU.
#pragma once

#include "kernels.h"

namespace flow2d {

// Shared planes of a CTA: its tw x th block with a one-cell ring, (tw+2) x (th+2) floats; the ring is written by the
// neighbouring CTAs only.  cluster_plane(N) bounds that for every block the planner makes (tw, th >= kClusterMinBlock,
// tw * th <= N): (tw+2)(th+2) <= N + 2 (8 + N/8) + 4.
__host__ __device__ constexpr int cluster_plane(int n) { return n + n / 4 + 20; }

struct ClusterCell {
  int ac, al, ar, au, ad;   // float offsets within a plane of the own CTA: own cell, left, right, up, down
  int push_h_rank, push_h;  // CTA of the cluster and offset there that mirror this cell (left / right edge); rank -1 = none
  int push_v_rank, push_v;  // the same for the top / bottom edge
  int gx, gy;               // the cell's pixel of the level (may lie outside the image)
  int mine;                 // 0 = the thread has no cell (it only takes part in the barriers)
  int live;                 // the cell is a pixel of the image
  int out;                  // the cell belongs to the output tile of the cluster
};

// geometry of a launch: ClusterGeom + the fields of SolveArgs that place the cluster's region on the level
struct ClusterLevel {
  int w, h;        // the level
  int ow, oh;      // output tile of one cluster
  int halo;        // region origin = tile origin - halo (0: the region covers the level)
  int y0, y1;      // rows of the level this launch produces
};

__host__ __device__ inline ClusterCell cluster_cell(const ClusterGeom& cg, const ClusterLevel& lv, int rank, int cl, int t) {
  ClusterCell c;
  const int cly = cl / cg.ncx, clx = cl - cly * cg.ncx;         // which region of the level
  const int cby = rank / cg.cx, cbx = rank - cby * cg.cx;       // this CTA's block of the region
  const int tw = cg.tw, th = cg.th, sw = tw + 2;                // block, and the row stride of its planes
  c.mine = t < tw * th ? 1 : 0;
  const int ly = c.mine ? t / tw : 0, lx = c.mine ? t - ly * tw : 0;
  const int rw = cg.cx * tw, rh = cg.cy * th;                   // the region
  const int rx = cbx * tw + lx, ry = cby * th + ly;             // this thread's cell in it
  const int ox0 = clx * lv.ow, oy0 = lv.y0 + cly * lv.oh;       // output tile of the cluster
  const int ox1 = ox0 + lv.ow < lv.w ? ox0 + lv.ow : lv.w, oy1 = oy0 + lv.oh < lv.y1 ? oy0 + lv.oh : lv.y1;
  const int gx = ox0 - lv.halo + rx, gy = oy0 - lv.halo + ry;
  c.gx = gx; c.gy = gy;
  // offset of block cell (x, y), x in [-1, tw], y in [-1, th] (the ring)
#define FLOW2D_AT(x, y) (((y) + 1) * sw + ((x) + 1))
  // Neighbours.  At the image border the mirrored neighbour (-1 -> 1, n -> n-2) is the opposite one; a cell on the
  // region's own edge has no outer neighbour (it is never exact in pass mode; in whole-level mode that edge is the
  // image's or lies beyond it) and reads the opposite one as well, so that every value read is a computed one.  Across
  // a CTA edge inside the region the neighbour is the ring cell that the CTA next door keeps up to date.
  if (c.mine) {
    c.ac = FLOW2D_AT(lx, ly);
    c.al = FLOW2D_AT((gx == 0 || rx == 0) ? lx + 1 : lx - 1, ly);
    c.ar = FLOW2D_AT((gx == lv.w - 1 || rx == rw - 1) ? lx - 1 : lx + 1, ly);
    c.au = FLOW2D_AT(lx, (gy == 0 || ry == 0) ? ly + 1 : ly - 1);
    c.ad = FLOW2D_AT(lx, (gy == lv.h - 1 || ry == rh - 1) ? ly - 1 : ly + 1);
  } else {
    c.ac = c.al = c.ar = c.au = c.ad = 0;  // parked on a corner of the ring, which nothing reads
  }
  // Where this cell is pushed to: the image of a cell of the leftmost column is in column tw of the CTA to the left, ...
  c.push_h_rank = c.push_v_rank = -1;
  c.push_h = c.push_v = 0;
  if (c.mine) {
    if (lx == 0 && cbx > 0) { c.push_h_rank = rank - 1; c.push_h = FLOW2D_AT(tw, ly); }
    else if (lx == tw - 1 && cbx < cg.cx - 1) { c.push_h_rank = rank + 1; c.push_h = FLOW2D_AT(-1, ly); }
    if (ly == 0 && cby > 0) { c.push_v_rank = rank - cg.cx; c.push_v = FLOW2D_AT(lx, th); }
    else if (ly == th - 1 && cby < cg.cy - 1) { c.push_v_rank = rank + cg.cx; c.push_v = FLOW2D_AT(lx, -1); }
  }
#undef FLOW2D_AT
  c.live = (c.mine && gx >= 0 && gx < lv.w && gy >= 0 && gy < lv.h) ? 1 : 0;
  c.out = (c.mine && gx >= ox0 && gx < ox1 && gy >= oy0 && gy < oy1) ? 1 : 0;
  return c;
}

}  // namespace flow2d

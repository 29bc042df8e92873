// flow2d_api.cu -- handle, level scheduler and the C ABI of include/flow2d.h.
//
// Host-side orchestration of the hot path; follows OpticalFlow2D::ComputeFlow
// (src/optical_flow/optical_flow_2d.cpp:142-569) stage for stage, but everything is enqueued on one
// stream without any host synchronisation inside the pyramid (the reference blocks on
// cuStreamSynchronize after every inner sweep, cuda_operation_solve_2d.cpp:291).
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <initializer_list>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/flow2d.h"
#include "kernels.h"
#include "solve_cluster_geom.h"

using namespace flow2d;

namespace {

constexpr int kPitchAlign = 128;  // floats; = the 512-byte pitch granularity the reference gets from cuMemAllocPitch

enum Container {
  C_IN0, C_IN1,        // uploaded frames
  C_BLUR0, C_BLUR1,    // presmoothed frames
  C_RES0, C_RES1,      // frames resampled to the current level
  C_WARPED,            // frame 1 registered by the current flow
  C_U, C_V, C_U2, C_V2,
  C_DU0, C_DV0, C_DU1, C_DV1,
  C_PHI, C_KSI,
  C_FX, C_FY, C_FT,
  C_TMP0, C_TMP1,      // x-pass results of the resampler
  C_OUT_U, C_OUT_V,    // final flow of flow2d_compute (host API)
  C_IN2, C_IN3,        // landing pair of the uploads: flow2d_compute_async uploads call n+1 while call n computes
  C_J0, C_J1, C_J2, C_J3, C_J4,  // gradient mode only
  C_COUNT
};

}  // namespace

struct flow2d_handle {
  int device = 0;
  size_t W = 0, H = 0, pitch = 0;  // pitch in floats
  int constancy = FLOW2D_GREY;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  // host API: the frames of a call are uploaded on a stream of their own into a landing pair, so the upload of the next
  // call overlaps the computation of the current one (a sequence through one handle, several handles per GPU)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_in_ready[2] = {nullptr, nullptr}, ev_in_free[2] = {nullptr, nullptr};
  unsigned long long async_calls = 0;
  float* pool = nullptr;
  float* c[C_COUNT] = {};
  long long launches = 0;
  long long kind_launches[FLOW2D_KERNEL_KINDS] = {};
  double* d_residuals = nullptr;  // 2 doubles per pyramid level (sums of r_u^2, r_v^2), flow2d_params.report_residuals
  int residual_levels = 0;        // levels of the last compute that recorded a residual
  int residual_px[FLOW2D_MAX_LEVELS] = {};
  int levels_run = 0;
  float device_ms = 0.f;
  unsigned long long* timing = nullptr;  // debug: phase stamps of solve_pass (flow2d_debug_timing)
  // captured level schedules, one per (buffers, parameters) combination, least recently used one evicted:
  // a caller that rotates a few frame / flow containers (a frame ring, several pairs per handle) replays
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    unsigned char key[192] = {};
    long long launches = 0;
    long long kind_launches[FLOW2D_KERNEL_KINDS] = {};
    int levels = 0, residual_levels = 0;
    int residual_px[FLOW2D_MAX_LEVELS] = {};
    int iter_levels = 0, iter_default = 0;
    unsigned long long last_use = 0;
  };
  static constexpr int kGraphSlots = 8;
  GraphEntry graphs[kGraphSlots];
  unsigned long long graph_clock = 0;
  long long graph_captures = 0, graph_replays = 0;
  int sm_count = 148;
  bool pass3 = false;             // the TMA-staged persistent pass (solve_pass3.cu) is usable on this device / driver
  // thread-block clusters for mid-size levels (solve_cluster.cu).  cluster_active[i][j]: clusters of 2^(i+1) CTAs x
  // (256 << j) threads the device holds at once (0 = not launchable here)
  int cluster_whole = -1;         // whole-level mode: FLOW2D_CLUSTER = 0 off / 1 levels whose blocks fit 256 threads per CTA
                                  // (<= 4096 px) / 2 every level that fits a cluster (<= 16384 px); unset = -1: by the number
                                  // of handles alive on the device (cluster_mode below)
  bool counted = false;           // this handle is included in g_live_handles
  int cluster_pass = 0;           // pass mode: FLOW2D_CLUSTER_PASS = 0 / 1 (2-4: forced, see run_solve; unset: kClusterPassDefault)
  bool cluster_compact = false;   // FLOW2D_CLUSTER_COMPACT: fewest CTAs instead of shortest sweeps
  int cluster_active[4][3] = {};
  // opt-in extensions (flow2d_params.scheme / omega / data_term / residual_tolerance / cascaded_restriction); allocated on first use
  float* ext_pool = nullptr;      // six tensor planes of the extension data terms
  float* J6[6] = {};
  float* pyr_pool = nullptr;      // cascaded restriction: levels 1.. of both frame pyramids, row after row at the handle's pitch
  size_t pyr_rows = 0;            // rows per frame the pool holds
  // Both frame pyramids are flow-independent: they are restricted on a stream of their own, ahead of the level loop
  // (which waits, level by level, for an event), into pyr_pool; the coarse levels' long, latency-bound restriction
  // kernels then overlap the coarse levels' solves instead of sitting between them
  cudaStream_t side_stream = nullptr;
  std::vector<cudaEvent_t> pyr_events;  // [0]: fork, [l]: level l of both pyramids is ready
  int* d_stop = nullptr;          // FLOW2D_MAX_LEVELS stop words, then FLOW2D_MAX_LEVELS iteration counts
  double* d_partials = nullptr;   // per-CTA partial sums of the convergence test
  unsigned* d_counter = nullptr;
  int iter_levels = 0, iter_default = 0;  // levels of the last compute / their outer iteration count when not stopped early
  // per-level timers (flow2d_params.report_level_times): events at the start of a level, before and after its solve
  std::vector<cudaEvent_t> level_events;
  int timed_levels = 0;
  // row-slab decomposition (flow2d_slab_connect): this handle is rank `slab_rank` of `slab_world`
  int slab_rank = 0, slab_world = 1;
  size_t slab_min_rows = 64;
  unsigned char* mailbox = nullptr;          // own mailbox (its own cudaMalloc: CUDA IPC exports whole allocations)
  size_t mailbox_bytes = 0, mailbox_rows = 0;
  unsigned char* peer[2] = {nullptr, nullptr};  // mapped mailboxes of rank-1 (above) and rank+1 (below)
  void* imported[2] = {nullptr, nullptr};    // cudaIpcOpenMemHandle mappings to close at destroy
  unsigned long long slab_epoch = 0;         // exchanges so far (same sequence on every rank)
  unsigned* slab_counter = nullptr;
  long long slab_exchanges = 0, slab_bytes_sent = 0;
  int slab_levels = 0;
  std::string err;
};

namespace {

int fail(flow2d_handle* h, int code, const char* fmt, ...) {
  if (h) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    h->err = buf;
  }
  return code;
}

#define CU_TRY(h, call)                                                                        \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail((h), FLOW2D_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                         \
  } while (0)

const char* const kKindNames[FLOW2D_KERNEL_KINDS] = {"blur", "resample", "warp", "derivatives", "grad_tensor", "solve_pass",
                                                      "solve_pass(resident)", "solve_small_pass", "solve_tiny", "add_median",
                                                      "add", "residual", "solve_ext", "solve_cluster"};

int check_launch(flow2d_handle* h, int kind, int kernels) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, FLOW2D_ERR_CUDA, "launch of %s failed: %s", kKindNames[kind], cudaGetErrorString(e));
  h->launches += kernels;
  h->kind_launches[kind] += kernels;
  return FLOW2D_OK;
}
void reset_launch_counts(flow2d_handle* h) {
  h->launches = 0;
  for (auto& k : h->kind_launches) k = 0;
}

// NVTX ranges around what the host enqueues per stage and level (the reference has two event timers and
// nothing else).  Header-only NVTX3: a no-op costing a few nanoseconds unless a profiler is attached.  Kernels of a
// replayed CUDA graph carry no host range; run with FLOW2D_NO_GRAPH=1 under nsys / ncu --nvtx to see them per level.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

#define TRY(expr)              \
  do {                         \
    int rc_ = (expr);          \
    if (rc_ != FLOW2D_OK) return rc_; \
  } while (0)

// cuda_operation_convolution_2d.cpp:83-112
int gauss_taps(flow2d_handle* h, float sigma, GaussTaps* t) {
  const float pixel_size = 1.0f;
  const size_t precision = 3;
  const size_t radius = (size_t)((float)precision * sigma / pixel_size);
  if (radius > (size_t)kMaxGaussRadius)
    return fail(h, FLOW2D_ERR_UNSUPPORTED, "gaussian_sigma %g needs a kernel radius of %zu > %d", (double)sigma, radius,
                kMaxGaussRadius);
  const int r = (int)radius;
  t->radius = r;
  for (int i = -r; i <= r; i++) {
    const float arg = -((float)(i * i) * pixel_size * pixel_size);
    t->c[i + r] = (float)(1.0 / ((double)sigma * std::sqrt(2.0 * 3.1415926)) *
                          std::exp((double)arg / (2.0 * (double)sigma * (double)sigma)));
  }
  float sum = 0.0f;
  for (int i = 0; i < 2 * r + 1; i++) sum = sum + t->c[i];
  for (int i = 0; i < 2 * r + 1; i++) t->c[i] = t->c[i] / sum;
  return FLOW2D_OK;
}

// cuda_operation_median_2d.cpp:100-111 -> 1, 3, 5, 7 or an error
int normalise_median(flow2d_handle* h, size_t radius, int* out) {
  const size_t asked = radius;
  if (radius == 1) {  // copy; tested before the even-value rule, like upstream
    *out = 1;
    return FLOW2D_OK;
  }
  if (radius % 2 == 0 && radius > 0) radius -= 1;
  if (radius == 3 || radius == 5 || radius == 7) {
    *out = (int)radius;
    return FLOW2D_OK;
  }
  radius = asked;
  return fail(h, FLOW2D_ERR_UNSUPPORTED, "median_radius %zu is not supported (1 = off, 3, 5, 7; even values are decremented)",
              radius);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Every kernel addresses containers with float4 loads / stores: a caller pointer that is only 4-byte aligned (a
// sub-view of a larger buffer) must be refused here, not fault on the device.
int check_aligned(flow2d_handle* h, const char* stage, std::initializer_list<const void*> ptrs) {
  for (const void* q : ptrs)
    if (!aligned16(q)) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "%s: device containers must be 16-byte aligned", stage);
  return FLOW2D_OK;
}

int check_level(flow2d_handle* h, size_t w, size_t hh) {
  if (w < 2 || hh < 2 || w > h->W || hh > h->H)
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "level size %zux%zu outside [2x2, %zux%zu]", w, hh, h->W, h->H);
  return FLOW2D_OK;
}

LevelGeom geom(const flow2d_handle* h, size_t w, size_t hh, float hx, float hy) {
  LevelGeom g;
  g.w = (int)w; g.h = (int)hh; g.pitch = (int)h->pitch; g.hx = hx; g.hy = hy;
  return g;
}

// ---- row-slab decomposition ---------------------------------------------------------------------------------------
// Rank r owns rows [Y0, Y1) of a slabbed level.  Everything a level computes is computed on the own rows plus the
// margins the next stage needs; only rows of neighbours ever move:
//   solve      du, dv exact on [Y0 - ghost, Y1 + ghost) after an exchange; a pass loses S+1 rows at each cut edge, so the
//              ghost rows are refreshed from the two neighbours every `outers_per_exchange` outer iterations
//   inputs     u, v, fx, fy, ft on own +- (ghost + S + 1) rows (what the passes read), warped frame one row more
//   median     u + du on own +- median/2 rows: the solve ends with that much validity to spare (no extra exchange)
//   next level its prolongation reads own +- ~30 rows of this level's flow: one exchange of u, v rows after the median
// Both frames (and their restrictions) are complete on every rank: the warp gathers rows it cannot know in advance.
constexpr int kSlabHeaderBytes = 256;
constexpr size_t kSlabRowsCap = 128;  // rows per message the mailbox can hold

struct SolvePlan {
  int S = 1, npass = 1;
  bool pad2 = false;                // solve_pass3: x halo rounded up to 2 columns (68-wide staging boxes) instead of to 4
  bool slabbed = false;
  int Y0 = 0, Y1 = 0;               // own rows
  int ghost = 0;                    // rows of du/dv received from each neighbour per exchange
  int outers_per_exchange = 4;
  int keep = 0;                     // rows beyond the own rows on which the increment must still be exact at the end
  int in_margin = 0;                // u, v, fx, fy, ft are needed on own +- in_margin rows
};

void slab_own_rows(int h, int rank, int world, int* y0, int* y1) {
  const int base = h / world, extra = h % world;
  *y0 = rank * base + (rank < extra ? rank : extra);
  *y1 = *y0 + base + (rank < extra ? 1 : 0);
}

// Which generation of the tiled pass a handle uses (A/B switches: FLOW2D_SOLVE_V1, FLOW2D_SOLVE_V2), and the x halo of its
// region: S+1 columns rounded up to the 4 of an aligned float4 strip, or (solve_pass3: TMA boxes start anywhere) to 2.
int tiled_generation(const flow2d_handle* h) {
  static const bool v1 = std::getenv("FLOW2D_SOLVE_V1") != nullptr, v2 = std::getenv("FLOW2D_SOLVE_V2") != nullptr;
  if (v1) return 1;
  return (h->constancy == FLOW2D_GRADIENT || v2 || !h->pass3) ? 2 : 3;
}
int tiled_halo_x(bool pad2, int sweeps) { return pad2 ? ((sweeps + 2) & ~1) : ((sweeps + 1 <= 4) ? 4 : 8); }

// Time model of one tile of a tiled pass in us, fitted on B200 (profiles/r02/pass3_ab): setup + hand-over of a pass that
// computes phi / ksi (`first`) or reloads them, the 64-bit staging reads of a region that starts off a 16-byte boundary,
// and 0.65 us per sweep.  solve_pass2: 4.9 / 3.7.
double tile_us(int gen, bool first, int halo_x, int sweeps) {
  if (gen != 3) return (first ? 4.9 : 3.7) + 0.65 * sweeps;
  return (first ? 4.6 : 3.2) + ((halo_x & 2) ? (first ? 0.9 : 1.1) : 0.0) + 0.65 * sweeps;
}

SolvePlan plan_solve(const flow2d_handle* h, const LevelGeom& g, const flow2d_params* p, int median, bool allow_slab) {
  SolvePlan pl;
  const int inner = (int)p->inner_iterations_count;
  int S = p->sweeps_per_pass;
  pl.pad2 = tiled_generation(h) == 3 && std::getenv("FLOW2D_P3_HALO4") == nullptr;
  if (S <= 0) {
    // Pick the sweeps per pass that minimise the modelled time of one outer iteration (region 64 x 48, halo S+1 rows and
    // S+1 columns rounded up to 4, or -- solve_pass3 -- to 2 with the unaligned staging boxes; 0.7 us per launch).
    // Two models, A/B-measured on one box (profiles/r02/pass3_ab/geometry_ab.txt, C4 = 1024^2 pairs):
    //   default      fractional waves, 2-column rounding wherever it differs: best for several pairs sharing the GPU, where
    //                other streams fill the SMs a level leaves idle (batch 111.3 Mpix/s, one pair alone 15.46 ms)
    //   FLOW2D_P3_MODEL=1  whole rounds of one tile per SM, rounding chosen per level: best for one pair at a time
    //                (15.24 ms; batch 109.1 Mpix/s)
    //   FLOW2D_P3_HALO4=1  never the 2-column rounding (batch 108.7 Mpix/s, 15.66 ms)
    static const bool halo4 = std::getenv("FLOW2D_P3_HALO4") != nullptr;
    static const bool rounds_model = std::getenv("FLOW2D_P3_MODEL") != nullptr;
    const int gen = tiled_generation(h);
    double best = 1e300;
    if (rounds_model) {
      for (int s_ = 1; s_ <= FLOW2D_MAX_SWEEPS_PER_PASS; ++s_) {
        for (int pad = 0; pad < ((gen == 3 && !halo4) ? 2 : 1); ++pad) {
          const int np = (inner + s_ - 1) / s_;
          const int hx = tiled_halo_x(pad != 0, s_);
          if (pad && !(hx & 2)) continue;  // same geometry as pad = 0
          const int ow = kSolveLW - 2 * hx, oh = kSolveLH - 2 * (s_ + 1);
          const long long tiles = (long long)((g.w + ow - 1) / ow) * ((g.h + oh - 1) / oh);
          const double rounds = (double)((tiles + h->sm_count - 1) / h->sm_count);
          const double cost = rounds * (tile_us(gen, true, hx, s_) + (np - 1) * tile_us(gen, false, hx, s_)) + 0.7 * np;
          if (cost < best) { best = cost; S = s_; pl.pad2 = pad != 0; }
        }
      }
    } else {
      pl.pad2 = gen == 3 && !halo4;
      for (int s_ = 1; s_ <= FLOW2D_MAX_SWEEPS_PER_PASS; ++s_) {
        const int np = (inner + s_ - 1) / s_;
        const int hx = tiled_halo_x(pl.pad2, s_);
        const int ow = kSolveLW - 2 * hx, oh = kSolveLH - 2 * (s_ + 1);
        double waves = (double)((g.w + ow - 1) / ow) * ((g.h + oh - 1) / oh) / (double)h->sm_count;
        if (waves < 1.0) waves = 1.0;
        // (a region that starts off a 16-byte boundary pays for its 64-bit staging reads: + 0.9 / 1.1 us per pass)
        const double first = 4.9 + ((hx & 2) ? 0.9 : 0.0), later = 3.7 + ((hx & 2) ? 1.1 : 0.0);
        const double cost = waves * (first + (np - 1) * later + 0.65 * inner);
        if (cost < best) { best = cost; S = s_; }
      }
    }
  }
  if (S > FLOW2D_MAX_SWEEPS_PER_PASS) S = FLOW2D_MAX_SWEEPS_PER_PASS;
  if (S > inner && inner > 0) S = inner;
  pl.S = S;
  pl.npass = inner > 0 ? (inner + S - 1) / S : 1;
  pl.Y0 = 0; pl.Y1 = g.h;
  if (allow_slab && h->slab_world > 1 && inner > 0 && p->outer_iterations_count > 0) {
    // rows lost per outer iteration: every pass s_q + 1 (the sweeps are spread evenly over the passes)
    int per_outer = 0, left = inner;
    for (int q = 0; q < pl.npass; ++q) {
      const int s_ = (left + (pl.npass - q) - 1) / (pl.npass - q);
      left -= s_;
      per_outer += s_ + 1;
    }
    pl.keep = median / 2;
    int k = 4;
    while (k > 1 && (size_t)(per_outer * k + pl.keep) > kSlabRowsCap) --k;
    pl.outers_per_exchange = k;
    pl.ghost = per_outer * k + pl.keep;
    const int base = g.h / h->slab_world;
    // the next level's prolongation takes up to ~ (ghost + S + 3) rows of this level from the neighbours
    if ((size_t)base >= h->slab_min_rows && base >= 2 * (pl.ghost + S + 4) && (size_t)pl.ghost <= kSlabRowsCap) {
      pl.slabbed = true;
      slab_own_rows(g.h, h->slab_rank, h->slab_world, &pl.Y0, &pl.Y1);
      pl.in_margin = pl.ghost + S + 1;
    }
  }
  return pl;
}

// One message to / from each neighbour: rows [up0, up0+upn) of (fa, fb) go to the rank above, [dn0, dn0+dnn) to the rank
// below; rows [rup0, +rupn) arrive from above and [rdn0, +rdnn) from below (same row indices on both sides).
int slab_exchange(flow2d_handle* h, float* fa, float* fb, const LevelGeom& g, int up0, int upn, int dn0, int dnn, int rup0,
                  int rupn, int rdn0, int rdnn) {
  if (h->slab_rank == 0) upn = rupn = 0;
  if (h->slab_rank == h->slab_world - 1) dnn = rdnn = 0;
  if ((size_t)upn > h->mailbox_rows || (size_t)dnn > h->mailbox_rows || (size_t)rupn > h->mailbox_rows || (size_t)rdnn > h->mailbox_rows)
    return fail(h, FLOW2D_ERR_UNSUPPORTED, "slab exchange of %d rows exceeds the mailbox (%zu rows)", upn > dnn ? upn : dnn, h->mailbox_rows);
  if ((upn > 0 && !h->peer[0]) || (dnn > 0 && !h->peer[1])) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "slab: neighbour mailbox not connected");
  const unsigned long long e = ++h->slab_epoch;
  const size_t plane = h->mailbox_rows * h->pitch;  // floats per receive buffer
  auto recv_buf = [&](unsigned char* box, int from, int field) {
    return reinterpret_cast<float*>(box + kSlabHeaderBytes) + ((size_t)((from * 2 + (int)(e & 1)) * 2 + field)) * plane;
  };
  auto flag_of = [&](unsigned char* box, int from) { return reinterpret_cast<unsigned long long*>(box) + from; };
  SlabPush ps;
  std::memset(&ps, 0, sizeof ps);
  ps.field[0] = fa; ps.field[1] = fb;
  ps.counter = h->slab_counter;
  ps.w = g.w; ps.pitch = g.pitch;
  if (upn > 0) {  // to the rank above: it receives "from below" (index 1)
    ps.dir[0].dst[0] = recv_buf(h->peer[0], 1, 0); ps.dir[0].dst[1] = recv_buf(h->peer[0], 1, 1);
    ps.dir[0].flag = flag_of(h->peer[0], 1); ps.dir[0].epoch = e; ps.dir[0].row0 = up0; ps.dir[0].rows = upn;
  }
  if (dnn > 0) {  // to the rank below: it receives "from above" (index 0)
    ps.dir[1].dst[0] = recv_buf(h->peer[1], 0, 0); ps.dir[1].dst[1] = recv_buf(h->peer[1], 0, 1);
    ps.dir[1].flag = flag_of(h->peer[1], 0); ps.dir[1].epoch = e; ps.dir[1].row0 = dn0; ps.dir[1].rows = dnn;
  }
  launch_slab_push(h->stream, ps);
  SlabUnpack up;
  std::memset(&up, 0, sizeof up);
  up.field[0] = fa; up.field[1] = fb;
  up.error = h->slab_counter + 1;
  up.w = g.w; up.pitch = g.pitch;
  if (rupn > 0) {
    up.dir[0].src[0] = recv_buf(h->mailbox, 0, 0); up.dir[0].src[1] = recv_buf(h->mailbox, 0, 1);
    up.dir[0].flag = flag_of(h->mailbox, 0); up.dir[0].epoch = e; up.dir[0].row0 = rup0; up.dir[0].rows = rupn;
  }
  if (rdnn > 0) {
    up.dir[1].src[0] = recv_buf(h->mailbox, 1, 0); up.dir[1].src[1] = recv_buf(h->mailbox, 1, 1);
    up.dir[1].flag = flag_of(h->mailbox, 1); up.dir[1].epoch = e; up.dir[1].row0 = rdn0; up.dir[1].rows = rdnn;
  }
  launch_slab_unpack(h->stream, up);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(h, FLOW2D_ERR_CUDA, "slab exchange launch failed: %s", cudaGetErrorString(err));
  ++h->slab_exchanges;
  h->slab_bytes_sent += (long long)(upn + dnn) * g.w * 2 * 4;
  return FLOW2D_OK;
}

// ---- opt-in extensions --------------------------------------------------------------------------------------------
// the relaxation itself differs from the reference's (another ordering, a relaxation factor, a tensor data term):
// such levels are solved by the kernels of solve_ext.cu
bool ext_solver(const flow2d_params* p) {
  return p->scheme != FLOW2D_SCHEME_JACOBI || (p->omega != 0.f && p->omega != 1.f) || p->data_term != FLOW2D_TERM_DEFAULT;
}
bool early_exit(const flow2d_params* p) { return p->residual_tolerance > 0.f; }
bool any_extension(const flow2d_params* p) { return ext_solver(p) || early_exit(p) || p->cascaded_restriction != 0; }

void invalidate_graphs(flow2d_handle* h) {
  for (auto& g : h->graphs)
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
}

// rows of levels 1 .. levels-1 of one frame pyramid
size_t pyramid_rows(const flow2d_handle* h, const flow2d_params* p, int levels, size_t* row_of /* [levels] or null */) {
  size_t rows = 0;
  for (int l = 1; l < levels; l++) {
    size_t cw, ch; float hx, hy;
    flow2d_level_geometry(h->W, h->H, p->warp_scale_factor, l, &cw, &ch, &hx, &hy);
    if (row_of) row_of[l] = rows;
    rows += ch;
  }
  return rows;
}
int levels_of(const flow2d_handle* h, const flow2d_params* p) {
  const size_t max_level = flow2d_max_warp_level(h->W, h->H, p->warp_scale_factor);
  return (int)(p->warp_levels_count < max_level ? p->warp_levels_count : max_level);
}

// Device memory of the extensions; must run OUTSIDE a stream capture (compute_on_device calls it first).
int ensure_ext(flow2d_handle* h, const flow2d_params* p) {
  if (ext_solver(p) && !h->ext_pool) {
    const size_t csize = h->pitch * h->H;
    if (cudaMalloc(&h->ext_pool, csize * 6 * sizeof(float)) != cudaSuccess) {
      (void)cudaGetLastError();
      return fail(h, FLOW2D_ERR_OUT_OF_MEMORY, "allocation of the tensor planes failed");
    }
    CU_TRY(h, cudaMemset(h->ext_pool, 0, csize * 6 * sizeof(float)));
    for (int k = 0; k < 6; k++) h->J6[k] = h->ext_pool + csize * k;
  }
  if (early_exit(p) && !h->d_stop) {
    const size_t ctas = ((h->W + 31) / 32) * ((h->H + 7) / 8);
    if (cudaMalloc(&h->d_stop, sizeof(int) * 2 * FLOW2D_MAX_LEVELS) != cudaSuccess ||
        cudaMalloc(&h->d_partials, sizeof(double) * 2 * ctas) != cudaSuccess ||
        cudaMalloc(&h->d_counter, sizeof(unsigned)) != cudaSuccess) {
      (void)cudaGetLastError();
      return fail(h, FLOW2D_ERR_OUT_OF_MEMORY, "allocation of the convergence-test buffers failed");
    }
    CU_TRY(h, cudaMemset(h->d_stop, 0, sizeof(int) * 2 * FLOW2D_MAX_LEVELS));
    CU_TRY(h, cudaMemset(h->d_counter, 0, sizeof(unsigned)));
  }
  static const bool no_side = std::getenv("FLOW2D_NO_SIDE_PYRAMID") != nullptr;  // A/B switch: restrict inside the level loop
  const bool side = !no_side && !p->cascaded_restriction && levels_of(h, p) > 1;
  if (side) {
    if (!h->side_stream) CU_TRY(h, cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    while ((int)h->pyr_events.size() < levels_of(h, p) + 1) {
      cudaEvent_t e = nullptr;
      CU_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->pyr_events.push_back(e);
    }
  }
  if (p->cascaded_restriction || side) {
    const size_t rows = pyramid_rows(h, p, levels_of(h, p), nullptr);
    if (rows > h->pyr_rows) {
      CU_TRY(h, cudaStreamSynchronize(h->stream));
      invalidate_graphs(h);  // captured schedules point into the old pool
      if (h->pyr_pool) cudaFree(h->pyr_pool);
      h->pyr_pool = nullptr; h->pyr_rows = 0;
      if (cudaMalloc(&h->pyr_pool, rows * h->pitch * 2 * sizeof(float)) != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(h, FLOW2D_ERR_OUT_OF_MEMORY, "allocation of the frame pyramids failed");
      }
      CU_TRY(h, cudaMemset(h->pyr_pool, 0, rows * h->pitch * 2 * sizeof(float)));
      h->pyr_rows = rows;
    }
  }
  return FLOW2D_OK;
}

// The convergence test after outer iteration `outer_done` (1-based) of level slot `slot`: sets the level's stop word
// when both RMS residuals are <= residual_tolerance.  `which`: 1 = the increment is in the result pair, 2 = scratch pair.
int enqueue_decide(flow2d_handle* h, const LevelGeom& g, const float* const* J, bool grad, const float* u, const float* v,
                   const float* du, const float* dv, const float* phi, const float* ksi, const flow2d_params* p, int slot,
                   int which, int outer_done) {
  ResidualDecide d;
  d.partials = h->d_partials; d.counter = h->d_counter; d.sums = nullptr;
  d.stop = h->d_stop + slot; d.iterations = h->d_stop + FLOW2D_MAX_LEVELS + slot;
  d.tol = p->residual_tolerance; d.which = which; d.outer_done = outer_done;
  launch_residual_decide(h->stream, h->c[C_FX], h->c[C_FY], h->c[C_FT], J, grad, u, v, du, dv, phi, ksi, g, p->equation_alpha, d);
  return check_launch(h, FLOW2D_K_RESIDUAL, 1);
}

// The solve of one level with the kernels of solve_ext.cu (see there): tensor planes in
// h->J6, one launch per sweep (red-black: per colour).  Result in du_a/dv_a.
int run_solve_ext(flow2d_handle* h, const LevelGeom& g, const float* u, const float* v, float* du_a, float* dv_a, float* du_b,
                  float* dv_b, float* phi, float* ksi, const flow2d_params* p, int slot) {
  const int outer = (int)p->outer_iterations_count, inner = (int)p->inner_iterations_count;
  cudaStream_t st = h->stream;
  CU_TRY(h, cudaMemset2DAsync(du_a, h->pitch * 4, 0, (size_t)g.w * 4, g.h, st));
  CU_TRY(h, cudaMemset2DAsync(dv_a, h->pitch * 4, 0, (size_t)g.w * 4, g.h, st));
  if (outer == 0 || inner == 0) return FLOW2D_OK;
  const bool rb = p->scheme == FLOW2D_SCHEME_RED_BLACK, early = early_exit(p);
  const float omega = p->omega == 0.f ? 1.f : p->omega;
  const int every = p->residual_check_every > 0 ? p->residual_check_every : 1;
  const int* stop = early ? h->d_stop + slot : nullptr;
  ExtTensor J;
  for (int k = 0; k < 6; k++) J.p[k] = h->J6[k];
  // Jacobi alternates between the two pairs: start where the last sweep ends in du_a/dv_a
  float *cu = du_a, *cv = dv_a, *ou = du_b, *ov = dv_b;
  if (!rb && (((long long)outer * inner) & 1)) {
    CU_TRY(h, cudaMemset2DAsync(du_b, h->pitch * 4, 0, (size_t)g.w * 4, g.h, st));
    CU_TRY(h, cudaMemset2DAsync(dv_b, h->pitch * 4, 0, (size_t)g.w * 4, g.h, st));
    std::swap(cu, ou); std::swap(cv, ov);
  }
  for (int o = 0; o < outer; ++o) {
    launch_ext_phi_ksi(st, J, u, v, cu, cv, phi, ksi, g, p->equation_smoothness, p->equation_data, stop);
    TRY(check_launch(h, FLOW2D_K_EXT, 1));
    for (int j = 0; j < inner; ++j) {
      if (rb) {
        launch_ext_sweep(st, J, u, v, cu, cv, phi, ksi, cu, cv, g, p->equation_alpha, omega, 0, stop);
        launch_ext_sweep(st, J, u, v, cu, cv, phi, ksi, cu, cv, g, p->equation_alpha, omega, 1, stop);
        TRY(check_launch(h, FLOW2D_K_EXT, 2));
      } else {
        launch_ext_sweep(st, J, u, v, cu, cv, phi, ksi, ou, ov, g, p->equation_alpha, omega, -1, stop);
        TRY(check_launch(h, FLOW2D_K_EXT, 1));
        std::swap(cu, ou); std::swap(cv, ov);
      }
    }
    if (early && (o + 1) % every == 0)
      TRY(enqueue_decide(h, g, h->J6, true, u, v, cu, cv, phi, ksi, p, slot, cu == du_a ? 1 : 2, o + 1));
  }
  if (early) {
    launch_ext_pick(st, stop, du_b, dv_b, du_a, dv_a, g);
    TRY(check_launch(h, FLOW2D_K_EXT, 1));
  }
  return FLOW2D_OK;
}

// handles alive per device (this process): several handles on one GPU = several frame pairs in flight at once
constexpr int kMaxCountedDevices = 64;
std::atomic<int> g_live_handles[kMaxCountedDevices];
int live_handles(int device) { return device >= 0 && device < kMaxCountedDevices ? g_live_handles[device].load() : 1; }

// ---- thread-block clusters for mid-size levels (solve_cluster.cu) ---------------------------------------------------
// Defaults of the two modes; FLOW2D_CLUSTER / FLOW2D_CLUSTER_PASS override them per handle (A/B measurements, tests).
// Measured on B200 (profiles/r02/cluster_ab/, same box; one level of the 1024^2 pyramid = 40 x 5 iterations):
//   blocks of <= 256 threads (levels of 1 025 .. 4 096 px)  138-154 us against 142-151 us with one solve_small_pass launch per
//       outer iteration; 512 threads (.. 8 192 px) 188-193 us against 146 us; 1024 threads (.. 16 384 px) 252-262 us against
//       151-181 us: the barrier.cluster (MEMBAR.ALL.GPU + UCGABAR + CCTL.IVALL, ~0.31 us with 16 CTAs; ncu: 60 % of the
//       kernel's stall samples are the membar) costs per sweep what the relaunch costs per outer iteration
//   whole flows, every level that fits on a cluster (FLOW2D_CLUSTER=2) against none (=0):
//       16 pairs on 8 handles   C4 116.2 -> 117.5 Mpix/s device, 114.9 -> 116.8 end to end; C1b 84.1 -> 86.5, 81.7 -> 85.4
//       one pair at a time      C4 14.98 -> 15.54 ms, C1b 7.14 -> 7.87 ms
//       (only the levels up to 4 096 px, =1: batch within 1 % either way, one pair 1-3 % slower)
//     i.e. one cluster of <= 16 CTAs for all 40 outer iterations takes far less SM time than 40 launches of 16-81 CTAs with
//     their redundant rings -- which is what counts when other pairs' kernels wait for SMs -- but not less time
//   pass mode: 1.5-3.3x slower than solve_small_pass / the tiled pass on every level -> off
//   => default (FLOW2D_CLUSTER unset): every level that fits a cluster when at least kClusterAutoHandles handles are alive
//      on the device (frame pairs in flight at once: what counts is SM time), none for a lone handle (what counts is latency)
constexpr int kClusterWholeDefault = -1;
constexpr int kClusterPassDefault = 0;
constexpr int kClusterAutoHandles = 4;
int cluster_mode(const flow2d_handle* h) {
  if (h->cluster_whole >= 0) return h->cluster_whole;
  return live_handles(h->device) >= kClusterAutoHandles ? 2 : 0;
}

// Time of one barrier-to-barrier phase of the cluster kernel in us (a sweep, or one of the three set-up phases of an
// outer iteration): the cluster barrier + the issue time of the CTA's warps.  Fitted to the measurements above
// (8 / 16 / 32 warps per CTA: 0.43 / 0.59 / 0.80 us).
double cluster_phase_us(int threads) { return 0.31 + 0.0155 * ((threads + 31) / 32); }

struct ClusterPlan {
  ClusterGeom cg;
  int threads = 0;   // threads per CTA (whole warps covering the block); 0 = no cluster shape fits
  int active = 0;    // clusters of that shape the device holds at once
  double phase_us = 0.0;
};

// The cluster shape for a region of rw x rh cells: cx * cy in {2, 4, 8, 16} CTAs with blocks of ceil(rw / cx) x
// ceil(rh / cy) >= 8 x 8 cells, at most 1024 per CTA.  Fewer threads per CTA make shorter sweeps (the CTA's warps issue
// one after the other), more CTAs occupy more SMs: `compact` weighs the second, the default the first.
ClusterPlan plan_cluster(const int (*cluster_active)[3], int rw, int rh, bool compact) {
  ClusterPlan best;
  double best_cost = 1e300;
  for (int i = 0; i < 4; i++) {
    const int csize = 2 << i;
    for (int cx = 1; cx <= csize; cx *= 2) {
      const int cy = csize / cx;
      int tw = (rw + cx - 1) / cx, th = (rh + cy - 1) / cy;
      if (tw < kClusterMinBlock) tw = kClusterMinBlock;
      if (th < kClusterMinBlock) th = kClusterMinBlock;
      const int cells = tw * th;
      if (cells > 1024) continue;
      const int j = cells <= 256 ? 0 : cells <= 512 ? 1 : 2;
      const int active = cluster_active[i][j];
      if (active <= 0) continue;
      const int threads = (cells + 31) / 32 * 32;  // whole warps; the kernel's shared planes are sized by the bucket j
      const double phase = cluster_phase_us(threads);
      // tie-breakers: little padding (cells beyond the region), squarish blocks (fewer edge cells to push)
      const double cost = phase * (1.0 + (compact ? 0.3 : 0.02) * csize) + 1e-6 * ((double)cells * csize - (double)rw * rh) +
                          1e-5 * (tw + th);
      if (cost < best_cost) {
        best_cost = cost;
        best.cg.cx = cx; best.cg.cy = cy; best.cg.tw = tw; best.cg.th = th; best.cg.ncx = 1;
        best.threads = threads;
        best.active = active;
        best.phase_us = phase;
      }
    }
  }
  return best;
}

// CudaOperationSolve2D::Execute (cuda_operation_solve_2d.cpp:229-299) on top of the solve kernels.
// The result is left in du_a/dv_a; du_b/dv_b are scratch.  fx,fy,ft (and J in gradient mode) must
// hold the derivative planes of this level (on the own rows +- plan.in_margin when the level is slabbed).
int run_solve(flow2d_handle* h, const LevelGeom& g, const float* u, const float* v, float* du_a, float* dv_a,
              float* du_b, float* dv_b, float* phi, float* ksi, bool want_phi, const flow2d_params* p,
              const SolvePlan& pl, int slot = 0) {
  const int outer = (int)p->outer_iterations_count, inner = (int)p->inner_iterations_count;
  // convergence test (flow2d_params.residual_tolerance): every pass reads the level's stop word first; the levels that
  // would run inside one CTA (solve_tiny, resident) go through the pass kernels so that the test sees every iteration
  const bool early = early_exit(p) && !pl.slabbed;
  const int every = p->residual_check_every > 0 ? p->residual_check_every : 1;
  if (outer == 0 || inner == 0) {
    // no sweep runs: the increment stays at its initial zero (cuda_operation_solve_2d.cpp:229-232)
    CU_TRY(h, cudaMemset2DAsync(du_a, h->pitch * 4, 0, (size_t)g.w * 4, g.h, h->stream));
    CU_TRY(h, cudaMemset2DAsync(dv_a, h->pitch * 4, 0, (size_t)g.w * 4, g.h, h->stream));
    return FLOW2D_OK;
  }
  const bool grad = h->constancy == FLOW2D_GRADIENT;
  SolveArgs a;
  std::memset(&a, 0, sizeof a);
  a.fx = h->c[C_FX]; a.fy = h->c[C_FY]; a.ft = h->c[C_FT];
  for (int i = 0; i < 5; i++) a.J[i] = h->c[C_J0 + i];
  a.u = u; a.v = v;
  a.w = g.w; a.h = g.h; a.pitch = g.pitch;
  a.hx = g.hx; a.hy = g.hy;
  a.alpha = p->equation_alpha; a.e_smooth = p->equation_smoothness; a.e_data = p->equation_data;
  a.hx_2 = a.alpha / (a.hx * a.hx);
  a.hy_2 = a.alpha / (a.hy * a.hy);
  a.timing = h->timing;
  a.y0 = 0; a.y1 = g.h;
  // e_smooth = 0 or e_data = 0 (legal) makes the argument of sqrt exactly zero on every flat cell, outside the range
  // the branch-free sqrt / rcp of the one-pixel kernels cover: take their plain IEEE variant from the start instead of
  // computing every outer iteration twice
  a.exact = (a.e_smooth * a.e_smooth < 0x1p-100f || a.e_data * a.e_data < 0x1p-100f) ? 1 : 0;
  a.stop = early ? h->d_stop + slot : nullptr;
  const bool slabbed = pl.slabbed;

  // tiny levels (<= 1024 pixels): one CTA, one thread per pixel, all outer iterations in the kernel
  if (!slabbed && !early && p->resident_levels >= 0 && p->resident_levels != 2 && solve_tiny_fits(g.w, g.h)) {
    a.du_in = a.dv_in = nullptr;
    a.phi_in = a.ksi_in = nullptr;
    a.du_out = du_a; a.dv_out = dv_a;
    a.phi_out = want_phi ? phi : nullptr; a.ksi_out = want_phi ? ksi : nullptr;
    a.sweeps = inner; a.outer = outer;
    launch_solve_tiny(h->stream, a, grad);
    return check_launch(h, FLOW2D_K_SOLVE_TINY, 1);
  }
  // mid-size levels (up to 16 x 1024 pixels): one thread-block cluster, one thread per pixel, halos through distributed
  // shared memory, all outer iterations in the kernel
  const int cmode = cluster_mode(h);
  if (cmode && !slabbed && !early && p->resident_levels == 0 && g.w >= 2 && g.h >= 2 &&
      (long long)g.w * g.h <= (long long)kClusterMaxCtas * 1024) {
    const ClusterPlan cp = plan_cluster(h->cluster_active, g.w, g.h, h->cluster_compact || p->throughput_mode != 0);
    if (cp.threads && (cp.threads <= 256 || cmode >= 2)) {
      a.du_in = a.dv_in = nullptr;
      a.phi_in = a.ksi_in = nullptr;
      a.du_out = du_a; a.dv_out = dv_a;
      a.phi_out = want_phi ? phi : nullptr; a.ksi_out = want_phi ? ksi : nullptr;
      a.sweeps = inner; a.outer = outer;
      a.ow = cp.cg.cx * cp.cg.tw; a.oh = cp.cg.cy * cp.cg.th; a.halo_x = a.halo_y = 0;
      a.pdl = 0;
      launch_solve_cluster(h->stream, a, grad, cp.cg, cp.threads, 1);
      return check_launch(h, FLOW2D_K_SOLVE_CLUSTER, 1);
    }
  }
  // resident mode: the whole level (plus a one-cell apron) fits one CTA's region
  const bool fits = g.w + 4 + 1 <= kSolveLW && g.h + 1 + 1 <= kSolveLH;
  // ... unless one frame pair at a time is being solved and an outer iteration is a single one-thread-per-
  // pixel pass: a handful of 32x32 CTAs, one launch per outer iteration, then beat one CTA that carries four
  // pixels per thread through all of them (measured on the rub pair: ~240 us against 420-520 us per level).
  // With several handles sharing the GPU the resident CTA wins again: it occupies one SM instead of nine.
  const bool small_pass_instead = p->resident_levels == 0 && !p->throughput_mode && outer > 1 &&
                                  (p->sweeps_per_pass == 0 || p->sweeps_per_pass >= inner) &&
                                  inner <= FLOW2D_MAX_SWEEPS_PER_PASS && kSmallTS - 2 * (inner + 1) >= 4;
  if (!slabbed && !early && fits && p->resident_levels >= 0 && !small_pass_instead) {
    a.du_in = a.dv_in = nullptr;
    a.phi_in = a.ksi_in = nullptr;
    a.du_out = du_a; a.dv_out = dv_a;
    a.phi_out = want_phi ? phi : nullptr; a.ksi_out = want_phi ? ksi : nullptr;
    a.sweeps = inner; a.outer = outer;
    a.ow = kSolveLW; a.oh = kSolveLH; a.halo_x = 4; a.halo_y = 1;
    launch_solve_pass(h->stream, a, grad, 1, 1, g.h + 2);  // image rows + one apron row above and below
    return check_launch(h, FLOW2D_K_SOLVE_RESIDENT, 1);
  }

  const int npass = pl.npass;
  const long long total = (long long)outer * npass;
  float* bufs[2][2] = {{du_a, dv_a}, {du_b, dv_b}};
  long long pass = 0;
  static const bool no_pdl = std::getenv("FLOW2D_NO_PDL") != nullptr;  // A/B switch for measurements
  float *cur_du = nullptr, *cur_dv = nullptr;
  a.outer = 1;

  // Row-slab decomposition: this rank produces rows [Y0, Y1) of the level (+- keep).  A pass can only be exact where
  // its input increment was exact S+1 rows further out, so the rows a rank works on shrink by S+1 per pass from both
  // cut edges (never at the true image border) until the ghost rows are refreshed from the neighbours.  Exchanges
  // happen between outer iterations only (phi / ksi of a multi-pass outer iteration are not exchanged).
  const int Y0 = pl.Y0, Y1 = pl.Y1, ghost = pl.ghost;
  // rows that must stay exact to the end: the own rows and what the median reads beyond them
  const int K0 = slabbed ? (Y0 - pl.keep > 0 ? Y0 - pl.keep : 0) : 0, K1 = slabbed ? (Y1 + pl.keep < g.h ? Y1 + pl.keep : g.h) : g.h;
  int va = 0, vb = g.h;  // rows on which the current increment is exact on this rank
  if (slabbed) {
    va = Y0 - ghost > 0 ? Y0 - ghost : 0;
    vb = Y1 + ghost < g.h ? Y1 + ghost : g.h;
  }
  bool after_exchange = false;
  for (int o = 0; o < outer; ++o) {
    if (slabbed && o > 0) {
      // would this outer iteration still cover the rows that have to stay exact?
      int sa = va, sb = vb, left = inner;
      for (int q = 0; q < npass; ++q) {
        const int s = (left + (npass - q) - 1) / (npass - q);
        left -= s;
        if (sa > 0) sa += s + 1;
        if (sb < g.h) sb -= s + 1;
      }
      if (sa > K0 || sb < K1) {
        // ghost rows of the increment: [Y0, Y0+ghost) to the rank above, [Y1-ghost, Y1) to the rank below
        TRY(slab_exchange(h, cur_du, cur_dv, g, Y0, ghost, Y1 - ghost, ghost, Y0 - ghost, ghost, Y1, ghost));
        va = Y0 - ghost > 0 ? Y0 - ghost : 0;
        vb = Y1 + ghost < g.h ? Y1 + ghost : g.h;
        after_exchange = true;
      }
    }
    int left = inner;
    for (int q = 0; q < npass; ++q, ++pass) {
      const int s = (left + (npass - q) - 1) / (npass - q);  // spread the sweeps evenly over the passes
      left -= s;
      const int dst = (int)((total - 1 - pass) & 1);  // the last pass writes buffer 0 = du_a/dv_a
      a.du_in = cur_du; a.dv_in = cur_dv;
      a.du_out = bufs[dst][0]; a.dv_out = bufs[dst][1];
      const bool first = (q == 0);
      a.phi_in = first ? nullptr : phi; a.ksi_in = first ? nullptr : ksi;
      const bool check = early && (o + 1) % every == 0;  // the convergence test needs this iteration's phi, ksi
      const bool store_phi = first && (npass > 1 || (want_phi && o == outer - 1) || check);
      a.phi_out = store_phi ? phi : nullptr; a.ksi_out = store_phi ? ksi : nullptr;
      a.sweeps = s;
      a.halo_y = s + 1;
      a.halo_x = tiled_halo_x(pl.pad2, s);
      a.ow = kSolveLW - 2 * a.halo_x;
      a.oh = kSolveLH - 2 * a.halo_y;
      if (slabbed && pass > 0) {  // the very first pass starts from du = dv = 0, exact everywhere
        if (va > 0) va += s + 1;
        if (vb < g.h) vb -= s + 1;
      }
      a.y0 = va; a.y1 = vb;
      a.pdl = (pass > 0 && !no_pdl && !after_exchange) ? 1 : 0;  // the first pass follows the derivatives kernel
      after_exchange = false;
      // Mid-size levels cannot fill the GPU with 64x48 regions; there the latency of one CTA is what
      // counts and the one-thread-per-pixel pass is faster as long as its many more CTAs still fit a few
      // waves.  Time model fitted to tools/level_timing.py on B200 (us per outer iteration, launch gap included):
      //   tiled 64x48 tiles   0.7 + (4.9 + 0.65 * sweeps) * max(1, tiles / 148)   (solve_pass2; 11.1 per wave with the first generation)
      //   ts x ts regions     1.0 + ceil(regions / 148) * (2.5 + 1.6 * ts^2 / 1024)
      bool small = false;
      if (npass == 1 && p->resident_levels != 2 && p->resident_levels != -1 && !p->throughput_mode) {
        const long long n_big = (long long)((g.w + a.ow - 1) / a.ow) * ((vb - va + a.oh - 1) / a.oh);
        const long long sms = h->sm_count;
        double t_best = 0.7 + (4.9 + 0.65 * s) * (n_big > sms ? (double)n_big / (double)sms : 1.0);
        int ts_best = 0;
        static const int kRegion[3] = {32, 24, 16};
        for (int ts : kRegion) {
          const int so = ts - 2 * (s + 1);
          if (so < 4) continue;
          const long long n = (long long)((g.w + so - 1) / so) * ((vb - va + so - 1) / so);
          const double t = 1.0 + (double)((n + sms - 1) / sms) * (2.5 + 1.6 * (ts * ts) / 1024.0);
          if (t < t_best) { t_best = t; ts_best = ts; }
        }
        // the same pass on a grid of clusters (FLOW2D_CLUSTER_PASS): 128x128 (16 x 1024 threads) or 128x64 (16 x 512)
        // regions with the S+1 halo, 82 % / 74 % of the cells are results instead of 39 % of a 32x32 region
        ClusterPlan cp_best;
        int cl_n = 0;
        if (h->cluster_pass && !slabbed) {
          // FLOW2D_CLUSTER_PASS: 1 = where the model says so, 2 = always (shape by the model), 3 / 4 = always 128x64 / 128x128
          static const int kShape[2][2] = {{128, 64}, {128, 128}};
          double t_cl = 1e300;
          for (int k = 0; k < 2; ++k) {
            if ((h->cluster_pass == 3 && k != 0) || (h->cluster_pass == 4 && k != 1)) continue;
            const int cw = kShape[k][0] - 2 * (s + 1), ch = kShape[k][1] - 2 * (s + 1);
            const ClusterPlan cp = plan_cluster(h->cluster_active, kShape[k][0], kShape[k][1], false);
            if (!cp.threads || cp.cg.cx * cp.cg.tw != kShape[k][0] || cp.cg.cy * cp.cg.th != kShape[k][1]) continue;
            const long long n = (long long)((g.w + cw - 1) / cw) * ((vb - va + ch - 1) / ch);
            const double t = 1.0 + (double)((n + cp.active - 1) / cp.active) * (1.3 + (3 + s) * cp.phase_us);
            if (t < t_cl) { t_cl = t; cp_best = cp; cl_n = (int)n; }
          }
          if (cl_n && h->cluster_pass == 1 && !(t_cl < t_best)) cl_n = 0;
        }
        if (cl_n) {
          small = true;
          a.halo_x = a.halo_y = s + 1;
          a.ow = cp_best.cg.cx * cp_best.cg.tw - 2 * (s + 1);
          a.oh = cp_best.cg.cy * cp_best.cg.th - 2 * (s + 1);
          cp_best.cg.ncx = (g.w + a.ow - 1) / a.ow;
          launch_solve_cluster(h->stream, a, grad, cp_best.cg, cp_best.threads, cl_n);
          TRY(check_launch(h, FLOW2D_K_SOLVE_CLUSTER, 1));
        } else if (ts_best) {
          small = true;
          const int so = ts_best - 2 * (s + 1);
          a.halo_x = a.halo_y = s + 1;
          a.ow = a.oh = so;
          launch_solve_small_pass(h->stream, a, grad, (g.w + so - 1) / so, (vb - va + so - 1) / so, ts_best);
          TRY(check_launch(h, FLOW2D_K_SOLVE_SMALL_PASS, 1));
        }
      }
      if (!small) {
        const int tx = (g.w + a.ow - 1) / a.ow, ty = (vb - va + a.oh - 1) / a.oh;
        const int gen = tiled_generation(h);
        if (gen == 1) launch_solve_pass(h->stream, a, grad, tx, ty);
        else if (gen == 2) launch_solve_pass2(h->stream, a, grad, tx, ty);
        else if (!launch_solve_pass3(h->stream, a, tx, ty, h->sm_count))
          return fail(h, FLOW2D_ERR_CUDA, "solve_pass3: tensor map or launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        TRY(check_launch(h, FLOW2D_K_SOLVE_PASS, 1));
      }
      cur_du = a.du_out; cur_dv = a.dv_out;
    }
    if (early && (o + 1) % every == 0) {
      const float* J[5] = {h->c[C_J0], h->c[C_J1], h->c[C_J2], h->c[C_J3], h->c[C_J4]};
      TRY(enqueue_decide(h, g, J, grad, u, v, cur_du, cur_dv, phi, ksi, p, slot, cur_du == du_a ? 1 : 2, o + 1));
      after_exchange = true;  // the next pass follows the test kernel: no programmatic dependent launch
    }
  }
  if (early) {
    launch_ext_pick(h->stream, h->d_stop + slot, du_b, dv_b, du_a, dv_a, g);
    TRY(check_launch(h, FLOW2D_K_EXT, 1));
  }
  return FLOW2D_OK;
}

// Derivative planes (and, for the extension data terms, the tensor planes) of a level, then its solve.
int run_level_solve(flow2d_handle* h, const LevelGeom& g, const float* f0, const float* f1w, const float* u, const float* v,
                    float* du_a, float* dv_a, float* du_b, float* dv_b, float* phi, float* ksi, bool want_phi,
                    const flow2d_params* p, const SolvePlan& pl, int slot, int D0, int D1);

// y0, y1: rows on which the solve needs fx, fy, ft (y1 <= y0: the whole level).  Gradient constancy takes central
// differences of them (inside the reference's 16x8 tiles), so the derivative planes cover one row more on each side.
int run_derivatives(flow2d_handle* h, const LevelGeom& g, const float* f0, const float* f1w, int y0 = 0, int y1 = 0) {
  const bool grad = h->constancy == FLOW2D_GRADIENT;
  if (y1 <= y0) { y0 = 0; y1 = g.h; }
  const int e = grad ? 1 : 0;
  const int d0 = y0 - e > 0 ? y0 - e : 0, d1 = y1 + e < g.h ? y1 + e : g.h;
  launch_derivatives(h->stream, f0, f1w, h->c[C_FX], h->c[C_FY], h->c[C_FT], g, d0, d1);
  TRY(check_launch(h, FLOW2D_K_DERIVATIVES, 1));
  if (grad) {
    float* J[5] = {h->c[C_J0], h->c[C_J1], h->c[C_J2], h->c[C_J3], h->c[C_J4]};
    launch_grad_tensor(h->stream, h->c[C_FX], h->c[C_FY], h->c[C_FT], J, g, y0, y1);
    TRY(check_launch(h, FLOW2D_K_GRAD_TENSOR, 1));
  }
  return FLOW2D_OK;
}

int run_level_solve(flow2d_handle* h, const LevelGeom& g, const float* f0, const float* f1w, const float* u, const float* v,
                    float* du_a, float* dv_a, float* du_b, float* dv_b, float* phi, float* ksi, bool want_phi,
                    const flow2d_params* p, const SolvePlan& pl, int slot, int D0, int D1) {
  if (!ext_solver(p)) {
    TRY(run_derivatives(h, g, f0, f1w, D0, D1));
    return run_solve(h, g, u, v, du_a, dv_a, du_b, dv_b, phi, ksi, want_phi, p, pl, slot);
  }
  // extension data terms: fx, fy, ft of the frames (or of log(1 + frame); the resampler's scratch planes are free at
  // this point of a level), then the six tensor planes
  if (p->data_term == FLOW2D_TERM_LOG_GRADIENT) {
    launch_ext_log(h->stream, f0, h->c[C_TMP0], g);
    launch_ext_log(h->stream, f1w, h->c[C_TMP1], g);
    TRY(check_launch(h, FLOW2D_K_EXT, 2));
    f0 = h->c[C_TMP0]; f1w = h->c[C_TMP1];
  }
  launch_derivatives(h->stream, f0, f1w, h->c[C_FX], h->c[C_FY], h->c[C_FT], g);
  TRY(check_launch(h, FLOW2D_K_DERIVATIVES, 1));
  ExtTensor J;
  for (int k = 0; k < 6; k++) J.p[k] = h->J6[k];
  launch_ext_tensor(h->stream, h->c[C_FX], h->c[C_FY], h->c[C_FT], J, g, p->data_term, p->gamma);
  TRY(check_launch(h, FLOW2D_K_EXT, 1));
  return run_solve_ext(h, g, u, v, du_a, dv_a, du_b, dv_b, phi, ksi, p, slot);
}

int validate_params(flow2d_handle* h, const flow2d_params* p, int* median) {
  if (!p) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "params is null");
  if (!(p->warp_scale_factor > 0.f) || !(p->warp_scale_factor < 1.f))
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT,
                "warp_scale_factor %g must be in (0,1): the reference runs no level at all for >= 1 "
                "(optical_flow_base_2d.cpp:43-58); use warp_levels_count = 1 for a single level",
                (double)p->warp_scale_factor);
  if (p->warp_levels_count < 1) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "warp_levels_count must be >= 1");
  if (p->sweeps_per_pass < 0 || p->sweeps_per_pass > FLOW2D_MAX_SWEEPS_PER_PASS)
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "sweeps_per_pass must be 0..%d", FLOW2D_MAX_SWEEPS_PER_PASS);
  if (p->outer_iterations_count > 1000000 || p->inner_iterations_count > 1000000)
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "iteration counts out of range");
  // opt-in extensions
  if (p->scheme != FLOW2D_SCHEME_JACOBI && p->scheme != FLOW2D_SCHEME_RED_BLACK)
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "scheme must be FLOW2D_SCHEME_JACOBI or FLOW2D_SCHEME_RED_BLACK");
  if (!(p->omega >= 0.f) || !(p->omega < 2.f)) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "omega %g must be in [0, 2)", (double)p->omega);
  if (p->data_term < FLOW2D_TERM_DEFAULT || p->data_term > FLOW2D_TERM_COMBINED)
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "data_term must be one of FLOW2D_TERM_*");
  if (!(p->gamma >= 0.f) || !(p->gamma < 1e30f)) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "gamma must be finite and >= 0");
  if (!(p->residual_tolerance >= 0.f)) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "residual_tolerance must be >= 0");
  if (h && ext_solver(p) && h->constancy != FLOW2D_GREY)
    return fail(h, FLOW2D_ERR_UNSUPPORTED, "scheme / omega / data_term need a FLOW2D_GREY handle (use data_term = FLOW2D_TERM_GRADIENT "
                                            "for gradient constancy)");
  return normalise_median(h, p->median_radius, median);
}

// The pyramid.  frame_0 / frame_1 / out_u / out_v are device containers.
// With `slab` (flow2d_compute_slab_device on a connected handle) the large levels are slabbed by rows: this rank then
// produces its own rows of every stage (plus the margins the next stage needs) and of the final flow.
int enqueue_pyramid(flow2d_handle* h, const float* frame_0, const float* frame_1, float* out_u, float* out_v,
                    const flow2d_params* p, bool slab = false) {
  int median = 1;
  TRY(validate_params(h, p, &median));
  const size_t W = h->W, H = h->H;
  cudaStream_t st = h->stream;
  h->levels_run = 0;
  const bool slab_on = slab && h->slab_world > 1;
  if (slab_on && any_extension(p))
    return fail(h, FLOW2D_ERR_UNSUPPORTED, "the opt-in extensions (scheme, omega, data_term, residual_tolerance, cascaded_restriction) "
                                            "are not available in row-slab mode");
  if (!slab_on && any_extension(p) &&
      ((ext_solver(p) && !h->ext_pool) || (early_exit(p) && !h->d_stop) ||
       (p->cascaded_restriction && h->pyr_rows < pyramid_rows(h, p, levels_of(h, p), nullptr))))
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "internal: extension buffers not allocated");
  const bool residuals = p->report_residuals != 0 && !slab_on;
  if (early_exit(p)) CU_TRY(h, cudaMemsetAsync(h->d_stop, 0, sizeof(int) * 2 * FLOW2D_MAX_LEVELS, st));
  h->iter_levels = 0;
  h->iter_default = (int)p->outer_iterations_count;
  h->residual_levels = 0;
  h->slab_levels = 0;
  if (residuals) CU_TRY(h, cudaMemsetAsync(h->d_residuals, 0, sizeof(double) * 2 * FLOW2D_MAX_LEVELS, st));

  const bool level_times = p->report_level_times != 0 && !slab_on;
  h->timed_levels = 0;
  auto stamp_level = [&](int slot) {  // event `slot` (0: level start, 1: solve start, 2: solve end) of the current level
    if (!level_times || h->timed_levels >= FLOW2D_MAX_LEVELS) return;
    const size_t i = (size_t)h->timed_levels * 3 + slot;
    while (h->level_events.size() <= i) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreate(&e) != cudaSuccess) { (void)cudaGetLastError(); return; }
      h->level_events.push_back(e);
    }
    cudaEventRecord(h->level_events[i], st);
  };

  // presmoothing (optical_flow_2d.cpp:218-246)
  const float* frame[2] = {frame_0, frame_1};
  if (p->gaussian_sigma > 0.0f) {
    NvtxRange r_blur("flow2d: presmoothing");
    GaussTaps taps;
    TRY(gauss_taps(h, p->gaussian_sigma, &taps));
    for (int i = 0; i < 2; i++) {
      launch_blur(st, frame[i], h->c[C_BLUR0 + i], (int)W, (int)H, (int)h->pitch, taps);
      TRY(check_launch(h, FLOW2D_K_BLUR, 1));
      frame[i] = h->c[C_BLUR0 + i];
    }
  }

  const size_t max_level = flow2d_max_warp_level(W, H, p->warp_scale_factor);
  int level = (int)(p->warp_levels_count < max_level ? p->warp_levels_count : max_level) - 1;
  float *u = h->c[C_U], *v = h->c[C_V], *u2 = h->c[C_U2], *v2 = h->c[C_V2];
  size_t pw = 0, ph = 0;
  const int e = h->constancy == FLOW2D_GRADIENT ? 1 : 0;
  bool prev_slabbed = false;

  // cascaded restriction (extension): both frame pyramids once, level l from level l-1; the reference -- and the
  // default path below -- restricts every level from the full-resolution frames (optical_flow_2d.cpp:279-305)
  size_t pyr_row_of[FLOW2D_MAX_LEVELS + 64] = {};
  const bool cascade = p->cascaded_restriction != 0 && level >= 1 && level < FLOW2D_MAX_LEVELS + 63;
  auto pyr_level = [&](int frame_i, int l) { return h->pyr_pool + ((size_t)frame_i * h->pyr_rows + pyr_row_of[l]) * h->pitch; };
  if (cascade) {
    pyramid_rows(h, p, level + 1, pyr_row_of);
    size_t lw = W, lh = H;
    const float* src[2] = {frame[0], frame[1]};
    for (int l = 1; l <= level; l++) {
      size_t cw, ch; float hx, hy;
      flow2d_level_geometry(W, H, p->warp_scale_factor, l, &cw, &ch, &hx, &hy);
      ResampleJob jb[2];
      for (int i = 0; i < 2; i++) jb[i] = ResampleJob{src[i], h->c[C_TMP0 + i], pyr_level(i, l), (int)lw, (int)lh, (int)cw, (int)ch, 0, 0, 0, 0, 0};
      launch_resample_batch(st, jb, 2, (int)h->pitch);
      TRY(check_launch(h, FLOW2D_K_RESAMPLE, 2));
      for (int i = 0; i < 2; i++) src[i] = pyr_level(i, l);
      lw = cw; lh = ch;
    }
  }

  // default: every level of both pyramids from the full-resolution frames (optical_flow_2d.cpp:279-305), on the side
  // stream, coarsest level first (the order in which the loop below consumes them)
  static const bool no_side = std::getenv("FLOW2D_NO_SIDE_PYRAMID") != nullptr;
  const bool side = !no_side && !slab_on && !cascade && level >= 1 && level < FLOW2D_MAX_LEVELS + 63 && h->side_stream &&
                    (int)h->pyr_events.size() > level && h->pyr_rows >= pyramid_rows(h, p, level + 1, nullptr);
  if (side) {
    NvtxRange r_pyr("flow2d: frame pyramids (side stream)");
    pyramid_rows(h, p, level + 1, pyr_row_of);
    CU_TRY(h, cudaEventRecord(h->pyr_events[0], st));
    CU_TRY(h, cudaStreamWaitEvent(h->side_stream, h->pyr_events[0], 0));
    for (int l = level; l >= 1; --l) {
      size_t cw, ch; float hx, hy;
      flow2d_level_geometry(W, H, p->warp_scale_factor, l, &cw, &ch, &hx, &hy);
      ResampleJob jb[2];
      for (int i = 0; i < 2; i++) jb[i] = ResampleJob{frame[i], h->c[C_RES0 + i], pyr_level(i, l), (int)W, (int)H, (int)cw, (int)ch, 0, 0, 0, 0, 0};
      launch_resample_batch(h->side_stream, jb, 2, (int)h->pitch);
      TRY(check_launch(h, FLOW2D_K_RESAMPLE, 2));
      CU_TRY(h, cudaEventRecord(h->pyr_events[l], h->side_stream));
    }
  }

  while (level >= 0) {
    size_t cw, ch;
    float hx, hy;
    flow2d_level_geometry(W, H, p->warp_scale_factor, level, &cw, &ch, &hx, &hy);
    const LevelGeom g = geom(h, cw, ch, hx, hy);
    char range_name[64];
    std::snprintf(range_name, sizeof range_name, "flow2d: level %d (%zux%zu)", level, cw, ch);
    NvtxRange r_level(range_name);
    stamp_level(0);
    const SolvePlan pl = plan_solve(h, g, p, median, slab_on);
    if (prev_slabbed && !pl.slabbed) return fail(h, FLOW2D_ERR_UNSUPPORTED, "slab: level %dx%d cannot be slabbed after a slabbed coarser level", g.w, g.h);
    // rows of this level on which the stages work: D = derivative planes, Wr = warped frame and flow
    int D0 = 0, D1 = g.h, W0 = 0, W1 = g.h;
    if (pl.slabbed) {
      ++h->slab_levels;
      D0 = pl.Y0 - pl.in_margin > 0 ? pl.Y0 - pl.in_margin : 0;
      D1 = pl.Y1 + pl.in_margin < g.h ? pl.Y1 + pl.in_margin : g.h;
      W0 = D0 - 1 - e > 0 ? D0 - 1 - e : 0;
      W1 = D1 + 1 + e < g.h ? D1 + 1 + e : g.h;
    }

    // frames of this level: always restricted from the full-resolution frames (279-305);
    // flow of this level: zero, or prolongated from the previous level (308-341).
    // Both resamplings share one x launch and one y launch (the increment containers are free at this point
    // and serve as the x-pass scratch of the flow).
    const float* fr[2] = {frame[0], frame[1]};
    ResampleJob jobs[4];
    int njobs = 0;
    if (level != 0 && (cascade || side)) {
      fr[0] = pyr_level(0, level); fr[1] = pyr_level(1, level);
      if (side) CU_TRY(h, cudaStreamWaitEvent(st, h->pyr_events[level], 0));
    } else if (level != 0) {
      // (frame 1 is needed wherever the flow may point: all rows; frame 0 on the rows that are warped / differentiated)
      jobs[njobs++] = ResampleJob{frame[0], h->c[C_TMP0], h->c[C_RES0], (int)W, (int)H, g.w, g.h, W0, W1, 0, 0, 0};
      jobs[njobs++] = ResampleJob{frame[1], h->c[C_TMP1], h->c[C_RES1], (int)W, (int)H, g.w, g.h, 0, 0, 0, 0, 0};
      fr[0] = h->c[C_RES0]; fr[1] = h->c[C_RES1];
    }
    if (pw == 0) {
      CU_TRY(h, cudaMemset2DAsync(u, h->pitch * 4, 0, cw * 4, ch, st));
      CU_TRY(h, cudaMemset2DAsync(v, h->pitch * 4, 0, cw * 4, ch, st));
    } else {
      jobs[njobs++] = ResampleJob{u, h->c[C_DU1], u2, (int)pw, (int)ph, g.w, g.h, W0, W1, 0, 0, 0};
      jobs[njobs++] = ResampleJob{v, h->c[C_DV1], v2, (int)pw, (int)ph, g.w, g.h, W0, W1, 0, 0, 0};
      std::swap(u, u2); std::swap(v, v2);
    }
    if (njobs) {
      NvtxRange r_res("flow2d: resample");
      launch_resample_batch(st, jobs, njobs, g.pitch);
      TRY(check_launch(h, FLOW2D_K_RESAMPLE, 2));
    }
    // backward registration (344-363) and the level's derivative planes
    {
      NvtxRange r_warp("flow2d: warp");
      launch_warp(st, fr[0], fr[1], u, v, h->c[C_WARPED], g, W0, W1);
      TRY(check_launch(h, FLOW2D_K_WARP, 1));
    }
    stamp_level(1);
    // solve (366-406)
    const int slot = h->iter_levels < FLOW2D_MAX_LEVELS ? h->iter_levels : FLOW2D_MAX_LEVELS - 1;
    {
      NvtxRange r_solve("flow2d: solve");
      TRY(run_level_solve(h, g, fr[0], h->c[C_WARPED], u, v, h->c[C_DU0], h->c[C_DV0], h->c[C_DU1], h->c[C_DV1], h->c[C_PHI],
                          h->c[C_KSI], residuals, p, pl, slot, D0, D1));
    }
    stamp_level(2);
    if (level_times && h->timed_levels < FLOW2D_MAX_LEVELS) ++h->timed_levels;
    if (h->iter_levels < FLOW2D_MAX_LEVELS) ++h->iter_levels;
    if (residuals && h->residual_levels < FLOW2D_MAX_LEVELS && p->outer_iterations_count > 0 && p->inner_iterations_count > 0) {
      const bool ext = ext_solver(p);
      const float* J[5] = {h->c[C_J0], h->c[C_J1], h->c[C_J2], h->c[C_J3], h->c[C_J4]};
      if (ext) for (int k = 0; k < 5; k++) J[k] = h->J6[k];
      launch_residual(st, h->c[C_FX], h->c[C_FY], h->c[C_FT], J, ext || h->constancy == FLOW2D_GRADIENT, u, v, h->c[C_DU0],
                      h->c[C_DV0], h->c[C_PHI], h->c[C_KSI], g, p->equation_alpha, h->d_residuals + 2 * h->residual_levels);
      TRY(check_launch(h, FLOW2D_K_RESIDUAL, 1));
      h->residual_px[h->residual_levels++] = g.w * g.h;
    }
    // u += du, v += dv, median (409-449); the finest level writes the caller's flow containers
    float* nu = level == 0 ? out_u : u2;
    float* nv = level == 0 ? out_v : v2;
    {
      const float* a[2] = {u, v};
      const float* b[2] = {h->c[C_DU0], h->c[C_DV0]};
      float* out[2] = {nu, nv};
      NvtxRange r_med("flow2d: add + median");
      launch_add_median(st, a, b, out, 2, g.w, g.h, g.pitch, median, pl.Y0, pl.Y1);
      TRY(check_launch(h, FLOW2D_K_ADD_MEDIAN, 1));
      std::swap(u, u2); std::swap(v, v2);
    }
    // slabbed: the next level's prolongation reads rows of this level's flow that the neighbours own
    if (pl.slabbed && level > 0) {
      size_t nw, nh;
      float nhx, nhy;
      flow2d_level_geometry(W, H, p->warp_scale_factor, level - 1, &nw, &nh, &nhx, &nhy);
      const LevelGeom ng = geom(h, nw, nh, nhx, nhy);
      const SolvePlan npl = plan_solve(h, ng, p, median, true);
      if (!npl.slabbed) return fail(h, FLOW2D_ERR_UNSUPPORTED, "slab: finer level %dx%d not slabbed", ng.w, ng.h);
      // rows [lo_q, hi_q) of THIS level that rank q reads when it prolongates to its rows of the next level
      auto needs = [&](int q, int* lo, int* hi) {
        int z0, z1;
        slab_own_rows(ng.h, q, h->slab_world, &z0, &z1);
        const int m = npl.in_margin + 1 + e;
        const int o0 = z0 - m > 0 ? z0 - m : 0, o1 = z1 + m < ng.h ? z1 + m : ng.h;
        const float delta = (float)g.h / (float)ng.h;  // as in the resampler: rows of this level per row of the next
        *lo = (int)floorf((float)o0 * delta);
        const int t = (int)ceilf((float)o1 * delta);
        *hi = t < g.h ? t : g.h;
      };
      const int r = h->slab_rank;
      int lo, hi, up0 = 0, upn = 0, dn0 = 0, dnn = 0, rup0 = 0, rupn = 0, rdn0 = 0, rdnn = 0;
      needs(r, &lo, &hi);
      if (lo < pl.Y0) { rup0 = lo; rupn = pl.Y0 - lo; }
      if (hi > pl.Y1) { rdn0 = pl.Y1; rdnn = hi - pl.Y1; }
      if (r > 0) {  // what the rank above reads below its own rows = my first rows
        int l2, h2, a0, a1;
        needs(r - 1, &l2, &h2);
        slab_own_rows(g.h, r - 1, h->slab_world, &a0, &a1);
        if (h2 > a1) { up0 = a1; upn = h2 - a1; }
        if (rupn > a1 - a0) return fail(h, FLOW2D_ERR_UNSUPPORTED, "slab: halo exceeds the neighbour's rows");
      }
      if (r < h->slab_world - 1) {
        int l2, h2, b0, b1;
        needs(r + 1, &l2, &h2);
        slab_own_rows(g.h, r + 1, h->slab_world, &b0, &b1);
        if (l2 < b0) { dn0 = l2; dnn = b0 - l2; }
        if (rdnn > b1 - b0) return fail(h, FLOW2D_ERR_UNSUPPORTED, "slab: halo exceeds the neighbour's rows");
      }
      if (upn > pl.Y1 - pl.Y0 || dnn > pl.Y1 - pl.Y0) return fail(h, FLOW2D_ERR_UNSUPPORTED, "slab: halo exceeds the own rows");
      TRY(slab_exchange(h, nu, nv, g, up0, upn, dn0, dnn, rup0, rupn, rdn0, rdnn));
    }
    prev_slabbed = pl.slabbed;
    pw = cw; ph = ch;
    --level;
    ++h->levels_run;
  }
  return FLOW2D_OK;
}

// The level schedule of one (frames, flow, parameters) combination is captured into a CUDA graph
// the first time and replayed afterwards: a sequence of frame pairs through one handle costs one
// graph launch per pair instead of ~1 500 kernel launches of host work.
struct GraphKey {
  const void* ptr[4];
  flow2d_params p;
  const void* timing;
};

// launch = false: only make sure the schedule is in the graph cache (flow2d_prepare)
int compute_on_device(flow2d_handle* h, const float* frame_0, const float* frame_1, float* out_u, float* out_v,
                      const flow2d_params* p, bool launch = true) {
  int median = 1;
  TRY(validate_params(h, p, &median));
  if (p->gaussian_sigma > 0.0f) {
    GaussTaps taps;
    TRY(gauss_taps(h, p->gaussian_sigma, &taps));
  }
  TRY(ensure_ext(h, p));  // allocations of the opt-in extensions: never inside a stream capture
  static const bool no_graph = std::getenv("FLOW2D_NO_GRAPH") != nullptr;  // A/B switch for measurements
  // per-level timers are CUDA events between the stages: plain enqueue (the reference's timed path is not a graph either)
  if (no_graph || p->report_level_times) return launch ? enqueue_pyramid(h, frame_0, frame_1, out_u, out_v, p) : FLOW2D_OK;

  GraphKey key;
  std::memset(&key, 0, sizeof key);
  key.ptr[0] = frame_0; key.ptr[1] = frame_1; key.ptr[2] = out_u; key.ptr[3] = out_v;
  // field by field: padding bytes of the caller's struct must not take part in the comparison
  key.p.warp_levels_count = p->warp_levels_count; key.p.warp_scale_factor = p->warp_scale_factor;
  key.p.outer_iterations_count = p->outer_iterations_count; key.p.inner_iterations_count = p->inner_iterations_count;
  key.p.equation_alpha = p->equation_alpha; key.p.equation_smoothness = p->equation_smoothness;
  key.p.equation_data = p->equation_data; key.p.median_radius = p->median_radius;
  key.p.gaussian_sigma = p->gaussian_sigma; key.p.sweeps_per_pass = p->sweeps_per_pass;
  key.p.resident_levels = p->resident_levels;
  key.p.throughput_mode = p->throughput_mode;
  key.p.report_residuals = p->report_residuals;
  key.p.scheme = p->scheme; key.p.omega = p->omega; key.p.data_term = p->data_term; key.p.gamma = p->gamma;
  key.p.residual_tolerance = p->residual_tolerance; key.p.residual_check_every = p->residual_check_every;
  key.p.cascaded_restriction = p->cascaded_restriction;
  key.timing = h->timing;
  static_assert(sizeof(GraphKey) <= sizeof(flow2d_handle::GraphEntry::key), "graph key storage too small");
  flow2d_handle::GraphEntry* slot = nullptr;
  for (auto& g : h->graphs)
    if (g.exec && std::memcmp(&key, g.key, sizeof key) == 0) slot = &g;
  if (slot && !launch) return FLOW2D_OK;
  if (slot) {
    NvtxRange r_replay("flow2d: graph replay");
    CU_TRY(h, cudaGraphLaunch(slot->exec, h->stream));
    slot->last_use = ++h->graph_clock;
    ++h->graph_replays;
    h->launches = slot->launches;
    for (int k = 0; k < FLOW2D_KERNEL_KINDS; k++) h->kind_launches[k] = slot->kind_launches[k];
    h->levels_run = slot->levels;
    h->residual_levels = slot->residual_levels;
    std::memcpy(h->residual_px, slot->residual_px, sizeof h->residual_px);
    h->iter_levels = slot->iter_levels; h->iter_default = slot->iter_default;
    return FLOW2D_OK;
  }
  if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    (void)cudaGetLastError();
    return launch ? enqueue_pyramid(h, frame_0, frame_1, out_u, out_v, p) : FLOW2D_OK;  // e.g. the stream is already capturing
  }
  reset_launch_counts(h);
  const int rc = enqueue_pyramid(h, frame_0, frame_1, out_u, out_v, p);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  if (rc != FLOW2D_OK || e != cudaSuccess || !graph) {
    (void)cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    if (rc != FLOW2D_OK) return rc;
    return launch ? enqueue_pyramid(h, frame_0, frame_1, out_u, out_v, p) : FLOW2D_OK;
  }
  cudaGraphExec_t exec = nullptr;
  if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
    (void)cudaGetLastError();
    cudaGraphDestroy(graph);
    return launch ? enqueue_pyramid(h, frame_0, frame_1, out_u, out_v, p) : FLOW2D_OK;
  }
  cudaGraphDestroy(graph);
  // a free slot, else the least recently used one (its graph may still be executing: destroying an exec that is
  // in flight is allowed, the driver defers the release)
  slot = &h->graphs[0];
  for (auto& g : h->graphs) {
    if (!g.exec) { slot = &g; break; }
    if (g.last_use < slot->last_use) slot = &g;
  }
  if (slot->exec) cudaGraphExecDestroy(slot->exec);
  slot->exec = exec;
  std::memcpy(slot->key, &key, sizeof key);
  slot->launches = h->launches;
  for (int k = 0; k < FLOW2D_KERNEL_KINDS; k++) slot->kind_launches[k] = h->kind_launches[k];
  slot->levels = h->levels_run;
  slot->residual_levels = h->residual_levels;
  std::memcpy(slot->residual_px, h->residual_px, sizeof h->residual_px);
  slot->iter_levels = h->iter_levels; slot->iter_default = h->iter_default;
  slot->last_use = ++h->graph_clock;
  ++h->graph_captures;
  if (launch) CU_TRY(h, cudaGraphLaunch(slot->exec, h->stream));
  return FLOW2D_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
#define STAGE_PROLOGUE(h)                                    \
  if (!(h)) return FLOW2D_ERR_INVALID_ARGUMENT;              \
  CU_TRY((h), cudaSetDevice((h)->device))

extern "C" {

const char* flow2d_version(void) { return "flow2d-b200 0.1.0 sm_100a"; }

void* flow2d_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  return p;
}

void flow2d_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

void flow2d_default_params(flow2d_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof *p);
  p->warp_levels_count = 50;     // main.cpp:70
  p->warp_scale_factor = 0.9f;   // main.cpp:71
  p->outer_iterations_count = 40;
  p->inner_iterations_count = 5;
  p->equation_alpha = 35.0f;
  p->equation_smoothness = 0.001f;
  p->equation_data = 0.001f;
  p->median_radius = 5;
  p->gaussian_sigma = 1.5f;
  p->sweeps_per_pass = 0;
  p->resident_levels = 0;
  p->throughput_mode = 0;
  p->report_residuals = 0;
  p->scheme = FLOW2D_SCHEME_JACOBI;
  p->omega = 0.f;
  p->data_term = FLOW2D_TERM_DEFAULT;
  p->gamma = 0.f;
  p->residual_tolerance = 0.f;
  p->residual_check_every = 0;
  p->cascaded_restriction = 0;
  p->report_level_times = 0;
}

size_t flow2d_max_warp_level(size_t width, size_t height, float scale_factor) {
  // optical_flow_base_2d.cpp:36-59, same fp32 expressions
  size_t r_width = 1, r_height = 1, level_counter = 1;
  while (scale_factor < 1.f) {
    const float scale = std::pow(scale_factor, static_cast<float>(level_counter));
    r_width = static_cast<size_t>(std::ceil(width * scale));
    r_height = static_cast<size_t>(std::ceil(height * scale));
    if (r_width < 4 || r_height < 4) break;
    ++level_counter;
  }
  if (r_width == 1 || r_height == 1) --level_counter;
  return level_counter;
}

int flow2d_level_geometry(size_t width, size_t height, float scale_factor, int level, size_t* cw, size_t* ch, float* hx,
                          float* hy) {
  if (!cw || !ch || !hx || !hy || level < 0) return FLOW2D_ERR_INVALID_ARGUMENT;
  // optical_flow_2d.cpp:268-272
  const float scale = std::pow(scale_factor, static_cast<float>(level));
  *cw = static_cast<size_t>(std::ceil(width * scale));
  *ch = static_cast<size_t>(std::ceil(height * scale));
  *hx = width / static_cast<float>(*cw);
  *hy = height / static_cast<float>(*ch);
  return FLOW2D_OK;
}

int flow2d_create(flow2d_handle** out, int device, size_t width, size_t height, int constancy) {
  if (!out) return FLOW2D_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (width < 4 || height < 4 || width > 65536 || height > 65536) return FLOW2D_ERR_INVALID_ARGUMENT;
  if (constancy != FLOW2D_GREY && constancy != FLOW2D_GRADIENT) return FLOW2D_ERR_UNSUPPORTED;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) {
    (void)cudaGetLastError();
    return FLOW2D_ERR_NO_DEVICE;
  }
  if (cudaSetDevice(device) != cudaSuccess) return FLOW2D_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FLOW2D_ERR_CUDA;
  if (prop.major != 10) return FLOW2D_ERR_NO_DEVICE;  // sm_100a only: no other code path exists

  flow2d_handle* h = new (std::nothrow) flow2d_handle;
  if (!h) return FLOW2D_ERR_OUT_OF_MEMORY;
  h->device = device;
  h->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
  h->W = width; h->H = height;
  h->pitch = (width + kPitchAlign - 1) / kPitchAlign * kPitchAlign;
  h->constancy = constancy;
  const int ncont = constancy == FLOW2D_GRADIENT ? C_COUNT : C_J0;
  const size_t csize = h->pitch * height;
  if (csize > (size_t)0x7fffffff) {  // kernels index a container with 32-bit offsets (25+ such containers exceed any HBM anyway)
    delete h;
    return FLOW2D_ERR_OUT_OF_MEMORY;
  }
  if (cudaMalloc(&h->pool, csize * ncont * sizeof(float)) != cudaSuccess) {
    (void)cudaGetLastError();
    delete h;
    return FLOW2D_ERR_OUT_OF_MEMORY;
  }
  for (int i = 0; i < ncont; i++) h->c[i] = h->pool + csize * i;
  if (cudaMalloc(&h->d_residuals, sizeof(double) * 2 * FLOW2D_MAX_LEVELS) != cudaSuccess) {
    (void)cudaGetLastError();
    flow2d_destroy(h);
    return FLOW2D_ERR_OUT_OF_MEMORY;
  }
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&h->ev_start) != cudaSuccess || cudaEventCreate(&h->ev_stop) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_in_ready[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_in_ready[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_in_free[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_in_free[1], cudaEventDisableTiming) != cudaSuccess ||
      solve_pass_configure() != cudaSuccess || solve_pass2_configure() != cudaSuccess || solve_pass3_configure() != cudaSuccess || cudaMemsetAsync(h->pool, 0, csize * ncont * sizeof(float), h->own_stream) != cudaSuccess ||
      cudaStreamSynchronize(h->own_stream) != cudaSuccess) {
    (void)cudaGetLastError();
    flow2d_destroy(h);
    return FLOW2D_ERR_CUDA;
  }
  h->stream = h->own_stream;
  h->pass3 = solve_pass3_available();
  {
    // A/B switches, read per handle (tests flip them between handles of one process)
    const char* e = std::getenv("FLOW2D_CLUSTER");
    h->cluster_whole = e ? std::atoi(e) : kClusterWholeDefault;
    e = std::getenv("FLOW2D_CLUSTER_PASS");
    h->cluster_pass = e ? std::atoi(e) : kClusterPassDefault;
    h->cluster_compact = std::getenv("FLOW2D_CLUSTER_COMPACT") != nullptr;
    int cmax = kClusterMaxCtas;
    if ((e = std::getenv("FLOW2D_CLUSTER_MAX")) != nullptr) cmax = std::atoi(e);
    if (h->cluster_whole != 0 || h->cluster_pass)
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < 3; j++) h->cluster_active[i][j] = (2 << i) <= cmax ? solve_cluster_max_active(2 << i, 256 << j) : 0;
  }
  if (device < kMaxCountedDevices) { ++g_live_handles[device]; h->counted = true; }
  *out = h;
  return FLOW2D_OK;
}

int flow2d_live_handles(int device) { return device >= 0 && device < kMaxCountedDevices ? g_live_handles[device].load() : 0; }

int flow2d_destroy(flow2d_handle* h) {
  if (!h) return FLOW2D_OK;
  if (h->counted) { --g_live_handles[h->device]; h->counted = false; }
  cudaSetDevice(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  for (int i = 0; i < 2; i++) {
    if (h->ev_in_ready[i]) cudaEventDestroy(h->ev_in_ready[i]);
    if (h->ev_in_free[i]) cudaEventDestroy(h->ev_in_free[i]);
  }
  for (auto& g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  if (h->ev_stop) cudaEventDestroy(h->ev_stop);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  for (void* m : h->imported)
    if (m) cudaIpcCloseMemHandle(m);
  if (h->mailbox) cudaFree(h->mailbox);
  if (h->slab_counter) cudaFree(h->slab_counter);
  if (h->pool) cudaFree(h->pool);
  if (h->d_residuals) cudaFree(h->d_residuals);
  if (h->ext_pool) cudaFree(h->ext_pool);
  if (h->pyr_pool) cudaFree(h->pyr_pool);
  if (h->d_stop) cudaFree(h->d_stop);
  if (h->d_partials) cudaFree(h->d_partials);
  if (h->d_counter) cudaFree(h->d_counter);
  for (cudaEvent_t e : h->level_events) cudaEventDestroy(e);
  for (cudaEvent_t e : h->pyr_events) cudaEventDestroy(e);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  delete h;
  return FLOW2D_OK;
}

const char* flow2d_last_error(const flow2d_handle* h) { return h ? h->err.c_str() : "null handle"; }
size_t flow2d_pitch_elems(const flow2d_handle* h) { return h ? h->pitch : 0; }
size_t flow2d_width(const flow2d_handle* h) { return h ? h->W : 0; }
size_t flow2d_height(const flow2d_handle* h) { return h ? h->H : 0; }

int flow2d_set_stream(flow2d_handle* h, void* cuda_stream) {
  if (!h) return FLOW2D_ERR_INVALID_ARGUMENT;
  h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return FLOW2D_OK;
}
void* flow2d_get_stream(const flow2d_handle* h) { return h ? static_cast<void*>(h->stream) : nullptr; }

int flow2d_last_stats(const flow2d_handle* h, long long* kernel_launches, int* levels_run, float* device_ms) {
  if (!h) return FLOW2D_ERR_INVALID_ARGUMENT;
  if (kernel_launches) *kernel_launches = h->launches;
  if (levels_run) *levels_run = h->levels_run;
  if (device_ms) *device_ms = h->device_ms;
  return FLOW2D_OK;
}

int flow2d_graph_stats(const flow2d_handle* h, long long* captures, long long* replays) {
  if (!h) return FLOW2D_ERR_INVALID_ARGUMENT;
  if (captures) *captures = h->graph_captures;
  if (replays) *replays = h->graph_replays;
  return FLOW2D_OK;
}

int flow2d_level_residuals(flow2d_handle* h, double* rms_u, double* rms_v, int capacity, int* levels) {
  STAGE_PROLOGUE(h);
  if (levels) *levels = h->residual_levels;
  if (h->residual_levels == 0 || capacity <= 0) return FLOW2D_OK;
  double sums[2 * FLOW2D_MAX_LEVELS];
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  CU_TRY(h, cudaMemcpy(sums, h->d_residuals, sizeof(double) * 2 * h->residual_levels, cudaMemcpyDeviceToHost));
  for (int i = 0; i < h->residual_levels && i < capacity; i++) {
    if (rms_u) rms_u[i] = std::sqrt(sums[2 * i] / h->residual_px[i]);
    if (rms_v) rms_v[i] = std::sqrt(sums[2 * i + 1] / h->residual_px[i]);
  }
  return FLOW2D_OK;
}

int flow2d_level_outer_iterations(flow2d_handle* h, int* iterations, int capacity, int* levels) {
  STAGE_PROLOGUE(h);
  if (levels) *levels = h->iter_levels;
  if (!iterations || capacity <= 0 || h->iter_levels == 0) return FLOW2D_OK;
  int used[FLOW2D_MAX_LEVELS] = {};
  if (h->d_stop) {
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    CU_TRY(h, cudaMemcpy(used, h->d_stop + FLOW2D_MAX_LEVELS, sizeof(int) * h->iter_levels, cudaMemcpyDeviceToHost));
  }
  for (int i = 0; i < h->iter_levels && i < capacity; i++) iterations[i] = used[i] > 0 ? used[i] : h->iter_default;
  return FLOW2D_OK;
}

int flow2d_level_times(flow2d_handle* h, float* level_ms, float* solve_ms, int capacity, int* levels) {
  STAGE_PROLOGUE(h);
  if (levels) *levels = h->timed_levels;
  if (h->timed_levels == 0 || capacity <= 0) return FLOW2D_OK;
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < h->timed_levels && i < capacity; i++) {
    // a level ends where the next one starts; the last one with its solve (the final add + median is not timed)
    float a = 0.f, b = 0.f;
    const bool last = i + 1 >= h->timed_levels;
    if (cudaEventElapsedTime(&a, h->level_events[3 * i], last ? h->level_events[3 * i + 2] : h->level_events[3 * (i + 1)]) != cudaSuccess ||
        cudaEventElapsedTime(&b, h->level_events[3 * i + 1], h->level_events[3 * i + 2]) != cudaSuccess) {
      (void)cudaGetLastError();
      return fail(h, FLOW2D_ERR_CUDA, "level timers not available");
    }
    if (level_ms) level_ms[i] = a;
    if (solve_ms) solve_ms[i] = b;
  }
  return FLOW2D_OK;
}

int flow2d_stage_residual(flow2d_handle* h, const float* d_frame_0, const float* d_frame_1_warped, const float* d_u,
                          const float* d_v, const float* d_du, const float* d_dv, const float* d_phi, const float* d_ksi,
                          size_t w, size_t hh, float hx, float hy, const flow2d_params* p, double* rms_u, double* rms_v) {
  STAGE_PROLOGUE(h);
  if (!d_frame_0 || !d_frame_1_warped || !d_u || !d_v || !d_du || !d_dv || !d_phi || !d_ksi || !p)
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "null pointer");
  if (w < 2 || hh < 2 || w > h->W || hh > h->H) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "level %zux%zu does not fit the handle", w, hh);
  TRY(check_aligned(h, "residual", {d_frame_0, d_frame_1_warped, d_u, d_v, d_du, d_dv, d_phi, d_ksi}));
  const LevelGeom g = geom(h, w, hh, hx, hy);
  TRY(run_derivatives(h, g, d_frame_0, d_frame_1_warped));
  CU_TRY(h, cudaMemsetAsync(h->d_residuals, 0, sizeof(double) * 2, h->stream));
  const float* J[5] = {h->c[C_J0], h->c[C_J1], h->c[C_J2], h->c[C_J3], h->c[C_J4]};
  launch_residual(h->stream, h->c[C_FX], h->c[C_FY], h->c[C_FT], J, h->constancy == FLOW2D_GRADIENT, d_u, d_v, d_du, d_dv,
                  d_phi, d_ksi, g, p->equation_alpha, h->d_residuals);
  TRY(check_launch(h, FLOW2D_K_RESIDUAL, 1));
  double sums[2];
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  CU_TRY(h, cudaMemcpy(sums, h->d_residuals, sizeof sums, cudaMemcpyDeviceToHost));
  h->residual_levels = 0;
  if (rms_u) *rms_u = std::sqrt(sums[0] / ((double)w * (double)hh));
  if (rms_v) *rms_v = std::sqrt(sums[1] / ((double)w * (double)hh));
  return FLOW2D_OK;
}

int flow2d_last_launch_counts(const flow2d_handle* h, long long* counts) {
  if (!h || !counts) return FLOW2D_ERR_INVALID_ARGUMENT;
  for (int k = 0; k < FLOW2D_KERNEL_KINDS; k++) counts[k] = h->kind_launches[k];
  return FLOW2D_OK;
}
int flow2d_cluster_shape(int rw, int rh, int compact, int shape[5]) {
  if (!shape || rw < 2 || rh < 2) return FLOW2D_ERR_INVALID_ARGUMENT;
  int all[4][3];
  for (auto& row : all)
    for (int& v : row) v = 1;
  const ClusterPlan cp = plan_cluster(all, rw, rh, compact != 0);
  if (!cp.threads) return FLOW2D_ERR_UNSUPPORTED;
  shape[0] = cp.cg.cx; shape[1] = cp.cg.cy; shape[2] = cp.cg.tw; shape[3] = cp.cg.th; shape[4] = cp.threads;
  return FLOW2D_OK;
}

int flow2d_debug_cluster_cell(const int geom[5], const int level[7], int rank, int cluster, int thread, int cell[15]) {
  if (!geom || !level || !cell) return FLOW2D_ERR_INVALID_ARGUMENT;
  ClusterGeom cg;
  cg.cx = geom[0]; cg.cy = geom[1]; cg.tw = geom[2]; cg.th = geom[3]; cg.ncx = geom[4];
  if (cg.cx < 1 || cg.cy < 1 || cg.cx * cg.cy > kClusterMaxCtas || cg.tw < 2 || cg.th < 2 || cg.tw * cg.th > 1024 || cg.ncx < 1 ||
      rank < 0 || rank >= cg.cx * cg.cy || cluster < 0 || thread < 0 || thread >= 1024)
    return FLOW2D_ERR_INVALID_ARGUMENT;
  ClusterLevel lv;
  lv.w = level[0]; lv.h = level[1]; lv.ow = level[2]; lv.oh = level[3]; lv.halo = level[4]; lv.y0 = level[5]; lv.y1 = level[6];
  const ClusterCell c = cluster_cell(cg, lv, rank, cluster, thread);
  const int v[15] = {c.ac, c.al, c.ar, c.au, c.ad, c.push_h_rank, c.push_h, c.push_v_rank, c.push_v, c.gx, c.gy, c.mine, c.live, c.out,
                     cluster_plane(cg.tw * cg.th <= 256 ? 256 : cg.tw * cg.th <= 512 ? 512 : 1024)};
  for (int i = 0; i < 15; i++) cell[i] = v[i];
  return FLOW2D_OK;
}

const char* flow2d_kernel_kind_name(int kind) { return kind >= 0 && kind < FLOW2D_KERNEL_KINDS ? kKindNames[kind] : ""; }

int flow2d_compute_device(flow2d_handle* h, const float* d_frame_0, const float* d_frame_1, float* d_flow_u,
                          float* d_flow_v, const flow2d_params* p) {
  if (!h) return FLOW2D_ERR_INVALID_ARGUMENT;
  if (!d_frame_0 || !d_frame_1 || !d_flow_u || !d_flow_v) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "null image pointer");
  if (!aligned16(d_frame_0) || !aligned16(d_frame_1) || !aligned16(d_flow_u) || !aligned16(d_flow_v))
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "device containers must be 16-byte aligned");
  CU_TRY(h, cudaSetDevice(h->device));
  reset_launch_counts(h);
  return compute_on_device(h, d_frame_0, d_frame_1, d_flow_u, d_flow_v, p);
}

int flow2d_compute_async(flow2d_handle* h, const float* frame_0, const float* frame_1, float* flow_u, float* flow_v,
                         const flow2d_params* p) {
  if (!h) return FLOW2D_ERR_INVALID_ARGUMENT;
  if (!frame_0 || !frame_1 || !flow_u || !flow_v) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "null image pointer");
  CU_TRY(h, cudaSetDevice(h->device));
  reset_launch_counts(h);
  const size_t row = h->W * sizeof(float), dpitch = h->pitch * sizeof(float);
  cudaStream_t st = h->stream, cs = h->copy_stream;
  // CopyData2DtoDevice (cuda_utils.cpp:66-84) on the upload stream into a landing pair; the handle's stream moves the
  // frames from there into its fixed input pair (two device-to-device copies, microseconds) and frees the landing pair at
  // once, so the upload of the next call runs while this one computes -- and the captured schedule, which reads the fixed
  // pair, is the same graph for every call
  float* land0 = h->c[C_IN2];
  float* land1 = h->c[C_IN3];
  ++h->async_calls;
  CU_TRY(h, cudaStreamWaitEvent(cs, h->ev_in_free[0], 0));
  CU_TRY(h, cudaEventRecord(h->ev_start, cs));
  CU_TRY(h, cudaMemcpy2DAsync(land0, dpitch, frame_0, row, row, h->H, cudaMemcpyHostToDevice, cs));
  CU_TRY(h, cudaMemcpy2DAsync(land1, dpitch, frame_1, row, row, h->H, cudaMemcpyHostToDevice, cs));
  CU_TRY(h, cudaEventRecord(h->ev_in_ready[0], cs));
  CU_TRY(h, cudaStreamWaitEvent(st, h->ev_in_ready[0], 0));
  CU_TRY(h, cudaMemcpy2DAsync(h->c[C_IN0], dpitch, land0, dpitch, row, h->H, cudaMemcpyDeviceToDevice, st));
  CU_TRY(h, cudaMemcpy2DAsync(h->c[C_IN1], dpitch, land1, dpitch, row, h->H, cudaMemcpyDeviceToDevice, st));
  CU_TRY(h, cudaEventRecord(h->ev_in_free[0], st));
  float* out_u = h->c[C_OUT_U];
  float* out_v = h->c[C_OUT_V];
  int rc = compute_on_device(h, h->c[C_IN0], h->c[C_IN1], out_u, out_v, p);
  if (rc != FLOW2D_OK) {
    cudaStreamSynchronize(st);
    return rc;
  }
  // CopyData2DFromDevice (cuda_utils.cpp:87-105)
  CU_TRY(h, cudaMemcpy2DAsync(flow_u, row, out_u, dpitch, row, h->H, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpy2DAsync(flow_v, row, out_v, dpitch, row, h->H, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaEventRecord(h->ev_stop, st));
  return FLOW2D_OK;
}

int flow2d_prepare(flow2d_handle* h, const flow2d_params* p) {
  if (!h || !p) return FLOW2D_ERR_INVALID_ARGUMENT;
  CU_TRY(h, cudaSetDevice(h->device));
  return compute_on_device(h, h->c[C_IN0], h->c[C_IN1], h->c[C_OUT_U], h->c[C_OUT_V], p, false);
}

int flow2d_synchronize(flow2d_handle* h) {
  if (!h) return FLOW2D_ERR_INVALID_ARGUMENT;
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  if (cudaEventQuery(h->ev_stop) == cudaSuccess && cudaEventQuery(h->ev_start) == cudaSuccess)
    (void)cudaEventElapsedTime(&h->device_ms, h->ev_start, h->ev_stop);
  (void)cudaGetLastError();
  return FLOW2D_OK;
}

int flow2d_compute(flow2d_handle* h, const float* frame_0, const float* frame_1, float* flow_u, float* flow_v,
                   const flow2d_params* p) {
  int rc = flow2d_compute_async(h, frame_0, frame_1, flow_u, flow_v, p);
  if (rc != FLOW2D_OK) return rc;
  return flow2d_synchronize(h);
}

// ---- one large frame on several GPUs ---------------------------------------------------------------------------
static int slab_alloc_mailbox(flow2d_handle* h) {
  if (h->mailbox) return FLOW2D_OK;
  h->mailbox_rows = kSlabRowsCap;
  h->mailbox_bytes = kSlabHeaderBytes + (size_t)8 * h->mailbox_rows * h->pitch * sizeof(float);  // 2 senders x 2 epochs x 2 fields
  if (cudaMalloc(&h->mailbox, h->mailbox_bytes) != cudaSuccess || cudaMalloc(&h->slab_counter, 2 * sizeof(unsigned)) != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(h, FLOW2D_ERR_OUT_OF_MEMORY, "slab mailbox allocation failed");
  }
  CU_TRY(h, cudaMemset(h->mailbox, 0, h->mailbox_bytes));
  CU_TRY(h, cudaMemset(h->slab_counter, 0, 2 * sizeof(unsigned)));
  return FLOW2D_OK;
}

int flow2d_slab_mailbox(flow2d_handle* h, void** d_mailbox, size_t* bytes) {
  STAGE_PROLOGUE(h);
  TRY(slab_alloc_mailbox(h));
  if (d_mailbox) *d_mailbox = h->mailbox;
  if (bytes) *bytes = h->mailbox_bytes;
  return FLOW2D_OK;
}

int flow2d_slab_export(flow2d_handle* h, unsigned char ipc_handle[64]) {
  STAGE_PROLOGUE(h);
  if (!ipc_handle) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "null handle buffer");
  TRY(slab_alloc_mailbox(h));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t m;
  CU_TRY(h, cudaIpcGetMemHandle(&m, h->mailbox));
  std::memcpy(ipc_handle, &m, 64);
  return FLOW2D_OK;
}

int flow2d_slab_import(flow2d_handle* h, const unsigned char ipc_handle[64], void** d_mailbox) {
  STAGE_PROLOGUE(h);
  if (!ipc_handle || !d_mailbox) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "null argument");
  cudaIpcMemHandle_t m;
  std::memcpy(&m, ipc_handle, 64);
  void* ptr = nullptr;
  CU_TRY(h, cudaIpcOpenMemHandle(&ptr, m, cudaIpcMemLazyEnablePeerAccess));
  for (auto& slot : h->imported)
    if (!slot) { slot = ptr; break; }
  *d_mailbox = ptr;
  return FLOW2D_OK;
}

int flow2d_slab_connect(flow2d_handle* h, int rank, int world, void* mailbox_above, void* mailbox_below, size_t min_rows_per_rank) {
  STAGE_PROLOGUE(h);
  if (world < 1 || rank < 0 || rank >= world) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "bad slab rank %d of %d", rank, world);
  if (world > 1 && ((rank > 0 && !mailbox_above) || (rank < world - 1 && !mailbox_below)))
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "slab: rank %d of %d needs the mailboxes of its neighbours", rank, world);
  TRY(slab_alloc_mailbox(h));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  CU_TRY(h, cudaMemset(h->mailbox, 0, kSlabHeaderBytes));  // flags start at epoch 0 (every rank connects before anyone computes)
  CU_TRY(h, cudaMemset(h->slab_counter, 0, 2 * sizeof(unsigned)));
  h->slab_rank = rank; h->slab_world = world;
  h->slab_min_rows = min_rows_per_rank ? min_rows_per_rank : 64;
  // same process, another device: the neighbour's mailbox must be mapped into this device (IPC imports already are)
  for (void* m : {mailbox_above, mailbox_below}) {
    if (!m) continue;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, m) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device != h->device) {
      const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(h, FLOW2D_ERR_CUDA, "slab: no peer access from device %d to device %d: %s", h->device, at.device, cudaGetErrorString(e));
    }
    (void)cudaGetLastError();
  }
  h->peer[0] = static_cast<unsigned char*>(mailbox_above);
  h->peer[1] = static_cast<unsigned char*>(mailbox_below);
  h->slab_epoch = 0;
  h->slab_exchanges = 0; h->slab_bytes_sent = 0;
  if (world > 1) {
    // a rank's stream spins on flags its neighbours set: no kernel may be loaded lazily (with a context
    // synchronisation) once the ranks are in flight
    preload_pyramid_kernels(); preload_median_kernels(); preload_solve_kernels(); preload_solve_pass2_kernels(); preload_solve_pass3_kernels(); preload_solve_cluster_kernels();
    preload_slab_kernels();
    (void)cudaGetLastError();
  }
  return FLOW2D_OK;
}

int flow2d_slab_rows(const flow2d_handle* h, size_t level_height, size_t* y0, size_t* y1) {
  if (!h || !y0 || !y1 || level_height == 0) return FLOW2D_ERR_INVALID_ARGUMENT;
  int a, b;
  slab_own_rows((int)level_height, h->slab_rank, h->slab_world, &a, &b);
  *y0 = (size_t)a; *y1 = (size_t)b;
  return FLOW2D_OK;
}

int flow2d_slab_stats(const flow2d_handle* h, long long* exchanges, long long* bytes_sent, int* levels_slabbed) {
  if (!h) return FLOW2D_ERR_INVALID_ARGUMENT;
  if (exchanges) *exchanges = h->slab_exchanges;
  if (bytes_sent) *bytes_sent = h->slab_bytes_sent;
  if (levels_slabbed) *levels_slabbed = h->slab_levels;
  return FLOW2D_OK;
}

int flow2d_slab_status(flow2d_handle* h) {
  STAGE_PROLOGUE(h);
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  if (!h->slab_counter) return FLOW2D_OK;
  unsigned err = 0;
  CU_TRY(h, cudaMemcpyAsync(&err, h->slab_counter + 1, sizeof err, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  if (err) return fail(h, FLOW2D_ERR_CUDA, "slab: rank %d timed out waiting for halo rows of a neighbour (reconnect every rank)", h->slab_rank);
  return FLOW2D_OK;
}

int flow2d_compute_slab_device(flow2d_handle* h, const float* d_frame_0, const float* d_frame_1, float* d_flow_u,
                               float* d_flow_v, const flow2d_params* p) {
  if (!h) return FLOW2D_ERR_INVALID_ARGUMENT;
  if (!d_frame_0 || !d_frame_1 || !d_flow_u || !d_flow_v) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "null image pointer");
  TRY(check_aligned(h, "compute", {d_frame_0, d_frame_1, d_flow_u, d_flow_v}));
  CU_TRY(h, cudaSetDevice(h->device));
  reset_launch_counts(h);
  return enqueue_pyramid(h, d_frame_0, d_frame_1, d_flow_u, d_flow_v, p, true);  // epochs are kernel arguments: no graph replay
}

int flow2d_stage_solve_slab(flow2d_handle* h, const float* d_frame_0, const float* d_frame_1, const float* d_flow_u,
                            const float* d_flow_v, float* d_flow_du, float* d_flow_dv, size_t w, size_t hh, float hx,
                            float hy, const flow2d_params* p, int* slabbed) {
  STAGE_PROLOGUE(h);
  if (!d_frame_0 || !d_frame_1 || !d_flow_u || !d_flow_v || !d_flow_du || !d_flow_dv || !p)
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "solve: null argument");
  if (p->sweeps_per_pass < 0 || p->sweeps_per_pass > FLOW2D_MAX_SWEEPS_PER_PASS)
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "sweeps_per_pass must be 0..%d", FLOW2D_MAX_SWEEPS_PER_PASS);
  TRY(check_level(h, w, hh));
  TRY(check_aligned(h, "solve", {d_frame_0, d_frame_1, d_flow_u, d_flow_v, d_flow_du, d_flow_dv}));
  const LevelGeom g = geom(h, w, hh, hx, hy);
  TRY(run_derivatives(h, g, d_frame_0, d_frame_1));
  const SolvePlan pl = plan_solve(h, g, p, 1, true);
  if (slabbed) *slabbed = pl.slabbed ? 1 : 0;
  return run_solve(h, g, d_flow_u, d_flow_v, d_flow_du, d_flow_dv, h->c[C_DU1], h->c[C_DV1], h->c[C_PHI], h->c[C_KSI],
                   false, p, pl);
}

// ---- one large frame on several GPUs of one process: N handles, N host threads ----
struct flow2d_slab_group {
  std::vector<flow2d_handle*> h;
  std::string err;
};

int flow2d_slab_group_create(flow2d_slab_group** out, const int* devices, int n, size_t width, size_t height, int constancy) {
  if (!out || !devices || n < 1) return FLOW2D_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  flow2d_slab_group* g = new (std::nothrow) flow2d_slab_group;
  if (!g) return FLOW2D_ERR_OUT_OF_MEMORY;
  int rc = FLOW2D_OK;
  for (int r = 0; r < n && rc == FLOW2D_OK; r++) {
    flow2d_handle* hd = nullptr;
    rc = flow2d_create(&hd, devices[r], width, height, constancy);
    if (rc == FLOW2D_OK) g->h.push_back(hd);
  }
  std::vector<void*> box(n, nullptr);
  for (int r = 0; r < n && rc == FLOW2D_OK; r++) rc = flow2d_slab_mailbox(g->h[r], &box[r], nullptr);
  for (int r = 0; r < n && rc == FLOW2D_OK; r++)
    rc = flow2d_slab_connect(g->h[r], r, n, r > 0 ? box[r - 1] : nullptr, r < n - 1 ? box[r + 1] : nullptr, 0);
  if (rc != FLOW2D_OK) {
    for (auto* hd : g->h) flow2d_destroy(hd);
    delete g;
    return rc;
  }
  *out = g;
  return FLOW2D_OK;
}

int flow2d_slab_group_destroy(flow2d_slab_group* g) {
  if (!g) return FLOW2D_OK;
  for (auto* hd : g->h) flow2d_destroy(hd);
  delete g;
  return FLOW2D_OK;
}

const char* flow2d_slab_group_last_error(const flow2d_slab_group* g) { return g ? g->err.c_str() : "null group"; }

int flow2d_slab_group_compute(flow2d_slab_group* g, const float* frame_0, const float* frame_1, float* flow_u, float* flow_v,
                              const flow2d_params* p, float* device_ms) {
  if (!g || !frame_0 || !frame_1 || !flow_u || !flow_v || !p) return FLOW2D_ERR_INVALID_ARGUMENT;
  const int n = (int)g->h.size();
  std::vector<int> rcs(n, FLOW2D_OK);
  std::vector<float> ms(n, 0.f);
  std::vector<std::thread> th;
  for (int r = 0; r < n; r++)
    th.emplace_back([&, r]() {
      flow2d_handle* h = g->h[r];
      auto run = [&]() -> int {
        CU_TRY(h, cudaSetDevice(h->device));
        const size_t row = h->W * sizeof(float), dpitch = h->pitch * sizeof(float);
        cudaStream_t st = h->stream;
        CU_TRY(h, cudaEventRecord(h->ev_start, st));
        // every rank needs both frames completely (the warp reads rows it cannot know in advance)
        CU_TRY(h, cudaMemcpy2DAsync(h->c[C_IN0], dpitch, frame_0, row, row, h->H, cudaMemcpyHostToDevice, st));
        CU_TRY(h, cudaMemcpy2DAsync(h->c[C_IN1], dpitch, frame_1, row, row, h->H, cudaMemcpyHostToDevice, st));
        reset_launch_counts(h);
        TRY(enqueue_pyramid(h, h->c[C_IN0], h->c[C_IN1], h->c[C_OUT_U], h->c[C_OUT_V], p, true));
        size_t y0 = 0, y1 = 0;
        flow2d_slab_rows(h, h->H, &y0, &y1);
        // ... and returns its own rows of the flow
        CU_TRY(h, cudaMemcpy2DAsync(flow_u + y0 * h->W, row, h->c[C_OUT_U] + y0 * h->pitch, dpitch, row, y1 - y0, cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaMemcpy2DAsync(flow_v + y0 * h->W, row, h->c[C_OUT_V] + y0 * h->pitch, dpitch, row, y1 - y0, cudaMemcpyDeviceToHost, st));
        CU_TRY(h, cudaEventRecord(h->ev_stop, st));
        TRY(flow2d_slab_status(h));
        (void)cudaEventElapsedTime(&ms[r], h->ev_start, h->ev_stop);
        return FLOW2D_OK;
      };
      rcs[r] = run();
    });
  for (auto& t : th) t.join();
  float worst = 0.f;
  for (int r = 0; r < n; r++) {
    worst = ms[r] > worst ? ms[r] : worst;
    if (rcs[r] != FLOW2D_OK) {
      g->err = "rank " + std::to_string(r) + ": " + g->h[r]->err;
      return rcs[r];
    }
  }
  if (device_ms) *device_ms = worst;
  return FLOW2D_OK;
}

// Debug aid (not part of the drop-in surface): solve_pass writes 8 globaltimer stamps per CTA of the
// LAST launch into this device buffer (null switches it off).
int flow2d_debug_timing(flow2d_handle* h, unsigned long long* d_stamps) {
  if (!h) return FLOW2D_ERR_INVALID_ARGUMENT;
  h->timing = d_stamps;
  return FLOW2D_OK;
}

// ---- per-stage API ----

int flow2d_stage_blur(flow2d_handle* h, const float* d_in, float* d_out, size_t w, size_t hh, float sigma) {
  STAGE_PROLOGUE(h);
  if (!d_in || !d_out || d_in == d_out) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "blur: bad buffers (in-place is refused)");
  TRY(check_level(h, w, hh));
  TRY(check_aligned(h, "blur", {d_in, d_out}));
  GaussTaps taps;
  TRY(gauss_taps(h, sigma, &taps));
  launch_blur(h->stream, d_in, d_out, (int)w, (int)hh, (int)h->pitch, taps);
  return check_launch(h, FLOW2D_K_BLUR, 1);
}

int flow2d_stage_resample(flow2d_handle* h, const float* d_in, size_t iw, size_t ih, float* d_out, size_t ow, size_t oh) {
  STAGE_PROLOGUE(h);
  if (!d_in || !d_out || d_in == d_out) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "resample: bad buffers (in-place is refused)");
  if (iw < 1 || ih < 1 || ow < 1 || oh < 1 || iw > h->W || ow > h->W || ih > h->H || oh > h->H)
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "resample: size outside the container");
  TRY(check_aligned(h, "resample", {d_in, d_out}));
  const float* in[2] = {d_in, d_in};
  float* tmp[2] = {h->c[C_TMP0], h->c[C_TMP0]};
  float* out[2] = {d_out, d_out};
  launch_resample(h->stream, in, tmp, out, 1, (int)iw, (int)ih, (int)ow, (int)oh, (int)h->pitch);
  return check_launch(h, FLOW2D_K_RESAMPLE, 2);
}

int flow2d_stage_warp(flow2d_handle* h, const float* d_frame_0, const float* d_frame_1, const float* d_flow_u,
                      const float* d_flow_v, float* d_out, size_t w, size_t hh, float hx, float hy) {
  STAGE_PROLOGUE(h);
  if (!d_frame_0 || !d_frame_1 || !d_flow_u || !d_flow_v || !d_out || d_out == d_frame_1 || d_out == d_frame_0)
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "warp: bad buffers (in-place is refused)");
  TRY(check_level(h, w, hh));
  TRY(check_aligned(h, "warp", {d_frame_0, d_frame_1, d_flow_u, d_flow_v, d_out}));
  launch_warp(h->stream, d_frame_0, d_frame_1, d_flow_u, d_flow_v, d_out, geom(h, w, hh, hx, hy));
  return check_launch(h, FLOW2D_K_WARP, 1);
}

int flow2d_stage_solve(flow2d_handle* h, const float* d_frame_0, const float* d_frame_1, const float* d_flow_u,
                       const float* d_flow_v, float* d_flow_du, float* d_flow_dv, float* d_phi, float* d_ksi, size_t w,
                       size_t hh, float hx, float hy, const flow2d_params* p) {
  STAGE_PROLOGUE(h);
  if (!d_frame_0 || !d_frame_1 || !d_flow_u || !d_flow_v || !d_flow_du || !d_flow_dv || !p)
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "solve: null buffer");
  if ((d_phi == nullptr) != (d_ksi == nullptr)) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "solve: phi and ksi go together");
  const void* ptrs[] = {d_frame_0, d_frame_1, d_flow_u, d_flow_v, d_flow_du, d_flow_dv, d_phi, d_ksi};
  for (const void* q : ptrs)
    if (!aligned16(q)) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "solve: containers must be 16-byte aligned");
  if (p->sweeps_per_pass < 0 || p->sweeps_per_pass > FLOW2D_MAX_SWEEPS_PER_PASS)
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "sweeps_per_pass must be 0..%d", FLOW2D_MAX_SWEEPS_PER_PASS);
  TRY(check_level(h, w, hh));
  int median = 1;
  flow2d_params q = *p;
  q.median_radius = 1; q.cascaded_restriction = 0;  // not used by this stage
  if (!(q.warp_scale_factor > 0.f && q.warp_scale_factor < 1.f)) q.warp_scale_factor = 0.9f;
  if (q.warp_levels_count < 1) q.warp_levels_count = 1;
  TRY(validate_params(h, &q, &median));
  TRY(ensure_ext(h, &q));
  if (early_exit(&q)) CU_TRY(h, cudaMemsetAsync(h->d_stop, 0, sizeof(int) * 2 * FLOW2D_MAX_LEVELS, h->stream));
  h->iter_levels = 1; h->iter_default = (int)q.outer_iterations_count;
  const LevelGeom g = geom(h, w, hh, hx, hy);
  const bool want_phi = d_phi != nullptr;
  return run_level_solve(h, g, d_frame_0, d_frame_1, d_flow_u, d_flow_v, d_flow_du, d_flow_dv, h->c[C_DU1], h->c[C_DV1],
                         want_phi ? d_phi : h->c[C_PHI], want_phi ? d_ksi : h->c[C_KSI], want_phi, &q,
                         plan_solve(h, g, &q, 1, false), 0, 0, 0);
}

int flow2d_stage_add(flow2d_handle* h, float* d_a, const float* d_b, size_t w, size_t hh) {
  STAGE_PROLOGUE(h);
  if (!d_a || !d_b) return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "add: null buffer");
  TRY(check_level(h, w, hh));
  TRY(check_aligned(h, "add", {d_a, d_b}));
  launch_add(h->stream, d_a, d_b, (int)w, (int)hh, (int)h->pitch);
  return check_launch(h, FLOW2D_K_ADD, 1);
}

int flow2d_stage_add_median(flow2d_handle* h, const float* d_a, const float* d_b, float* d_out, size_t w, size_t hh,
                            size_t radius) {
  STAGE_PROLOGUE(h);
  if (!d_a || !d_out || d_a == d_out || d_b == d_out)
    return fail(h, FLOW2D_ERR_INVALID_ARGUMENT, "median: bad buffers (in-place is refused)");
  TRY(check_level(h, w, hh));
  TRY(check_aligned(h, "median", {d_a, d_b, d_out}));
  int r = 1;
  TRY(normalise_median(h, radius, &r));
  const float* a[2] = {d_a, d_a};
  const float* b[2] = {d_b, d_b};
  float* out[2] = {d_out, d_out};
  launch_add_median(h->stream, a, d_b ? b : nullptr, out, 1, (int)w, (int)hh, (int)h->pitch, r);
  return check_launch(h, FLOW2D_K_ADD_MEDIAN, 1);
}

int flow2d_stage_median(flow2d_handle* h, const float* d_in, float* d_out, size_t w, size_t hh, size_t radius) {
  return flow2d_stage_add_median(h, d_in, nullptr, d_out, w, hh, radius);
}

}  // extern "C"

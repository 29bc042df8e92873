/*
 * flow2d.h -- C ABI of the B200-native 2D variational optical-flow solver.
 *
 * This is the drop-in boundary for the hot path of axruff/cuda-flow2d
 * (OpticalFlow2D::ComputeFlow and the six CudaOperation*2D operators under it).  Plain pointers
 * and sizes only; no C++ or torch types.  Every entry point names the reference interface it
 * replaces (paths relative to the reference repository).
 *
 * Conventions
 *   - All images are fp32, row-major.  HOST images are dense (row stride = width).  DEVICE images
 *     are "containers": the handle's pitch (flow2d_pitch_elems) floats per row, height rows;
 *     a pyramid level of size cw x ch occupies the top-left corner of a container, exactly like
 *     the reference's pitched containers (src/optical_flow/optical_flow_2d.cpp:84-140).
 *   - Every function returns FLOW2D_OK (0) or a negative flow2d_status; nothing throws or aborts.
 *     flow2d_last_error() gives the message of the last failure on that handle.
 *   - One handle = one device + one stream.  Handles are independent (one per GPU / host thread);
 *     a single handle is not thread-safe (same as the reference class).
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef FLOW2D_H_
#define FLOW2D_H_

#include <stddef.h>

#if defined(__GNUC__)
#define FLOW2D_API __attribute__((visibility("default")))
#else
#define FLOW2D_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct flow2d_handle flow2d_handle;

typedef enum flow2d_status {
  FLOW2D_OK = 0,
  FLOW2D_ERR_INVALID_ARGUMENT = -1, /* bad size / pointer / parameter (reference: printf + early return) */
  FLOW2D_ERR_CUDA = -2,             /* a CUDA runtime call failed (reference: CheckCudaError, cuda_utils.h:33) */
  FLOW2D_ERR_NO_DEVICE = -3,        /* no usable CUDA device (reference: main.cpp exit code 1) */
  FLOW2D_ERR_OUT_OF_MEMORY = -4,    /* container allocation failed (optical_flow_2d.cpp:112-137) */
  FLOW2D_ERR_UNSUPPORTED = -5       /* parameter outside what the path supports (e.g. median size 9) */
} flow2d_status;

/* src/data_types/data_structs.h:27 `enum class DataConstancy` */
typedef enum flow2d_constancy {
  FLOW2D_GREY = 0,     /* brightness constancy (the only value reachable from the reference CLI) */
  FLOW2D_GRADIENT = 1  /* gradient constancy, reproducing solve_2d_grad incl. its 16x8 tile artefact */
} flow2d_constancy;

/* The nine solver parameters of OpticalFlow2D::ComputeFlow, passed there by name through
 * OperationParameters (src/optical_flow/optical_flow_2d.cpp:160-168, src/main.cpp:192-201),
 * plus scheduling knobs that never change results. */
typedef struct flow2d_params {
  size_t warp_levels_count;      /* "warp_levels_count"      settings.xml Warping@levels      */
  float  warp_scale_factor;      /* "warp_scale_factor"      settings.xml Warping@scaling     */
  size_t outer_iterations_count; /* "outer_iterations_count" settings.xml Iterations@outer    */
  size_t inner_iterations_count; /* "inner_iterations_count" settings.xml Iterations@inner    */
  float  equation_alpha;         /* "equation_alpha"         settings.xml Model@alpha         */
  float  equation_smoothness;    /* "equation_smoothness"    settings.xml Model@e_smooth      */
  float  equation_data;          /* "equation_data"          settings.xml Model@e_data        */
  size_t median_radius;          /* "median_radius" (window diameter: 1 = off, 3, 5, 7; even values -1) */
  float  gaussian_sigma;         /* "gaussian_sigma"         settings.xml Model@sigma; <= 0 = off */
  /* scheduling (0 = automatic); results are identical for every value */
  int    sweeps_per_pass;        /* Jacobi sweeps fused into one solve_pass launch (1..FLOW2D_MAX_SWEEPS_PER_PASS) */
  int    resident_levels;        /* 0 auto (whole-level one-thread-per-pixel CTA for levels <= 1024 px, one-thread-per-pixel passes
                                    or 64x48 tiles above, chosen per level by a time model) / 2 = no one-thread-per-pixel kernels
                                    (levels <= 59x46 run in one resident solve_pass CTA) / -1 = always tiled */
  int    throughput_mode;        /* 0 (default) = mid-size levels use the one-thread-per-pixel pass, which trades redundant halo
                                    work for a 3x shorter dependent chain; 1 = they use 64x48 tiles / the resident CTA (least
                                    SM time).  Measured on B200 with the kernels of round 1, 0 is faster for one frame pair AND
                                    for 4-8 handles sharing the GPU; 1 is kept for many handles on a saturated GPU */
  int    report_residuals;       /* opt-in diagnostics (no reference counterpart, see flow2d_level_residuals): 1 = record
                                    the residual norm of every level's last linear system; results are unchanged */
  /* ---- opt-in EXTENSIONS beyond the reference.  All zero (flow2d_default_params) = the
   * reference's behaviour, bit for bit.  Anything else changes the result BY DESIGN and is specified by, and tested
   * against, oracle/flow2d_oracle_ext.c -- never against the reference, which has none of it:
   * fixed Jacobi counts without a relaxation factor (solve_2d.cu:361-374, cuda_operation_solve_2d.cpp:229-299), every
   * level restricted from the original frames (optical_flow_2d.cpp:279-305), derivative halos of the gradient / log
   * terms taken from the CUDA block (solve_2d.cu:391-669, 813-841), README.md:32-34's two data terms never combined. */
  int    scheme;                 /* FLOW2D_SCHEME_JACOBI (reference) or FLOW2D_SCHEME_RED_BLACK: Gauss-Seidel in red-black
                                    order (cells with even x+y first), in place */
  float  omega;                  /* relaxation factor: new = old + omega * (update - old); 0 or 1 = none.  With red-black
                                    ordering this is SOR (1 < omega < 2) */
  int    data_term;              /* FLOW2D_TERM_*: tensor data terms with true (mirrored) neighbour halos and the robust
                                    weight on the full tensor; needs a FLOW2D_GREY handle */
  float  gamma;                  /* FLOW2D_TERM_COMBINED: J = J_brightness + gamma * J_gradient */
  float  residual_tolerance;     /* > 0: convergence test -- a level stops iterating once both RMS residuals of its lagged
                                    system (see flow2d_level_residuals) are <= this; decided on the device, no host sync */
  int    residual_check_every;   /* test after every n-th outer iteration (<= 0: every one) */
  int    cascaded_restriction;   /* 1: level l of the two frame pyramids is restricted from level l-1 (once, before the
                                    level loop) instead of from the full-resolution frame */
  int    report_level_times;     /* opt-in diagnostics: 1 = time every level and its solve with CUDA events (the reference's
                                    per-level solve timer, cuda_operation_solve_2d.cpp:214-220, 302-311); the schedule is then
                                    enqueued kernel by kernel instead of replayed as a CUDA graph.  Results are unchanged */
} flow2d_params;

enum { FLOW2D_SCHEME_JACOBI = 0, FLOW2D_SCHEME_RED_BLACK = 1 };
enum {
  FLOW2D_TERM_DEFAULT = 0,       /* the handle's flow2d_constancy, as the reference computes it */
  FLOW2D_TERM_GRADIENT = 1,      /* gradient constancy, second derivatives over the true neighbours */
  FLOW2D_TERM_LOG_GRADIENT = 2,  /* the same on log(1 + f): DataConstancy::LogDerivatives done right */
  FLOW2D_TERM_COMBINED = 3       /* brightness + gamma * gradient constancy, jointly robustified */
};

#define FLOW2D_MAX_SWEEPS_PER_PASS 7
#define FLOW2D_MAX_LEVELS 256    /* residuals are recorded for at most this many levels (the coarsest ones) */

/* Fills *p with the reference's argv-form defaults (src/main.cpp:70-80). */
FLOW2D_API void flow2d_default_params(flow2d_params* p);

/* ---- lifetime: replaces OpticalFlow2D::Initialize / Destroy ------------------------------------
 * (src/optical_flow/optical_flow_2d.cpp:48-140, 571-590).  Allocates every device container,
 * stream and event for images of exactly width x height on CUDA device `device`. */
FLOW2D_API int flow2d_create(flow2d_handle** out, int device, size_t width, size_t height, int constancy);
FLOW2D_API int flow2d_destroy(flow2d_handle* h);
FLOW2D_API const char* flow2d_last_error(const flow2d_handle* h);

/* container geometry (reference: DataSize3.pitch / sizeof(float), cuMemAllocPitch) */
FLOW2D_API size_t flow2d_pitch_elems(const flow2d_handle* h);
FLOW2D_API size_t flow2d_width(const flow2d_handle* h);
FLOW2D_API size_t flow2d_height(const flow2d_handle* h);
/* The handle launches everything on this stream (a cudaStream_t).  NULL = handle-owned stream. */
FLOW2D_API int flow2d_set_stream(flow2d_handle* h, void* cuda_stream);
FLOW2D_API void* flow2d_get_stream(const flow2d_handle* h);

/* ---- the hot path: replaces OpticalFlow2D::ComputeFlow ------------------------------------------
 * (src/optical_flow/optical_flow_2d.cpp:142-569).
 * flow2d_compute: host in, host out (dense width*height floats each); blocks until u, v are
 * written, like the reference (H2D, solve, D2H).
 * flow2d_compute_device: frames and flow stay on the device as containers (pitch =
 * flow2d_pitch_elems); asynchronous on the handle's stream. */
FLOW2D_API int flow2d_compute(flow2d_handle* h, const float* frame_0, const float* frame_1,
                   float* flow_u, float* flow_v, const flow2d_params* p);
FLOW2D_API int flow2d_compute_device(flow2d_handle* h, const float* d_frame_0, const float* d_frame_1,
                          float* d_flow_u, float* d_flow_v, const flow2d_params* p);

/* Asynchronous form of flow2d_compute for pipelines that keep several handles busy at once (one
 * stream each): enqueues H2D, the solve and D2H and returns.  The host buffers must be page-locked
 * (flow2d_host_alloc) and stay untouched until flow2d_synchronize().  The two frames are uploaded
 * on an internal copy stream into a landing pair, starting at once -- their contents must be final
 * when the call is made -- so that the upload of call n+1 overlaps the solve of call n on the same
 * handle; the solve and the D2H are ordered on the handle's stream. */
FLOW2D_API int flow2d_compute_async(flow2d_handle* h, const float* frame_0, const float* frame_1,
                         float* flow_u, float* flow_v, const flow2d_params* p);
FLOW2D_API int flow2d_synchronize(flow2d_handle* h);
/* Optional: captures the schedule that flow2d_compute / flow2d_compute_async will replay for these parameters now (tens of
 * milliseconds of host time for a full pyramid) instead of inside the first call.  Handles are independent: a program that
 * owns many of them (K per GPU on N GPUs) prepares them from several host threads at once.  No reference counterpart (the
 * reference JIT-compiles its PTX in Initialize, optical_flow_2d.cpp:59-82). */
FLOW2D_API int flow2d_prepare(flow2d_handle* h, const flow2d_params* p);

/* Diagnostics of the last flow2d_compute*(): kernels launched, pyramid levels run, and (after
 * flow2d_compute only) the device time in ms between the first H2D and the last D2H -- the span
 * of the reference's "Total GPU computation time" (optical_flow_2d.cpp:179,548-554). */
FLOW2D_API int flow2d_last_stats(const flow2d_handle* h, long long* kernel_launches, int* levels_run, float* device_ms);

/* The same launch count split by kernel (index = FLOW2D_K_*), e.g. to name the dominant kernel of a
 * workload.  No reference counterpart (the reference launches blindly, cuda_operation_*.cpp). */
enum {
  FLOW2D_K_BLUR = 0, FLOW2D_K_RESAMPLE, FLOW2D_K_WARP, FLOW2D_K_DERIVATIVES, FLOW2D_K_GRAD_TENSOR, FLOW2D_K_SOLVE_PASS,
  FLOW2D_K_SOLVE_RESIDENT, FLOW2D_K_SOLVE_SMALL_PASS, FLOW2D_K_SOLVE_TINY, FLOW2D_K_ADD_MEDIAN, FLOW2D_K_ADD,
  FLOW2D_K_RESIDUAL,
  FLOW2D_K_EXT,  /* the kernels of the opt-in extensions (csrc/solve_ext.cu) */
  FLOW2D_K_SOLVE_CLUSTER,  /* a mid-size level solved on one thread-block cluster (csrc/solve_cluster.cu) */
  FLOW2D_KERNEL_KINDS
};
FLOW2D_API int flow2d_last_launch_counts(const flow2d_handle* h, long long* counts /* [FLOW2D_KERNEL_KINDS] */);
FLOW2D_API const char* flow2d_kernel_kind_name(int kind);

/* ---- scheduler introspection (used by the CPU tests; no reference counterpart) ------------------------------------
 * Mid-size levels are solved on one thread-block cluster (csrc/solve_cluster.cu): cx * cy CTAs with a tw x th block of
 * the level each, one thread per pixel.  flow2d_cluster_shape: the decomposition the scheduler picks for a region of
 * rw x rh cells when every cluster shape is launchable, shape = {cx, cy, tw, th, threads per CTA (whole warps)}; `compact` != 0 weighs
 * occupied SMs over sweep latency (throughput_mode).  FLOW2D_ERR_UNSUPPORTED = the region does not fit a cluster.
 * flow2d_debug_cluster_cell: what one thread of that kernel works on, computed by the very function the kernel calls
 * (csrc/solve_cluster_geom.h) -- geom = {cx, cy, tw, th, regions per row}, level = {w, h, ow, oh, halo, y0, y1},
 * cell = {own, left, right, up, down offsets in a shared plane, push rank / offset (horizontal), push rank / offset
 * (vertical), gx, gy, has a cell, live, output, floats per shared plane}. */
FLOW2D_API int flow2d_cluster_shape(int rw, int rh, int compact, int shape[5]);
/* Handles of this process alive on `device`.  With FLOW2D_CLUSTER unset the scheduler reads it when a schedule is made
 * (first call with a parameter set, or flow2d_prepare): from 4 handles on -- several frame pairs in flight on one GPU --
 * every level of 1 025 .. 16 384 px runs on a cluster (least SM time: C4 batch +1.7 %, C1b batch +4.5 % end to end); a lone
 * handle keeps one launch per outer iteration there (least latency).  Results are identical either way. */
FLOW2D_API int flow2d_live_handles(int device);
FLOW2D_API int flow2d_debug_cluster_cell(const int geom[5], const int level[7], int rank, int cluster, int thread, int cell[15]);

/* The level schedule of one (containers, parameters) combination is captured into a CUDA graph on first use and replayed
 * afterwards; a handle keeps the 8 most recently used graphs, so a caller that rotates a few containers (a frame ring)
 * never re-captures.  captures / replays count both since flow2d_create.  No reference counterpart. */
FLOW2D_API int flow2d_graph_stats(const flow2d_handle* h, long long* captures, long long* replays);

/* ---- opt-in convergence diagnostics -----------------------------------------
 * The reference runs fixed iteration counts and computes no norm (cuda_operation_solve_2d.cpp:229-299).
 * With flow2d_params.report_residuals = 1 every level additionally records the RMS residual of the
 * lagged linear system of its LAST outer iteration, after that iteration's inner sweeps:
 *   r_u = ksi*(-J13 - J12*dv - J11*du) + sum_n a_n*((u_n+du_n) - (u+du)),   r_v likewise
 * (the quantities of solve_2d.cu:333-374; zero at the fixed point of the Jacobi update).  One extra
 * kernel per level (warp-shuffle + atomic reduction in double precision); flow results are unchanged.
 * flow2d_level_residuals waits for the handle's stream and returns the values of the last
 * flow2d_compute*() call, coarsest level first.  Not recorded in the row-slab mode. */
FLOW2D_API int flow2d_level_residuals(flow2d_handle* h, double* rms_u, double* rms_v, int capacity, int* levels);
/* The same number for one level on caller containers (phi, ksi as returned by flow2d_stage_solve). Synchronous. */
FLOW2D_API int flow2d_stage_residual(flow2d_handle* h, const float* d_frame_0, const float* d_frame_1_warped,
                          const float* d_u, const float* d_v, const float* d_du, const float* d_dv,
                          const float* d_phi, const float* d_ksi, size_t w, size_t hh, float hx, float hy,
                          const flow2d_params* p, double* rms_u, double* rms_v);

/* With flow2d_params.report_level_times = 1: device time in ms of every level of the last flow2d_compute*() call (from its
 * first kernel to the first kernel of the next level) and of its solve alone, coarsest level first.  Replaces the per-level
 * "solve" timer of the reference (cuda_operation_solve_2d.cpp:214-220, 302-311, printed when !silent).  Waits for the stream. */
FLOW2D_API int flow2d_level_times(flow2d_handle* h, float* level_ms, float* solve_ms, int capacity, int* levels);

/* Outer iterations that ran per level in the last flow2d_compute*() call, coarsest level first (equal to
 * outer_iterations_count unless residual_tolerance ended a level early).  Waits for the handle's stream. */
FLOW2D_API int flow2d_level_outer_iterations(flow2d_handle* h, int* iterations, int capacity, int* levels);

/* ---- level table: replaces OpticalFlowBase2D::GetMaxWarpLevel and the per-level size formulas ---
 * (src/optical_flow/optical_flow_base_2d.cpp:36-59, src/optical_flow/optical_flow_2d.cpp:268-272).
 * Host-only integer/fp32 arithmetic, bit-exact with the reference. */
FLOW2D_API size_t flow2d_max_warp_level(size_t width, size_t height, float scale_factor);
FLOW2D_API int flow2d_level_geometry(size_t width, size_t height, float scale_factor, int level,
                          size_t* cw, size_t* ch, float* hx, float* hy);

/* ---- per-stage API on device containers -----------------------------------------------------------
 * One call per reference operator, so each kernel can be checked against the reference's on the same
 * buffers.  All pointers are device containers of this handle's geometry.  Asynchronous on the
 * handle's stream.  In-place use (input == output) is refused like in the reference, except for
 * flow2d_stage_add. */

/* CudaOperationConvolution2D::Execute (cuda_operation_convolution_2d.cpp:132-176): zero-padded
 * separable Gaussian, radius = (size_t)(3*sigma) <= 16. */
FLOW2D_API int flow2d_stage_blur(flow2d_handle* h, const float* d_in, float* d_out, size_t w, size_t h_, float sigma);
/* CudaOperationResample2D::Execute (cuda_operation_resample_2d.cpp:76-106): area resampling
 * (iw x ih) -> (ow x oh), x pass then y pass. */
FLOW2D_API int flow2d_stage_resample(flow2d_handle* h, const float* d_in, size_t iw, size_t ih,
                          float* d_out, size_t ow, size_t oh);
/* CudaOperationRegistration2D::Execute (cuda_operation_registration_2d.cpp:74-128). */
FLOW2D_API int flow2d_stage_warp(flow2d_handle* h, const float* d_frame_0, const float* d_frame_1,
                      const float* d_flow_u, const float* d_flow_v, float* d_out,
                      size_t w, size_t h_, float hx, float hy);
/* CudaOperationSolve2D::Execute (cuda_operation_solve_2d.cpp:106-315): zero du/dv, then
 * outer x (phi/ksi + inner x Jacobi sweep).  d_frame_1 is the WARPED second frame.  Results in
 * d_flow_du, d_flow_dv; d_phi / d_ksi (optional, may be NULL) receive the last outer iteration's
 * robust weights.  Uses p->outer/inner/alpha/smoothness/data and the scheduling knobs. */
FLOW2D_API int flow2d_stage_solve(flow2d_handle* h, const float* d_frame_0, const float* d_frame_1,
                       const float* d_flow_u, const float* d_flow_v,
                       float* d_flow_du, float* d_flow_dv, float* d_phi, float* d_ksi,
                       size_t w, size_t h_, float hx, float hy, const flow2d_params* p);
/* CudaOperationAdd2D::Execute (cuda_operation_add_2d.cpp:75-106): d_a += d_b. */
FLOW2D_API int flow2d_stage_add(flow2d_handle* h, float* d_a, const float* d_b, size_t w, size_t h_);
/* CudaOperationMedian2D::Execute (cuda_operation_median_2d.cpp:77-155): radius 1 = copy, even
 * values are decremented, 3/5/7 filter with mirrored borders; anything else is
 * FLOW2D_ERR_UNSUPPORTED (the reference prints an error and leaves d_out untouched). */
FLOW2D_API int flow2d_stage_median(flow2d_handle* h, const float* d_in, float* d_out, size_t w, size_t h_, size_t radius);
/* Fused 2 x add + 2 x median of one level (optical_flow_2d.cpp:409-449): out = median(a + b). */
FLOW2D_API int flow2d_stage_add_median(flow2d_handle* h, const float* d_a, const float* d_b, float* d_out,
                            size_t w, size_t h_, size_t radius);

/* ---- one large frame on several GPUs: row-slab decomposition -------------------------------------------
 * (no counterpart upstream: the reference is single-GPU, cuda_operation_solve_2d.cpp:168; SURVEY.md 8e,
 * BASELINE.json configs[4]).  One handle per GPU (one process each, or several in one process); every rank holds
 * both full frames.  Pyramid levels with at least min_rows_per_rank rows per rank are slabbed: a rank then computes
 * its own rows of EVERY stage of the level (prolongation, warp, derivatives, robust iterations, add + median) plus
 * the few rows of margin the next stage reads, and only rows of neighbours ever move:
 *   - ghost rows of the increment (du, dv), refreshed from the two neighbour ranks every few outer iterations (a solve
 *     pass with S sweeps is exact S+1 rows less far from a cut edge than its input was), and
 *   - once per level, the rows of the new flow (u, v) that the neighbours' prolongation to the next level reads.
 * The smaller levels and the restriction of the two frames are computed by every rank.  The Jacobi scheme makes the
 * result bit-identical to a single GPU.
 * Transport: each handle owns a MAILBOX in device memory; the neighbours write rows straight into it over NVLink
 * (peer-mapped stores from a copy kernel) followed by a system-scope flag, and the receiver's stream waits for the
 * flag in a kernel.  No host synchronisation, no collective library on the data path.  Wiring:
 *   same process:   flow2d_slab_mailbox() of the neighbours' handles (the library enables peer access on connect)
 *   one process per GPU: flow2d_slab_export() -> 64 bytes to the neighbours (any channel) -> flow2d_slab_import()
 * then flow2d_slab_connect() on every rank, then flow2d_compute_slab_device() on every rank (each on its own host
 * thread / process: a rank's stream waits for its neighbours' kernels).  On return (after the stream has drained) rows
 * [y0, y1) = flow2d_slab_rows(handle, height) of d_flow_u / d_flow_v hold this rank's part of the flow. */
FLOW2D_API int flow2d_slab_mailbox(flow2d_handle* h, void** d_mailbox, size_t* bytes);
FLOW2D_API int flow2d_slab_export(flow2d_handle* h, unsigned char ipc_handle[64]);
FLOW2D_API int flow2d_slab_import(flow2d_handle* h, const unsigned char ipc_handle[64], void** d_mailbox);
/* mailbox_above = mailbox of rank-1 (NULL for rank 0), mailbox_below = mailbox of rank+1 (NULL for the last rank), both
 * as addresses valid on this handle's device.  min_rows_per_rank: 0 = default (64; a level also needs twice the ghost rows per rank).  world = 1 switches slabbing off. */
FLOW2D_API int flow2d_slab_connect(flow2d_handle* h, int rank, int world, void* mailbox_above, void* mailbox_below,
                        size_t min_rows_per_rank);
FLOW2D_API int flow2d_slab_rows(const flow2d_handle* h, size_t level_height, size_t* y0, size_t* y1);
/* exchanges and payload bytes sent by this rank since flow2d_slab_connect; levels slabbed in the last compute */
FLOW2D_API int flow2d_slab_stats(const flow2d_handle* h, long long* exchanges, long long* bytes_sent, int* levels_slabbed);
FLOW2D_API int flow2d_compute_slab_device(flow2d_handle* h, const float* d_frame_0, const float* d_frame_1,
                               float* d_flow_u, float* d_flow_v, const flow2d_params* p);
/* Waits for the handle's stream; FLOW2D_ERR_CUDA if a wait for a neighbour's rows gave up after 20 s (a rank that never
 * sent: the results of this and the neighbouring ranks are then invalid and every rank must connect again).  Host code
 * that drives several ranks from one process must not synchronise the whole DEVICE while ranks are in flight: a rank's
 * stream waits for kernels its neighbours may not have enqueued yet. */
FLOW2D_API int flow2d_slab_status(flow2d_handle* h);
/* flow2d_stage_solve on a slabbed level (all inputs complete on every rank; the result is exact on the rank's own rows
 * of the level +- median_radius / 2).  *slabbed tells whether the level was large enough to be slabbed. */
FLOW2D_API int flow2d_stage_solve_slab(flow2d_handle* h, const float* d_frame_0, const float* d_frame_1,
                            const float* d_flow_u, const float* d_flow_v, float* d_flow_du, float* d_flow_dv,
                            size_t w, size_t h_, float hx, float hy, const flow2d_params* p, int* slabbed);

/* The same for callers that drive all GPUs from ONE process (the command line: cuda-flow2d --slab): N handles on the
 * listed devices, wired mailbox to mailbox with peer access, one host thread per rank inside flow2d_slab_group_compute.
 * Host images in (dense width*height floats, uploaded to every GPU), host flow out (every rank downloads its own rows);
 * blocks until the flow is complete.  *device_ms (optional): slowest rank, first upload to last download. */
typedef struct flow2d_slab_group flow2d_slab_group;
FLOW2D_API int flow2d_slab_group_create(flow2d_slab_group** out, const int* devices, int n, size_t width, size_t height, int constancy);
FLOW2D_API int flow2d_slab_group_compute(flow2d_slab_group* g, const float* frame_0, const float* frame_1, float* flow_u,
                              float* flow_v, const flow2d_params* p, float* device_ms);
FLOW2D_API int flow2d_slab_group_destroy(flow2d_slab_group* g);
FLOW2D_API const char* flow2d_slab_group_last_error(const flow2d_slab_group* g);

/* Debug aid, not part of the drop-in surface: every solve_pass CTA writes 8 %globaltimer stamps
 * (entry, loads+tensor, phi, weights, sweeps, stores) into d_stamps (device memory, >= 8 * CTAs of the
 * largest launch); NULL switches it off. */
FLOW2D_API int flow2d_debug_timing(flow2d_handle* h, unsigned long long* d_stamps);

/* Page-locked host memory for frames / flow fields, so the H2D / D2H copies of flow2d_compute are
 * plain DMA (the reference's optional ALLOCATE_PINNED_MEMORY path, src/data_types/data2d.cpp:52-60).
 * flow2d_host_alloc returns NULL when there is no device or the allocation fails. */
FLOW2D_API void* flow2d_host_alloc(size_t bytes);
FLOW2D_API void flow2d_host_free(void* p);

/* Library identification: "flow2d-b200 <version> sm_100a". */
FLOW2D_API const char* flow2d_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FLOW2D_H_ */

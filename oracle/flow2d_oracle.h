/*
 * flow2d_oracle.h -- CPU oracle for the cuda-flow2d hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C, fp32, literal restatement of the reference's ComputeFlow path
 * (axruff/cuda-flow2d).  It exists to check the sm_100a kernels; it is never linked
 * into, imported by or called from the product library.  Only tests/, the smoke check
 * in __graft_entry__.py and bench.py's cpu_baseline leg may use it.
 *
 * Parity status: PINNED.  The reference ships no tests, golden vectors or fixtures for this path
 * (SURVEY.md F6), so this restatement is pinned against the reference's OWN CUDA build
 * (oracle/_ref, built by oracle/Makefile from the sources under /root/reference and run on a
 * B200): operator by operator through the reference's CudaOperation*2D::Execute on identical
 * buffers (oracle/ref_harness.cpp) and end to end on the bundled rub pair with both reference
 * parameter sets.  Result: bit-exact (tests/test_golden_cpu.py on the committed outputs in
 * tests/golden/, tests/test_reference_gpu.py live on the GPU box).
 *
 * Every function cites the reference file:line it follows.  Floating-point expression
 * trees (which mul/add pairs are fused) follow the reference kernels as compiled by
 * nvcc 12.9 -ptx + ptxas -arch=sm_100 (see DESIGN.md "Arithmetic contract").
 */
#ifndef FLOW2D_ORACLE_H_
#define FLOW2D_ORACLE_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_GREY = 0, ORACLE_GRADIENT = 1 };

typedef struct oracle_params {
  size_t warp_levels_count;      /* main.cpp:70 */
  float  warp_scale_factor;      /* main.cpp:71 */
  size_t outer_iterations_count; /* main.cpp:72 */
  size_t inner_iterations_count; /* main.cpp:73 */
  float  equation_alpha;         /* main.cpp:74 */
  float  equation_smoothness;    /* main.cpp:75 */
  float  equation_data;          /* main.cpp:76 */
  size_t median_radius;          /* main.cpp:77 (really the window diameter) */
  float  gaussian_sigma;         /* main.cpp:78 */
  int    constancy;              /* data_structs.h:27; ORACLE_GREY | ORACLE_GRADIENT */
} oracle_params;

/* optical_flow_base_2d.cpp:36-59 */
size_t oracle_max_warp_level(size_t width, size_t height, float scale_factor);
/* optical_flow_2d.cpp:268-272 */
void oracle_level_geometry(size_t W, size_t H, float scale_factor, int level,
                           size_t* cw, size_t* ch, float* hx, float* hy);

/* cuda_operation_convolution_2d.cpp:83-112; taps must hold 2*radius+1 floats (<= 51). Returns radius. */
int oracle_gauss_taps(float sigma, float* taps);
/* convolution_2d.cu:74-261 (zero padding, rows then columns) */
void oracle_blur(const float* in, float* out, size_t w, size_t h, size_t pitch, float sigma);

/* resample_2d.cu:34-118 + cuda_operation_resample_2d.cpp:99-105 (x pass then y pass) */
void oracle_resample(const float* in, size_t iw, size_t ih, float* out, size_t ow, size_t oh, size_t pitch);
/* index part of resample_2d.cu:44-51, exposed for bit-exact index tests */
void oracle_resample_cells(size_t in_n, size_t out_n, size_t x, int* left_i, int* right_i);

/* registration_2d.cu:34-74 */
void oracle_warp(const float* f0, const float* f1, const float* u, const float* v,
                 size_t w, size_t h, size_t pitch, float hx, float hy, float* out);

/* solve_2d.cu:43-198 */
void oracle_phi_ksi(const float* f0, const float* f1, const float* u, const float* v,
                    const float* du, const float* dv, size_t w, size_t h, size_t pitch,
                    float hx, float hy, float e_smooth, float e_data, float* phi, float* ksi);
/* solve_2d.cu:200-377 (one Jacobi sweep, Grey) */
void oracle_sweep_grey(const float* f0, const float* f1, const float* u, const float* v,
                       const float* du, const float* dv, const float* phi, const float* ksi,
                       size_t w, size_t h, size_t pitch, float hx, float hy, float alpha,
                       float* du_out, float* dv_out);
/* solve_2d.cu:683-953 (one Jacobi sweep, Gradient; 16x8 tile replicate artefact emulated, F5) */
void oracle_sweep_grad(const float* f0, const float* f1, const float* u, const float* v,
                       const float* du, const float* dv, const float* phi, const float* ksi,
                       size_t w, size_t h, size_t pitch, float hx, float hy, float alpha,
                       float* du_out, float* dv_out);
/* cuda_operation_solve_2d.cpp:106-315: memset du,dv; outer x (phi/ksi + inner x sweep + swap).
 * du/dv/tmp_du/tmp_dv are caller buffers; the result is left in du, dv (pointer swaps are
 * resolved by copying). */
void oracle_solve_level(const float* f0, const float* f1, const float* u, const float* v,
                        float* du, float* dv, float* phi, float* ksi, float* tmp_du, float* tmp_dv,
                        size_t w, size_t h, size_t pitch, float hx, float hy,
                        const oracle_params* p);

/* EXTENSION beyond the reference (which computes no norm): RMS residual of the lagged linear system of
 * jacobi_update (solve_2d.cu:333-374) for given phi, ksi and increment; weights in fp32 as the sweep forms
 * them, everything else in double.  Checks flow2d_stage_residual / flow2d_level_residuals. */
void oracle_residual(const float* f0, const float* f1, const float* u, const float* v, const float* du,
                     const float* dv, const float* phi, const float* ksi, size_t w, size_t h, size_t pitch,
                     float hx, float hy, float alpha, int constancy, double* rms_u, double* rms_v);

/* add_2d.cu:33-46 */
void oracle_add(float* a, const float* b, size_t w, size_t h, size_t pitch);
/* median_2d.cu:87-299 + cuda_operation_median_2d.cpp:100-111. Returns 0 if a filter/copy ran,
 * 1 if the radius is unsupported (the reference then leaves `out` untouched). */
int oracle_median(const float* in, float* out, size_t w, size_t h, size_t pitch, size_t radius);

/* optical_flow_2d.cpp:142-569: the whole path. f0,f1,u,v are dense W*H row-major. Returns 0. */
int oracle_compute_flow(const float* f0, const float* f1, size_t W, size_t H,
                        const oracle_params* p, float* u, float* v);

/* ---- EXTENSIONS beyond the reference (flow2d_oracle_ext.c): a specification for the opt-in features of SURVEY.md 8(f)
 * ranks 3-4, never reference parity.  Checks flow2d_params.{scheme, omega, data_term, gamma, residual_tolerance,
 * residual_check_every, cascaded_restriction}. */
enum { ORACLE_SCHEME_JACOBI = 0, ORACLE_SCHEME_RED_BLACK = 1 };
enum { ORACLE_TERM_DEFAULT = 0, ORACLE_TERM_GRADIENT = 1, ORACLE_TERM_LOG_GRADIENT = 2, ORACLE_TERM_COMBINED = 3 };
typedef struct oracle_ext {
  int   scheme;                /* ORACLE_SCHEME_* */
  float omega;                 /* relaxation factor; 0 or 1 = none */
  int   data_term;             /* ORACLE_TERM_* */
  float gamma;                 /* weight of the gradient tensor in ORACLE_TERM_COMBINED */
  float residual_tolerance;    /* > 0: a level ends when both RMS residuals are <= this */
  int   residual_check_every;  /* test after every n-th outer iteration (<= 0: 1) */
  int   cascaded_restriction;  /* 1: level l of the frame pyramid is restricted from level l-1 */
} oracle_ext;
void oracle_ext_tensor(const float* f0, const float* f1w, size_t w, size_t h, size_t pitch, float hx, float hy,
                       int data_term, float gamma, float* const J[6]);
void oracle_ext_phi_ksi(const float* const J[6], const float* u, const float* v, const float* du, const float* dv,
                        size_t w, size_t h, size_t pitch, float hx, float hy, float e_smooth, float e_data, float* phi,
                        float* ksi);
void oracle_ext_residual(const float* const J[6], const float* u, const float* v, const float* du, const float* dv,
                         const float* phi, const float* ksi, size_t w, size_t h, size_t pitch, float hx, float hy,
                         float alpha, double* rms_u, double* rms_v);
/* returns the outer iterations that ran; du, dv, phi, ksi are pitch*h caller buffers */
int oracle_ext_solve_level(const float* f0, const float* f1w, const float* u, const float* v, float* du, float* dv,
                           float* phi, float* ksi, size_t w, size_t h, size_t pitch, float hx, float hy,
                           const oracle_params* p, const oracle_ext* e);
/* returns the levels run; outer_used[cap] (optional): outer iterations per level, coarsest first */
int oracle_ext_compute_flow(const float* f0, const float* f1, size_t W, size_t H, const oracle_params* p,
                            const oracle_ext* e, float* u, float* v, int* outer_used, int cap);

/* number of OpenMP threads the oracle will use (1 if built without OpenMP) */
int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif

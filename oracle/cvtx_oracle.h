#ifndef CVTX_ORACLE_H
#define CVTX_ORACLE_H
/*
 * cvtx_oracle.h -- CPU ORACLE for the cvortex all-pairs (M2M) hot path.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  It is never
 * linked into, imported by or called from the product (libcvortex.so /
 * cvortex_b200/); the product fails loudly when its CUDA path is unavailable.
 *
 * Two restatements of the reference's OpenMP CPU algorithm, each function
 * citing the reference file:line it follows:
 *   *_f32 : same arithmetic as the reference -- FP32 pair maths in the
 *           reference's operation order, one FP64 accumulator per output
 *           component, final cast and scale as the reference does.
 *   *_f64 : the same formulas (incl. the Abramowitz-Stegun erf polynomial and
 *           the coincident-pair rule) evaluated entirely in FP64; arbitrates
 *           between the FP32 reference and the FP32 GPU kernels.
 *
 * PARITY IS PINNED: tests/test_oracle.py checks this file against (a) the 24
 * regularisation known-answer values of reference test/testvortfunc.h:37-67,
 * (b) the 31 structural S2S checks of reference test/testparticle.h:48-109 and
 * (c) outputs of the reference's own CPU path compiled here from
 * /root/reference (oracle/_ref, see oracle/Makefile) -- live when that .so is
 * present, and through fixtures under tests/golden/ generated from it by
 * tests/golden/make_golden.py.
 *
 * Array conventions (flat, row-major, float32 in / float32 or float64 out):
 *   P3D particle row : x y z wx wy wz vol         (7 floats, = cvtx_P3D)
 *   P2D particle row : x y gamma area             (4 floats, = cvtx_P2D)
 *   F3D filament row : ax ay az bx by bz gamma    (7 floats, = cvtx_F3D)
 *   points           : 3 (or 2) floats per row
 */
#ifdef __cplusplus
extern "C" {
#endif

/* Regularisation ids; the order is the one the reference's tests use
 * (reference src/VortFunc.cpp:201-251). */
enum {
	CVTX_ORACLE_SINGULAR = 0,
	CVTX_ORACLE_WINCKELMANS = 1,
	CVTX_ORACLE_PLANETARY = 2,
	CVTX_ORACLE_GAUSSIAN = 3
};

/* Regularisation scalars, FP32 as the reference computes them. */
float cvtx_oracle_g3d(int reg, float rho);
float cvtx_oracle_zeta3d(int reg, float rho);
float cvtx_oracle_eta3d(int reg, float rho);
float cvtx_oracle_g2d(int reg, float rho);
float cvtx_oracle_eta2d(int reg, float rho);
/* FP64 versions of the same formulas. */
double cvtx_oracle_g3d_f64(int reg, double rho);
double cvtx_oracle_zeta3d_f64(int reg, double rho);
double cvtx_oracle_eta3d_f64(int reg, double rho);
double cvtx_oracle_g2d_f64(int reg, double rho);
double cvtx_oracle_eta2d_f64(int reg, double rho);

/* Single-pair functions (reference S2S), FP32; out = 3 (or 2 / 1) floats. */
void cvtx_oracle_P3D_S2S_vel(const float *src7, const float *pt3, int reg, float sigma, float *out3);
void cvtx_oracle_P3D_S2S_dvort(const float *src7, const float *tgt7, int reg, float sigma, float *out3);
void cvtx_oracle_P3D_S2S_visc_dvort(const float *src7, const float *tgt7, int reg, float sigma, float nu, float *out3);
void cvtx_oracle_P2D_S2S_vel(const float *src4, const float *pt2, int reg, float sigma, float *out2);
void cvtx_oracle_P2D_S2S_visc_dvort(const float *src4, const float *tgt4, int reg, float sigma, float nu, float *out1);
void cvtx_oracle_F3D_S2S_vel(const float *fil7, const float *pt3, float *out3);
void cvtx_oracle_F3D_S2S_dvort(const float *fil7, const float *tgt7, float *out3);

/* M2M, reference arithmetic (FP32 pairs, FP64 accumulation). */
void cvtx_oracle_P3D_M2M_vel_f32(const float *src7, int n, const float *pts3, int m, float *out3, int reg, float sigma);
void cvtx_oracle_P3D_M2M_dvort_f32(const float *src7, int n, const float *tgt7, int m, float *out3, int reg, float sigma);
void cvtx_oracle_P3D_M2M_visc_dvort_f32(const float *src7, int n, const float *tgt7, int m, float *out3, int reg, float sigma, float nu);
void cvtx_oracle_P3D_M2M_vort_f32(const float *src7, int n, const float *pts3, int m, float *out3, int reg, float sigma);
void cvtx_oracle_P2D_M2M_vel_f32(const float *src4, int n, const float *pts2, int m, float *out2, int reg, float sigma);
void cvtx_oracle_P2D_M2M_visc_dvort_f32(const float *src4, int n, const float *tgt4, int m, float *out1, int reg, float sigma, float nu);
void cvtx_oracle_F3D_M2M_vel_f32(const float *fil7, int n, const float *pts3, int m, float *out3);
void cvtx_oracle_F3D_M2M_dvort_f32(const float *fil7, int n, const float *tgt7, int m, float *out3);

/* Dense filament influence matrix (reference cvtx_F3D_inf_mtrx), out[i * n + j]. */
void cvtx_oracle_F3D_inf_mtrx_f32(const float *fil7, int n, const float *pts3, const float *dirs3, int m, float *out);
void cvtx_oracle_F3D_inf_mtrx_f64(const float *fil7, int n, const float *pts3, const float *dirs3, int m, double *out);

/* M2M, all-FP64 evaluation of the same formulas; outputs are doubles. */
void cvtx_oracle_P3D_M2M_vel_f64(const float *src7, int n, const float *pts3, int m, double *out3, int reg, float sigma);
void cvtx_oracle_P3D_M2M_dvort_f64(const float *src7, int n, const float *tgt7, int m, double *out3, int reg, float sigma);
void cvtx_oracle_P3D_M2M_visc_dvort_f64(const float *src7, int n, const float *tgt7, int m, double *out3, int reg, float sigma, float nu);
void cvtx_oracle_P3D_M2M_vort_f64(const float *src7, int n, const float *pts3, int m, double *out3, int reg, float sigma);
void cvtx_oracle_P2D_M2M_vel_f64(const float *src4, int n, const float *pts2, int m, double *out2, int reg, float sigma);
void cvtx_oracle_P2D_M2M_visc_dvort_f64(const float *src4, int n, const float *tgt4, int m, double *out1, int reg, float sigma, float nu);
void cvtx_oracle_F3D_M2M_vel_f64(const float *fil7, int n, const float *pts3, int m, double *out3);
void cvtx_oracle_F3D_M2M_dvort_f64(const float *fil7, int n, const float *tgt7, int m, double *out3);

/* ---- redistribution onto a grid and relaxation (cvtx_oracle_remesh.c) ----
 * The steps either side of the all-pairs sums in a time step.  Sequential, reference
 * arithmetic; pinned against the reference's own implementation (see that file). */
enum {
	CVTX_ORACLE_LAMBDA0 = 0,
	CVTX_ORACLE_LAMBDA1 = 1,
	CVTX_ORACLE_LAMBDA2 = 2,
	CVTX_ORACLE_LAMBDA3 = 3,
	CVTX_ORACLE_M4P = 4
};
float cvtx_oracle_redist(int which, float U);          /* reference src/RedistFunc.cpp:36-96 */
float cvtx_oracle_redist_radius(int which);
/* Returns the number of particles created; writes them to out (capacity max_out rows) when
 * have_out is non-zero (reference passes output_particles == NULL to ask for the count). */
int cvtx_oracle_P3D_redistribute(const float *src7, int n, float *out7, int max_out, int have_out, int which, float h, float negligible);
int cvtx_oracle_P2D_redistribute(const float *src4, int n, float *out4, int max_out, int have_out, int which, float h, float negligible);
/* out7 = src7 with relaxed vorticity (reference cvtx_P3D_pedrizzetti_relaxation). */
void cvtx_oracle_P3D_pedrizzetti(const float *src7, int n, float fdt, int reg, float sigma, float *out7);

/* Number of OpenMP threads the M2M loops will use. */
int cvtx_oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif

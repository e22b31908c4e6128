/*
 * cvtx_oracle.c -- CPU ORACLE for the cvortex all-pairs hot path.
 * TEST INFRASTRUCTURE ONLY: see the header of cvtx_oracle.h for who may load
 * this and how its parity with the reference is pinned.
 *
 * The arithmetic lives in cvtx_oracle_body.inc, instantiated here once in the
 * reference's precision (FP32 pairs, FP64 sums) and once entirely in FP64.
 * Build: `make -C oracle port` (gcc -O2 -fopenmp -ffp-contract=off).
 */
#include <math.h>
#include <omp.h>
#include "cvtx_oracle.h"

/* ---------------- FP32 instantiation: the reference's arithmetic --------- */
#define R float
#define OUT_T float
#define RS(name) name##_f32
#define M_POW_ powf
#define M_SQRT_ sqrtf
#define M_EXP_ expf
#define M_FABS_ fabsf
#define M_ACOS_ acosf
#include "cvtx_oracle_body.inc"
#undef R
#undef OUT_T
#undef RS
#undef M_POW_
#undef M_SQRT_
#undef M_EXP_
#undef M_FABS_
#undef M_ACOS_

/* ---------------- FP64 instantiation: the arbiter ------------------------ */
#define R double
#define OUT_T double
#define RS(name) name##_f64
#define M_POW_ pow
#define M_SQRT_ sqrt
#define M_EXP_ exp
#define M_FABS_ fabs
#define M_ACOS_ acos
#include "cvtx_oracle_body.inc"
#undef R
#undef OUT_T
#undef RS
#undef M_POW_
#undef M_SQRT_
#undef M_EXP_
#undef M_FABS_
#undef M_ACOS_

/* ---------------- exported scalar / single-pair entry points ------------- */
float cvtx_oracle_g3d(int reg, float rho) { return g3d_f32(reg, rho); }
float cvtx_oracle_zeta3d(int reg, float rho) { return zeta3d_f32(reg, rho); }
float cvtx_oracle_eta3d(int reg, float rho) { return eta3d_f32(reg, rho); }
float cvtx_oracle_g2d(int reg, float rho) { return g2d_f32(reg, rho); }
float cvtx_oracle_eta2d(int reg, float rho) { return eta2d_f32(reg, rho); }
double cvtx_oracle_g3d_f64(int reg, double rho) { return g3d_f64(reg, rho); }
double cvtx_oracle_zeta3d_f64(int reg, double rho) { return zeta3d_f64(reg, rho); }
double cvtx_oracle_eta3d_f64(int reg, double rho) { return eta3d_f64(reg, rho); }
double cvtx_oracle_g2d_f64(int reg, double rho) { return g2d_f64(reg, rho); }
double cvtx_oracle_eta2d_f64(int reg, double rho) { return eta2d_f64(reg, rho); }

static void put3(float *out, vec3_f32 v) { out[0] = v.v[0]; out[1] = v.v[1]; out[2] = v.v[2]; }

/* reference src/P3D.cpp:74-84: inner * 1/(4 pi) */
void cvtx_oracle_P3D_S2S_vel(const float *src7, const float *pt3, int reg, float sigma, float *out3)
{
	vec3_f32 u = p3d_vel_pair_f32(src7, pt3, reg, 1.f / fabsf(sigma));
	put3(out3, scl3_f32(u, 1.f / (4.f * 3.14159265359f)));
}
void cvtx_oracle_P3D_S2S_dvort(const float *src7, const float *tgt7, int reg, float sigma, float *out3)
{
	put3(out3, p3d_dvort_pair_f32(src7, tgt7, reg, sigma));
}
void cvtx_oracle_P3D_S2S_visc_dvort(const float *src7, const float *tgt7, int reg, float sigma, float nu, float *out3)
{
	put3(out3, p3d_visc_pair_f32(src7, tgt7, reg, sigma, nu));
}
/* reference src/P2D.cpp:71-81: inner * 1/(2 pi) */
void cvtx_oracle_P2D_S2S_vel(const float *src4, const float *pt2, int reg, float sigma, float *out2)
{
	float ux, uy;
	p2d_vel_pair_f32(src4, pt2, reg, 1.f / fabsf(sigma), &ux, &uy);
	const float scale = 1.f / (2.f * acosf(-1.f));
	out2[0] = ux * scale;
	out2[1] = uy * scale;
}
void cvtx_oracle_P2D_S2S_visc_dvort(const float *src4, const float *tgt4, int reg, float sigma, float nu, float *out1)
{
	out1[0] = p2d_visc_pair_f32(src4, tgt4, reg, sigma, nu);
}
void cvtx_oracle_F3D_S2S_vel(const float *fil7, const float *pt3, float *out3)
{
	put3(out3, f3d_vel_pair_f32(fil7, pt3));
}
void cvtx_oracle_F3D_S2S_dvort(const float *fil7, const float *tgt7, float *out3)
{
	put3(out3, f3d_dvort_pair_f32(fil7, tgt7));
}

int cvtx_oracle_num_threads(void) { return omp_get_max_threads(); }

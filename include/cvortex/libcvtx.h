#ifndef CVTX_LIBCVTX_H
#define CVTX_LIBCVTX_H
/*
 * cvortex/libcvtx.h -- public C ABI of libcvortex, B200 build.
 *
 * The contract is the one of the reference's include/cvortex/libcvtx.h (v0.3.8):
 * the same 52 unmangled symbols, the same POD layouts, the same calling
 * conventions, so existing C callers and Julia's CVortex.jl (ccall) link and
 * run unchanged.  This file is a fresh write-up of that contract, with each
 * group annotated by what stands behind it in this build:
 *
 *   [B200]  runs on the sm_100a kernels of cvortex_b200/csrc when at least one
 *           accelerator is enabled and the cvtx_VortFunc carries one of the
 *           four built-in kernel names; a CUDA failure is reported on stderr
 *           and aborts -- it is never replaced by a silent CPU result.
 *   [host]  small scalar host code (single pairs, one-to-many, many-to-one,
 *           and the all-pairs ops when the caller disabled every accelerator
 *           or passed a user-defined cvtx_VortFunc).
 *   [B200/remesh]  redistribution onto a grid: the node build (spread, sort by
 *           node, per-node sums) runs on the first enabled accelerator when the
 *           cvtx_RedistFunc is one of the five built-ins, on the host otherwise
 *           (same bits either way); the pruning of weak nodes runs on the host.
 *           Relaxation is one cvtx_P3D_M2M_vort call plus a per-particle blend.
 *
 * Attribution: the API restated here -- names, signatures, struct layouts -- is that of
 * cvortex by H. J. A. Bird (https://github.com/hjabird/cvortex), MIT License,
 * Copyright (c) 2018 HJA Bird.  The implementation behind it in this tree is new.
 *
 * Struct sizes relied on across the ABI (static_asserted in the library):
 *   cvtx_P3D 28, cvtx_F3D 28, cvtx_P2D 16, cvtx_VortFunc 80 (LP64),
 *   bsv_V3f 12, bsv_V2f 8.
 */

#ifndef CVTX_EXPORT
# ifdef _WIN32
#  define CVTX_EXPORT __declspec(dllimport)
# else
#  define CVTX_EXPORT
# endif
#endif

#ifdef __cplusplus
extern "C" {
#endif

#include <bsv/bsv.h>

/* ------------------------------------------------------------------ types */

/* 3D vortex particle (reference libcvtx.h:53-57) */
typedef struct {
	bsv_V3f coord;
	bsv_V3f vorticity;
	float volume;
} cvtx_P3D;

/* 3D straight singular vortex filament, start -> end (reference :60-63) */
typedef struct {
	bsv_V3f start, end;
	float strength;
} cvtx_F3D;

/* 2D vortex particle (reference :66-70) */
typedef struct {
	bsv_V2f coord;
	float vorticity;
	float area;
} cvtx_P2D;

/* Regularisation: g, zeta, eta in Winckelmans' naming, as host function
 * pointers, plus the key that selects the accelerated kernel
 * ("singular", "winckelmans", "planetary", "gaussian"; anything else means
 * "only the function pointers can evaluate this") (reference :86-94). */
typedef struct {
	float (*g_3D)(float rho);
	float (*g_2D)(float rho);
	float (*zeta_3D)(float rho);
	void (*combined_3D)(float rho, float *g, float *zeta);
	float (*eta_3D)(float rho);
	float (*eta_2D)(float rho);
	char cl_kernel_name_ext[32];
} cvtx_VortFunc;

/* Redistribution interpolant (reference :96-99) */
typedef struct {
	float (*func)(float U);
	float radius;
} cvtx_RedistFunc;

/* ------------------------------------------- library / accelerator control
 * [B200] accelerator k is CUDA device k (reference :102-110,
 * src/accelerators.cpp:39-118).  After cvtx_initialise() accelerator 0 is
 * enabled, as in the reference (src/opencl_acc.cpp:192-201); enabling more
 * shards the targets of every all-pairs call across them. */
CVTX_EXPORT void cvtx_initialise();
CVTX_EXPORT void cvtx_finalise();
CVTX_EXPORT const char *cvtx_information();
CVTX_EXPORT int cvtx_num_accelerators();
CVTX_EXPORT int cvtx_num_enabled_accelerators();
CVTX_EXPORT const char *cvtx_accelerator_name(int accelerator_id);
CVTX_EXPORT int cvtx_accelerator_enabled(int accelerator_id);
CVTX_EXPORT void cvtx_accelerator_enable(int accelerator_id);
CVTX_EXPORT void cvtx_accelerator_disable(int accelerator_id);

/* ------------------------------------------------ regularisations [host] */
CVTX_EXPORT const cvtx_VortFunc cvtx_VortFunc_singular(void);
CVTX_EXPORT const cvtx_VortFunc cvtx_VortFunc_winckelmans(void);
CVTX_EXPORT const cvtx_VortFunc cvtx_VortFunc_planetary(void);
CVTX_EXPORT const cvtx_VortFunc cvtx_VortFunc_gaussian(void);

/* ---------------------------------------- redistribution kernels [host] */
CVTX_EXPORT const cvtx_RedistFunc cvtx_RedistFunc_lambda0(void);
CVTX_EXPORT const cvtx_RedistFunc cvtx_RedistFunc_lambda1(void);
CVTX_EXPORT const cvtx_RedistFunc cvtx_RedistFunc_lambda2(void);
CVTX_EXPORT const cvtx_RedistFunc cvtx_RedistFunc_lambda3(void);
CVTX_EXPORT const cvtx_RedistFunc cvtx_RedistFunc_m4p(void);

/* ============================================================ 3D particles */

/* ---- all pairs, many sources on many targets: THE HOT PATH [B200] ----
 * result_array has num_mes / num_induced entries and is overwritten. */
CVTX_EXPORT void cvtx_P3D_M2M_vel(
	const cvtx_P3D **array_start, const int num_particles,
	const bsv_V3f *mes_start, const int num_mes,
	bsv_V3f *result_array,
	const cvtx_VortFunc *kernel, float regularisation_radius);

CVTX_EXPORT void cvtx_P3D_M2M_dvort(
	const cvtx_P3D **array_start, const int num_particles,
	const cvtx_P3D **induced_start, const int num_induced,
	bsv_V3f *result_array,
	const cvtx_VortFunc *kernel, float regularisation_radius);

CVTX_EXPORT void cvtx_P3D_M2M_visc_dvort(
	const cvtx_P3D **array_start, const int num_particles,
	const cvtx_P3D **induced_start, const int num_induced,
	bsv_V3f *result_array,
	const cvtx_VortFunc *kernel, float regularisation_radius,
	float kinematic_visc);

CVTX_EXPORT void cvtx_P3D_M2M_vort(
	const cvtx_P3D **array_start, const int num_particles,
	const bsv_V3f *mes_start, const int num_mes,
	bsv_V3f *result_array,
	const cvtx_VortFunc *kernel, float regularisation_radius);

/* ---- single pair [host] ---- */
CVTX_EXPORT bsv_V3f cvtx_P3D_S2S_vel(
	const cvtx_P3D *self, const bsv_V3f mes_point,
	const cvtx_VortFunc *kernel, float regularisation_radius);
CVTX_EXPORT bsv_V3f cvtx_P3D_S2S_dvort(
	const cvtx_P3D *self, const cvtx_P3D *induced_particle,
	const cvtx_VortFunc *kernel, float regularisation_radius);
CVTX_EXPORT bsv_V3f cvtx_P3D_S2S_visc_dvort(
	const cvtx_P3D *self, const cvtx_P3D *induced_particle,
	const cvtx_VortFunc *kernel, float regularisation_radius,
	float kinematic_visc);
CVTX_EXPORT bsv_V3f cvtx_P3D_S2S_vort(
	const cvtx_P3D *self, const bsv_V3f mes_point,
	const cvtx_VortFunc *kernel, float regularisation_radius);

/* ---- one source on many targets [host] ---- */
CVTX_EXPORT void cvtx_P3D_S2M_vel(
	const cvtx_P3D *self, const bsv_V3f *mes_start, const int num_mes,
	bsv_V3f *result_array,
	const cvtx_VortFunc *kernel, float regularisation_radius);
CVTX_EXPORT void cvtx_P3D_S2M_dvort(
	const cvtx_P3D *self, const cvtx_P3D **induced_start, const int num_induced,
	bsv_V3f *result_array,
	const cvtx_VortFunc *kernel, float regularisation_radius);
CVTX_EXPORT void cvtx_P3D_S2M_visc_dvort(
	const cvtx_P3D *self, const cvtx_P3D **induced_start, const int num_induced,
	bsv_V3f *result_array,
	const cvtx_VortFunc *kernel, float regularisation_radius,
	float kinematic_visc);
CVTX_EXPORT void cvtx_P3D_S2M_vort(
	const cvtx_P3D *self, const bsv_V3f *mes_start, const int num_mes,
	bsv_V3f *result_array,
	const cvtx_VortFunc *kernel, float regularisation_radius);

/* ---- many sources on one target [host] ---- */
CVTX_EXPORT bsv_V3f cvtx_P3D_M2S_vel(
	const cvtx_P3D **array_start, const int num_particles,
	const bsv_V3f mes_point,
	const cvtx_VortFunc *kernel, float regularisation_radius);
CVTX_EXPORT bsv_V3f cvtx_P3D_M2S_dvort(
	const cvtx_P3D **array_start, const int num_particles,
	const cvtx_P3D *induced_particle,
	const cvtx_VortFunc *kernel, float regularisation_radius);
CVTX_EXPORT bsv_V3f cvtx_P3D_M2S_visc_dvort(
	const cvtx_P3D **array_start, const int num_particles,
	const cvtx_P3D *induced_particle,
	const cvtx_VortFunc *kernel, float regularisation_radius,
	float kinematic_visc);
CVTX_EXPORT bsv_V3f cvtx_P3D_M2S_vort(
	const cvtx_P3D **array_start, const int num_particles,
	const bsv_V3f mes_point,
	const cvtx_VortFunc *kernel, float regularisation_radius);

/* ---- redistribution / relaxation [B200/remesh] ----
 * output_particles == NULL asks for the count only.  When more than
 * max_output_particles survive negligible_vort, the strongest are kept. */
CVTX_EXPORT int cvtx_P3D_redistribute_on_grid(
	const cvtx_P3D **input_array_start, const int n_input_particles,
	cvtx_P3D *output_particles, int max_output_particles,
	const cvtx_RedistFunc *redistributor,
	float grid_density, float negligible_vort);
CVTX_EXPORT void cvtx_P3D_pedrizzetti_relaxation(
	cvtx_P3D **input_array_start, const int n_input_particles,
	float fdt,
	const cvtx_VortFunc *kernel, float regularisation_radius);

/* ============================================================ 3D filaments */

/* ---- all pairs [B200] ---- */
CVTX_EXPORT void cvtx_F3D_M2M_vel(
	const cvtx_F3D **array_start, const int num_filaments,
	const bsv_V3f *mes_start, const int num_mes,
	bsv_V3f *result_array);
CVTX_EXPORT void cvtx_F3D_M2M_dvort(
	const cvtx_F3D **array_start, const int num_filaments,
	const cvtx_P3D **induced_start, const int num_induced,
	bsv_V3f *result_array);

/* ---- single pair, many on one [host] ---- */
CVTX_EXPORT bsv_V3f cvtx_F3D_S2S_vel(const cvtx_F3D *self, const bsv_V3f mes_point);
CVTX_EXPORT bsv_V3f cvtx_F3D_S2S_dvort(const cvtx_F3D *self, const cvtx_P3D *induced_particle);
CVTX_EXPORT bsv_V3f cvtx_F3D_M2S_vel(
	const cvtx_F3D **array_start, const int num_filaments, const bsv_V3f mes_point);
CVTX_EXPORT bsv_V3f cvtx_F3D_M2S_dvort(
	const cvtx_F3D **array_start, const int num_filaments, const cvtx_P3D *induced_particle);

/* ---- dense influence matrix, result_matrix[i * num_filaments + j] [host] ---- */
CVTX_EXPORT void cvtx_F3D_inf_mtrx(
	const cvtx_F3D **array_start, const int num_filaments,
	const bsv_V3f *mes_start, const bsv_V3f *dir_start, const int num_mes,
	float *result_matrix);

/* ============================================================ 2D particles */

/* ---- all pairs [B200] ---- */
CVTX_EXPORT void cvtx_P2D_M2M_vel(
	const cvtx_P2D **array_start, const int num_particles,
	const bsv_V2f *mes_start, const int num_mes,
	bsv_V2f *result_array,
	const cvtx_VortFunc *kernel, float regularisation_radius);
CVTX_EXPORT void cvtx_P2D_M2M_visc_dvort(
	const cvtx_P2D **array_start, const int num_particles,
	const cvtx_P2D **induced_start, const int num_induced,
	float *result_array,
	const cvtx_VortFunc *kernel, float regularisation_radius,
	float kinematic_visc);

/* ---- single pair, one on many, many on one [host] ---- */
CVTX_EXPORT bsv_V2f cvtx_P2D_S2S_vel(
	const cvtx_P2D *self, const bsv_V2f mes_point,
	const cvtx_VortFunc *kernel, float regularisation_radius);
CVTX_EXPORT void cvtx_P2D_S2M_vel(
	const cvtx_P2D *self, const bsv_V2f *mes_start, const int num_mes,
	bsv_V2f *result_array,
	const cvtx_VortFunc *kernel, float regularisation_radius);
CVTX_EXPORT bsv_V2f cvtx_P2D_M2S_vel(
	const cvtx_P2D **array_start, const int num_particles,
	const bsv_V2f mes_point,
	const cvtx_VortFunc *kernel, float regularisation_radius);
CVTX_EXPORT float cvtx_P2D_S2S_visc_dvort(
	const cvtx_P2D *self, const cvtx_P2D *induced_particle,
	const cvtx_VortFunc *kernel, float regularisation_radius,
	float kinematic_visc);
CVTX_EXPORT void cvtx_P2D_S2M_visc_dvort(
	const cvtx_P2D *self, const cvtx_P2D **induced_start, const int num_induced,
	float *result_array,
	const cvtx_VortFunc *kernel, float regularisation_radius,
	float kinematic_visc);
CVTX_EXPORT float cvtx_P2D_M2S_visc_dvort(
	const cvtx_P2D **array_start, const int num_particles,
	const cvtx_P2D *induced_particle,
	const cvtx_VortFunc *kernel, float regularisation_radius,
	float kinematic_visc);

/* ---- redistribution; returns the number of particles created [B200/remesh] ---- */
CVTX_EXPORT int cvtx_P2D_redistribute_on_grid(
	const cvtx_P2D **input_array_start, const int num_particles,
	cvtx_P2D *output_particles, int num_output_particles,
	const cvtx_RedistFunc *redistributor,
	float grid_density, float negligible_vort);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* CVTX_LIBCVTX_H */

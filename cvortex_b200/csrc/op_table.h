// op_table.h -- the closed set of (op, regularisation) pairs the B200 backend
// implements, the packed source formats, and a compile-time dispatcher from
// the run-time (op id, regularisation id) to the matching policy of
// pair_math.cuh.
//
// The reference selects its OpenCL kernel by string concatenation,
// "cvtx_nb_P3D_vel_" + kernel->cl_kernel_name_ext (reference
// src/ocl_P3D.cpp:162); here the same key picks a template instantiation.
#pragma once
#include <string.h>
#include "pair_math.cuh"

namespace cvtx {

enum OpId {
	OP_P3D_VEL = 0,     // cvtx_P3D_M2M_vel         libcvtx.h:213-220
	OP_P3D_DVORT = 1,   // cvtx_P3D_M2M_dvort       libcvtx.h:222-229
	OP_P3D_VISC = 2,    // cvtx_P3D_M2M_visc_dvort  libcvtx.h:231-239
	OP_P3D_VORT = 3,    // cvtx_P3D_M2M_vort        libcvtx.h:241-248
	OP_P2D_VEL = 4,     // cvtx_P2D_M2M_vel         libcvtx.h:329-336
	OP_P2D_VISC = 5,    // cvtx_P2D_M2M_visc_dvort  libcvtx.h:362-370
	OP_F3D_VEL = 6,     // cvtx_F3D_M2M_vel         libcvtx.h:285-290
	OP_F3D_DVORT = 7,   // cvtx_F3D_M2M_dvort       libcvtx.h:292-297
	OP_P3D_VEL_DVORT = 8,   // fused vel + dvort on particle targets (thin ABI only)
	OP_COUNT = 9
};

enum SrcKind { SRC_P3D = 0, SRC_P2D = 1, SRC_F3D = 2 };

inline SrcKind src_kind(int op) {
	if (op == OP_P3D_VEL_DVORT) return SRC_P3D;
	return op <= OP_P3D_VORT ? SRC_P3D : (op <= OP_P2D_VISC ? SRC_P2D : SRC_F3D);
}
inline int src_cols(int op) { return src_kind(op) == SRC_P2D ? 4 : 7; }   // floats per raw source row
inline bool op_is_filament(int op) { return op == OP_F3D_VEL || op == OP_F3D_DVORT; }

// Is (op, reg) a combination the reference accelerates?  visc ops only exist
// for the two regularisations with an eta (reference src/nbody.cl:498-510,
// :593-605); filaments are singular only.
inline bool op_supported(int op, int reg) {
	if (op < 0 || op >= OP_COUNT) return false;
	if (op_is_filament(op)) return true;
	if (reg < 0 || reg > 3) return false;
	if (op == OP_P3D_VISC || op == OP_P2D_VISC) return reg == REG_WINCKELMANS || reg == REG_GAUSSIAN;
	return true;
}

// The dispatch key of the reference ABI: cvtx_VortFunc::cl_kernel_name_ext
// (reference src/VortFunc.cpp:210,223,236,249).  -1 = not one of ours (a user
// regularisation: only its function pointers can evaluate it).
inline int reg_from_name(const char *name) {
	if (!strcmp(name, "singular")) return REG_SINGULAR;
	if (!strcmp(name, "winckelmans")) return REG_WINCKELMANS;
	if (!strcmp(name, "planetary")) return REG_PLANETARY;
	if (!strcmp(name, "gaussian")) return REG_GAUSSIAN;
	return -1;
}

// ---- packed source records (16-byte aligned, one or two float4 per source) --
// P3D: a = {x, y, z, vol}     b = {wx, wy, wz, 0}
// P2D: a = {x, y, Gamma, area}                      (a cvtx_P2D verbatim)
// F3D: a = {ax, ay, az, G/4pi} b = {bx, by, bz, 3 G/(4 pi |r0|)} c = {r0, |r0|^2},  r0 = end - start (FP32)
// Rows beyond n (padding up to a whole tile) carry zero strength, which makes
// every pair formula contribute exactly 0: particles are all-zero records; the
// padding filament is the unit segment (0,0,0)-(1,0,0) with zero strength, NOT a
// zero-length one, so that its terms are finite zeros for every target off the
// x axis and the optimistic pair loop (pair_math.cuh, GUARDS) does not have to
// re-evaluate the last chain of every call.
CVTX_HD void pack_source(int kind, const float *row, f4 &a, f4 &b, f4 &c) {
	c.x = c.y = c.z = c.w = 0.0f;
	if (kind == SRC_P3D) {
		a.x = row[0]; a.y = row[1]; a.z = row[2]; a.w = row[6];
		b.x = row[3]; b.y = row[4]; b.z = row[5]; b.w = 0.0f;
	} else if (kind == SRC_P2D) {
		a.x = row[0]; a.y = row[1]; a.z = row[2]; a.w = row[3];
		b.x = b.y = b.z = b.w = 0.0f;
	} else {
		const float t1 = row[6] / (4.0f * 3.14159265359f);          // strength/(4 pi_f), reference src/F3D.cpp:66
		const float rx = row[3] - row[0], ry = row[4] - row[1], rz = row[5] - row[2];
		a.x = row[0]; a.y = row[1]; a.z = row[2]; a.w = t1;
		b.x = row[3]; b.y = row[4]; b.z = row[5];
		b.w = (3.0f / sqrtf(rx * rx + ry * ry + rz * rz)) * t1;      // (3/|r0|) t1, reference src/F3D.cpp:70
		c.x = rx; c.y = ry; c.z = rz; c.w = rx * rx + ry * ry + rz * rz;
	}
}

CVTX_HD void pad_source(int kind, f4 &a, f4 &b, f4 &c) {
	a.x = a.y = a.z = a.w = 0.0f;
	b.x = b.y = b.z = b.w = 0.0f;
	c.x = c.y = c.z = c.w = 0.0f;
	if (kind == SRC_F3D) { b.x = 1.0f; c.x = 1.0f; c.w = 1.0f; }
}
inline int src_records(int op) { return src_kind(op) == SRC_P2D ? 1 : (src_kind(op) == SRC_F3D ? 3 : 2); }   // float4 records per packed source

// Call f.template run<Policy>() for the policy of (op, reg).  Returns false
// for an unsupported combination.
template <class F> inline bool dispatch_op(int op, int reg, F &f) {
	if (!op_supported(op, reg)) return false;
#define CVTX_REG_CASES(POLICY)                                                       \
	switch (reg) {                                                                   \
	case REG_SINGULAR:    f.template run<POLICY<REG_SINGULAR>>(); return true;       \
	case REG_WINCKELMANS: f.template run<POLICY<REG_WINCKELMANS>>(); return true;    \
	case REG_PLANETARY:   f.template run<POLICY<REG_PLANETARY>>(); return true;      \
	default:              f.template run<POLICY<REG_GAUSSIAN>>(); return true;       \
	}
#define CVTX_ETA_CASES(POLICY)                                                       \
	if (reg == REG_WINCKELMANS) f.template run<POLICY<REG_WINCKELMANS>>();           \
	else f.template run<POLICY<REG_GAUSSIAN>>();                                     \
	return true;
	switch (op) {
	case OP_P3D_VEL:   CVTX_REG_CASES(P3DVel)
	case OP_P3D_DVORT: CVTX_REG_CASES(P3DDvort)
	case OP_P3D_VISC:  CVTX_ETA_CASES(P3DVisc)
	case OP_P3D_VORT:  CVTX_REG_CASES(P3DVort)
	case OP_P2D_VEL:   CVTX_REG_CASES(P2DVel)
	case OP_P2D_VISC:  CVTX_ETA_CASES(P2DVisc)
	case OP_F3D_VEL:   f.template run<F3DVel>(); return true;
	case OP_F3D_DVORT: f.template run<F3DDvort>(); return true;
	default:           CVTX_REG_CASES(P3DVelDvort)
	}
#undef CVTX_REG_CASES
#undef CVTX_ETA_CASES
}

}  // namespace cvtx

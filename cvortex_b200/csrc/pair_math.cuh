// pair_math.cuh -- per-pair arithmetic of every all-pairs (M2M) op, one policy
// struct per (op, regularisation), specialised at compile time.
//
// This is the B200 re-derivation of what the reference evaluates per pair in
//   src/P3D.cpp:51-144, src/P2D.cpp:49-69,167-194, src/F3D.cpp:34-85 with the
//   regularisation scalars of src/VortFunc.cpp:64-199
// (the reference's OpenCL reading of the same maths is src/nbody.cl:38-728).
// It is NOT a transcription: every formula is rewritten so that
//   * sigma powers, 1/4pi, nu ... leave the pair loop (they scale the finished
//     sums once, in FP64);
//   * target-only factors leave it too, using bilinearity:
//       visc      : w_s V_t - w_t V_s = (w_s - w_t) V_t - w_t (V_s - V_t), so
//                   dw_t = V_t sum eta (w_s - w_t) - w_t sum eta (V_s - V_t)
//                   (differences taken per pair: for a smooth field the plain
//                   hoisting V_t sum(eta w_s) - w_t sum(eta V_s) cancels
//                   catastrophically; measured 8x worse than the reference)
//       F3D dvort : sum_s (B w_t + A x w_t) = (sum B) w_t + (sum A) x w_t
//     P3D dvort deliberately keeps c = w_t x w_s per pair for the same reason;
//   * Winckelmans terms are division-free and regular at r = 0
//       g/r^3          = (rho^2+2.5)(rho^2+1)^-5/2 / sigma^3
//       (3g/rho^3 - zeta)/r^2 = (3rho^2+10.5)(rho^2+1)^-7/2 / sigma^2
//     so one MUFU.RSQ serves the whole pair;
//   * the only special-function work is MUFU rsqrt / rcp / ex2 (approx, ftz).
// The Gaussian g uses the same Abramowitz-Stegun 7.1.26 polynomial as the
// reference (src/VortFunc.cpp:164-173), because erff() differs from it by up
// to 1e-3 relative on near pairs.
//
// Coincident-pair rule (reference: contribution is exactly zero when the two
// positions compare equal, src/P3D.cpp:58,94,127, src/P2D.cpp:57,178): here a
// pair is dropped when r^2 == 0.  For formulas that are regular at r = 0 and
// carry a factor `rad` the rule holds with no test at all.
//
// Each policy exposes
//   NSRC4   float4 records per packed source (1 or 2)
//   TCOLS   floats per raw target row in global memory
//   NTGT    target floats kept in registers during the pair loop
//   NACC    FP32 running sums per target
//   NOUT    output floats per target
//   CHAIN   sources per FP32 running-sum chain before it is flushed into the
//           FP64 accumulator (0 = one chain per shared-memory tile).  The
//           viscous ops use short chains: their pair terms cancel to second
//           order over a smooth field (that is what PSE computes), so a long
//           FP32 chain would lose what the reference keeps by summing every
//           pair in double (src/P3D.cpp:283-295, src/P2D.cpp:222-230)
//   LANE_OPS, SFU_OPS   algorithmic FP32 lane-ops / MUFU ops per pair of THIS
//                       formulation (FMA = 1 lane-op; compares/selects not
//                       counted) -- the roofline denominators, see DESIGN.md
//   load_target(), pair(), finish()
// and is usable from host code too (tests/hostcheck compiles this header with
// g++ to validate the algebra against the oracle without a GPU; MUFU ops are
// then replaced by libm).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define CVTX_HD __host__ __device__ __forceinline__
namespace cvtx { typedef float4 f4; }
#else
#define CVTX_HD inline
namespace cvtx { struct f4 { float x, y, z, w; }; }
#endif

namespace cvtx {

enum Reg { REG_SINGULAR = 0, REG_WINCKELMANS = 1, REG_PLANETARY = 2, REG_GAUSSIAN = 3 };

// Pair-loop constants (FP32, meaning depends on the op) and the FP64 factors
// that scale the finished sums.  Filled on the host by make_consts().
struct PairConsts {
	float c0, c1, c2, c3;
	double s0, s1;
};

// ---- MUFU wrappers --------------------------------------------------------
CVTX_HD float mufu_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
	float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
	return 1.0f / sqrtf(x);
#endif
}
CVTX_HD float mufu_rcp(float x) {
#if defined(__CUDA_ARCH__)
	float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
	return 1.0f / x;
#endif
}
CVTX_HD float mufu_ex2(float x) {
#if defined(__CUDA_ARCH__)
	float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
	return exp2f(x);
#endif
}

static constexpr double kPi = 3.14159265359;              // CVTX_PI_F, reference src/P3D.cpp:47 (as a float literal there)
static constexpr double kSqrt2OverPi = 0.7978845608028654;  // reference src/VortFunc.cpp:49
static constexpr double kRecipSqrt2 = 0.7071067811865475;   // reference src/VortFunc.cpp:50
static constexpr double kLog2e = 1.4426950408889634;

// Abramowitz & Stegun 7.1.26, coefficients as in reference src/VortFunc.cpp:165-166.
// Returns s with  g_gauss3D(rho) = 1 - e * s,  e = exp(-rho^2/2):
//   erf(z) ~ 1 - poly(t) e,  t = 1/(1 + p z),  z = rho/sqrt2
//   g = erf(z) - rho sqrt(2/pi) e = 1 - e (poly(t) + rho sqrt(2/pi))
CVTX_HD float gauss_tail(float r, float k_t, float k_c) {
	const float t = mufu_rcp(fmaf(r, k_t, 1.0f));
	float p = fmaf(t, 1.061405429f, -1.453152027f);
	p = fmaf(t, p, 1.421413741f);
	p = fmaf(t, p, -0.284496736f);
	p = fmaf(t, p, 0.254829592f);
	return fmaf(r, k_c, p * t);
}

// ---------------------------------------------------------------------------
// 3D kernels:   A(r2)  ~ g(rho)/r^3           (velocity / first stretching term)
//               Bn(r2) = -(3 g/rho^3 - zeta)/r^2 in the SAME units as A, so that
//                        one accumulator set takes  A c + Bn (rad.c) rad
// Units are chosen per regularisation so the loop needs the fewest lane-ops;
// the matching sigma power sits in PairConsts::s0 (scaleA).
// ---------------------------------------------------------------------------
template <int REG> struct Reg3D;

template <> struct Reg3D<REG_WINCKELMANS> {
	// c0 = 1/sigma^2, c1 = -3/sigma^4, c2 = -10.5/sigma^2.
	// A = sigma^3 g/r^3 = g/rho^3.  Bn = -(3 rho^2 + 10.5)(rho^2+1)^-7/2 / sigma^2.
	static constexpr int A_OPS = 6, AB_OPS = 9, SFU = 1;
	CVTX_HD static float A(float r2, const PairConsts &k) {
		const float a = fmaf(r2, k.c0, 1.0f), b = fmaf(r2, k.c0, 2.5f);
		const float ra = mufu_rsqrt(a), ra2 = ra * ra, ra4 = ra2 * ra2;
		return b * (ra4 * ra);
	}
	CVTX_HD static void AB(float r2, const PairConsts &k, float &A_, float &B_) {
		const float a = fmaf(r2, k.c0, 1.0f), b = fmaf(r2, k.c0, 2.5f), b2 = fmaf(r2, k.c1, k.c2);
		const float ra = mufu_rsqrt(a), ra2 = ra * ra, ra4 = ra2 * ra2, ra5 = ra4 * ra;
		// the self pair must give exactly 0: c = w_t x w_t formed with FMAs is
		// only zero to rounding, and A(0) = 2.5 would amplify that residue
		A_ = r2 > 0.0f ? b * ra5 : 0.0f;
		B_ = b2 * (ra5 * ra2);
	}
	static void consts(PairConsts &k, double s) {
		k.c0 = (float)(1.0 / (s * s)); k.c1 = (float)(-3.0 / (s * s * s * s)); k.c2 = (float)(-10.5 / (s * s));
	}
	static double scaleA(double s) { return 1.0 / (s * s * s); }
};

template <> struct Reg3D<REG_SINGULAR> {
	// A = 1/r^3, Bn = -3/r^5; both dropped at r = 0.
	static constexpr int A_OPS = 2, AB_OPS = 4, SFU = 1;
	CVTX_HD static float A(float r2, const PairConsts &) {
		const float ri = mufu_rsqrt(r2);
		return r2 > 0.0f ? ri * ri * ri : 0.0f;
	}
	CVTX_HD static void AB(float r2, const PairConsts &, float &A_, float &B_) {
		const float ri = mufu_rsqrt(r2), ri2 = ri * ri, ri3 = ri2 * ri;
		const bool ok = r2 > 0.0f;
		A_ = ok ? ri3 : 0.0f;
		B_ = ok ? ri3 * (ri2 * -3.0f) : 0.0f;
	}
	static void consts(PairConsts &, double) {}
	static double scaleA(double) { return 1.0; }
};

template <> struct Reg3D<REG_PLANETARY> {
	// rho < 1: g = rho^3, zeta = 3  ->  A = 1/sigma^3, Bn = 0;  else singular.
	// c0 = sigma^2, c1 = 1/sigma^3.
	static constexpr int A_OPS = 2, AB_OPS = 4, SFU = 1;
	CVTX_HD static float A(float r2, const PairConsts &k) {
		const float ri = mufu_rsqrt(r2);
		return r2 < k.c0 ? k.c1 : ri * ri * ri;
	}
	CVTX_HD static void AB(float r2, const PairConsts &k, float &A_, float &B_) {
		const float ri = mufu_rsqrt(r2), ri2 = ri * ri, ri3 = ri2 * ri;
		const bool in = r2 < k.c0;
		A_ = in ? (r2 > 0.0f ? k.c1 : 0.0f) : ri3;        // exact 0 for the self pair, as above
		B_ = in ? 0.0f : ri3 * (ri2 * -3.0f);
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(s * s); k.c1 = (float)(1.0 / (s * s * s)); }
	static double scaleA(double) { return 1.0; }
};

template <> struct Reg3D<REG_GAUSSIAN> {
	// c0 = p/(sqrt2 sigma), c1 = -log2(e)/(2 sigma^2), c2 = sqrt(2/pi)/sigma,
	// c3 = sqrt(2/pi)/sigma^3.   A = g/r^3,  Bn = (c3 e - 3A)/r^2.
	static constexpr int A_OPS = 13, AB_OPS = 16, SFU = 3;
	CVTX_HD static float A(float r2, const PairConsts &k) {
		const float ri = mufu_rsqrt(r2), r = r2 * ri;
		const float e = mufu_ex2(r2 * k.c1);
		const float s = gauss_tail(r, k.c0, k.c2);
		const float g = fmaf(-e, s, 1.0f);
		return r2 > 0.0f ? g * (ri * ri * ri) : 0.0f;
	}
	CVTX_HD static void AB(float r2, const PairConsts &k, float &A_, float &B_) {
		const float ri = mufu_rsqrt(r2), r = r2 * ri, ri2 = ri * ri;
		const float e = mufu_ex2(r2 * k.c1);
		const float s = gauss_tail(r, k.c0, k.c2);
		const float g = fmaf(-e, s, 1.0f);
		const float a = g * (ri2 * ri);
		const float h = fmaf(-3.0f, a, k.c3 * e);
		const bool ok = r2 > 0.0f;
		A_ = ok ? a : 0.0f;
		B_ = ok ? h * ri2 : 0.0f;
	}
	static void consts(PairConsts &k, double s) {
		k.c0 = (float)(0.3275911 * kRecipSqrt2 / s);
		k.c1 = (float)(-0.5 * kLog2e / (s * s));
		k.c2 = (float)(kSqrt2OverPi / s);
		k.c3 = (float)(kSqrt2OverPi / (s * s * s));
	}
	static double scaleA(double) { return 1.0; }
};

// ===========================================================================
// cvtx_P3D_M2M_vel      u_t = -(1/4pi) sum_s [g(rho)/r^3] (rad x w_s)
// reference: src/P3D.cpp:51-72 (pair), :230-251 (sum), :343-366 (entry)
// source  a = {x, y, z, vol}   b = {wx, wy, wz, 0}
// ===========================================================================
template <int REG> struct P3DVel {
	static constexpr int NSRC4 = 2, TCOLS = 3, NTGT = 3, NACC = 3, NOUT = 3, CHAIN = 0;
	static constexpr int LANE_OPS = 15 + Reg3D<REG>::A_OPS, SFU_OPS = Reg3D<REG>::SFU;
	CVTX_HD static void load_target(const float *row, float *tg) { tg[0] = row[0]; tg[1] = row[1]; tg[2] = row[2]; }
	CVTX_HD static void pair(const float *tg, const f4 a, const f4 b, float *acc, const PairConsts &k) {
		const float dx = tg[0] - a.x, dy = tg[1] - a.y, dz = tg[2] - a.z;
		const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
		const float K = Reg3D<REG>::A(r2, k);
		const float cx = fmaf(dy, b.z, -(dz * b.y));
		const float cy = fmaf(dz, b.x, -(dx * b.z));
		const float cz = fmaf(dx, b.y, -(dy * b.x));
		acc[0] = fmaf(K, cx, acc[0]);
		acc[1] = fmaf(K, cy, acc[1]);
		acc[2] = fmaf(K, cz, acc[2]);
	}
	CVTX_HD static void finish(const float *, const double *acc, double *out, const PairConsts &k) {
		out[0] = acc[0] * k.s0; out[1] = acc[1] * k.s0; out[2] = acc[2] * k.s0;
	}
	static PairConsts make_consts(float sigma, float) {
		PairConsts k = {}; const double s = fabs((double)sigma);      // reference uses 1/fabsf(sigma), src/P3D.cpp:239
		Reg3D<REG>::consts(k, s);
		k.s0 = -Reg3D<REG>::scaleA(s) / (4.0 * kPi);
		return k;
	}
};

// ===========================================================================
// cvtx_P3D_M2M_dvort    dw_t = 1/(4 pi sigma^3) sum_s [ A' c - (3A' - zeta)(rad.c) rad / r^2 ]
//                       c = w_t x w_s,  A' = g/rho^3
// reference: src/P3D.cpp:86-114 (pair), :253-273 (sum), :387-410 (entry)
// c is formed per pair (it is needed for rad.c anyway), NOT hoisted as
// w_t x sum(A w_s): neighbouring particles of a smooth field have nearly
// parallel vorticity, and the hoisted sum would cancel catastrophically there.
// ===========================================================================
template <int REG> struct P3DDvort {
	static constexpr int NSRC4 = 2, TCOLS = 7, NTGT = 6, NACC = 3, NOUT = 3, CHAIN = 0;
	static constexpr int LANE_OPS = 22 + Reg3D<REG>::AB_OPS, SFU_OPS = Reg3D<REG>::SFU;
	CVTX_HD static void load_target(const float *row, float *tg) {
		for (int i = 0; i < 6; ++i) tg[i] = row[i];
	}
	CVTX_HD static void pair(const float *tg, const f4 a, const f4 b, float *acc, const PairConsts &k) {
		const float dx = tg[0] - a.x, dy = tg[1] - a.y, dz = tg[2] - a.z;
		const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
		float A, Bn;
		Reg3D<REG>::AB(r2, k, A, Bn);
		const float cx = fmaf(tg[4], b.z, -(tg[5] * b.y));
		const float cy = fmaf(tg[5], b.x, -(tg[3] * b.z));
		const float cz = fmaf(tg[3], b.y, -(tg[4] * b.x));
		const float trip = fmaf(dz, cz, fmaf(dy, cy, dx * cx));
		const float s = Bn * trip;
		acc[0] = fmaf(s, dx, fmaf(A, cx, acc[0]));
		acc[1] = fmaf(s, dy, fmaf(A, cy, acc[1]));
		acc[2] = fmaf(s, dz, fmaf(A, cz, acc[2]));
	}
	CVTX_HD static void finish(const float *, const double *acc, double *out, const PairConsts &k) {
		out[0] = acc[0] * k.s0; out[1] = acc[1] * k.s0; out[2] = acc[2] * k.s0;
	}
	static PairConsts make_consts(float sigma, float) {
		PairConsts k = {}; const double s = fabs((double)sigma);
		Reg3D<REG>::consts(k, s);
		// rho uses |sigma| (src/P3D.cpp:99) but the prefactor keeps the sign of
		// sigma^3 (powf(sigma, 3), src/P3D.cpp:102).  A' = g/rho^3 = sigma^3 scaleA A.
		const double sign = sigma < 0 ? -1.0 : 1.0;
		k.s0 = sign * Reg3D<REG>::scaleA(s) / (4.0 * kPi);
		return k;
	}
};

// ===========================================================================
// cvtx_P3D_M2M_visc_dvort   dw_t = (2 nu/sigma^2) sum_s (w_s V_t - w_t V_s) eta(rho)
// reference: src/P3D.cpp:116-144 (pair), :275-296 (sum), :432-456 (entry)
// running sums: sum eta (w_s - w_t) (3), sum eta (V_s - V_t) (1); only
// Winckelmans / Gaussian have an eta (src/VortFunc.cpp:101-108, :245).
// ===========================================================================
template <int REG> struct Eta3D;
template <> struct Eta3D<REG_WINCKELMANS> {   // eta = 52.5 (rho^2+1)^-9/2
	static constexpr int OPS = 5, SFU = 1;
	CVTX_HD static float eta(float r2, const PairConsts &k) {
		const float ra = mufu_rsqrt(fmaf(r2, k.c0, 1.0f));
		const float ra2 = ra * ra, ra4 = ra2 * ra2, ra8 = ra4 * ra4;
		return ra8 * ra;
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(1.0 / (s * s)); }
	static double scale() { return 52.5; }
};
template <> struct Eta3D<REG_GAUSSIAN> {      // eta = sqrt(2/pi) exp(-rho^2/2)
	static constexpr int OPS = 1, SFU = 1;
	CVTX_HD static float eta(float r2, const PairConsts &k) { return mufu_ex2(r2 * k.c0); }
	static void consts(PairConsts &k, double s) { k.c0 = (float)(-0.5 * kLog2e / (s * s)); }
	static double scale() { return kSqrt2OverPi; }
};

template <int REG> struct P3DVisc {
	static constexpr int NSRC4 = 2, TCOLS = 7, NTGT = 7, NACC = 4, NOUT = 3, CHAIN = 8;
	static constexpr int LANE_OPS = 14 + Eta3D<REG>::OPS, SFU_OPS = Eta3D<REG>::SFU;
	CVTX_HD static void load_target(const float *row, float *tg) {
		for (int i = 0; i < 7; ++i) tg[i] = row[i];
	}
	CVTX_HD static void pair(const float *tg, const f4 a, const f4 b, float *acc, const PairConsts &k) {
		const float dx = tg[0] - a.x, dy = tg[1] - a.y, dz = tg[2] - a.z;
		const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
		float eta = Eta3D<REG>::eta(r2, k);
		eta = r2 > 0.0f ? eta : 0.0f;                 // coincident pair contributes nothing
		acc[0] = fmaf(eta, b.x - tg[3], acc[0]);
		acc[1] = fmaf(eta, b.y - tg[4], acc[1]);
		acc[2] = fmaf(eta, b.z - tg[5], acc[2]);
		acc[3] = fmaf(eta, a.w - tg[6], acc[3]);
	}
	CVTX_HD static void finish(const float *row, const double *acc, double *out, const PairConsts &k) {
		const double vt = row[6];
		out[0] = k.s0 * (vt * acc[0] - (double)row[3] * acc[3]);
		out[1] = k.s0 * (vt * acc[1] - (double)row[4] * acc[3]);
		out[2] = k.s0 * (vt * acc[2] - (double)row[5] * acc[3]);
	}
	static PairConsts make_consts(float sigma, float nu) {
		PairConsts k = {}; const double s = fabs((double)sigma);
		Eta3D<REG>::consts(k, s);
		k.s0 = 2.0 * (double)nu / (s * s) * Eta3D<REG>::scale();      // 2 nu / powf(sigma, 2), src/P3D.cpp:133
		return k;
	}
};

// ===========================================================================
// cvtx_P3D_M2M_vort     w(x_t) = 1/(4 pi sigma^3) sum_s zeta(rho) w_s   over the
// sources inside the 5-sigma box around x_t (the CPU reference's cutoff,
// src/P3D.cpp:298-322; entry :476-498).
// ===========================================================================
template <int REG> struct Zeta3D;
template <> struct Zeta3D<REG_SINGULAR> {
	static constexpr int OPS = 0, SFU = 0;
	CVTX_HD static float zeta(float, const PairConsts &) { return 0.0f; }
	static void consts(PairConsts &, double) {}
	static double scale() { return 0.0; }
};
template <> struct Zeta3D<REG_WINCKELMANS> {  // zeta = 7.5 (rho^2+1)^-7/2
	static constexpr int OPS = 5, SFU = 1;
	CVTX_HD static float zeta(float r2, const PairConsts &k) {
		const float ra = mufu_rsqrt(fmaf(r2, k.c0, 1.0f));
		const float ra2 = ra * ra, ra4 = ra2 * ra2;
		return (ra4 * ra2) * ra;
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(1.0 / (s * s)); }
	static double scale() { return 7.5; }
};
template <> struct Zeta3D<REG_PLANETARY> {    // zeta = rho < 1 ? 3 : 0
	static constexpr int OPS = 0, SFU = 0;
	CVTX_HD static float zeta(float r2, const PairConsts &k) { return r2 < k.c0 ? 1.0f : 0.0f; }
	static void consts(PairConsts &k, double s) { k.c0 = (float)(s * s); }
	static double scale() { return 3.0; }
};
template <> struct Zeta3D<REG_GAUSSIAN> {     // zeta = sqrt(2/pi) exp(-rho^2/2)
	static constexpr int OPS = 1, SFU = 1;
	CVTX_HD static float zeta(float r2, const PairConsts &k) { return mufu_ex2(r2 * k.c0); }
	static void consts(PairConsts &k, double s) { k.c0 = (float)(-0.5 * kLog2e / (s * s)); }
	static double scale() { return kSqrt2OverPi; }
};

template <int REG> struct P3DVort {
	static constexpr int NSRC4 = 2, TCOLS = 3, NTGT = 3, NACC = 3, NOUT = 3, CHAIN = 0;
	static constexpr int LANE_OPS = 9 + Zeta3D<REG>::OPS, SFU_OPS = Zeta3D<REG>::SFU;
	CVTX_HD static void load_target(const float *row, float *tg) { tg[0] = row[0]; tg[1] = row[1]; tg[2] = row[2]; }
	CVTX_HD static void pair(const float *tg, const f4 a, const f4 b, float *acc, const PairConsts &k) {
		const float dx = a.x - tg[0], dy = a.y - tg[1], dz = a.z - tg[2];
		const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
		float z = Zeta3D<REG>::zeta(r2, k);
		const bool in = fabsf(dx) < k.c3 && fabsf(dy) < k.c3 && fabsf(dz) < k.c3;   // c3 = 5 sigma
		z = in ? z : 0.0f;
		acc[0] = fmaf(z, b.x, acc[0]);
		acc[1] = fmaf(z, b.y, acc[1]);
		acc[2] = fmaf(z, b.z, acc[2]);
	}
	CVTX_HD static void finish(const float *, const double *acc, double *out, const PairConsts &k) {
		out[0] = acc[0] * k.s0; out[1] = acc[1] * k.s0; out[2] = acc[2] * k.s0;
	}
	static PairConsts make_consts(float sigma, float) {
		PairConsts k = {}; const double s = fabs((double)sigma);
		Zeta3D<REG>::consts(k, s);
		k.c3 = 5.0f * sigma;                                            // cutoff = 5.f * sigma, src/P3D.cpp:307
		k.s0 = Zeta3D<REG>::scale() / (4.0 * kPi * (double)sigma * (double)sigma * (double)sigma);
		return k;
	}
};

// ===========================================================================
// cvtx_P2D_M2M_vel      u_t = (1/2pi) sum_s g(rho) Gamma_s (rad_y, -rad_x)/r^2
// reference: src/P2D.cpp:49-69 (pair), :99-119 (sum), :141-162 (entry)
// source a = {x, y, Gamma, area}  (the 16-byte cvtx_P2D itself)
// running sums: sum K Gamma rad_y,  sum K Gamma rad_x   (sign fixed in finish)
// ===========================================================================
template <int REG> struct Reg2D;
template <> struct Reg2D<REG_SINGULAR> {      // K = 1/r^2
	static constexpr int OPS = 0, SFU = 1;
	CVTX_HD static float K(float r2, const PairConsts &) { const float ir = mufu_rcp(r2); return r2 > 0.0f ? ir : 0.0f; }
	static void consts(PairConsts &, double) {}
	static double scale(double) { return 1.0; }
};
template <> struct Reg2D<REG_WINCKELMANS> {   // K = sigma^2 g/r^2 = (rho^2+2)/(rho^2+1)^2
	static constexpr int OPS = 4, SFU = 1;
	CVTX_HD static float K(float r2, const PairConsts &k) {
		const float a = fmaf(r2, k.c0, 1.0f), b = fmaf(r2, k.c0, 2.0f);
		const float ia = mufu_rcp(a);
		return b * (ia * ia);
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(1.0 / (s * s)); }
	static double scale(double s) { return 1.0 / (s * s); }
};
template <> struct Reg2D<REG_PLANETARY> {     // K = rho < 1 ? 1/sigma^2 : 1/r^2
	static constexpr int OPS = 0, SFU = 1;
	CVTX_HD static float K(float r2, const PairConsts &k) { const float ir = mufu_rcp(r2); return r2 < k.c0 ? k.c1 : ir; }
	static void consts(PairConsts &k, double s) { k.c0 = (float)(s * s); k.c1 = (float)(1.0 / (s * s)); }
	static double scale(double) { return 1.0; }
};
template <> struct Reg2D<REG_GAUSSIAN> {      // K = (1 - exp(-rho^2/2))/r^2
	static constexpr int OPS = 2, SFU = 2;
	CVTX_HD static float K(float r2, const PairConsts &k) {
		const float e = mufu_ex2(r2 * k.c0), ir = mufu_rcp(r2);
		return r2 > 0.0f ? fmaf(-e, ir, ir) : 0.0f;
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(-0.5 * kLog2e / (s * s)); }
	static double scale(double) { return 1.0; }
};

template <int REG> struct P2DVel {
	static constexpr int NSRC4 = 1, TCOLS = 2, NTGT = 2, NACC = 2, NOUT = 2, CHAIN = 0;
	static constexpr int LANE_OPS = 7 + Reg2D<REG>::OPS, SFU_OPS = Reg2D<REG>::SFU;
	CVTX_HD static void load_target(const float *row, float *tg) { tg[0] = row[0]; tg[1] = row[1]; }
	CVTX_HD static void pair(const float *tg, const f4 a, const f4, float *acc, const PairConsts &k) {
		const float dx = tg[0] - a.x, dy = tg[1] - a.y;
		const float r2 = fmaf(dy, dy, dx * dx);
		const float kg = Reg2D<REG>::K(r2, k) * a.z;
		acc[0] = fmaf(kg, dy, acc[0]);
		acc[1] = fmaf(kg, dx, acc[1]);
	}
	CVTX_HD static void finish(const float *, const double *acc, double *out, const PairConsts &k) {
		out[0] = acc[0] * k.s0; out[1] = -(acc[1] * k.s0);
	}
	static PairConsts make_consts(float sigma, float) {
		PairConsts k = {}; const double s = fabs((double)sigma);
		Reg2D<REG>::consts(k, s);
		k.s0 = Reg2D<REG>::scale(s) / (2.0 * 3.14159265358979323846);   // 1/(2 acosf(-1)), src/P2D.cpp:118
		return k;
	}
};

// ===========================================================================
// cvtx_P2D_M2M_visc_dvort   dG_t = (2 nu/sigma^2) sum_s (G_s A_t - G_t A_s) eta2(rho)
// reference: src/P2D.cpp:167-194 (pair), :214-231 (sum), :252-276 (entry)
// ===========================================================================
template <int REG> struct Eta2D;
template <> struct Eta2D<REG_WINCKELMANS> {   // eta = 24 exp(4/a^3)/a^4, a = rho^2+1 (as coded, src/VortFunc.cpp:124-131)
	static constexpr int OPS = 8, SFU = 2;
	CVTX_HD static float eta(float r2, const PairConsts &k) {
		const float a = fmaf(r2, k.c0, 1.0f);
		float ia = mufu_rcp(a);
		// one Newton step: MUFU.RCP's ~1 ulp error is raised to the 4th power below, and the
		// PSE sum over a smooth field cancels to second order (measured on B200: 1.4e-5 from
		// FP64 without the step, against 8.6e-6 for the reference, on the 50x50 lattice test)
		ia = fmaf(ia, fmaf(-a, ia, 1.0f), ia);
		const float ia2 = ia * ia, ia3 = ia2 * ia;
		return mufu_ex2(ia3 * 5.770780163555854f) * (ia2 * ia2);       // 4 log2(e)
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(1.0 / (s * s)); }
	static double scale() { return 24.0; }
};
template <> struct Eta2D<REG_GAUSSIAN> {      // eta = exp(-rho^2/2), src/VortFunc.cpp:196-199
	static constexpr int OPS = 1, SFU = 1;
	CVTX_HD static float eta(float r2, const PairConsts &k) { return mufu_ex2(r2 * k.c0); }
	static void consts(PairConsts &k, double s) { k.c0 = (float)(-0.5 * kLog2e / (s * s)); }
	static double scale() { return 1.0; }
};

template <int REG> struct P2DVisc {
	static constexpr int NSRC4 = 1, TCOLS = 4, NTGT = 4, NACC = 2, NOUT = 1, CHAIN = 8;
	static constexpr int LANE_OPS = 8 + Eta2D<REG>::OPS, SFU_OPS = Eta2D<REG>::SFU;
	CVTX_HD static void load_target(const float *row, float *tg) { tg[0] = row[0]; tg[1] = row[1]; tg[2] = row[2]; tg[3] = row[3]; }
	CVTX_HD static void pair(const float *tg, const f4 a, const f4, float *acc, const PairConsts &k) {
		const float dx = tg[0] - a.x, dy = tg[1] - a.y;
		const float r2 = fmaf(dy, dy, dx * dx);
		float eta = Eta2D<REG>::eta(r2, k);
		eta = r2 > 0.0f ? eta : 0.0f;
		acc[0] = fmaf(eta, a.z - tg[2], acc[0]);      // sum eta (G_s - G_t)
		acc[1] = fmaf(eta, a.w - tg[3], acc[1]);      // sum eta (A_s - A_t)
	}
	CVTX_HD static void finish(const float *row, const double *acc, double *out, const PairConsts &k) {
		out[0] = k.s0 * ((double)row[3] * acc[0] - (double)row[2] * acc[1]);
	}
	static PairConsts make_consts(float sigma, float nu) {
		PairConsts k = {}; const double s = fabs((double)sigma);
		Eta2D<REG>::consts(k, s);
		k.s0 = 2.0 * (double)nu / (s * s) * Eta2D<REG>::scale();
		return k;
	}
};

// ===========================================================================
// cvtx_F3D_M2M_vel      straight singular filament a->b on a point x:
//   r1 = x-a, r2 = x-b, r0 = r1-r2, c = r1 x r2,
//   u = c [G/(4 pi |c|^2)] [r1.r0/|r1| - r2.r0/|r2|],  dropped unless both
//   bracketed factors are finite.
// reference: src/F3D.cpp:34-54 (pair), :87-107 (sum), :162-180 (entry)
// source a = {ax, ay, az, G/4pi}   b = {bx, by, bz, 3 G/(4 pi |b-a|)}
// r0 is formed per pair as r1 - r2, like the reference: the subtraction is exact
// (|r1| ~ |r2|), so r0 stays consistent with the rounded r1, r2 and the
// cancelling difference r1.r0/|r1| - r2.r0/|r2| keeps the reference's accuracy
// (a per-source r0 = b - a measured 2.4x worse against FP64).
// ===========================================================================
struct F3DVel {
	static constexpr int NSRC4 = 2, TCOLS = 3, NTGT = 3, NACC = 3, NOUT = 3, CHAIN = 0;
	static constexpr int LANE_OPS = 37, SFU_OPS = 3;
	CVTX_HD static void load_target(const float *row, float *tg) { tg[0] = row[0]; tg[1] = row[1]; tg[2] = row[2]; }
	CVTX_HD static void pair(const float *tg, const f4 a, const f4 b, float *acc, const PairConsts &) {
		const float px = tg[0] - a.x, py = tg[1] - a.y, pz = tg[2] - a.z;        // r1
		const float qx = tg[0] - b.x, qy = tg[1] - b.y, qz = tg[2] - b.z;        // r2
		const float ox = px - qx, oy = py - qy, oz = pz - qz;                    // r0 = r1 - r2
		const float cx = fmaf(py, qz, -(pz * qy));
		const float cy = fmaf(pz, qx, -(px * qz));
		const float cz = fmaf(px, qy, -(py * qx));
		const float c2 = fmaf(cz, cz, fmaf(cy, cy, cx * cx));
		const float n1 = fmaf(pz, pz, fmaf(py, py, px * px));
		const float n2 = fmaf(qz, qz, fmaf(qy, qy, qx * qx));
		const float d1 = fmaf(pz, oz, fmaf(py, oy, px * ox));
		const float d2 = fmaf(qz, oz, fmaf(qy, oy, qx * ox));
		const float t1 = a.w * mufu_rcp(c2);
		const float t2 = fmaf(d1, mufu_rsqrt(n1), -(d2 * mufu_rsqrt(n2)));
		const bool ok = (fabsf(t1) <= 3.40282346e38f) && (fabsf(t2) <= 3.40282346e38f);
		const float kk = ok ? t1 * t2 : 0.0f;
		acc[0] = fmaf(kk, cx, acc[0]);
		acc[1] = fmaf(kk, cy, acc[1]);
		acc[2] = fmaf(kk, cz, acc[2]);
	}
	CVTX_HD static void finish(const float *, const double *acc, double *out, const PairConsts &) {
		out[0] = acc[0]; out[1] = acc[1]; out[2] = acc[2];
	}
	static PairConsts make_consts(float, float) { PairConsts k = {}; k.s0 = 1.0; return k; }
};

// ===========================================================================
// cvtx_F3D_M2M_dvort    filament on particle (x_t, w_t):
//   dw = B w_t + A x w_t,  A = -r0 t1 t212/|r1 x r0|^2,  B = (3/|r0|) t1 t222
//   t212 = r0.r1/|r1| - r0.r2/|r2|,  t222 = |r0 x r1| (1/|r1| - 1/|r2|)
//   dropped when A, B (hence t212, t222) are NaN.
// reference: src/F3D.cpp:56-85 (pair), :109-128 (sum), :182-202 (entry)
// running sums: sum A (3), sum B (1); w_t applied once in finish().
// ===========================================================================
struct F3DDvort {
	static constexpr int NSRC4 = 2, TCOLS = 7, NTGT = 3, NACC = 4, NOUT = 3, CHAIN = 0;
	static constexpr int LANE_OPS = 43, SFU_OPS = 3;
	CVTX_HD static void load_target(const float *row, float *tg) { tg[0] = row[0]; tg[1] = row[1]; tg[2] = row[2]; }
	CVTX_HD static void pair(const float *tg, const f4 a, const f4 b, float *acc, const PairConsts &) {
		const float px = tg[0] - a.x, py = tg[1] - a.y, pz = tg[2] - a.z;        // r1
		const float qx = tg[0] - b.x, qy = tg[1] - b.y, qz = tg[2] - b.z;        // r2
		const float ox = px - qx, oy = py - qy, oz = pz - qz;                    // r0 = r1 - r2 (exact)
		const float xx = fmaf(py, oz, -(pz * oy));                               // X = r1 x r0
		const float xy = fmaf(pz, ox, -(px * oz));
		const float xz = fmaf(px, oy, -(py * ox));
		const float x2 = fmaf(xz, xz, fmaf(xy, xy, xx * xx));
		const float n1 = fmaf(pz, pz, fmaf(py, py, px * px));
		const float n2 = fmaf(qz, qz, fmaf(qy, qy, qx * qx));
		const float d1 = fmaf(pz, oz, fmaf(py, oy, px * ox));
		const float d2 = fmaf(qz, oz, fmaf(qy, oy, qx * ox));
		const float rs1 = mufu_rsqrt(n1), rs2 = mufu_rsqrt(n2), rsx = mufu_rsqrt(x2);
		const float t212 = fmaf(d1, rs1, -(d2 * rs2));
		const float sA = -(a.w * t212) * (rsx * rsx);
		const float t222 = (x2 * rsx) * (rs1 - rs2);
		const float Bv = b.w * t222;
		const bool ok = (sA == sA) && (Bv == Bv);
		const float sa = ok ? sA : 0.0f, bv = ok ? Bv : 0.0f;
		acc[0] = fmaf(sa, ox, acc[0]);
		acc[1] = fmaf(sa, oy, acc[1]);
		acc[2] = fmaf(sa, oz, acc[2]);
		acc[3] += bv;
	}
	CVTX_HD static void finish(const float *row, const double *acc, double *out, const PairConsts &) {
		const double wx = row[3], wy = row[4], wz = row[5];
		out[0] = acc[3] * wx + (acc[1] * wz - acc[2] * wy);
		out[1] = acc[3] * wy + (acc[2] * wx - acc[0] * wz);
		out[2] = acc[3] * wz + (acc[0] * wy - acc[1] * wx);
	}
	static PairConsts make_consts(float, float) { PairConsts k = {}; k.s0 = 1.0; return k; }
};

}  // namespace cvtx

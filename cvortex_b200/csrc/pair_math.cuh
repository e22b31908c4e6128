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
//   * target-only factors leave it too where that is numerically safe:
//       F3D dvort : sum_s (B w_t + A x w_t) = (sum B) w_t + (sum A) x w_t
//     and stay per pair where it is not (measured):
//       visc      : (w_s V_t - w_t V_s) is formed per pair.  Both hoisted forms,
//                   V_t sum(eta w_s) - w_t sum(eta V_s) and
//                   V_t sum eta (w_s - w_t) - w_t sum eta (V_s - V_t), end in a
//                   subtraction that cancels (smooth fields / random strengths)
//                   and lose up to 100x per target against the reference;
//       P3D dvort : c = w_t x w_s is formed per pair for the same reason;
//   * Winckelmans terms are division-free and regular at r = 0
//       g/r^3          = (rho^2+2.5)(rho^2+1)^-5/2 / sigma^3
//       (3g/rho^3 - zeta)/r^2 = (3rho^2+10.5)(rho^2+1)^-7/2 / sigma^2
//     so one MUFU.RSQ serves the whole pair;
//   * the only special-function work is MUFU rsqrt / rcp / ex2 (approx, ftz).
// The Gaussian g uses the same Abramowitz-Stegun 7.1.26 polynomial as the
// reference (src/VortFunc.cpp:164-173), because erff() differs from it by up
// to 1e-3 relative on near pairs.
//
// LANES.  Every formula is written once over Vec<W>, W targets side by side:
// W = 1 is plain FP32; W = 2 maps onto Blackwell's packed FP32x2 instructions
// (PTX add/mul/fma.f32x2 -> SASS FADD2 / FMUL2 / FFMA2).  One FFMA2 does two
// FMAs for one issue slot, and its scalar operands broadcast for free
// (`R25.F32`), which is exactly the shape of this loop: two targets of a
// thread against the same source record.  The FP32 pipe still retires 128
// lane-ops per SM per clock, but the kernel stops being ISSUE-bound: the
// scalar form of Winckelmans vel needs 21 FP32 + 1 MUFU + 0.6 other issue
// slots per pair (FP32 pipe <= 93 % busy at best); packed it needs 11.8.
// MUFU, compares and selects stay per lane.
//
// Coincident-pair rule (reference: contribution is exactly zero when the two
// positions compare equal, src/P3D.cpp:58,94,127, src/P2D.cpp:57,178): here a
// pair is dropped when r^2 == 0.  For formulas that are regular at r = 0 and
// carry a factor `rad` the rule holds with no test at all.
//
// GUARDS and the optimistic path.  The test costs a compare and one or two
// selects per pair, and on sm_100 those issue to the 16-lane ALU pipe: measured
// over all ops, every ALU instruction in the pair loop costs about as much as two
// FP32 lane-ops.  Where the UNguarded formula is singular
// -- rsqrt(0) / rcp(0) = inf, so the dropped pair's term comes out as inf or NaN
// -- the guard is redundant as a detector: pair<W, false>() omits it, the kernel
// looks at the FP32 running sums once per chain (256 sources) and, if any of
// them is not finite, discards the chain and evaluates it again with
// pair<W, true>().  A non-finite value can never be absorbed by later
// additions, so every chain in which a guard would have fired is re-evaluated,
// and in all other chains the two forms execute the same instructions on the
// same values: results are bit-identical with the guarded path (the tests pin
// this).  Policies advertise it as OPTIMISTIC; guards whose unguarded value is
// finite (Winckelmans stretching, the viscous ops, the planetary core, the
// vorticity box cutoff) are real selections and stay in both forms.
//
// Each policy exposes
//   NSRC4   float4 records per packed source (1 or 2)
//   TCOLS   floats per raw target row in global memory
//   NTGT    target values kept in registers during the pair loop
//   NACC    FP32 running sums per target
//   NOUT    output floats per target
//   CHAIN   sources per FP32 running-sum chain before it is flushed into the
//           FP64 accumulator (0 = one chain per 256-source shared-memory
//           tile, which every op uses today: with the pair terms formed as
//           the reference forms them, a 256-long FP32 chain measured no
//           worse than 8-long ones even where the sum cancels to second
//           order; the knob stays for ops that need it)
//   PREF_T  targets per thread that measured fastest on B200 for large problems
//           (profiles/sweep_ops_r1.txt); the planner falls back to smaller
//           tiles when there are too few targets to fill the chip
//   LANE_OPS, SFU_OPS   algorithmic FP32 lane-ops / MUFU ops per pair of THIS
//                       formulation (FMA = 1 lane-op; compares/selects not
//                       counted) -- the roofline denominators, see DESIGN.md
//   VW8, OPT8, VW4, OPT4   how m2m_kernel is instantiated for this policy in its two large-problem
//           geometries (8 targets per thread x 128 threads, 4 x 256): lanes per Vec and the M2M_* option
//           bits, read off tools/kernel_ab's table (profiles/kernel_ab_r2.txt)
//   OPTIMISTIC  pair<W, false>() exists and differs from pair<W, true>() (see GUARDS)
//   HYBRID      filament policies: fast<W, MODE>() + exact() instead of pair<>() (see FILAMENTS)
//   load_target(row, tg[]), pair<W, G>(tg, a, b, acc, k), finish(row, acc, out, k)
// and is usable from host code too (tests/hostcheck compiles this header with
// g++ to validate the algebra against the oracle without a GPU; MUFU ops are
// then replaced by libm and FP32x2 by two scalar lanes).
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

// ---- MUFU wrappers (scalar) --------------------------------------------------
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

// max(|a|, |b|, |c|), NaN if any input is: one FMNMX3 on sm_100 instead of three compares, so that
// `max3_abs(..) < c` is  |a| < c && |b| < c && |c| < c  for every input including NaNs
CVTX_HD float max3_abs(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
	float y; asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(fabsf(a)), "f"(fabsf(b)), "f"(fabsf(c))); return y;
#else
	if (a != a || b != b || c != c) return NAN;
	return fmaxf(fmaxf(fabsf(a), fabsf(b)), fabsf(c));
#endif
}

// ---- Vec<W>: W FP32 lanes -----------------------------------------------------
template <int W> struct Vec;

template <> struct Vec<1> {
	float x;
	static constexpr int LANES = 1;
	CVTX_HD float lane(int) const { return x; }
	CVTX_HD void set(int, float s) { x = s; }
};

#if !defined(__CUDACC__)
struct f2 { float x, y; };
#else
typedef float2 f2;
#endif

template <> struct Vec<2> {
	f2 v;
	static constexpr int LANES = 2;
	CVTX_HD float lane(int i) const { return i ? v.y : v.x; }
	CVTX_HD void set(int i, float s) { if (i) v.y = s; else v.x = s; }
};

// W = 4, 8: W/2 packed pairs side by side.  An operation on a Vec<W> is W/2 packed instructions in a row, so a
// formula written once comes out stage by stage across the pairs -- independent instructions next to each
// other in the source order.  ptxas schedules around the order it is given: with one pair after the other
// (W = 2 in a loop over the thread's pairs) it was seen to serialise the Horner chain of the Gaussian g and
// the rsqrt -> powers chains behind one another ("schedules").
template <int W> struct Vec {
	Vec<2> p[W / 2];
	static constexpr int LANES = W;
	CVTX_HD float lane(int i) const { return p[i >> 1].lane(i & 1); }
	CVTX_HD void set(int i, float s) { p[i >> 1].set(i & 1, s); }
};

template <int W> CVTX_HD Vec<W> bc(float s) {
	Vec<W> r;
	for (int i = 0; i < W; ++i) r.set(i, s);
	return r;
}

// fused a*b + c, a*b, a + b, a - b, -a  (FFMA2 / FMUL2 / FADD2 when W = 2;
// negations fold into the instructions' operand modifiers)
// (the scalar forms use the _rn intrinsics on the device so that nvcc never contracts a
// product into a following sum: the guarded and the optimistic pair forms, and the
// W = 1 and W = 2 forms, must round identically)
#if defined(__CUDA_ARCH__)
CVTX_HD Vec<1> vfma(Vec<1> a, Vec<1> b, Vec<1> c) { Vec<1> r; r.x = __fmaf_rn(a.x, b.x, c.x); return r; }
CVTX_HD Vec<1> vmul(Vec<1> a, Vec<1> b) { Vec<1> r; r.x = __fmul_rn(a.x, b.x); return r; }
CVTX_HD Vec<1> vadd(Vec<1> a, Vec<1> b) { Vec<1> r; r.x = __fadd_rn(a.x, b.x); return r; }
CVTX_HD Vec<1> vsub(Vec<1> a, Vec<1> b) { Vec<1> r; r.x = __fsub_rn(a.x, b.x); return r; }
#else
CVTX_HD Vec<1> vfma(Vec<1> a, Vec<1> b, Vec<1> c) { Vec<1> r; r.x = fmaf(a.x, b.x, c.x); return r; }
CVTX_HD Vec<1> vmul(Vec<1> a, Vec<1> b) { Vec<1> r; r.x = a.x * b.x; return r; }
CVTX_HD Vec<1> vadd(Vec<1> a, Vec<1> b) { Vec<1> r; r.x = a.x + b.x; return r; }
CVTX_HD Vec<1> vsub(Vec<1> a, Vec<1> b) { Vec<1> r; r.x = a.x - b.x; return r; }
#endif
CVTX_HD Vec<1> vneg(Vec<1> a) { Vec<1> r; r.x = -a.x; return r; }
CVTX_HD Vec<2> vneg(Vec<2> a) { Vec<2> r; r.v.x = -a.v.x; r.v.y = -a.v.y; return r; }
#if defined(__CUDA_ARCH__)
CVTX_HD Vec<2> vfma(Vec<2> a, Vec<2> b, Vec<2> c) { Vec<2> r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
CVTX_HD Vec<2> vmul(Vec<2> a, Vec<2> b) { Vec<2> r; r.v = __fmul2_rn(a.v, b.v); return r; }
CVTX_HD Vec<2> vadd(Vec<2> a, Vec<2> b) { Vec<2> r; r.v = __fadd2_rn(a.v, b.v); return r; }
CVTX_HD Vec<2> vsub(Vec<2> a, Vec<2> b) { Vec<2> r; r.v = __fadd2_rn(a.v, make_float2(-b.v.x, -b.v.y)); return r; }
#else
CVTX_HD Vec<2> vfma(Vec<2> a, Vec<2> b, Vec<2> c) { Vec<2> r; r.v.x = fmaf(a.v.x, b.v.x, c.v.x); r.v.y = fmaf(a.v.y, b.v.y, c.v.y); return r; }
CVTX_HD Vec<2> vmul(Vec<2> a, Vec<2> b) { Vec<2> r; r.v.x = a.v.x * b.v.x; r.v.y = a.v.y * b.v.y; return r; }
CVTX_HD Vec<2> vadd(Vec<2> a, Vec<2> b) { Vec<2> r; r.v.x = a.v.x + b.v.x; r.v.y = a.v.y + b.v.y; return r; }
CVTX_HD Vec<2> vsub(Vec<2> a, Vec<2> b) { Vec<2> r; r.v.x = a.v.x - b.v.x; r.v.y = a.v.y - b.v.y; return r; }
#endif

template <int W> CVTX_HD Vec<W> vfma(Vec<W> a, Vec<W> b, Vec<W> c) { Vec<W> r; for (int i = 0; i < W / 2; ++i) r.p[i] = vfma(a.p[i], b.p[i], c.p[i]); return r; }
template <int W> CVTX_HD Vec<W> vmul(Vec<W> a, Vec<W> b) { Vec<W> r; for (int i = 0; i < W / 2; ++i) r.p[i] = vmul(a.p[i], b.p[i]); return r; }
template <int W> CVTX_HD Vec<W> vadd(Vec<W> a, Vec<W> b) { Vec<W> r; for (int i = 0; i < W / 2; ++i) r.p[i] = vadd(a.p[i], b.p[i]); return r; }
template <int W> CVTX_HD Vec<W> vsub(Vec<W> a, Vec<W> b) { Vec<W> r; for (int i = 0; i < W / 2; ++i) r.p[i] = vsub(a.p[i], b.p[i]); return r; }
template <int W> CVTX_HD Vec<W> vneg(Vec<W> a) { Vec<W> r; for (int i = 0; i < W / 2; ++i) r.p[i] = vneg(a.p[i]); return r; }

// scalar-operand forms (the scalar broadcasts inside the instruction)
template <int W> CVTX_HD Vec<W> vfma(Vec<W> a, float b, Vec<W> c) { return vfma(a, bc<W>(b), c); }
template <int W> CVTX_HD Vec<W> vfma(Vec<W> a, Vec<W> b, float c) { return vfma(a, b, bc<W>(c)); }
template <int W> CVTX_HD Vec<W> vfma(Vec<W> a, float b, float c) { return vfma(a, bc<W>(b), bc<W>(c)); }
template <int W> CVTX_HD Vec<W> vmul(Vec<W> a, float b) { return vmul(a, bc<W>(b)); }
template <int W> CVTX_HD Vec<W> vsub(Vec<W> a, float b) { return vsub(a, bc<W>(b)); }
template <int W> CVTX_HD Vec<W> vsub(float a, Vec<W> b) { return vsub(bc<W>(a), b); }
// a*b - c
template <int W> CVTX_HD Vec<W> vfms(Vec<W> a, Vec<W> b, Vec<W> c) { return vfma(a, b, vneg(c)); }
template <int W> CVTX_HD Vec<W> vfms(Vec<W> a, float b, Vec<W> c) { return vfma(a, bc<W>(b), vneg(c)); }

// per-lane operations (MUFU, compares, selects have no packed form)
template <int W> CVTX_HD Vec<W> vrsqrt(Vec<W> a) { Vec<W> r; for (int i = 0; i < W; ++i) r.set(i, mufu_rsqrt(a.lane(i))); return r; }
template <int W> CVTX_HD Vec<W> vrcp(Vec<W> a) { Vec<W> r; for (int i = 0; i < W; ++i) r.set(i, mufu_rcp(a.lane(i))); return r; }
template <int W> CVTX_HD Vec<W> vex2(Vec<W> a) { Vec<W> r; for (int i = 0; i < W; ++i) r.set(i, mufu_ex2(a.lane(i))); return r; }
// c == 0 ? 0 : v     (the coincident-pair rule; c = r^2 >= 0).  Written as the rule itself, not as c > 0: a NaN
// r^2 -- a NaN coordinate in a source or a target -- must keep its term, as the reference's exact-equality test
// does (bsv_V3f_isequal: NaN != NaN), so that a diverging simulation shows up as NaN and not as a finite result.
template <int W> CVTX_HD Vec<W> keep_if_pos(Vec<W> c, Vec<W> v) {
	Vec<W> r;
	for (int i = 0; i < W; ++i) r.set(i, c.lane(i) == 0.0f ? 0.0f : v.lane(i));
	return r;
}
// the same rule where the unguarded v is inf / NaN whenever the rule would fire: G = false
// (optimistic path) hands v through and leaves the detection to the kernel
template <bool G, int W> CVTX_HD Vec<W> drop_if_coincident(Vec<W> c, Vec<W> v) {
	if (G) return keep_if_pos(c, v);
	return v;
}
// c < thr ? a : b
template <int W> CVTX_HD Vec<W> pick_if_less(Vec<W> c, float thr, Vec<W> a, Vec<W> b) {
	Vec<W> r;
	for (int i = 0; i < W; ++i) r.set(i, c.lane(i) < thr ? a.lane(i) : b.lane(i));
	return r;
}

// per-regularisation pick for the TUNE constants below: (singular, winckelmans, planetary, gaussian)
constexpr int by_reg(int reg, int s, int w, int p, int g) { return reg == REG_SINGULAR ? s : (reg == REG_WINCKELMANS ? w : (reg == REG_PLANETARY ? p : g)); }

static constexpr double kPi = 3.14159265359;              // CVTX_PI_F, reference src/P3D.cpp:47 (as a float literal there)
static constexpr double kSqrt2OverPi = 0.7978845608028654;  // reference src/VortFunc.cpp:49
static constexpr double kRecipSqrt2 = 0.7071067811865475;   // reference src/VortFunc.cpp:50
static constexpr double kLog2e = 1.4426950408889634;

// Abramowitz & Stegun 7.1.26, coefficients as in reference src/VortFunc.cpp:165-166.
// Returns s with  g_gauss3D(rho) = 1 - e * s,  e = exp(-rho^2/2):
//   erf(z) ~ 1 - poly(t) e,  t = 1/(1 + p z),  z = rho/sqrt2
//   g = erf(z) - rho sqrt(2/pi) e = 1 - e (poly(t) + rho sqrt(2/pi))
template <int W> CVTX_HD Vec<W> gauss_tail(Vec<W> r, float k_t, float k_c) {
	const Vec<W> t = vrcp(vfma(r, k_t, 1.0f));
	Vec<W> p = vfma(t, 1.061405429f, -1.453152027f);
	p = vfma(t, p, 1.421413741f);
	p = vfma(t, p, -0.284496736f);
	p = vfma(t, p, 0.254829592f);
	return vfma(r, k_c, vmul(p, t));
}

// ---------------------------------------------------------------------------
// 3D kernels:   A(r2)  ~ g(rho)/r^3           (velocity / first stretching term)
//               Bn(r2) = -(3 g/rho^3 - zeta)/r^2 in the SAME units as A, so that
//                        one accumulator set takes  A c + Bn (rad.c) rad.
//                        Bn is handed back as two factors B1 * B2 and the caller
//                        forms (B1 * (rad.c)) * B2: r^-5 alone leaves the FP32
//                        range for r < 2e-8 or r > 4e7, the reassociated product
//                        does not (the reference works in rho and is scale-free)
// Units are chosen per regularisation so the loop needs the fewest lane-ops;
// the matching sigma power sits in PairConsts::s0 (scaleA).
// ---------------------------------------------------------------------------
template <int REG> struct Reg3D;

template <> struct Reg3D<REG_WINCKELMANS> {
	// c0 = 1/sigma^2, c1 = -3/sigma^4, c2 = -10.5/sigma^2.
	// A = sigma^3 g/r^3 = g/rho^3.  Bn = -(3 rho^2 + 10.5)(rho^2+1)^-7/2 / sigma^2.
	static constexpr int A_OPS = 6, AB_OPS = 9, SFU = 1;
	static constexpr bool POISONS = false;      // regular at r = 0: nothing to detect, the one guard below is a real selection
	template <int W, bool G> CVTX_HD static Vec<W> A(Vec<W> r2, const PairConsts &k) {
		const Vec<W> a = vfma(r2, k.c0, 1.0f), b = vfma(r2, k.c0, 2.5f);
		const Vec<W> ra = vrsqrt(a), ra2 = vmul(ra, ra), ra4 = vmul(ra2, ra2);
		return vmul(b, vmul(ra4, ra));
	}
	template <int W, bool G> CVTX_HD static void AB(Vec<W> r2, const PairConsts &k, Vec<W> &A_, Vec<W> &B1, Vec<W> &B2) {
		const Vec<W> a = vfma(r2, k.c0, 1.0f), b = vfma(r2, k.c0, 2.5f), b2 = vfma(r2, k.c1, k.c2);
		const Vec<W> ra = vrsqrt(a), ra2 = vmul(ra, ra), ra4 = vmul(ra2, ra2), ra5 = vmul(ra4, ra);
		// the self pair must give exactly 0: c = w_t x w_t formed with FMAs is
		// only zero to rounding, and A(0) = 2.5 would amplify that residue
		A_ = keep_if_pos(r2, vmul(b, ra5));
		B1 = b2;
		B2 = vmul(ra5, ra2);
	}
	static void consts(PairConsts &k, double s) {
		k.c0 = (float)(1.0 / (s * s)); k.c1 = (float)(-3.0 / (s * s * s * s)); k.c2 = (float)(-10.5 / (s * s));
	}
	static double scaleA(double s) { return 1.0 / (s * s * s); }
};

template <> struct Reg3D<REG_SINGULAR> {
	// A = 1/r^3, Bn = -3/r^5; both dropped at r = 0.
	static constexpr int A_OPS = 2, AB_OPS = 4, SFU = 1;
	static constexpr bool POISONS = true;       // unguarded at r = 0: rsqrt(0) = inf in every factor
	template <int W, bool G> CVTX_HD static Vec<W> A(Vec<W> r2, const PairConsts &) {
		const Vec<W> ri = vrsqrt(r2);
		return drop_if_coincident<G>(r2, vmul(vmul(ri, ri), ri));
	}
	template <int W, bool G> CVTX_HD static void AB(Vec<W> r2, const PairConsts &, Vec<W> &A_, Vec<W> &B1, Vec<W> &B2) {
		const Vec<W> ri = vrsqrt(r2), ri2 = vmul(ri, ri), ri3 = vmul(ri2, ri);
		A_ = drop_if_coincident<G>(r2, ri3);
		B1 = drop_if_coincident<G>(r2, vmul(ri2, -3.0f));
		B2 = A_;
	}
	static void consts(PairConsts &, double) {}
	static double scaleA(double) { return 1.0; }
};

template <> struct Reg3D<REG_PLANETARY> {
	// rho < 1: g = rho^3, zeta = 3  ->  A = 1/sigma^3, Bn = 0;  else singular.
	// c0 = sigma^2, c1 = 1/sigma^3.
	static constexpr int A_OPS = 2, AB_OPS = 4, SFU = 1;
	static constexpr bool POISONS = false;      // the core test is a selection between two finite values
	template <int W, bool G> CVTX_HD static Vec<W> A(Vec<W> r2, const PairConsts &k) {
		const Vec<W> ri = vrsqrt(r2);
		return pick_if_less(r2, k.c0, bc<W>(k.c1), vmul(vmul(ri, ri), ri));
	}
	template <int W, bool G> CVTX_HD static void AB(Vec<W> r2, const PairConsts &k, Vec<W> &A_, Vec<W> &B1, Vec<W> &B2) {
		const Vec<W> ri = vrsqrt(r2), ri2 = vmul(ri, ri), ri3 = vmul(ri2, ri);
		A_ = pick_if_less(r2, k.c0, keep_if_pos(r2, bc<W>(k.c1)), ri3);        // exact 0 for the self pair, as above
		B1 = pick_if_less(r2, k.c0, bc<W>(0.0f), vmul(ri2, -3.0f));
		B2 = A_;                                                               // finite in the core, where B1 = 0
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(s * s); k.c1 = (float)(1.0 / (s * s * s)); }
	static double scaleA(double) { return 1.0; }
};

template <> struct Reg3D<REG_GAUSSIAN> {
	// c0 = p/(sqrt2 sigma), c1 = -log2(e)/(2 sigma^2), c2 = sqrt(2/pi)/sigma,
	// c3 = sqrt(2/pi)/sigma^3.   A = g/r^3,  Bn = (c3 e - 3A)/r^2.
	static constexpr int A_OPS = 13, AB_OPS = 16, SFU = 3;
	static constexpr bool POISONS = true;       // unguarded at r = 0: r = 0 * rsqrt(0) = NaN, carried by every factor
	template <int W, bool G> CVTX_HD static Vec<W> A(Vec<W> r2, const PairConsts &k) {
		const Vec<W> ri = vrsqrt(r2), r = vmul(r2, ri);
		const Vec<W> e = vex2(vmul(r2, k.c1));
		const Vec<W> s = gauss_tail(r, k.c0, k.c2);
		const Vec<W> g = vfma(vneg(e), s, 1.0f);
		return drop_if_coincident<G>(r2, vmul(g, vmul(vmul(ri, ri), ri)));
	}
	template <int W, bool G> CVTX_HD static void AB(Vec<W> r2, const PairConsts &k, Vec<W> &A_, Vec<W> &B1, Vec<W> &B2) {
		const Vec<W> ri = vrsqrt(r2), r = vmul(r2, ri), ri2 = vmul(ri, ri);
		const Vec<W> e = vex2(vmul(r2, k.c1));
		const Vec<W> s = gauss_tail(r, k.c0, k.c2);
		const Vec<W> g = vfma(vneg(e), s, 1.0f);
		const Vec<W> a = vmul(g, vmul(ri2, ri));
		A_ = drop_if_coincident<G>(r2, a);
		B1 = vfma(A_, -3.0f, vmul(e, k.c3));                                   // from the guarded A: finite at r = 0
		B2 = drop_if_coincident<G>(r2, ri2);
	}
	static void consts(PairConsts &k, double s) {
		k.c0 = (float)(0.3275911 * kRecipSqrt2 / s);
		k.c1 = (float)(-0.5 * kLog2e / (s * s));
		k.c2 = (float)(kSqrt2OverPi / s);
		k.c3 = (float)(kSqrt2OverPi / (s * s * s));
	}
	static double scaleA(double) { return 1.0; }
};

// rad = target - source position, r^2 (6 lane-ops); shared by the 3D particle ops
template <int W> struct Rad3 { Vec<W> x, y, z, r2; };
template <int W> CVTX_HD Rad3<W> rad3(const Vec<W> *tg, const f4 a) {
	Rad3<W> d;
	d.x = vsub(tg[0], a.x); d.y = vsub(tg[1], a.y); d.z = vsub(tg[2], a.z);
	d.r2 = vfma(d.z, d.z, vfma(d.y, d.y, vmul(d.x, d.x)));
	return d;
}

// ===========================================================================
// cvtx_P3D_M2M_vel      u_t = -(1/4pi) sum_s [g(rho)/r^3] (rad x w_s)
// reference: src/P3D.cpp:51-72 (pair), :230-251 (sum), :343-366 (entry)
// source  a = {x, y, z, vol}   b = {wx, wy, wz, 0}
// ===========================================================================
template <int REG> struct P3DVel {
	static constexpr int NSRC4 = 2, TCOLS = 3, NTGT = 3, NACC = 3, NOUT = 3, CHAIN = 0, PREF_T = 8;
	static constexpr int LANE_OPS = 15 + Reg3D<REG>::A_OPS, SFU_OPS = Reg3D<REG>::SFU;
	static constexpr bool OPTIMISTIC = Reg3D<REG>::POISONS;      // K = inf / NaN meets finite factors: every sum is poisoned
	static constexpr bool HYBRID = false;
	// TUNE (profiles/kernel_ab_r2.txt): lanes per Vec and M2M_* option bits of the 8 x 128 and 4 x 256 geometries
	static constexpr int VW8 = by_reg(REG, 8, 2, 8, 2), OPT8 = by_reg(REG, 0, 1, 0, 3);
	static constexpr int VW4 = by_reg(REG, 2, 2, 4, 2), OPT4 = by_reg(REG, 1, 1, 1, 1);
	CVTX_HD static void load_target(const float *row, float *tg) { tg[0] = row[0]; tg[1] = row[1]; tg[2] = row[2]; }
	template <int W, bool G = true> CVTX_HD static void pair(const Vec<W> *tg, const f4 a, const f4 b, Vec<W> *acc, const PairConsts &k) {
		const Rad3<W> d = rad3(tg, a);
		const Vec<W> K = Reg3D<REG>::template A<W, G>(d.r2, k);
		const Vec<W> cx = vfms(d.y, b.z, vmul(d.z, b.y));
		const Vec<W> cy = vfms(d.z, b.x, vmul(d.x, b.z));
		const Vec<W> cz = vfms(d.x, b.y, vmul(d.y, b.x));
		acc[0] = vfma(K, cx, acc[0]);
		acc[1] = vfma(K, cy, acc[1]);
		acc[2] = vfma(K, cz, acc[2]);
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
	// (4 targets per thread x 256 threads measured 0.5 - 1.7 % faster than 8 x 128 for all but the Gaussian, profiles/kernel_ab_r2.txt)
	static constexpr int NSRC4 = 2, TCOLS = 7, NTGT = 6, NACC = 3, NOUT = 3, CHAIN = 0, PREF_T = REG == REG_GAUSSIAN ? 8 : 4;
	static constexpr int LANE_OPS = 22 + Reg3D<REG>::AB_OPS, SFU_OPS = Reg3D<REG>::SFU;
	static constexpr bool OPTIMISTIC = Reg3D<REG>::POISONS;      // A = inf / NaN enters all three sums through fma(A, c, .)
	static constexpr bool HYBRID = false;
	// TUNE (profiles/kernel_ab_r2.txt): lanes per Vec and M2M_* option bits of the 8 x 128 and 4 x 256 geometries
	static constexpr int VW8 = by_reg(REG, 8, 8, 4, 8), OPT8 = by_reg(REG, 0, 0, 1, 0);
	static constexpr int VW4 = by_reg(REG, 4, 2, 4, 2), OPT4 = by_reg(REG, 1, 1, 0, 1);
	CVTX_HD static void load_target(const float *row, float *tg) {
		for (int i = 0; i < 6; ++i) tg[i] = row[i];
	}
	template <int W, bool G = true> CVTX_HD static void pair(const Vec<W> *tg, const f4 a, const f4 b, Vec<W> *acc, const PairConsts &k) {
		const Rad3<W> d = rad3(tg, a);
		Vec<W> A, B1, B2;
		Reg3D<REG>::template AB<W, G>(d.r2, k, A, B1, B2);
		const Vec<W> cx = vfms(tg[4], b.z, vmul(tg[5], b.y));
		const Vec<W> cy = vfms(tg[5], b.x, vmul(tg[3], b.z));
		const Vec<W> cz = vfms(tg[3], b.y, vmul(tg[4], b.x));
		const Vec<W> trip = vfma(d.z, cz, vfma(d.y, cy, vmul(d.x, cx)));
		const Vec<W> s = vmul(vmul(B1, trip), B2);
		acc[0] = vfma(s, d.x, vfma(A, cx, acc[0]));
		acc[1] = vfma(s, d.y, vfma(A, cy, acc[1]));
		acc[2] = vfma(s, d.z, vfma(A, cz, acc[2]));
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
// Fused cvtx_P3D_M2M_vel + cvtx_P3D_M2M_dvort on the SAME particle targets, one
// pass over the sources (SURVEY 7 step 4 / 8d config 2 "fused pass"): a time
// stepper needs u(x_t) and dw_t for the same particles, and the two sums share
// rad, r^2, the regularisation factor A = g/r^3 and the MUFU work.  Not an
// entry point of the reference ABI (there vel takes bare points); reachable
// through the thin ABI as CVTX_B200_P3D_VEL_DVORT.  out row = {u (3), dw (3)}.
// ===========================================================================
template <int REG> struct P3DVelDvort {
	static constexpr int NSRC4 = 2, TCOLS = 7, NTGT = 6, NACC = 6, NOUT = 6, CHAIN = 0, PREF_T = 4;
	static constexpr int LANE_OPS = 31 + Reg3D<REG>::AB_OPS, SFU_OPS = Reg3D<REG>::SFU;
	static constexpr bool OPTIMISTIC = Reg3D<REG>::POISONS;
	static constexpr bool HYBRID = false;
	// TUNE (profiles/kernel_ab_r2.txt): lanes per Vec and M2M_* option bits of the 8 x 128 and 4 x 256 geometries
	static constexpr int VW8 = by_reg(REG, 8, 8, 2, 2), OPT8 = by_reg(REG, 3, 3, 1, 1);
	static constexpr int VW4 = by_reg(REG, 4, 2, 2, 2), OPT4 = by_reg(REG, 1, 1, 3, 1);
	CVTX_HD static void load_target(const float *row, float *tg) {
		for (int i = 0; i < 6; ++i) tg[i] = row[i];
	}
	template <int W, bool G = true> CVTX_HD static void pair(const Vec<W> *tg, const f4 a, const f4 b, Vec<W> *acc, const PairConsts &k) {
		const Rad3<W> d = rad3(tg, a);
		Vec<W> A, B1, B2;
		Reg3D<REG>::template AB<W, G>(d.r2, k, A, B1, B2);
		// velocity: A (rad x w_s)          (A is zeroed at r = 0, where rad x w_s = 0 anyway)
		const Vec<W> ux = vfms(d.y, b.z, vmul(d.z, b.y));
		const Vec<W> uy = vfms(d.z, b.x, vmul(d.x, b.z));
		const Vec<W> uz = vfms(d.x, b.y, vmul(d.y, b.x));
		acc[0] = vfma(A, ux, acc[0]);
		acc[1] = vfma(A, uy, acc[1]);
		acc[2] = vfma(A, uz, acc[2]);
		// stretching: A c + Bn (rad.c) rad,  c = w_t x w_s
		const Vec<W> cx = vfms(tg[4], b.z, vmul(tg[5], b.y));
		const Vec<W> cy = vfms(tg[5], b.x, vmul(tg[3], b.z));
		const Vec<W> cz = vfms(tg[3], b.y, vmul(tg[4], b.x));
		const Vec<W> trip = vfma(d.z, cz, vfma(d.y, cy, vmul(d.x, cx)));
		const Vec<W> s = vmul(vmul(B1, trip), B2);
		acc[3] = vfma(s, d.x, vfma(A, cx, acc[3]));
		acc[4] = vfma(s, d.y, vfma(A, cy, acc[4]));
		acc[5] = vfma(s, d.z, vfma(A, cz, acc[5]));
	}
	CVTX_HD static void finish(const float *, const double *acc, double *out, const PairConsts &k) {
		out[0] = acc[0] * k.s1; out[1] = acc[1] * k.s1; out[2] = acc[2] * k.s1;
		out[3] = acc[3] * k.s0; out[4] = acc[4] * k.s0; out[5] = acc[5] * k.s0;
	}
	static PairConsts make_consts(float sigma, float nu) {
		PairConsts k = P3DDvort<REG>::make_consts(sigma, nu);     // s0: stretching scale (signed sigma^3)
		k.s1 = P3DVel<REG>::make_consts(sigma, nu).s0;            // s1: velocity scale -scaleA/(4 pi)
		return k;
	}
};

// ===========================================================================
// cvtx_P3D_M2M_visc_dvort   dw_t = (2 nu/sigma^2) sum_s (w_s V_t - w_t V_s) eta(rho)
// reference: src/P3D.cpp:116-144 (pair), :275-296 (sum), :432-456 (entry)
// running sums: the three components of the result itself; only
// Winckelmans / Gaussian have an eta (src/VortFunc.cpp:101-108, :245).
// ===========================================================================
template <int REG> struct Eta3D;
template <> struct Eta3D<REG_WINCKELMANS> {   // eta = 52.5 (rho^2+1)^-9/2
	static constexpr int OPS = 5, SFU = 1;
	template <int W> CVTX_HD static Vec<W> eta(Vec<W> r2, const PairConsts &k) {
		const Vec<W> ra = vrsqrt(vfma(r2, k.c0, 1.0f));
		const Vec<W> ra2 = vmul(ra, ra), ra4 = vmul(ra2, ra2), ra8 = vmul(ra4, ra4);
		return vmul(ra8, ra);
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(1.0 / (s * s)); }
	static double scale() { return 52.5; }
};
template <> struct Eta3D<REG_GAUSSIAN> {      // eta = sqrt(2/pi) exp(-rho^2/2)
	static constexpr int OPS = 1, SFU = 1;
	template <int W> CVTX_HD static Vec<W> eta(Vec<W> r2, const PairConsts &k) { return vex2(vmul(r2, k.c0)); }
	static void consts(PairConsts &k, double s) { k.c0 = (float)(-0.5 * kLog2e / (s * s)); }
	static double scale() { return kSqrt2OverPi; }
};

template <int REG> struct P3DVisc {
	static constexpr int NSRC4 = 2, TCOLS = 7, NTGT = 7, NACC = 3, NOUT = 3, CHAIN = 0, PREF_T = REG == REG_GAUSSIAN ? 4 : 8;
	static constexpr int LANE_OPS = 15 + Eta3D<REG>::OPS, SFU_OPS = Eta3D<REG>::SFU;
	static constexpr bool OPTIMISTIC = false;      // eta(0) is finite: the coincident-pair test is a real selection
	static constexpr bool HYBRID = false;
	// TUNE (profiles/kernel_ab_r2.txt): lanes per Vec and M2M_* option bits of the 8 x 128 and 4 x 256 geometries
	static constexpr int VW8 = by_reg(REG, 2, 8, 2, 2), OPT8 = by_reg(REG, 0, 1, 0, 1);
	static constexpr int VW4 = by_reg(REG, 2, 2, 2, 2), OPT4 = by_reg(REG, 0, 1, 0, 1);
	CVTX_HD static void load_target(const float *row, float *tg) {
		for (int i = 0; i < 7; ++i) tg[i] = row[i];
	}
	template <int W, bool G = true> CVTX_HD static void pair(const Vec<W> *tg, const f4 a, const f4 b, Vec<W> *acc, const PairConsts &k) {
		const Rad3<W> d = rad3(tg, a);
		const Vec<W> eta = keep_if_pos(d.r2, Eta3D<REG>::eta(d.r2, k));   // coincident pair contributes nothing
		// w_s V_t - w_t V_s per pair, one product rounded and one FMA (the reference rounds both
		// products and the sum, src/P3D.cpp:136-140)
		acc[0] = vfma(eta, vfms(tg[6], b.x, vmul(tg[3], a.w)), acc[0]);
		acc[1] = vfma(eta, vfms(tg[6], b.y, vmul(tg[4], a.w)), acc[1]);
		acc[2] = vfma(eta, vfms(tg[6], b.z, vmul(tg[5], a.w)), acc[2]);
	}
	CVTX_HD static void finish(const float *, const double *acc, double *out, const PairConsts &k) {
		out[0] = k.s0 * acc[0]; out[1] = k.s0 * acc[1]; out[2] = k.s0 * acc[2];
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
	template <int W> CVTX_HD static Vec<W> zeta(Vec<W>, const PairConsts &) { return bc<W>(0.0f); }
	static void consts(PairConsts &, double) {}
	static double scale() { return 0.0; }
};
template <> struct Zeta3D<REG_WINCKELMANS> {  // zeta = 7.5 (rho^2+1)^-7/2
	static constexpr int OPS = 5, SFU = 1;
	template <int W> CVTX_HD static Vec<W> zeta(Vec<W> r2, const PairConsts &k) {
		const Vec<W> ra = vrsqrt(vfma(r2, k.c0, 1.0f));
		const Vec<W> ra2 = vmul(ra, ra), ra4 = vmul(ra2, ra2);
		return vmul(vmul(ra4, ra2), ra);
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(1.0 / (s * s)); }
	static double scale() { return 7.5; }
};
template <> struct Zeta3D<REG_PLANETARY> {    // zeta = rho < 1 ? 3 : 0
	static constexpr int OPS = 0, SFU = 0;
	template <int W> CVTX_HD static Vec<W> zeta(Vec<W> r2, const PairConsts &k) {
		return pick_if_less(r2, k.c0, bc<W>(1.0f), bc<W>(0.0f));
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(s * s); }
	static double scale() { return 3.0; }
};
template <> struct Zeta3D<REG_GAUSSIAN> {     // zeta = sqrt(2/pi) exp(-rho^2/2)
	static constexpr int OPS = 1, SFU = 1;
	template <int W> CVTX_HD static Vec<W> zeta(Vec<W> r2, const PairConsts &k) { return vex2(vmul(r2, k.c0)); }
	static void consts(PairConsts &k, double s) { k.c0 = (float)(-0.5 * kLog2e / (s * s)); }
	static double scale() { return kSqrt2OverPi; }
};

template <int REG> struct P3DVort {
	static constexpr int NSRC4 = 2, TCOLS = 3, NTGT = 3, NACC = 3, NOUT = 3, CHAIN = 0, PREF_T = 8;
	static constexpr int LANE_OPS = 9 + Zeta3D<REG>::OPS, SFU_OPS = Zeta3D<REG>::SFU;
	static constexpr bool OPTIMISTIC = false;      // the box cutoff selects between finite values
	static constexpr bool HYBRID = false;
	// TUNE (profiles/kernel_ab_r2.txt): lanes per Vec and M2M_* option bits of the 8 x 128 and 4 x 256 geometries
	static constexpr int VW8 = by_reg(REG, 2, 2, 8, 4), OPT8 = by_reg(REG, 0, 1, 0, 0);
	static constexpr int VW4 = by_reg(REG, 2, 2, 2, 4), OPT4 = by_reg(REG, 0, 1, 1, 1);
	CVTX_HD static void load_target(const float *row, float *tg) { tg[0] = row[0]; tg[1] = row[1]; tg[2] = row[2]; }
	template <int W, bool G = true> CVTX_HD static void pair(const Vec<W> *tg, const f4 a, const f4 b, Vec<W> *acc, const PairConsts &k) {
		const Rad3<W> d = rad3(tg, a);
		const Vec<W> zeta = Zeta3D<REG>::zeta(d.r2, k);
		Vec<W> z;
		for (int i = 0; i < W; ++i) {                          // c3 = 5 sigma: the reference's box cutoff
			const bool in = max3_abs(d.x.lane(i), d.y.lane(i), d.z.lane(i)) < k.c3;
			z.set(i, in ? zeta.lane(i) : 0.0f);
		}
		acc[0] = vfma(z, b.x, acc[0]);
		acc[1] = vfma(z, b.y, acc[1]);
		acc[2] = vfma(z, b.z, acc[2]);
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
	static constexpr bool POISONS = true;       // unguarded at r = 0: rcp(0) = inf, times Gamma, times dx = dy = 0
	template <int W, bool G> CVTX_HD static Vec<W> K(Vec<W> r2, const PairConsts &) { return drop_if_coincident<G>(r2, vrcp(r2)); }
	static void consts(PairConsts &, double) {}
	static double scale(double) { return 1.0; }
};
template <> struct Reg2D<REG_WINCKELMANS> {   // K = sigma^2 g/r^2 = (rho^2+2)/(rho^2+1)^2
	static constexpr int OPS = 4, SFU = 1;
	static constexpr bool POISONS = false;
	template <int W, bool G> CVTX_HD static Vec<W> K(Vec<W> r2, const PairConsts &k) {
		const Vec<W> a = vfma(r2, k.c0, 1.0f), b = vfma(r2, k.c0, 2.0f);
		const Vec<W> ia = vrcp(a);
		return vmul(b, vmul(ia, ia));
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(1.0 / (s * s)); }
	static double scale(double s) { return 1.0 / (s * s); }
};
template <> struct Reg2D<REG_PLANETARY> {     // K = rho < 1 ? 1/sigma^2 : 1/r^2
	static constexpr int OPS = 0, SFU = 1;
	static constexpr bool POISONS = false;
	template <int W, bool G> CVTX_HD static Vec<W> K(Vec<W> r2, const PairConsts &k) {
		return pick_if_less(r2, k.c0, bc<W>(k.c1), vrcp(r2));
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(s * s); k.c1 = (float)(1.0 / (s * s)); }
	static double scale(double) { return 1.0; }
};
template <> struct Reg2D<REG_GAUSSIAN> {      // K = (1 - exp(-rho^2/2))/r^2
	static constexpr int OPS = 2, SFU = 2;
	static constexpr bool POISONS = true;       // unguarded at r = 0: fma(-1, inf, inf) = NaN
	template <int W, bool G> CVTX_HD static Vec<W> K(Vec<W> r2, const PairConsts &k) {
		const Vec<W> e = vex2(vmul(r2, k.c0)), ir = vrcp(r2);
		return drop_if_coincident<G>(r2, vfma(vneg(e), ir, ir));
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(-0.5 * kLog2e / (s * s)); }
	static double scale(double) { return 1.0; }
};

template <int REG> struct P2DVel {
	static constexpr int NSRC4 = 1, TCOLS = 2, NTGT = 2, NACC = 2, NOUT = 2, CHAIN = 0, PREF_T = 4;
	static constexpr int LANE_OPS = 7 + Reg2D<REG>::OPS, SFU_OPS = Reg2D<REG>::SFU;
	static constexpr bool OPTIMISTIC = Reg2D<REG>::POISONS;
	static constexpr bool HYBRID = false;
	// TUNE (profiles/kernel_ab_r2.txt): lanes per Vec and M2M_* option bits of the 8 x 128 and 4 x 256 geometries
	static constexpr int VW8 = by_reg(REG, 2, 8, 4, 2), OPT8 = by_reg(REG, 0, 2, 1, 2);
	static constexpr int VW4 = by_reg(REG, 4, 2, 2, 2), OPT4 = by_reg(REG, 3, 1, 1, 0);
	CVTX_HD static void load_target(const float *row, float *tg) { tg[0] = row[0]; tg[1] = row[1]; }
	template <int W, bool G = true> CVTX_HD static void pair(const Vec<W> *tg, const f4 a, const f4, Vec<W> *acc, const PairConsts &k) {
		const Vec<W> dx = vsub(tg[0], a.x), dy = vsub(tg[1], a.y);
		const Vec<W> r2 = vfma(dy, dy, vmul(dx, dx));
		const Vec<W> kg = vmul(Reg2D<REG>::template K<W, G>(r2, k), a.z);
		acc[0] = vfma(kg, dy, acc[0]);
		acc[1] = vfma(kg, dx, acc[1]);
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
	static constexpr int OPS = 6, SFU = 2;
	template <int W> CVTX_HD static Vec<W> eta(Vec<W> r2, const PairConsts &k) {
		const Vec<W> ia = vrcp(vfma(r2, k.c0, 1.0f));
		const Vec<W> ia2 = vmul(ia, ia), ia3 = vmul(ia2, ia);
		return vmul(vex2(vmul(ia3, 5.770780163555854f)), vmul(ia2, ia2));       // 4 log2(e)
	}
	static void consts(PairConsts &k, double s) { k.c0 = (float)(1.0 / (s * s)); }
	static double scale() { return 24.0; }
};
template <> struct Eta2D<REG_GAUSSIAN> {      // eta = exp(-rho^2/2), src/VortFunc.cpp:196-199
	static constexpr int OPS = 1, SFU = 1;
	template <int W> CVTX_HD static Vec<W> eta(Vec<W> r2, const PairConsts &k) { return vex2(vmul(r2, k.c0)); }
	static void consts(PairConsts &k, double s) { k.c0 = (float)(-0.5 * kLog2e / (s * s)); }
	static double scale() { return 1.0; }
};

template <int REG> struct P2DVisc {
	static constexpr int NSRC4 = 1, TCOLS = 4, NTGT = 4, NACC = 1, NOUT = 1, CHAIN = 0, PREF_T = REG == REG_WINCKELMANS ? 4 : 8;
	static constexpr int LANE_OPS = 7 + Eta2D<REG>::OPS, SFU_OPS = Eta2D<REG>::SFU;
	static constexpr bool OPTIMISTIC = false;      // eta(0) is finite
	static constexpr bool HYBRID = false;
	// TUNE (profiles/kernel_ab_r2.txt): lanes per Vec and M2M_* option bits of the 8 x 128 and 4 x 256 geometries
	static constexpr int VW8 = by_reg(REG, 2, 8, 2, 2), OPT8 = by_reg(REG, 0, 0, 0, 0);
	static constexpr int VW4 = by_reg(REG, 2, 2, 2, 2), OPT4 = by_reg(REG, 0, 0, 0, 1);
	CVTX_HD static void load_target(const float *row, float *tg) { tg[0] = row[0]; tg[1] = row[1]; tg[2] = row[2]; tg[3] = row[3]; }
	template <int W, bool G = true> CVTX_HD static void pair(const Vec<W> *tg, const f4 a, const f4, Vec<W> *acc, const PairConsts &k) {
		const Vec<W> dx = vsub(tg[0], a.x), dy = vsub(tg[1], a.y);
		const Vec<W> r2 = vfma(dy, dy, vmul(dx, dx));
		const Vec<W> eta = keep_if_pos(r2, Eta2D<REG>::eta(r2, k));
		acc[0] = vfma(eta, vfms(tg[3], a.z, vmul(tg[2], a.w)), acc[0]);      // eta (G_s A_t - G_t A_s)
	}
	CVTX_HD static void finish(const float *, const double *acc, double *out, const PairConsts &k) {
		out[0] = k.s0 * acc[0];
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
// cancelling difference r1.r0/|r1| - r2.r0/|r2| keeps the reference's accuracy.
// ===========================================================================
// u x v the way the reference's bsv_V3f_cross rounds it: both products rounded, then one rounded
// subtraction (9 lane-ops, not the 6 of fma(uy, vz, -(uz vy))).  The filament formulas divide by
// |r1 x r2|^2 (|r1 x r0|^2) and drop the pair when that is 0, and whether it IS 0 for a point on the
// filament's own line -- a segment's midpoint, the next node of a straight vortex line -- is decided
// by this rounding: with equal products the reference gets an exact 0 and drops the pair, while a
// fused cross product returns the rounding residue of one product (1e-9 ... 1e-16 relative), i.e. a
// finite 1/|c|^2 of 1e18+ and a "velocity" of 1e8 where the reference returns 0.
// `one` must be a RUN-TIME 1.0f: ptxas (12.9) fuses a packed product into a following packed sum or
// difference even when both carry .rn (and folds a literal 1.0 first), so the subtraction is issued
// as fma(p, one, -m), which it cannot take apart.
template <int W> CVTX_HD void cross_rounded(Vec<W> ux, Vec<W> uy, Vec<W> uz, Vec<W> vx, Vec<W> vy, Vec<W> vz, float one,
                                            Vec<W> &cx, Vec<W> &cy, Vec<W> &cz) {
	cx = vfma(vmul(uy, vz), one, vneg(vmul(uz, vy)));
	cy = vfma(vmul(uz, vx), one, vneg(vmul(ux, vz)));
	cz = vfma(vmul(ux, vy), one, vneg(vmul(uy, vx)));
}

// ---------------------------------------------------------------------------
// FILAMENTS, second version: three tiers per (target, 32-source sub-chain).
//
// Where the FP32 error of the reference's filament formulas comes from, for a segment of
// length l seen from distance r at angle theta: (i) c = r1 x r2 is a cross product of nearly
// parallel vectors, relative error ~ 1e-7 r/(l sin theta), and c/|c|^2 carries it three times;
// (ii) t2 = r1.r0/|r1| - r2.r0/|r2| cancels to ~ l^2 sin^2(theta)/r, relative error
// ~ 1e-7 r/(l sin^2 theta).  Both grow as the segments get short -- the FP32 reference itself
// sits 1e-6 ... 1e-5 from FP64 on short segments -- and (ii) makes it noise next to the axis.
// With |c|^2 = (|r1||r2| - r1.r2)(|r1||r2| + r1.r2) and
// t2 = (|r1| + |r2|)(|r1||r2| - r1.r2)/(|r1||r2|) the singular factor cancels exactly:
//     Kf := t2/|c|^2 = (1/|r1| + 1/|r2|) / (|r1||r2| + r1.r2)
// a sum of positive terms wherever r1.r2 > 0, i.e. outside the sphere whose diameter is the
// segment; and c = r1 x r2 = r0 x r1 with the per-SOURCE r0 = b - a (rounded once) is a cross
// product of non-parallel vectors.  Likewise 1/|r1| - 1/|r2| = -(|r1|^2 - |r2|^2) / ((|r1| + |r2|)|r1||r2|)
// with |r1|^2 - |r2|^2 = r0.(r1 + r2) = 2 r0.r1 - l^2.  That is the FAST form (`fast<W, F3D_NEW>`): ~2e-7
// from FP64 at any segment length, 34 / 41 lane-ops instead of 40 / 45.
// It is not usable (a) inside that sphere, where |r1||r2| + r1.r2 cancels and the reference's form
// is the accurate one, and (b) within ~1e-3 rad of the filament's axis, where the reference's result
// is decided by how ITS operations round (exact zeros for collinear points, or garbage / |c|) and a
// drop-in replacement has to return what the reference returns.  Each fast pair therefore also
// feeds a per-TARGET running minimum `flag` of  min(r1.r2, |c|^2 - tau l^2 |r1|^2)  (NaN-propagating);
// after every sub-chain of F3D_SUB = 32 sources the kernel looks at it and, for the targets whose
// flag is not positive, throws the sub-chain's sums away and evaluates those 32 pairs again with
// `exact()`: the reference's own operations in the reference's order with IEEE division and square
// root -- op for op what src/F3D.cpp:34-85 does, so points on a vortex line get the reference's bits.
// The decision is per target and per fixed sub-chain of the source order, never per thread: a result
// does not depend on which other targets share a thread, on the launch geometry or on the shard.
// A call whose filaments are long against the region they live in (a vortex ring, a lifting line:
// most filament sets in practice) would take the slow tier all the time.  For those the fast form
// selects per pair (`fast<W, F3D_WIDE>`): outside the sphere 1/(|r1||r2| + r1.r2) as above, inside it
// the same quantity as (|r1||r2| - r1.r2)/|c|^2 -- a sum of positive terms THERE over the cross
// product's own square -- so nothing cancels anywhere off the axis and only the axis test feeds the
// flag; one more MUFU and a select per pair.  (The first version of this kernel evaluated the
// reference's formula with MUFU in this regime: its t2 cancels for short segments and amplified the
// 1 - 2 ulp of MUFU.RSQ to 2 - 3x the reference's own distance from FP64 -- 4e-5 on one short filament
// seen from afar.)  Which of the two is a property of the SOURCES alone (f3d_pick_mode, decided on the
// device while packing), so every shard of a multi-GPU call makes the same choice.
// ---------------------------------------------------------------------------
enum F3DMode { F3D_NEW = 0, F3D_WIDE = 1 };
constexpr int F3D_SUB = 32;

// running per-lane minimum that keeps a NaN once it has seen one (one FMNMX3.NAN on sm_100)
CVTX_HD float min3_nan(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
	float y; asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y;
#else
	if (a != a || b != b || c != c) return NAN;
	return fminf(fminf(a, b), c);
#endif
}
CVTX_HD float min2_nan(float a, float b) {
#if defined(__CUDA_ARCH__)
	float y; asm("min.NaN.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b)); return y;
#else
	if (a != a || b != b) return NAN;
	return fminf(a, b);
#endif
}

// single rounded FP32 operations, immune to FMA contraction (the host build uses -ffp-contract=off)
#if defined(__CUDA_ARCH__)
CVTX_HD float rn_mul(float a, float b) { return __fmul_rn(a, b); }
CVTX_HD float rn_add(float a, float b) { return __fadd_rn(a, b); }
CVTX_HD float rn_sub(float a, float b) { return __fsub_rn(a, b); }
CVTX_HD float rn_div(float a, float b) { return __fdiv_rn(a, b); }
CVTX_HD float rn_sqrt(float a) { return __fsqrt_rn(a); }
#else
CVTX_HD float rn_mul(float a, float b) { return a * b; }
CVTX_HD float rn_add(float a, float b) { return a + b; }
CVTX_HD float rn_sub(float a, float b) { return a - b; }
CVTX_HD float rn_div(float a, float b) { return a / b; }
CVTX_HD float rn_sqrt(float a) { return sqrtf(a); }
#endif
struct V3r { float x, y, z; };
CVTX_HD V3r r_sub(V3r a, V3r b) { V3r r = {rn_sub(a.x, b.x), rn_sub(a.y, b.y), rn_sub(a.z, b.z)}; return r; }
CVTX_HD float r_dot(V3r a, V3r b) { return rn_add(rn_add(rn_mul(a.x, b.x), rn_mul(a.y, b.y)), rn_mul(a.z, b.z)); }
CVTX_HD float r_abs(V3r a) { return rn_sqrt(r_dot(a, a)); }
CVTX_HD V3r r_cross(V3r a, V3r b) {
	V3r r = {rn_sub(rn_mul(a.y, b.z), rn_mul(a.z, b.y)), rn_sub(rn_mul(a.z, b.x), rn_mul(a.x, b.z)),
	         rn_sub(rn_mul(a.x, b.y), rn_mul(a.y, b.x))};
	return r;
}

// Per-call choice of the fast form from the filaments alone: the expected share of (filament, point)
// pairs with the point inside the filament's sphere, for points spread over the filaments' own
// bounding box (every extent at least the longest filament).  A warp that meets one such pair
// re-evaluates a 32-source sub-chain in scalar code, ~3x the cost of the sub-chain itself, for
// 32 lanes x 8 targets x 32 sources = 8192 pairs: at p = 4e-6 that is +10 %.
// A small filament set takes the cancellation-free form whatever its shape: its bounding box says nothing about
// where the points are (one short filament seen from afar is the extreme), the form is 10x more accurate there
// than any that rounds r1 x r2, and the worst case -- every pair inside a sphere, every sub-chain re-evaluated --
// is ten times the cost of a launch that is microseconds long.
constexpr double kF3DSmallSet = 1024.0;
CVTX_HD int f3d_pick_mode(double sum_len3, double n, const float *lo, const float *hi, float max_len) {
	if (!(n > 0.0)) return F3D_WIDE;
	if (n <= kF3DSmallSet) return F3D_NEW;
	double vol = 1.0;
	for (int i = 0; i < 3; ++i) {
		const double e = (double)hi[i] - (double)lo[i];
		vol *= e > (double)max_len ? e : (double)max_len;
	}
	if (!(vol > 0.0)) return F3D_WIDE;
	const double p = 0.5235987755982988 * sum_len3 / (n * vol);      // (pi/6) <l^3> / V
	return p < 4e-6 ? F3D_NEW : F3D_WIDE;
}

// ===========================================================================
// cvtx_F3D_M2M_vel      straight singular filament a->b on a point x:
//   r1 = x-a, r2 = x-b, r0 = r1-r2, c = r1 x r2,
//   u = c [G/(4 pi |c|^2)] [r1.r0/|r1| - r2.r0/|r2|],  dropped unless both
//   bracketed factors are finite.
// reference: src/F3D.cpp:34-54 (pair), :87-107 (sum), :162-180 (entry)
// source  a = {ax, ay, az, G/4pi}   b = {bx, by, bz, 3 G/(4 pi |b-a|)}   c = {b - a, |b-a|^2}
// ===========================================================================
struct F3DVel {
	static constexpr int NSRC4 = 3, TCOLS = 3, NTGT = 3, NACC = 3, NOUT = 3, CHAIN = 0, PREF_T = 4;
	static constexpr int LANE_OPS = 34, SFU_OPS = 3;           // of fast<W, F3D_NEW>; the F3D_WIDE form: 39 / 4
	static constexpr bool OPTIMISTIC = false, HYBRID = true;
	// TUNE (profiles/kernel_ab_f3d_r2.txt): 4 targets per thread, 850 against 840 Gpair/s with 8; 5 = shared-memory
	// accumulators + the reference-arithmetic tier once per chain (M2M_DEFER_EXACT)
	static constexpr int VW8 = 8, OPT8 = 5, VW4 = 2, OPT4 = 5;
	CVTX_HD static void load_target(const float *row, float *tg) { tg[0] = row[0]; tg[1] = row[1]; tg[2] = row[2]; }

	template <int W, int MODE> CVTX_HD static void fast(const Vec<W> *tg, const f4 a, const f4 b, const f4 c, Vec<W> *acc,
	                                                   Vec<W> &flag, const PairConsts &k) {
		const Vec<W> px = vsub(tg[0], a.x), py = vsub(tg[1], a.y), pz = vsub(tg[2], a.z);          // r1
		const Vec<W> n1 = vfma(pz, pz, vfma(py, py, vmul(px, px)));
		const float tl = c.w * k.c1;                                                                // tau l^2
		if (MODE == F3D_NEW) {
			const Vec<W> qx = vsub(px, c.x), qy = vsub(py, c.y), qz = vsub(pz, c.z);               // r2 = r1 - r0
			const Vec<W> n2 = vfma(qz, qz, vfma(qy, qy, vmul(qx, qx)));
			const Vec<W> d12 = vfma(pz, qz, vfma(py, qy, vmul(px, qx)));                          // r1 . r2
			const Vec<W> cx = vfms(pz, c.y, vmul(py, c.z));                                       // r0 x r1 = r1 x r2
			const Vec<W> cy = vfms(px, c.z, vmul(pz, c.x));
			const Vec<W> cz = vfms(py, c.x, vmul(px, c.y));
			const Vec<W> c2 = vfma(cz, cz, vfma(cy, cy, vmul(cx, cx)));
			const Vec<W> rs1 = vrsqrt(n1), rs2 = vrsqrt(n2);
			const Vec<W> md = vfma(vmul(n1, rs1), vmul(n2, rs2), d12);                            // |r1||r2| + r1.r2
			const Vec<W> kk = vmul(vmul(vadd(rs1, rs2), a.w), vrcp(md));                          // G/4pi Kf
			acc[0] = vfma(kk, cx, acc[0]);
			acc[1] = vfma(kk, cy, acc[1]);
			acc[2] = vfma(kk, cz, acc[2]);
			const Vec<W> margin = vfma(n1, -tl, c2);                                              // > 0: off the axis
			for (int i = 0; i < W; ++i) flag.set(i, min3_nan(flag.lane(i), d12.lane(i), margin.lane(i)));
		} else {
			// the reference's geometry -- r2 = x - b, c = r1 x r2 rounded as it rounds it -- so that next to an end
			// point and next to the axis the error of c is the reference's own; the scalar factor t2/|c|^2 without
			// its cancellation: 1/(|r1||r2| + r1.r2) outside the sphere, (|r1||r2| - r1.r2)/|c|^2 inside
			const Vec<W> qx = vsub(tg[0], b.x), qy = vsub(tg[1], b.y), qz = vsub(tg[2], b.z);     // r2
			const Vec<W> n2 = vfma(qz, qz, vfma(qy, qy, vmul(qx, qx)));
			const Vec<W> d12 = vfma(pz, qz, vfma(py, qy, vmul(px, qx)));                          // r1 . r2
			Vec<W> cx, cy, cz;
			cross_rounded(px, py, pz, qx, qy, qz, k.c0, cx, cy, cz);                              // c = r1 x r2
			const Vec<W> c2 = vfma(cz, cz, vfma(cy, cy, vmul(cx, cx)));
			const Vec<W> rs1 = vrsqrt(n1), rs2 = vrsqrt(n2);
			// (|r1||r2| +- r1.r2 as explicit FMAs: ptxas fuses a packed product into a following packed sum and leaves
			// the scalar one alone -- the T = 1 geometry would round differently from the others)
			const Vec<W> q1 = vmul(n1, rs1), q2 = vmul(n2, rs2);
			const Vec<W> out = vrcp(vfma(q1, q2, d12)), in = vmul(vfms(q1, q2, d12), vrcp(c2));
			// a target ON an end point has c = 0 and n1 = 0 (or r2 = 0): margin = 0 or -tl n1, never positive
			const Vec<W> margin = vfma(n1, -tl, c2);
			Vec<W> im;
			for (int i = 0; i < W; ++i) {
				im.set(i, d12.lane(i) > 0.0f ? out.lane(i) : in.lane(i));
				flag.set(i, min2_nan(flag.lane(i), margin.lane(i)));
			}
			const Vec<W> kk = vmul(vmul(vadd(rs1, rs2), a.w), im);                                // G/4pi Kf
			acc[0] = vfma(kk, cx, acc[0]);
			acc[1] = vfma(kk, cy, acc[1]);
			acc[2] = vfma(kk, cz, acc[2]);
		}
	}

	// One pair in the reference's own arithmetic (src/F3D.cpp:34-54): raw filament row, raw target row.
	CVTX_HD static void exact(const float *fil, const float *tgt, float *acc) {
		const V3r x = {tgt[0], tgt[1], tgt[2]}, a = {fil[0], fil[1], fil[2]}, b = {fil[3], fil[4], fil[5]};
		const V3r r1 = r_sub(x, a), r2 = r_sub(x, b), r0 = r_sub(r1, r2), c = r_cross(r1, r2);
		const float nc = r_abs(c);
		const float t1 = rn_div(fil[6], rn_mul(4.0f * 3.14159265359f, rn_mul(nc, nc)));      // powf(|c|, 2)
		const float t21 = rn_div(r_dot(r1, r0), r_abs(r1)), t22 = rn_div(r_dot(r2, r0), r_abs(r2));
		const float t2 = rn_sub(t21, t22);
		if (fabsf(t1) <= 3.40282346e38f && fabsf(t2) <= 3.40282346e38f) {
			const float s = rn_mul(t1, t2);
			acc[0] = rn_add(acc[0], rn_mul(c.x, s));
			acc[1] = rn_add(acc[1], rn_mul(c.y, s));
			acc[2] = rn_add(acc[2], rn_mul(c.z, s));
		}
	}
	CVTX_HD static void finish(const float *, const double *acc, double *out, const PairConsts &) {
		out[0] = acc[0]; out[1] = acc[1]; out[2] = acc[2];
	}
	// c0: cross_rounded's `one`; c1: tau, the axis test |c|^2 < tau l^2 |r1|^2 (sin(theta) < 3e-5)
	static PairConsts make_consts(float, float) { PairConsts k = {}; k.c0 = 1.0f; k.c1 = 1e-9f; k.s0 = 1.0; return k; }
};

// ===========================================================================
// cvtx_F3D_M2M_dvort    filament on particle (x_t, w_t):
//   dw = B w_t + A x w_t,  A = -r0 t1 t212/|r1 x r0|^2,  B = (3/|r0|) t1 t222
//   t212 = r0.r1/|r1| - r0.r2/|r2|,  t222 = |r0 x r1| (1/|r1| - 1/|r2|)
//   dropped when the result, t212 or t222 is NaN.
// reference: src/F3D.cpp:56-85 (pair), :109-128 (sum), :182-202 (entry)
// running sums: sum A (3), sum B (1); w_t applied once in finish().
// fast form: A = -r0 t1 Kf (Kf as above),  B = -(3 t1/l) |X| (2 r0.r1 - l^2) / ((|r1| + |r2|)|r1||r2|),
//            |X|^2 = l^2 |r1|^2 - (r0.r1)^2
// ===========================================================================
struct F3DDvort {
	static constexpr int NSRC4 = 3, TCOLS = 7, NTGT = 3, NACC = 4, NOUT = 3, CHAIN = 0, PREF_T = 4;
	static constexpr int LANE_OPS = 41, SFU_OPS = 4;           // of fast<W, F3D_NEW>; the F3D_WIDE form: 55 / 5
	static constexpr bool OPTIMISTIC = false, HYBRID = true;
	// TUNE (profiles/kernel_ab_f3d_r2.txt): 4 targets per thread, 705 against 672 Gpair/s with 8
	static constexpr int VW8 = 8, OPT8 = 5, VW4 = 2, OPT4 = 5;
	CVTX_HD static void load_target(const float *row, float *tg) { tg[0] = row[0]; tg[1] = row[1]; tg[2] = row[2]; }

	template <int W, int MODE> CVTX_HD static void fast(const Vec<W> *tg, const f4 a, const f4 b, const f4 c, Vec<W> *acc,
	                                                   Vec<W> &flag, const PairConsts &k) {
		const Vec<W> px = vsub(tg[0], a.x), py = vsub(tg[1], a.y), pz = vsub(tg[2], a.z);          // r1
		const Vec<W> n1 = vfma(pz, pz, vfma(py, py, vmul(px, px)));
		if (MODE == F3D_NEW) {
			const Vec<W> qx = vsub(px, c.x), qy = vsub(py, c.y), qz = vsub(pz, c.z);               // r2 = r1 - r0
			const Vec<W> n2 = vfma(qz, qz, vfma(qy, qy, vmul(qx, qx)));
			const Vec<W> d12 = vfma(pz, qz, vfma(py, qy, vmul(px, qx)));                          // r1 . r2
			const Vec<W> d1 = vfma(pz, c.z, vfma(py, c.y, vmul(px, c.x)));                        // r0 . r1
			const Vec<W> ln = vmul(n1, c.w);                                                      // l^2 |r1|^2
			const Vec<W> x2 = vfma(vneg(d1), d1, ln);                                             // |r0 x r1|^2
			const Vec<W> rs1 = vrsqrt(n1), rs2 = vrsqrt(n2), rsx = vrsqrt(x2);
			const Vec<W> q1 = vmul(n1, rs1), q2 = vmul(n2, rs2);
			const Vec<W> md = vfma(q1, q2, d12), qs = vadd(q1, q2);
			const Vec<W> R = vrcp(vmul(md, qs)), P = vmul(rs1, rs2);
			const Vec<W> iq = vmul(R, md);                                                        // 1/(|r1| + |r2|)
			const Vec<W> sa = vmul(vmul(qs, P), vmul(vmul(qs, R), -a.w));                         // -t1 Kf
			acc[0] = vfma(sa, c.x, acc[0]);
			acc[1] = vfma(sa, c.y, acc[1]);
			acc[2] = vfma(sa, c.z, acc[2]);
			const Vec<W> s12 = vfma(d1, 2.0f, -c.w);                                              // r0 . (r1 + r2)
			const Vec<W> bv = vmul(vmul(vmul(x2, rsx), P), vmul(s12, iq));
			acc[3] = vfma(bv, -b.w, acc[3]);
			const Vec<W> margin = vfma(ln, -k.c1, x2);
			for (int i = 0; i < W; ++i) flag.set(i, min3_nan(flag.lane(i), d12.lane(i), margin.lane(i)));
		} else {
			// the reference's geometry (r2 = x - b, r0 = r1 - r2, X = r1 x r0 rounded as it rounds it), the scalar
			// factors without their cancellations (see F3DVel)
			const Vec<W> qx = vsub(tg[0], b.x), qy = vsub(tg[1], b.y), qz = vsub(tg[2], b.z);     // r2
			const Vec<W> ox = vsub(px, qx), oy = vsub(py, qy), oz = vsub(pz, qz);                 // r0 = r1 - r2
			const Vec<W> n2 = vfma(qz, qz, vfma(qy, qy, vmul(qx, qx)));
			const Vec<W> d12 = vfma(pz, qz, vfma(py, qy, vmul(px, qx)));                          // r1 . r2
			const Vec<W> d1 = vfma(pz, oz, vfma(py, oy, vmul(px, ox)));                           // r0 . r1
			Vec<W> xx, xy, xz;
			cross_rounded(px, py, pz, ox, oy, oz, k.c0, xx, xy, xz);                              // X = r1 x r0
			const Vec<W> x2 = vfma(xz, xz, vfma(xy, xy, vmul(xx, xx)));
			const Vec<W> rs1 = vrsqrt(n1), rs2 = vrsqrt(n2), rsx = vrsqrt(x2);
			const Vec<W> q1 = vmul(n1, rs1), q2 = vmul(n2, rs2);
			const Vec<W> qs = vadd(q1, q2), P = vmul(rs1, rs2);
			const Vec<W> iq = vrcp(qs);                                                           // 1/(|r1| + |r2|)
			const Vec<W> out = vrcp(vfma(q1, q2, d12)), in = vmul(vfms(q1, q2, d12), vmul(rsx, rsx));      // t212/(|X|^2 (1/|r1| + 1/|r2|))
			const Vec<W> margin = vfma(vmul(n1, c.w), -k.c1, x2);
			Vec<W> im;
			for (int i = 0; i < W; ++i) {
				im.set(i, d12.lane(i) > 0.0f ? out.lane(i) : in.lane(i));
				flag.set(i, min2_nan(flag.lane(i), margin.lane(i)));
			}
			const Vec<W> sa = vmul(vmul(qs, P), vmul(im, -a.w));                                  // -t1 Kf
			acc[0] = vfma(sa, ox, acc[0]);
			acc[1] = vfma(sa, oy, acc[1]);
			acc[2] = vfma(sa, oz, acc[2]);
			const Vec<W> s12 = vfma(d1, 2.0f, -c.w);                                              // r0 . (r1 + r2) = |r1|^2 - |r2|^2
			const Vec<W> bv = vmul(vmul(vmul(x2, rsx), P), vmul(s12, iq));                        // -t222
			acc[3] = vfma(bv, -b.w, acc[3]);
		}
	}

	// One pair in the reference's own arithmetic (src/F3D.cpp:56-85): the pair is dropped on the
	// reference's test (result, t222 or t212 NaN), otherwise its A and B join the running sums.
	CVTX_HD static void exact(const float *fil, const float *tgt, float *acc) {
		const V3r x = {tgt[0], tgt[1], tgt[2]}, w = {tgt[3], tgt[4], tgt[5]};
		const V3r a = {fil[0], fil[1], fil[2]}, b = {fil[3], fil[4], fil[5]};
		const V3r r1 = r_sub(x, a), r2 = r_sub(x, b), r0 = r_sub(r1, r2);
		const float t1 = rn_div(fil[6], 4.0f * 3.14159265359f);
		const float nx = r_abs(r_cross(r1, r0));
		const float den = -rn_mul(nx, nx);                                                       // -powf(|r1 x r0|, 2)
		const V3r t211 = {rn_div(r0.x, den), rn_div(r0.y, den), rn_div(r0.z, den)};
		const float n1 = r_abs(r1), n2 = r_abs(r2);
		const float t2121 = rn_div(r_dot(r0, r1), n1), t2122 = rn_div(-r_dot(r0, r2), n2);
		const float t221 = rn_div(3.0f, r_abs(r0));
		const float ny = r_abs(r_cross(r0, r1));
		const float t2221 = rn_div(ny, n1), t2222 = rn_div(-ny, n2);
		const float t222 = rn_add(t2221, t2222), t212 = rn_add(t2121, t2122);
		const float sc = rn_mul(t1, t212);
		const V3r A = {rn_mul(t211.x, sc), rn_mul(t211.y, sc), rn_mul(t211.z, sc)};
		const float B = rn_mul(rn_mul(t221, t1), t222);
		const V3r cr = r_cross(A, w);
		const float rx = rn_add(rn_mul(w.x, B), cr.x), ry = rn_add(rn_mul(w.y, B), cr.y), rz = rn_add(rn_mul(w.z, B), cr.z);
		if (rx == rx && ry == ry && rz == rz && t222 == t222 && t212 == t212) {
			acc[0] = rn_add(acc[0], A.x);
			acc[1] = rn_add(acc[1], A.y);
			acc[2] = rn_add(acc[2], A.z);
			acc[3] = rn_add(acc[3], B);
		}
	}
	CVTX_HD static void finish(const float *row, const double *acc, double *out, const PairConsts &) {
		const double wx = row[3], wy = row[4], wz = row[5];
		out[0] = acc[3] * wx + (acc[1] * wz - acc[2] * wy);
		out[1] = acc[3] * wy + (acc[2] * wx - acc[0] * wz);
		out[2] = acc[3] * wz + (acc[0] * wy - acc[1] * wx);
	}
	// c0: cross_rounded's `one`; c1: tau of the axis test |X|^2 < tau l^2 |r1|^2 (sin(theta) < 1e-3: the
	// fast form's |X|^2 = l^2 |r1|^2 - (r0.r1)^2 resolves sin^2(theta) to ~3e-7)
	static PairConsts make_consts(float, float) { PairConsts k = {}; k.c0 = 1.0f; k.c1 = 1e-6f; k.s0 = 1.0; return k; }
};

}  // namespace cvtx

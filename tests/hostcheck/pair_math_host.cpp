// pair_math_host.cpp -- TEST-ONLY host build of the device pair arithmetic.
//
// Compiles cvortex_b200/csrc/pair_math.cuh + op_table.h with g++ (MUFU ops
// replaced by libm) and runs the *same summation structure* the CUDA kernel
// uses -- sources packed into float4 records and zero-padded to whole tiles,
// FP32 partial sums per chain (Policy::CHAIN sources, or a 256-source tile) flushed into FP64, Policy::finish() --
// so the algebraic rewrites (bilinear hoisting, division-free Winckelmans,
// packed filament records, coincident-pair rule) can be validated against the
// oracle on a machine without a GPU.  It is not part of the product and is
// never a fallback for it: it lives under tests/ and is built by
// tests/conftest.py into tests/_build/.
#include <vector>
#include <cstring>
#include "../../cvortex_b200/csrc/op_table.h"

using namespace cvtx;

namespace {
int g_guarded_only = 0;      // 1: every chain in the guarded form (what cvtx_b200_guarded_only(1) does on the device)
int g_f3d_mode = -1;         // -1: f3d_pick_mode decides (as on the device); 0 / 1 pin the filament fast form
long g_reevaluated = 0;      // chains the optimistic form had to hand back / filament sub-chains sent to exact()
int g_last_f3d_mode = -1;

// The per-call choice of the filament fast form, from the raw rows (device: pack_sources_kernel + f3d_mode_kernel).
int pick_f3d_mode(const float *src, int n) {
	if (g_f3d_mode == 0 || g_f3d_mode == 1) return g_f3d_mode;
	double sum = 0.0; float mx = 0.0f, lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
	for (long j = 0; j < n; ++j) {
		f4 a, b, c;
		pack_source(SRC_F3D, src + 7 * j, a, b, c);
		const float len = sqrtf(c.w);
		sum += (double)len * len * len; mx = fmaxf(mx, len);
		const float A[3] = {a.x, a.y, a.z}, Bv[3] = {b.x, b.y, b.z};
		for (int d = 0; d < 3; ++d) { lo[d] = fminf(lo[d], fminf(A[d], Bv[d])); hi[d] = fmaxf(hi[d], fmaxf(A[d], Bv[d])); }
	}
	return f3d_pick_mode(sum, (double)n, lo, hi, mx);
}

// W = lanes per Vec: 1 = scalar FP32, 2 = the packed FP32x2 form the kernel uses for even T
template <int W> struct Runner {
	const float *src; int n; const float *tgt; int m; float *out; int op; float sigma, nu;
	template <class P> void run() {
		const int S = 256;
		const int kind = src_kind(op), cols = src_cols(op);
		const long npad = ((long)n + S - 1) / S * S;
		std::vector<f4> A(npad), B(npad), C(npad);
		for (long j = n; j < npad; ++j) pad_source(kind, A[j], B[j], C[j]);
		for (long j = 0; j < n; ++j) pack_source(kind, src + cols * j, A[j], B[j], C[j]);
		const PairConsts k = P::make_consts(sigma, nu);
		int mode = F3D_WIDE;
		if constexpr (P::HYBRID) { mode = pick_f3d_mode(src, n); g_last_f3d_mode = mode; }
#pragma omp parallel for schedule(static)
		for (long i0 = 0; i0 < m; i0 += W) {
			Vec<W> tg[P::NTGT];
			const float *rows[W];
			for (int l = 0; l < W; ++l) {
				const long i = i0 + l < m ? i0 + l : m - 1;        // odd tail: repeat the last target
				rows[l] = tgt + (long)P::TCOLS * i;
				float one[P::NTGT];
				P::load_target(rows[l], one);
				for (int c = 0; c < P::NTGT; ++c) tg[c].set(l, one[c]);
			}
			double dacc[W][P::NACC];
			for (int l = 0; l < W; ++l) for (int c = 0; c < P::NACC; ++c) dacc[l][c] = 0.0;
#ifdef HOSTCHECK_CHAIN
			const int chain = HOSTCHECK_CHAIN;      // experiment: one chain length for every op
#else
			const int chain = npad / S < 64 ? 32 : S;   // the kernel's grain (device_api.cu make_plan: kSmallSourceTiles)
#endif
			for (long t0 = 0; t0 < npad; t0 += chain) {
				Vec<W> acc[P::NACC];
				for (int c = 0; c < P::NACC; ++c) acc[c] = bc<W>(0.0f);
				if constexpr (P::HYBRID) {
					// the kernel's filament tiers (m2m_kernel.cuh): fast form over 32 sources with a per-target flag,
					// exact() for the targets whose flag fired
					for (long s0 = t0; s0 < t0 + chain; s0 += F3D_SUB) {
						Vec<W> sub[P::NACC], flag = bc<W>(3.0e38f);
						for (int c = 0; c < P::NACC; ++c) sub[c] = bc<W>(0.0f);
						for (long j = s0; j < s0 + F3D_SUB; ++j) {
							if (mode == F3D_NEW) P::template fast<W, F3D_NEW>(tg, A[j], B[j], C[j], sub, flag, k);
							else P::template fast<W, F3D_WIDE>(tg, A[j], B[j], C[j], sub, flag, k);
						}
						for (int l = 0; l < W; ++l) {
							if (flag.lane(l) > 0.0f) continue;
							float e[P::NACC];
							for (int c = 0; c < P::NACC; ++c) e[c] = 0.0f;
							for (long j = s0; j < s0 + F3D_SUB && j < n; ++j) P::exact(src + 7 * j, rows[l], e);
							for (int c = 0; c < P::NACC; ++c) sub[c].set(l, e[c]);
#pragma omp atomic
							++g_reevaluated;
						}
						for (int c = 0; c < P::NACC; ++c) acc[c] = vadd(acc[c], sub[c]);
					}
				} else {
					// the kernel's optimistic chain (m2m_kernel.cuh): unguarded pair form, one finiteness
					// check over the running sums, guarded re-evaluation of the chain if it fails
					bool guarded = true;
					if (P::OPTIMISTIC && !g_guarded_only) {
						for (long j = t0; j < t0 + chain; ++j) P::template pair<W, false>(tg, A[j], B[j], acc, k);
						Vec<W> chk = acc[0];
						for (int c = 1; c < P::NACC; ++c) chk = vadd(chk, acc[c]);
						const float s = W == 2 ? chk.lane(0) + chk.lane(1) : chk.lane(0);
						guarded = !(fabsf(s) <= 3.40282346e38f);
						if (guarded) {
							for (int c = 0; c < P::NACC; ++c) acc[c] = bc<W>(0.0f);
#pragma omp atomic
							++g_reevaluated;
						}
					}
					if (guarded)
						for (long j = t0; j < t0 + chain; ++j) P::template pair<W, true>(tg, A[j], B[j], acc, k);
				}
				for (int l = 0; l < W; ++l) for (int c = 0; c < P::NACC; ++c) dacc[l][c] += (double)acc[c].lane(l);
			}
			for (int l = 0; l < W && i0 + l < m; ++l) {
				double res[P::NOUT];
				P::finish(rows[l], dacc[l], res, k);
				for (int c = 0; c < P::NOUT; ++c) out[(long)P::NOUT * (i0 + l) + c] = (float)res[c];
			}
		}
	}
};
struct Meta {
	int *v;
	template <class P> void run() { v[0] = P::LANE_OPS; v[1] = P::SFU_OPS; v[2] = P::TCOLS; v[3] = P::NOUT; v[4] = P::NACC; v[5] = P::NSRC4; }
};
struct Optimistic {
	int v;
	template <class P> void run() { v = P::OPTIMISTIC ? 1 : 0; }
};
}  // namespace

extern "C" int hostcheck_m2m(int op, int reg, const float *src, int n, const float *tgt, int m,
                             float *out, float sigma, float nu)
{
	Runner<2> r = {src, n, tgt, m, out, op, sigma, nu};
	return dispatch_op(op, reg, r) ? 0 : -1;
}

extern "C" int hostcheck_m2m_scalar(int op, int reg, const float *src, int n, const float *tgt, int m,
                                    float *out, float sigma, float nu)
{
	Runner<1> r = {src, n, tgt, m, out, op, sigma, nu};
	return dispatch_op(op, reg, r) ? 0 : -1;
}

extern "C" void hostcheck_guarded_only(int on) { g_guarded_only = on; }
extern "C" void hostcheck_f3d_mode(int mode) { g_f3d_mode = mode; }
extern "C" int hostcheck_last_f3d_mode(void) { return g_last_f3d_mode; }
extern "C" long hostcheck_reevaluated(int reset)
{
	const long v = g_reevaluated;
	if (reset) g_reevaluated = 0;
	return v;
}
extern "C" int hostcheck_optimistic(int op, int reg)
{
	Optimistic q = {0};
	return dispatch_op(op, reg, q) ? q.v : -1;
}

extern "C" int hostcheck_meta(int op, int reg, int *six)
{
	Meta q = {six};
	return dispatch_op(op, reg, q) ? 0 : -1;
}

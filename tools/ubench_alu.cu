// tools/ubench_alu.cu -- what do compares and selects cost next to packed FP32 work?
// Not part of the product; run on a B200 through gpurun:
//     make -C tools && gpurun -- ./tools/ubench_alu [n]
// The coincident-pair / non-finite guards of the pair loop are FSETP + FSEL per lane, which issue
// to the 16-lane ALU pipe.  Part 1 times a stream of 21 FFMA2 (the packed Winckelmans mix of
// tools/ubench.cu) with 0, 1, 2, 3 (FSETP + FSEL) per lane added, and with FMUL.SAT instead
// (a compare-free way to form a 0/1 mask on the FMA pipe), to put a number on "one ALU
// instruction costs k FP32 lane-ops".  Part 2 runs the production pair kernel of the ops that have
// an optimistic form (pair_math.cuh, GUARDS) with M2MArgs::exact_only = 1 and 0 on the same
// inputs, independent targets and self-interaction, and checks that the outputs are the same bits.
// Part 3 is an EXPERIMENT that exists only here (queued experiment 1): the ops whose
// guard is a real selection (viscous ops, Winckelmans stretching: eta(0), A(0) are finite, so nothing
// poisons the sums) run the pair loop unguarded while keeping a running min(r^2) per target lane -- one
// FMNMX per pair instead of FSETP + FSEL -- and re-evaluate the chain with the production guarded form
// when the minimum is 0.  Timed against the production kernel on the same inputs, bits compared.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../cvortex_b200/csrc/m2m_kernel.cuh"

using namespace cvtx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// MODE 0: K x (FSETP + FSEL) per lane and round;  MODE 1: K x FMUL.SAT per lane and round
template <int K, int MODE>
__global__ void __launch_bounds__(256) guard_mix(float *out, int iters, float a, float b, float thr) {
	float2 x[7], z[3];
#pragma unroll
	for (int i = 0; i < 7; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
#pragma unroll
	for (int i = 0; i < 3; ++i) z[i] = make_float2(0.f, 0.f);
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 8; ++r) {
#pragma unroll
			for (int k = 0; k < 3; ++k)
#pragma unroll
				for (int i = 0; i < 7; ++i) x[i] = __ffma2_rn(x[i], make_float2(a, a), make_float2(b, b));   // 21 FFMA2
#pragma unroll
			for (int s = 0; s < K; ++s) {
				if (MODE == 0) {
					z[s].x = x[s].x > thr ? x[s + 3].x : z[s].x;
					z[s].y = x[s].y > thr ? x[s + 3].y : z[s].y;
				} else {
					z[s].x = __saturatef(x[s].x * thr);
					z[s].y = __saturatef(x[s].y * thr);
					x[s + 3].x += z[s].x * 1e-30f; x[s + 3].y += z[s].y * 1e-30f;   // keep it live (2 more lane-ops)
				}
			}
		}
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < 7; ++i) s += x[i].x + x[i].y;
#pragma unroll
	for (int i = 0; i < 3; ++i) s += z[i].x + z[i].y;
	if (s == 123.456f) out[0] = s;
}


// ---- part 3: running-min detection of coincident pairs (experiment) -------------------------------
template <int W> __device__ __forceinline__ Vec<W> vmin(Vec<W> a, Vec<W> b) {
	Vec<W> r;
#pragma unroll
	for (int i = 0; i < W; ++i) r.set(i, fminf(a.lane(i), b.lane(i)));
	return r;
}

template <int REG> struct P3DViscMin : P3DVisc<REG> {
	template <int W> __device__ __forceinline__ static void pair_min(const Vec<W> *tg, const f4 a, const f4 b, Vec<W> *acc, Vec<W> &rmin, const PairConsts &k) {
		const Rad3<W> d = rad3(tg, a);
		const Vec<W> eta = Eta3D<REG>::eta(d.r2, k);
		rmin = vmin(rmin, d.r2);
		acc[0] = vfma(eta, vfms(tg[6], b.x, vmul(tg[3], a.w)), acc[0]);
		acc[1] = vfma(eta, vfms(tg[6], b.y, vmul(tg[4], a.w)), acc[1]);
		acc[2] = vfma(eta, vfms(tg[6], b.z, vmul(tg[5], a.w)), acc[2]);
	}
};

template <int REG> struct P2DViscMin : P2DVisc<REG> {
	template <int W> __device__ __forceinline__ static void pair_min(const Vec<W> *tg, const f4 a, const f4, Vec<W> *acc, Vec<W> &rmin, const PairConsts &k) {
		const Vec<W> dx = vsub(tg[0], a.x), dy = vsub(tg[1], a.y);
		const Vec<W> r2 = vfma(dy, dy, vmul(dx, dx));
		const Vec<W> eta = Eta2D<REG>::eta(r2, k);
		rmin = vmin(rmin, r2);
		acc[0] = vfma(eta, vfms(tg[3], a.z, vmul(tg[2], a.w)), acc[0]);
	}
};

struct P3DDvortWMin : P3DDvort<REG_WINCKELMANS> {
	template <int W> __device__ __forceinline__ static void pair_min(const Vec<W> *tg, const f4 a, const f4 b, Vec<W> *acc, Vec<W> &rmin, const PairConsts &k) {
		const Rad3<W> d = rad3(tg, a);
		rmin = vmin(rmin, d.r2);
		// Reg3D<REG_WINCKELMANS>::AB without the selection on A
		const Vec<W> a1 = vfma(d.r2, k.c0, 1.0f), b1 = vfma(d.r2, k.c0, 2.5f), b2 = vfma(d.r2, k.c1, k.c2);
		const Vec<W> ra = vrsqrt(a1), ra2 = vmul(ra, ra), ra4 = vmul(ra2, ra2), ra5 = vmul(ra4, ra);
		const Vec<W> A = vmul(b1, ra5), B1 = b2, B2 = vmul(ra5, ra2);
		const Vec<W> cx = vfms(tg[4], b.z, vmul(tg[5], b.y));
		const Vec<W> cy = vfms(tg[5], b.x, vmul(tg[3], b.z));
		const Vec<W> cz = vfms(tg[3], b.y, vmul(tg[4], b.x));
		const Vec<W> trip = vfma(d.z, cz, vfma(d.y, cy, vmul(d.x, cx)));
		const Vec<W> s = vmul(vmul(B1, trip), B2);
		acc[0] = vfma(s, d.x, vfma(A, cx, acc[0]));
		acc[1] = vfma(s, d.y, vfma(A, cy, acc[1]));
		acc[2] = vfma(s, d.z, vfma(A, cz, acc[2]));
	}
};

// m2m_kernel (cvortex_b200/csrc/m2m_kernel.cuh) with the chain test replaced: P::pair_min() over the
// chain, then "is any lane's min(r^2) zero?" -> the production guarded form P::pair<W, true>().
template <class P, int T, int B, int MINB>
__global__ void __launch_bounds__(B, MINB) m2m_kernel_min(const M2MArgs args)
{
	constexpr int S = kSrcTile;
	constexpr uint32_t kTileBytes = S * sizeof(float4);
	__shared__ __align__(128) float4 tileA[2][S];
	__shared__ __align__(128) float4 tileB[2][P::NSRC4 == 2 ? S : 1];
	__shared__ __align__(8) uint64_t full[2];
	const int tid = threadIdx.x;
	const int tile0 = blockIdx.y * args.tiles_per_chunk;
	const int ntile = min(args.tiles_per_chunk, args.n_src_tiles - tile0);
	const float4 *gA = args.srcA + (size_t)tile0 * S;
	const float4 *gB = args.srcB + (size_t)tile0 * S;
	if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
	__syncthreads();
	if (tid == 0 && ntile > 0) {
		mbar_expect_tx(&full[0], kTileBytes * P::NSRC4);
		bulk_g2s(tileA[0], gA, kTileBytes, &full[0]);
		if (P::NSRC4 == 2) bulk_g2s(tileB[0], gB, kTileBytes, &full[0]);
	}
	constexpr int W = 2, NV = T / W;
	static_assert(T % 2 == 0, "packed lanes only");
	const long base = (long)blockIdx.x * (B * T) + tid;
	Vec<W> tg[NV][P::NTGT];
	double dacc[T][P::NACC];
#pragma unroll
	for (int t = 0; t < T; ++t) {
		long i = base + (long)t * B;
		i = i < args.n_tgt ? i : (long)args.n_tgt - 1;
		float one[P::NTGT];
		P::load_target(args.tgt + i * P::TCOLS, one);
#pragma unroll
		for (int c = 0; c < P::NTGT; ++c) tg[t / W][c].set(t % W, one[c]);
#pragma unroll
		for (int c = 0; c < P::NACC; ++c) dacc[t][c] = 0.0;
	}
	for (int it = 0; it < ntile; ++it) {
		const int buf = it & 1;
		if (tid == 0 && it + 1 < ntile) {
			mbar_expect_tx(&full[buf ^ 1], kTileBytes * P::NSRC4);
			bulk_g2s(tileA[buf ^ 1], gA + (size_t)(it + 1) * S, kTileBytes, &full[buf ^ 1]);
			if (P::NSRC4 == 2) bulk_g2s(tileB[buf ^ 1], gB + (size_t)(it + 1) * S, kTileBytes, &full[buf ^ 1]);
		}
		mbar_wait(&full[buf], (it >> 1) & 1);
		const float4 *sA = tileA[buf];
		const float4 *sB = tileB[P::NSRC4 == 2 ? buf : 0];
		Vec<W> acc[NV][P::NACC], rmin = bc<W>(3.0e38f);
#pragma unroll
		for (int v = 0; v < NV; ++v)
#pragma unroll
			for (int c = 0; c < P::NACC; ++c) acc[v][c] = bc<W>(0.0f);
#pragma unroll 8
		for (int j = 0; j < S; ++j) {
			const float4 a = sA[j];
			float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
			if (P::NSRC4 == 2) b = sB[j];
#pragma unroll
			for (int v = 0; v < NV; ++v) P::template pair_min<W>(tg[v], a, b, acc[v], rmin, args.k);
		}
		// (one running minimum for all of the thread's targets: any coincidence re-evaluates the chain)
		if (!(fminf(rmin.lane(0), rmin.lane(1)) > 0.0f)) {
#pragma unroll
			for (int v = 0; v < NV; ++v)
#pragma unroll
				for (int c = 0; c < P::NACC; ++c) acc[v][c] = bc<W>(0.0f);
#pragma unroll 8
			for (int j = 0; j < S; ++j) {
				const float4 a = sA[j];
				float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
				if (P::NSRC4 == 2) b = sB[j];
#pragma unroll
				for (int v = 0; v < NV; ++v) P::template pair<W, true>(tg[v], a, b, acc[v], args.k);
			}
		}
#pragma unroll
		for (int t = 0; t < T; ++t)
#pragma unroll
			for (int c = 0; c < P::NACC; ++c) dacc[t][c] += (double)acc[t / W][c].lane(t % W);
		__syncthreads();
	}
#pragma unroll
	for (int t = 0; t < T; ++t) {
		const long i = base + (long)t * B;
		if (i < args.n_tgt) {
			double res[P::NOUT];
			P::finish(args.tgt + i * P::TCOLS, dacc[t], res, args.k);
			double *dst = args.partial + ((size_t)blockIdx.y * args.n_tgt + i) * P::NOUT;
#pragma unroll
			for (int c = 0; c < P::NOUT; ++c) dst[c] = res[c];
		}
	}
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

struct Bench {
	float4 *A, *B; float *tgt_pts, *tgt_self, *tgt_self4; float *out[2]; double *partial; int n; double peak_lane;
	cudaEvent_t e0, e1;
};

template <class P, int T, int BLK, int MINB>
static void guarded_vs_optimistic(Bench &b, const char *name, bool self) {
	const int n = b.n, chunks = 8;
	const int n_tiles = n / kSrcTile;
	M2MArgs a = {};
	a.srcA = b.A; a.srcB = b.B; a.n_src_tiles = n_tiles;
	a.tiles_per_chunk = (n_tiles + chunks - 1) / chunks;
	const int gy = (n_tiles + a.tiles_per_chunk - 1) / a.tiles_per_chunk;
	a.tgt = self ? b.tgt_self : b.tgt_pts; a.n_tgt = n; a.partial = b.partial;
	a.k = P::make_consts(0.02f, 1.0f);
	const dim3 grid((n + BLK * T - 1) / (BLK * T), gy);
	float best[2] = {1e30f, 1e30f};
	const size_t nvals = (size_t)n * P::NOUT;
	for (int mode = 0; mode < 2; ++mode) {
		a.exact_only = mode == 0 ? 1 : 0;
		a.out = b.out[mode];
		for (int rep = 0; rep < 3; ++rep) {
			CK(cudaEventRecord(b.e0));
			m2m_kernel<P, T, BLK, MINB><<<grid, BLK>>>(a);
			CK(cudaEventRecord(b.e1));
			CK(cudaEventSynchronize(b.e1));
			CK(cudaGetLastError());
			const float ms = time_ms(b.e0, b.e1);
			if (ms < best[mode]) best[mode] = ms;
		}
		reduce_partials_kernel<<<(unsigned)((nvals + 255) / 256), 256>>>(b.partial, b.out[mode], (long)nvals, gy);
		CK(cudaDeviceSynchronize());
	}
	std::vector<float> h0(nvals), h1(nvals);
	CK(cudaMemcpy(h0.data(), b.out[0], sizeof(float) * nvals, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(h1.data(), b.out[1], sizeof(float) * nvals, cudaMemcpyDeviceToHost));
	const bool same = memcmp(h0.data(), h1.data(), sizeof(float) * nvals) == 0;
	const double pairs = (double)n * n;
	double r[2];
	for (int m = 0; m < 2; ++m) r[m] = pairs / (best[m] * 1e-3);
	printf("%-22s %-5s T=%d  guarded %8.3f ms %7.1f Gpair/s %5.1f%%   optimistic %8.3f ms %7.1f Gpair/s %5.1f%%   x%.3f  bits %s\n",
	       name, self ? "self" : "indep", T, best[0], r[0] * 1e-9, 100.0 * r[0] * P::LANE_OPS / b.peak_lane,
	       best[1], r[1] * 1e-9, 100.0 * r[1] * P::LANE_OPS / b.peak_lane, best[0] / best[1], same ? "identical" : "DIFFER");
	fflush(stdout);
}


// production kernel (guard = FSETP + FSEL per pair) vs m2m_kernel_min on the same inputs
template <class P, class PMIN, int T, int BLK, int MINB>
static void guard_vs_running_min(Bench &b, const char *name, bool self) {
	const int n = b.n, chunks = 8;
	const int n_tiles = n / kSrcTile;
	M2MArgs a = {};
	a.srcA = b.A; a.srcB = b.B; a.n_src_tiles = n_tiles;
	a.tiles_per_chunk = (n_tiles + chunks - 1) / chunks;
	const int gy = (n_tiles + a.tiles_per_chunk - 1) / a.tiles_per_chunk;
	a.tgt = self ? (P::TCOLS == 4 ? b.tgt_self4 : b.tgt_self) : b.tgt_pts; a.n_tgt = n; a.partial = b.partial;
	a.k = P::make_consts(0.02f, 1.0f);
	const dim3 grid((n + BLK * T - 1) / (BLK * T), gy);
	float best[2] = {1e30f, 1e30f};
	const size_t nvals = (size_t)n * P::NOUT;
	for (int mode = 0; mode < 2; ++mode) {
		a.out = b.out[mode];
		for (int rep = 0; rep < 3; ++rep) {
			CK(cudaEventRecord(b.e0));
			if (mode == 0) m2m_kernel<P, T, BLK, MINB><<<grid, BLK>>>(a);
			else m2m_kernel_min<PMIN, T, BLK, MINB><<<grid, BLK>>>(a);
			CK(cudaEventRecord(b.e1));
			CK(cudaEventSynchronize(b.e1));
			CK(cudaGetLastError());
			const float ms = time_ms(b.e0, b.e1);
			if (ms < best[mode]) best[mode] = ms;
		}
		reduce_partials_kernel<<<(unsigned)((nvals + 255) / 256), 256>>>(b.partial, b.out[mode], (long)nvals, gy);
		CK(cudaDeviceSynchronize());
	}
	std::vector<float> h0(nvals), h1(nvals);
	CK(cudaMemcpy(h0.data(), b.out[0], sizeof(float) * nvals, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(h1.data(), b.out[1], sizeof(float) * nvals, cudaMemcpyDeviceToHost));
	const bool same = memcmp(h0.data(), h1.data(), sizeof(float) * nvals) == 0;
	const double pairs = (double)n * n;
	double r[2];
	for (int m = 0; m < 2; ++m) r[m] = pairs / (best[m] * 1e-3);
	printf("%-22s %-5s T=%d  select/pair %8.3f ms %7.1f Gpair/s %5.1f%%   running min %8.3f ms %7.1f Gpair/s %5.1f%%   x%.3f  bits %s\n",
	       name, self ? "self" : "indep", T, best[0], r[0] * 1e-9, 100.0 * r[0] * P::LANE_OPS / b.peak_lane,
	       best[1], r[1] * 1e-9, 100.0 * r[1] * P::LANE_OPS / b.peak_lane, best[0] / best[1], same ? "identical" : "DIFFER");
	fflush(stdout);
}

template <int K, int MODE>
static void run_mix(float *dummy, int blocks, int iters, double peak_lane, cudaEvent_t e0, cudaEvent_t e1, const char *what) {
	float ms = 1e30f;
	for (int pass = 0; pass < 3; ++pass) {
		CK(cudaEventRecord(e0));
		guard_mix<K, MODE><<<blocks, 256>>>(dummy, iters, 1.0001f, 0.5f, MODE == 0 ? 0.25f : 1e30f);
		CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		const float t = time_ms(e0, e1);
		if (pass && t < ms) ms = t;
	}
	const double rounds = (double)blocks * 256 * iters * 8;            // thread-rounds
	const double ffma2_lane_ops = rounds * 42;
	// cycles of one SM sub-partition per warp-round: elapsed * clock / (warp-rounds per sub-partition)
	printf("21 FFMA2 + %d x %-14s per lane: %7.3f ms  FFMA2 work alone = %5.1f%% of nominal FP32 peak  (%.2f lane-op-times per round, 42 = free)\n",
	       K, what, ms, 100.0 * ffma2_lane_ops / (ms * 1e-3) / peak_lane, 42.0 * peak_lane * ms * 1e-3 / ffma2_lane_ops);
	fflush(stdout);
}

int main(int argc, char **argv) {
	int n = argc > 1 ? atoi(argv[1]) : 262144;
	n = (n + 2047) / 2048 * 2048;
	cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
	int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
	const double peak_lane = (double)prop.multiProcessorCount * 128.0 * khz * 1e3;
	printf("device: %s, %d SMs, max clock %.0f MHz, nominal FP32 peak %.2f T lane-op/s\n", prop.name, prop.multiProcessorCount, khz * 1e-3, peak_lane * 1e-12);
	float *dummy; CK(cudaMalloc(&dummy, 4096));
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	const int blocks = prop.multiProcessorCount * 8, iters = 4096;
	printf("\n== part 1: compares / selects next to packed FP32 work (per thread and round: 21 FFMA2 = 42 lane-ops)\n");
	run_mix<0, 0>(dummy, blocks, iters, peak_lane, e0, e1, "(FSETP+FSEL)");
	run_mix<1, 0>(dummy, blocks, iters, peak_lane, e0, e1, "(FSETP+FSEL)");
	run_mix<2, 0>(dummy, blocks, iters, peak_lane, e0, e1, "(FSETP+FSEL)");
	run_mix<3, 0>(dummy, blocks, iters, peak_lane, e0, e1, "(FSETP+FSEL)");
	run_mix<1, 1>(dummy, blocks, iters, peak_lane, e0, e1, "(FMUL.SAT+FFMA)");
	run_mix<2, 1>(dummy, blocks, iters, peak_lane, e0, e1, "(FMUL.SAT+FFMA)");

	// ---- synthetic cloud as in reference bench/bencharraysetup.c:43-58; self targets = the sources
	std::vector<float4> hA(n), hB(n);
	std::vector<float> hp((size_t)n * 7), hs((size_t)n * 7);
	srand(1234);
	auto rnd = []() { return 10.0f * (float)rand() / (float)RAND_MAX; };
	for (int i = 0; i < n; ++i) {
		hA[i] = make_float4(rnd(), rnd(), rnd(), 0.01f);
		hB[i] = make_float4(rnd(), rnd(), rnd(), 0.1f * rnd());      // w: only the filament ops read it ((3/|r0|) t1)
		for (int c = 0; c < 6; ++c) hp[(size_t)i * 7 + c] = rnd();
		hp[(size_t)i * 7 + 6] = 0.01f;
		hs[(size_t)i * 7 + 0] = hA[i].x; hs[(size_t)i * 7 + 1] = hA[i].y; hs[(size_t)i * 7 + 2] = hA[i].z;
		hs[(size_t)i * 7 + 3] = hB[i].x; hs[(size_t)i * 7 + 4] = hB[i].y; hs[(size_t)i * 7 + 5] = hB[i].z;
		hs[(size_t)i * 7 + 6] = 0.01f;
	}
	Bench b = {};
	b.n = n; b.peak_lane = peak_lane; b.e0 = e0; b.e1 = e1;
	CK(cudaMalloc(&b.A, sizeof(float4) * n)); CK(cudaMalloc(&b.B, sizeof(float4) * n));
	CK(cudaMalloc(&b.tgt_pts, sizeof(float) * 7 * n)); CK(cudaMalloc(&b.tgt_self, sizeof(float) * 7 * n));
	for (int m = 0; m < 2; ++m) CK(cudaMalloc(&b.out[m], sizeof(float) * 6 * n));
	CK(cudaMalloc(&b.partial, sizeof(double) * 6 * (size_t)n * 8));
	CK(cudaMemcpy(b.A, hA.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(b.B, hB.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(b.tgt_pts, hp.data(), sizeof(float) * 7 * n, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(b.tgt_self, hs.data(), sizeof(float) * 7 * n, cudaMemcpyHostToDevice));
	CK(cudaMalloc(&b.tgt_self4, sizeof(float4) * n));                       // the packed P2D records are their own target rows
	CK(cudaMemcpy(b.tgt_self4, hA.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));

	printf("\n== part 2: pair kernel, guarded form only vs optimistic chains, n = m = %d, 8 source chunks\n", n);
	// (rows of 7 floats serve every op: 3-column ops read a prefix with a stride of their own, so `self`
	// only coincides for the 7-column ops; the point ops get their coincidences from `indep` = none)
	guarded_vs_optimistic<P3DVel<REG_GAUSSIAN>, 8, 128, 2>(b, "P3D vel gaussian", false);
	guarded_vs_optimistic<P3DVel<REG_SINGULAR>, 8, 128, 2>(b, "P3D vel singular", false);
	guarded_vs_optimistic<P3DDvort<REG_GAUSSIAN>, 8, 128, 2>(b, "P3D dvort gaussian", false);
	guarded_vs_optimistic<P3DDvort<REG_GAUSSIAN>, 8, 128, 2>(b, "P3D dvort gaussian", true);
	guarded_vs_optimistic<P3DDvort<REG_SINGULAR>, 8, 128, 2>(b, "P3D dvort singular", true);
	guarded_vs_optimistic<P3DVelDvort<REG_GAUSSIAN>, 4, 256, 2>(b, "P3D vel+dvort gaussian", true);
	guarded_vs_optimistic<P2DVel<REG_SINGULAR>, 4, 256, 2>(b, "P2D vel singular", false);
	guarded_vs_optimistic<P2DVel<REG_GAUSSIAN>, 4, 256, 2>(b, "P2D vel gaussian", false);
	guarded_vs_optimistic<F3DVel, 8, 128, 2>(b, "F3D vel", false);
	guarded_vs_optimistic<F3DDvort, 8, 128, 2>(b, "F3D dvort", true);
	guarded_vs_optimistic<P3DDvort<REG_GAUSSIAN>, 4, 256, 2>(b, "P3D dvort gaussian", true);
	guarded_vs_optimistic<P3DVel<REG_GAUSSIAN>, 4, 256, 2>(b, "P3D vel gaussian", false);

	printf("\n== part 3 (experiment): guard as a select per pair (production) vs running min(r^2) + guarded re-evaluation\n");
	guard_vs_running_min<P3DVisc<REG_WINCKELMANS>, P3DViscMin<REG_WINCKELMANS>, 8, 128, 2>(b, "P3D visc winckelmans", true);
	guard_vs_running_min<P3DVisc<REG_GAUSSIAN>, P3DViscMin<REG_GAUSSIAN>, 8, 128, 2>(b, "P3D visc gaussian", true);
	guard_vs_running_min<P3DDvort<REG_WINCKELMANS>, P3DDvortWMin, 8, 128, 2>(b, "P3D dvort winckelmans", true);
	guard_vs_running_min<P2DVisc<REG_GAUSSIAN>, P2DViscMin<REG_GAUSSIAN>, 8, 128, 2>(b, "P2D visc gaussian", true);
	guard_vs_running_min<P2DVisc<REG_WINCKELMANS>, P2DViscMin<REG_WINCKELMANS>, 2, 256, 3>(b, "P2D visc winckelmans", true);
	printf("done\n");
	return 0;
}

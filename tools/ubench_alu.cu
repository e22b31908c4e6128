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

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

struct Bench {
	float4 *A, *B; float *tgt_pts, *tgt_self; float *out[2]; double *partial; int n; double peak_lane;
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
	printf("done\n");
	return 0;
}

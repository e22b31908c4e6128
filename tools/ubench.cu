// tools/ubench.cu -- design-space micro-benchmarks for the all-pairs kernel.
// Not part of the product; run on a B200 through gpurun:
//     make -C tools && gpurun -- ./tools/ubench [n]
// 1. measured FP32-FMA and MUFU issue peaks of the chip (the roofline
//    denominators of DESIGN.md section 4 are nominal: SMs x lanes x clock;
//    this records what the silicon sustains at its actual clock);
// 2. m2m_kernel<P3DVel<Winckelmans>> over targets-per-thread / block size /
//    occupancy / chunking variants;
// 3. every op family at the production configuration.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../cvortex_b200/csrc/m2m_kernel.cuh"

using namespace cvtx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// ---- pipe peaks -------------------------------------------------------------
template <int ILP>
__global__ void __launch_bounds__(256) ffma_peak(float *out, int iters, float a, float b) {
	float x[ILP];
#pragma unroll
	for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 16; ++r)
#pragma unroll
			for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < ILP; ++i) s += x[i];
	if (s == 123.456f) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) ffma2_peak(float *out, int iters, float a, float b) {
	float2 x[ILP];
#pragma unroll
	for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 16; ++r)
#pragma unroll
			for (int i = 0; i < ILP; ++i) x[i] = __ffma2_rn(x[i], make_float2(a, a), make_float2(b, b));
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
	if (s == 123.456f) out[0] = s;
}

// OP: 0 = FADD2, 1 = FMUL2, 2 = FADD (scalar), 3 = FMUL (scalar)
template <int OP>
__global__ void __launch_bounds__(256) packed_op_peak(float *out, int iters, float a) {
	float2 x[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 16; ++r)
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				if (OP == 0) x[i] = __fadd2_rn(x[i], make_float2(a, a));
				else if (OP == 1) x[i] = __fmul2_rn(x[i], make_float2(a, a));
				else if (OP == 2) { x[i].x = x[i].x + a; x[i].y = x[i].y + a; }
				else { x[i].x = x[i].x * a; x[i].y = x[i].y * a; }
			}
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
	if (s == 123.456f) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) mufu_peak(float *out, int iters) {
	float x[ILP];
#pragma unroll
	for (int i = 0; i < ILP; ++i) x[i] = 1.0f + threadIdx.x * 1e-3f + i;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 16; ++r)
#pragma unroll
			for (int i = 0; i < ILP; ++i) x[i] = mufu_rsqrt(x[i]);
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < ILP; ++i) s += x[i];
	if (s == 123.456f) out[0] = s;
}

// FFMA : MUFU at the Winckelmans-vel ratio (21 : 1) to see whether the two pipes overlap
__global__ void __launch_bounds__(256) mix_peak(float *out, int iters, float a, float b) {
	float x[8], y[2];
#pragma unroll
	for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
	y[0] = 1.5f; y[1] = 2.5f;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 8; ++r) {
#pragma unroll
			for (int k = 0; k < 5; ++k)
#pragma unroll
				for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);       // 40 FFMA
			x[0] = fmaf(x[0], a, b); x[1] = fmaf(x[1], a, b);              // +2 = 42
			y[0] = mufu_rsqrt(y[0] + 1.0f); y[1] = mufu_rsqrt(y[1] + 1.0f); // 2 MUFU (+2 FADD)
		}
	}
	float s = y[0] + y[1];
#pragma unroll
	for (int i = 0; i < 8; ++i) s += x[i];
	if (s == 123.456f) out[0] = s;
}

// 21 FFMA2 : 2 MUFU per two pairs -- the packed Winckelmans-vel mix.  If the pipes were
// independent the FMA pipe would stay 100 % busy (issue needs only 23 of 42 cycles).
__global__ void __launch_bounds__(256) mix2_peak(float *out, int iters, float a, float b) {
	float2 x[7];
	float y[2];
#pragma unroll
	for (int i = 0; i < 7; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
	y[0] = 1.5f; y[1] = 2.5f;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 8; ++r) {
#pragma unroll
			for (int k = 0; k < 3; ++k)
#pragma unroll
				for (int i = 0; i < 7; ++i) x[i] = __ffma2_rn(x[i], make_float2(a, a), make_float2(b, b));   // 21 FFMA2
			y[0] = mufu_rsqrt(fabsf(x[0].x) + 1.0f + y[0]); y[1] = mufu_rsqrt(fabsf(x[1].y) + 1.0f + y[1]); // 2 MUFU (+4 FADD)
		}
	}
	float s = y[0] + y[1];
#pragma unroll
	for (int i = 0; i < 7; ++i) s += x[i].x + x[i].y;
	if (s == 123.456f) out[0] = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

// ---- m2m variants -------------------------------------------------------------
struct Bench {
	float4 *A, *B; float *tgt; float *out; double *pieces; int *tickets; int n; int sms; double peak_lane;
	cudaEvent_t e0, e1;
};

// One geometry of the persistent pair kernel: grid = resident blocks, equal runs (m2m_kernel.cuh).
template <class P, int T, int BLK, int MINB>
static void run_variant(Bench &b, const char *name) {
	const int n = b.n;
	const int n_tiles = n / kSrcTile;
	int occ = 0;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, m2m_kernel<P, T, BLK, MINB>, BLK, 0));
	cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, m2m_kernel<P, T, BLK, MINB>));
	M2MArgs a = {};
	a.srcA = b.A; a.srcB = b.B; a.n_src = n; a.n_src_tiles = n_tiles; a.grain = kSrcTile;
	const long long tiles_t = (n + BLK * T - 1) / (BLK * T);
	a.total_grains = tiles_t * n_tiles;
	a.tgt = b.tgt; a.n_tgt = n; a.out = b.out; a.pieces = b.pieces; a.tickets = b.tickets;
	a.k = P::make_consts(0.02f, 1.0f);
	const int grid = occ * b.sms;
	float best = 1e30f;
	for (int rep = 0; rep < 3; ++rep) {
		CK(cudaEventRecord(b.e0));
		m2m_kernel<P, T, BLK, MINB><<<grid, BLK>>>(a);
		CK(cudaEventRecord(b.e1));
		CK(cudaEventSynchronize(b.e1));
		CK(cudaGetLastError());
		const float ms = time_ms(b.e0, b.e1);
		if (ms < best) best = ms;
	}
	const double pairs = (double)n * n;
	const double rate = pairs / (best * 1e-3);
	printf("%-28s T=%d B=%3d regs=%3d occ=%d grid=%d %8.3f ms  %8.1f Gpair/s  lane-ops %2d -> %5.1f%% of nominal FP32 peak\n",
	       name, T, BLK, fa.numRegs, occ, grid, best, rate * 1e-9, P::LANE_OPS,
	       100.0 * rate * P::LANE_OPS / b.peak_lane);
	fflush(stdout);
}

int main(int argc, char **argv) {
	int n = argc > 1 ? atoi(argv[1]) : 262144;
	n = (n + 1023) / 1024 * 1024;
	cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
	int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
	const double peak_lane = (double)prop.multiProcessorCount * 128.0 * khz * 1e3;
	printf("device: %s, %d SMs, max clock %.0f MHz, nominal FP32 peak %.2f T lane-op/s, nominal MUFU peak %.2f T op/s\n",
	       prop.name, prop.multiProcessorCount, khz * 1e-3, peak_lane * 1e-12, peak_lane / 8 * 1e-12);

	float *dummy; CK(cudaMalloc(&dummy, 4096));
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	const int blocks = prop.multiProcessorCount * 8, iters = 4096;
	for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms up
		CK(cudaEventRecord(e0)); ffma_peak<8><<<blocks, 256>>>(dummy, iters, 1.0001f, 0.5f); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		double ops = (double)blocks * 256 * iters * 16 * 8;
		if (pass) printf("FFMA peak (ILP 8, 8 blk/SM): %.2f T lane-op/s = %.1f%% of nominal\n", ops / (time_ms(e0, e1) * 1e-3) * 1e-12, 100 * ops / (time_ms(e0, e1) * 1e-3) / peak_lane);
		CK(cudaEventRecord(e0)); ffma_peak<4><<<blocks, 256>>>(dummy, iters, 1.0001f, 0.5f); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		ops = (double)blocks * 256 * iters * 16 * 4;
		if (pass) printf("FFMA peak (ILP 4, 8 blk/SM): %.2f T lane-op/s = %.1f%% of nominal\n", ops / (time_ms(e0, e1) * 1e-3) * 1e-12, 100 * ops / (time_ms(e0, e1) * 1e-3) / peak_lane);
		CK(cudaEventRecord(e0)); ffma2_peak<8><<<blocks, 256>>>(dummy, iters, 1.0001f, 0.5f); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		ops = (double)blocks * 256 * iters * 16 * 8 * 2;
		if (pass) printf("FFMA2 (packed f32x2) peak (ILP 8): %.2f T lane-op/s = %.1f%% of nominal, at half the issue slots\n", ops / (time_ms(e0, e1) * 1e-3) * 1e-12, 100 * ops / (time_ms(e0, e1) * 1e-3) / peak_lane);
		{
			const char *names[4] = {"FADD2", "FMUL2", "FADD", "FMUL"};
			for (int op = 0; op < 4; ++op) {
				CK(cudaEventRecord(e0));
				if (op == 0) packed_op_peak<0><<<blocks, 256>>>(dummy, iters, 1.0001f);
				else if (op == 1) packed_op_peak<1><<<blocks, 256>>>(dummy, iters, 1.0001f);
				else if (op == 2) packed_op_peak<2><<<blocks, 256>>>(dummy, iters, 1.0001f);
				else packed_op_peak<3><<<blocks, 256>>>(dummy, iters, 1.0001f);
				CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
				ops = (double)blocks * 256 * iters * 16 * 8 * 2;
				if (pass) printf("%s peak: %.2f T lane-op/s = %.1f%% of nominal\n", names[op], ops / (time_ms(e0, e1) * 1e-3) * 1e-12, 100 * ops / (time_ms(e0, e1) * 1e-3) / peak_lane);
			}
		}
		CK(cudaEventRecord(e0)); mufu_peak<8><<<blocks, 256>>>(dummy, iters / 4); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		ops = (double)blocks * 256 * (iters / 4) * 16 * 8;
		if (pass) printf("MUFU.RSQ peak: %.2f T op/s = %.1f%% of nominal (SMs x 16 x clock)\n", ops / (time_ms(e0, e1) * 1e-3) * 1e-12, 100 * ops / (time_ms(e0, e1) * 1e-3) / (peak_lane / 8));
		CK(cudaEventRecord(e0)); mix_peak<<<blocks, 256>>>(dummy, iters, 1.0001f, 0.5f); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		ops = (double)blocks * 256 * iters * 8 * 44;     // 42 FFMA + 2 FADD per 2 MUFU
		if (pass) printf("FP32:MUFU = 22:1 mix: %.2f T FP32 lane-op/s = %.1f%% of nominal (ideal if pipes overlap: 100%%, issue-limited: %.1f%%)\n",
		                 ops / (time_ms(e0, e1) * 1e-3) * 1e-12, 100 * ops / (time_ms(e0, e1) * 1e-3) / peak_lane, 100.0 * 44 / 46);
	}

	CK(cudaEventRecord(e0)); mix2_peak<<<blocks, 256>>>(dummy, iters, 1.0001f, 0.5f); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
	{
		const double ops = (double)blocks * 256 * iters * 8 * (42 + 4);
		printf("FFMA2:MUFU = 21:2 mix (packed): %.2f T FP32 lane-op/s = %.1f%% of nominal  (issue slots needed: %.0f%%)\n",
		       ops / (time_ms(e0, e1) * 1e-3) * 1e-12, 100 * ops / (time_ms(e0, e1) * 1e-3) / peak_lane, 100.0 * 27 / 46);
	}

	// ---- synthetic cloud as in reference bench/bencharraysetup.c:43-58
	std::vector<float4> hA(n), hB(n);
	std::vector<float> ht((size_t)n * 7);
	srand(1234);
	auto rnd = []() { return 10.0f * (float)rand() / (float)RAND_MAX; };
	for (int i = 0; i < n; ++i) {
		hA[i] = make_float4(rnd(), rnd(), rnd(), 0.01f);
		hB[i] = make_float4(rnd(), rnd(), rnd(), 0.f);
		for (int c = 0; c < 6; ++c) ht[(size_t)i * 7 + c] = rnd();
		ht[(size_t)i * 7 + 6] = 0.01f;
	}
	Bench b = {};
	b.n = n; b.sms = prop.multiProcessorCount; b.peak_lane = peak_lane; b.e0 = e0; b.e1 = e1;
	CK(cudaMalloc(&b.A, sizeof(float4) * n)); CK(cudaMalloc(&b.B, sizeof(float4) * n));
	CK(cudaMalloc(&b.tgt, sizeof(float) * 7 * n)); CK(cudaMalloc(&b.out, sizeof(float) * 3 * n));
	CK(cudaMalloc(&b.pieces, sizeof(double) * 2 * 6 * 2048 * (size_t)prop.multiProcessorCount * 8));
	CK(cudaMalloc(&b.tickets, sizeof(int) * (size_t)n)); CK(cudaMemset(b.tickets, 0, sizeof(int) * (size_t)n));
	CK(cudaMemcpy(b.A, hA.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(b.B, hB.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(b.tgt, ht.data(), sizeof(float) * 7 * n, cudaMemcpyHostToDevice));

	printf("\n== P3D vel Winckelmans, n = m = %d : geometry sweep (persistent kernel, equal runs)\n", n);
	typedef P3DVel<REG_WINCKELMANS> W;
	run_variant<W, 1, 128, 8>(b, "vel-W");
	run_variant<W, 2, 256, 3>(b, "vel-W");
	run_variant<W, 4, 128, 4>(b, "vel-W");
	run_variant<W, 4, 256, 2>(b, "vel-W");
	run_variant<W, 6, 128, 3>(b, "vel-W");
	run_variant<W, 8, 128, 2>(b, "vel-W");
	run_variant<W, 8, 128, 3>(b, "vel-W (168-reg cap)");
	run_variant<W, 8, 256, 1>(b, "vel-W");
	typedef P3DDvort<REG_GAUSSIAN> DG;
	run_variant<DG, 8, 128, 2>(b, "dvort-G");
	run_variant<DG, 6, 128, 3>(b, "dvort-G");
	run_variant<DG, 4, 128, 4>(b, "dvort-G (128-reg cap)");
	run_variant<P3DVel<REG_GAUSSIAN>, 8, 128, 2>(b, "vel-G");
	run_variant<P3DVel<REG_GAUSSIAN>, 6, 128, 3>(b, "vel-G");
	run_variant<P3DVel<REG_GAUSSIAN>, 4, 128, 4>(b, "vel-G (128-reg cap)");
	run_variant<P3DVort<REG_GAUSSIAN>, 8, 128, 2>(b, "vort-G");
	run_variant<P3DVort<REG_GAUSSIAN>, 8, 128, 3>(b, "vort-G (168-reg cap)");
	run_variant<P2DVisc<REG_GAUSSIAN>, 8, 128, 2>(b, "P2D visc-G");
	run_variant<P2DVisc<REG_GAUSSIAN>, 8, 128, 3>(b, "P2D visc-G (168-reg cap)");
	run_variant<P2DVisc<REG_GAUSSIAN>, 8, 256, 1>(b, "P2D visc-G");
	printf("done\n");
	return 0;
}

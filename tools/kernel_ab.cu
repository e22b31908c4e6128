// tools/kernel_ab.cu -- A/B harness for the pair kernel (not part of the product):
//   * the round-1 kernel (grid = target tiles x source chunks, FP64 partials), rebuilt from the
//     repository history into tools/_r1/ under its own namespace, as the control;
//   * this tree's persistent kernel in its vector widths (VW = 2 / 4 / 8 lanes per Vec) and with
//     shorter runs (grid = a multiple of the resident blocks);
// same packed sources, same targets, one line per variant, outputs compared through a checksum.
//     make -C tools kernel_ab && gpurun -- ./tools/kernel_ab [n] [filter]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../cvortex_b200/csrc/m2m_kernel.cuh"
#include "_r1/m2m_kernel.cuh"

using namespace cvtx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Ctx {
	float4 *A, *B, *C; float *raw; float *tgt; float *out; double *pieces; int *tickets; int *mode; double *partial;
	int n, sms; double peak_lane; cudaEvent_t e0, e1; const char *filter; std::vector<float> host;
};

static double checksum(Ctx &c, int nout) {
	c.host.resize((size_t)c.n * nout);
	CK(cudaMemcpy(c.host.data(), c.out, sizeof(float) * c.host.size(), cudaMemcpyDeviceToHost));
	double s = 0;
	for (size_t i = 0; i < c.host.size(); ++i) s += (double)c.host[i] * (double)((i % 7) + 1);
	return s;
}

static void report(Ctx &c, const char *name, const char *variant, int regs, int grid, float ms, int lane_ops, int nout) {
	const double rate = (double)c.n * c.n / (ms * 1e-3);
	printf("%-16s %-26s regs=%3d grid=%5d %8.3f ms %8.1f Gpair/s %5.1f%% FP32  sum %.9e\n", name, variant, regs, grid, ms,
	       rate * 1e-9, 100.0 * rate * lane_ops / c.peak_lane, checksum(c, nout));
	fflush(stdout);
}

static int g_grid_abs = 0, g_grain = 0, g_direct = 0;      // "small" mode: explicit grid, run-time grain, in-kernel packing

template <class P, int T, int BLK, int MINB, int VW, int OPT = 0, int GRAIN = 0>
static float run_new(Ctx &c, const char *name, int grid_mult, bool times = false, int reps = 3) {
	char variant[96];
	if (g_grid_abs) snprintf(variant, sizeof variant, "r2 T=%d B=%d/%d VW=%d O%d G%d grid=%d grain=%d%s", T, BLK, MINB, VW, OPT, GRAIN, g_grid_abs, g_grain ? g_grain : kSrcTile, g_direct ? " direct" : "");
	else snprintf(variant, sizeof variant, "r2 T=%d B=%d/%d VW=%d O%d G%d x%d", T, BLK, MINB, VW, OPT, GRAIN, grid_mult);
	if (c.filter && !strstr(name, c.filter) && !strstr(variant, c.filter)) return 0.f;
	auto kern = m2m_kernel<P, T, BLK, MINB, VW, OPT, GRAIN>;
	int occ = 0;
	const size_t dyn = m2m_smem_bytes<P, T, BLK, OPT>();
	if (dyn > 0) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLK, dyn));
	cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
	M2MArgs a = {};
	a.srcA = c.A; a.srcB = c.B; a.srcC = c.C; a.src_raw = c.raw; a.n_src = c.n; a.n_src_tiles = c.n / kSrcTile; a.grain = kSrcTile;
	const long long tiles_t = (c.n + BLK * T - 1) / (BLK * T);
	a.total_grains = tiles_t * a.n_src_tiles;
	if (g_grain) { a.grain = g_grain; a.total_grains *= kSrcTile / g_grain; }
	a.direct = g_direct;
	a.tgt = c.tgt; a.n_tgt = c.n; a.out = c.out; a.pieces = c.pieces; a.tickets = c.tickets; a.f3d_mode = c.mode;
	a.k = P::make_consts(0.02f, 1.0f);
	const int grid = g_grid_abs ? g_grid_abs : occ * c.sms * grid_mult;
	unsigned long long *bt = nullptr;
	if (times) { CK(cudaMalloc(&bt, sizeof(unsigned long long) * 3 * grid)); a.block_times = bt; }
	float best = 1e30f;
	for (int rep = 0; rep < reps; ++rep) {
		CK(cudaEventRecord(c.e0));
		kern<<<grid, BLK, dyn>>>(a);
		CK(cudaEventRecord(c.e1));
		CK(cudaEventSynchronize(c.e1));
		CK(cudaGetLastError());
		float ms; CK(cudaEventElapsedTime(&ms, c.e0, c.e1));
		if (ms < best) best = ms;
	}
	report(c, name, variant, fa.numRegs, grid, best, P::LANE_OPS, P::NOUT);
	if (times) {
		// per-SM busy time of the last launch: are the SMs equally fast?
		std::vector<unsigned long long> h(3 * (size_t)grid);
		CK(cudaMemcpy(h.data(), bt, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
		CK(cudaFree(bt));
		unsigned long long t0 = ~0ull, t1 = 0;
		for (int i = 0; i < grid; ++i) { if (h[3 * i + 1] < t0) t0 = h[3 * i + 1]; if (h[3 * i + 2] > t1) t1 = h[3 * i + 2]; }
		std::vector<double> dur(grid), sm_end(256, 0.0); std::vector<int> sm_blocks(256, 0);
		double dmin = 1e30, dmax = 0, dsum = 0;
		for (int i = 0; i < grid; ++i) {
			dur[i] = (double)(h[3 * i + 2] - h[3 * i + 1]) * 1e-6;
			dmin = dur[i] < dmin ? dur[i] : dmin; dmax = dur[i] > dmax ? dur[i] : dmax; dsum += dur[i];
			const int sm = (int)h[3 * i];
			sm_blocks[sm]++;
			const double e = (double)(h[3 * i + 2] - t0) * 1e-6;
			if (e > sm_end[sm]) sm_end[sm] = e;
		}
		int bmin = 1 << 30, bmax = 0; double emin = 1e30, emax = 0;
		for (int sm = 0; sm < 256; ++sm) if (sm_blocks[sm]) {
			bmin = sm_blocks[sm] < bmin ? sm_blocks[sm] : bmin; bmax = sm_blocks[sm] > bmax ? sm_blocks[sm] : bmax;
			emin = sm_end[sm] < emin ? sm_end[sm] : emin; emax = sm_end[sm] > emax ? sm_end[sm] : emax;
		}
		if (g_grid_abs) printf("    occupancy %d blocks/SM; ", occ);
		printf("    block times: span %.3f ms; per block min %.3f mean %.3f max %.3f ms; blocks per SM %d..%d; SMs finish at %.3f..%.3f ms\n",
		       (double)(t1 - t0) * 1e-6, dmin, dsum / grid, dmax, bmin, bmax, emin, emax);
		// slowest and fastest few blocks with their SM
		std::vector<int> idx(grid); for (int i = 0; i < grid; ++i) idx[i] = i;
		std::sort(idx.begin(), idx.end(), [&](int x, int y) { return dur[x] < dur[y]; });
		printf("    fastest:"); for (int k = 0; k < 6 && k < grid; ++k) printf(" b%d@sm%d %.3f", idx[k], (int)h[3 * idx[k]], dur[idx[k]]);
		printf("\n    slowest:"); for (int k = 0; k < 6 && k < grid; ++k) printf(" b%d@sm%d %.3f", idx[grid - 1 - k], (int)h[3 * idx[grid - 1 - k]], dur[idx[grid - 1 - k]]);
		printf("\n");
	}
	return best;
}

__global__ void reduce_r1(const double *__restrict__ partial, float *__restrict__ out, long n_vals, int n_chunks) {
	const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_vals) return;
	double s = 0.0;
	for (int k = 0; k < n_chunks; ++k) s += partial[(size_t)k * n_vals + i];
	out[i] = (float)s;
}

template <class P, int T, int BLK, int MINB>
static void run_r1(Ctx &c, const char *name, int chunks, int lane_ops, int reps = 3) {
	char variant[64];
	snprintf(variant, sizeof variant, "r1 T=%d B=%d chunks=%d", T, BLK, chunks);
	if (c.filter && !strstr(name, c.filter) && !strstr(variant, c.filter)) return;
	auto kern = cvtx_r1::m2m_kernel<P, T, BLK, MINB>;
	cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
	cvtx_r1::M2MArgs a = {};
	a.srcA = c.A; a.srcB = c.B; a.n_src_tiles = c.n / 256; a.tiles_per_chunk = (a.n_src_tiles + chunks - 1) / chunks;
	a.tgt = c.tgt; a.n_tgt = c.n; a.out = c.out; a.partial = c.partial;
	const cvtx_r1::PairConsts k = P::make_consts(0.02f, 1.0f);
	a.k = k;
	const dim3 grid((c.n + BLK * T - 1) / (BLK * T), chunks);
	float best = 1e30f;
	for (int rep = 0; rep < reps; ++rep) {
		CK(cudaEventRecord(c.e0));
		kern<<<grid, BLK>>>(a);
		if (chunks > 1) reduce_r1<<<(c.n * P::NOUT + 255) / 256, 256>>>(c.partial, c.out, (long)c.n * P::NOUT, chunks);
		CK(cudaEventRecord(c.e1));
		CK(cudaEventSynchronize(c.e1));
		CK(cudaGetLastError());
		float ms; CK(cudaEventElapsedTime(&ms, c.e0, c.e1));
		if (ms < best) best = ms;
	}
	report(c, name, variant, fa.numRegs, grid.x * grid.y, best, lane_ops, P::NOUT);
}

// every (vector width, accumulator placement) of the two large-problem geometries, against the round-1 kernel
template <class P, class P1>
static void family(Ctx &c, const char *name, int chunks) {
	run_r1<P1, 8, 128, 2>(c, name, chunks, P::LANE_OPS);
	run_r1<P1, 4, 256, 2>(c, name, chunks, P::LANE_OPS);
	run_new<P, 8, 128, 2, 2, 0>(c, name, 8);
	run_new<P, 8, 128, 2, 2, 1>(c, name, 8);
	run_new<P, 8, 128, 2, 4, 0>(c, name, 8);
	run_new<P, 8, 128, 2, 4, 1>(c, name, 8);
	run_new<P, 8, 128, 2, 8, 0>(c, name, 8);
	run_new<P, 8, 128, 2, 8, 1>(c, name, 8);
	run_new<P, 4, 256, 2, 2, 0>(c, name, 8);
	run_new<P, 4, 256, 2, 2, 1>(c, name, 8);
	run_new<P, 4, 256, 2, 4, 0>(c, name, 8);
	run_new<P, 4, 256, 2, 4, 1>(c, name, 8);
}
// every (vector width, accumulator placement, loop form) of the two large-problem geometries at the compile-time
// chain length, against the round-1 kernel: the table the per-policy TUNE_* constants of pair_math.cuh are read from
template <class P, class P1>
static void sweep(Ctx &c, const char *name, int chunks) {
	run_r1<P1, 8, 128, 2>(c, name, chunks, P::LANE_OPS);
	run_r1<P1, 4, 256, 2>(c, name, chunks, P::LANE_OPS);
#define ROW(T, B, VW) run_new<P, T, B, 2, VW, 0, 256>(c, name, 4); run_new<P, T, B, 2, VW, 1, 256>(c, name, 4); \
	run_new<P, T, B, 2, VW, 2, 256>(c, name, 4); run_new<P, T, B, 2, VW, 3, 256>(c, name, 4);
	ROW(8, 128, 2) ROW(8, 128, 4) ROW(8, 128, 8)
	ROW(4, 256, 2) ROW(4, 256, 4)
#undef ROW
}
template <class P>
static void sweep_f3d(Ctx &c, const char *name) {
#ifdef F3D_QUICK
	run_new<P, 4, 256, 2, 2, 1, 256>(c, name, 4);
	run_new<P, 4, 256, 2, 2, 5, 256>(c, name, 4);
	run_new<P, 4, 256, 2, 4, 1, 256>(c, name, 4);
	run_new<P, 4, 256, 2, 4, 5, 256>(c, name, 4);
	run_new<P, 8, 128, 2, 8, 1, 256>(c, name, 4);
	run_new<P, 8, 128, 2, 8, 5, 256>(c, name, 4);
	run_new<P, 2, 256, 3, 2, 1, 256>(c, name, 4);
	run_new<P, 2, 256, 3, 2, 5, 256>(c, name, 4);
	return;
#endif
#define ROW(T, B, VW) run_new<P, T, B, 2, VW, 0, 256>(c, name, 4); run_new<P, T, B, 2, VW, 1, 256>(c, name, 4);
	ROW(8, 128, 2) ROW(8, 128, 4) ROW(8, 128, 8)
	ROW(4, 256, 2) ROW(4, 256, 4)
#undef ROW
	run_new<P, 2, 256, 3, 2, 0, 256>(c, name, 4);
	run_new<P, 2, 256, 3, 2, 1, 256>(c, name, 4);
	run_new<P, 4, 128, 4, 2, 0, 256>(c, name, 4);
	run_new<P, 4, 128, 4, 4, 1, 256>(c, name, 4);
}
template <class P>
static void mult_sweep(Ctx &c, const char *name) {
	run_new<P, 8, 128, 2, 2, 0>(c, name, 1, true);
	run_new<P, 8, 128, 2, 2, 0>(c, name, 2);
	run_new<P, 8, 128, 2, 2, 0>(c, name, 3);
	run_new<P, 8, 128, 2, 2, 0>(c, name, 4, true);
	run_new<P, 8, 128, 2, 2, 0>(c, name, 8);
	run_new<P, 8, 128, 2, 2, 0>(c, name, 16);
	run_new<P, 8, 128, 2, 2, 0>(c, name, 32);
}

int main(int argc, char **argv) {
	int n = argc > 1 ? atoi(argv[1]) : 262144;
	n = (n + 2047) / 2048 * 2048;
	Ctx c = {};
	c.filter = argc > 2 ? argv[2] : nullptr;
	cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
	int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
	c.n = n; c.sms = prop.multiProcessorCount; c.peak_lane = (double)c.sms * 128.0 * khz * 1e3;
	CK(cudaEventCreate(&c.e0)); CK(cudaEventCreate(&c.e1));
	printf("device: %s, %d SMs, %.0f MHz; n = m = %d\n", prop.name, c.sms, khz * 1e-3, n);
	std::vector<float4> hA(n), hB(n);
	std::vector<float> ht((size_t)n * 7);
	srand(1234);
	auto rnd = []() { return 10.0f * (float)rand() / (float)RAND_MAX; };
	for (int i = 0; i < n; ++i) {
		hA[i] = make_float4(rnd(), rnd(), rnd(), 0.01f);
		hB[i] = make_float4(rnd(), rnd(), rnd(), 0.f);
		// particle targets = the sources themselves (self interaction, like the reference bench)
		ht[(size_t)i * 7 + 0] = hA[i].x; ht[(size_t)i * 7 + 1] = hA[i].y; ht[(size_t)i * 7 + 2] = hA[i].z;
		ht[(size_t)i * 7 + 3] = hB[i].x; ht[(size_t)i * 7 + 4] = hB[i].y; ht[(size_t)i * 7 + 5] = hB[i].z;
		ht[(size_t)i * 7 + 6] = 0.01f;
	}
	const int chunks = 8;
	CK(cudaMalloc(&c.A, sizeof(float4) * n)); CK(cudaMalloc(&c.B, sizeof(float4) * n)); CK(cudaMalloc(&c.C, sizeof(float4) * n));
	CK(cudaMalloc(&c.raw, sizeof(float) * 7 * n));
	CK(cudaMalloc(&c.tgt, sizeof(float) * 7 * n)); CK(cudaMalloc(&c.out, sizeof(float) * 6 * n));
	CK(cudaMalloc(&c.pieces, sizeof(double) * 2 * 6 * 2048 * (size_t)c.sms * 8 * 8));
	CK(cudaMalloc(&c.partial, sizeof(double) * 6 * (size_t)n * chunks));
	CK(cudaMalloc(&c.tickets, sizeof(int) * (size_t)n)); CK(cudaMemset(c.tickets, 0, sizeof(int) * (size_t)n));
	CK(cudaMalloc(&c.mode, 64)); CK(cudaMemset(c.mode, 0, 64));
	CK(cudaMemcpy(c.A, hA.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(c.B, hB.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
	CK(cudaMemset(c.C, 0, sizeof(float4) * n));
	CK(cudaMemcpy(c.tgt, ht.data(), sizeof(float) * 7 * n, cudaMemcpyHostToDevice));

	if (c.filter && strstr(c.filter, "small")) {
		// the 10k x 10k regime: which geometry / grain / grid, and where the launch's time goes
		c.filter = nullptr;
		std::vector<float> rows((size_t)n * 7);
		for (int i = 0; i < n; ++i) { float *r = &rows[(size_t)i * 7]; r[0] = hA[i].x; r[1] = hA[i].y; r[2] = hA[i].z; r[3] = hB[i].x; r[4] = hB[i].y; r[5] = hB[i].z; r[6] = 0.01f; }
		CK(cudaMemcpy(c.raw, rows.data(), sizeof(float) * 7 * n, cudaMemcpyHostToDevice));
		typedef P3DVel<REG_WINCKELMANS> PW;
		const int grids[3] = {c.sms, 2 * c.sms, 4 * c.sms};
		for (int direct = 0; direct < 2; ++direct) for (int gr = 0; gr < 2; ++gr) for (int gi = 0; gi < 3; ++gi) {
			g_direct = direct; g_grain = gr ? 32 : 0; g_grid_abs = grids[gi];
			const bool t = gi == 0;
			run_new<PW, 1, 128, 8, 1, 0, 0>(c, "vel-w", 1, t, 5);
			run_new<PW, 2, 256, 3, 2, 0, 0>(c, "vel-w", 1, t, 5);
			run_new<PW, 4, 256, 2, PW::VW4, PW::OPT4, 0>(c, "vel-w", 1, t, 5);
			run_new<PW, 8, 128, 2, PW::VW8, PW::OPT8, 0>(c, "vel-w", 1, t, 5);
			if (!direct && !gr) {
				run_new<PW, 4, 256, 2, PW::VW4, PW::OPT4, 256>(c, "vel-w", 1, t, 5);
				run_new<PW, 8, 128, 2, PW::VW8, PW::OPT8, 256>(c, "vel-w", 1, t, 5);
			}
		}
		printf("done\n");
		return 0;
	}
	if (c.filter && strstr(c.filter, "guards")) {
		// the ops whose coincident-pair guard is a real selection, in their tuned instances (A/B of pair_math.cuh's
		// CVTX_GUARD_BY_MASK: build this tool once with -DCVTX_GUARD_BY_MASK=0 and once with =1)
		c.filter = nullptr;
		run_new<P3DVisc<REG_WINCKELMANS>, 8, 128, 2, 8, 1, 256>(c, "visc-winckelmans", 4);
		run_new<P3DVisc<REG_GAUSSIAN>, 4, 256, 2, 2, 1, 256>(c, "visc-gaussian", 4);
		run_new<P3DVisc<REG_GAUSSIAN>, 8, 128, 2, 2, 1, 256>(c, "visc-gaussian", 4);
		run_new<P2DVisc<REG_GAUSSIAN>, 8, 128, 2, 2, 0, 256>(c, "p2dvisc-gaussian", 4);
		run_new<P2DVisc<REG_GAUSSIAN>, 4, 256, 2, 2, 1, 256>(c, "p2dvisc-gaussian", 4);
		run_new<P2DVisc<REG_WINCKELMANS>, 4, 256, 2, 2, 0, 256>(c, "p2dvisc-winckelmans", 4);
		run_new<P2DVisc<REG_WINCKELMANS>, 8, 128, 2, 8, 0, 256>(c, "p2dvisc-winckelmans", 4);
		run_new<P3DDvort<REG_WINCKELMANS>, 4, 256, 2, 2, 1, 256>(c, "dvort-winckelmans", 4);
		run_new<P3DDvort<REG_WINCKELMANS>, 8, 128, 2, 8, 0, 256>(c, "dvort-winckelmans", 4);
		run_new<P3DVelDvort<REG_WINCKELMANS>, 4, 256, 2, 2, 1, 256>(c, "veldvort-winckelmans", 4);
		printf("done\n");
		return 0;
	}
	if (c.filter && strstr(c.filter, "f3d")) {
		// filaments: start uniform in the box, end = start + U(-0.1, 0.1)^3, packed on the host
		std::vector<float> rows((size_t)n * 7);
		std::vector<float4> fa(n), fb(n), fc(n);
		for (int i = 0; i < n; ++i) {
			float *r = &rows[(size_t)i * 7];
			for (int k = 0; k < 3; ++k) { r[k] = rnd(); r[3 + k] = r[k] + 0.02f * (rnd() - 5.0f); }
			r[6] = rnd();
			pack_source(SRC_F3D, r, fa[i], fb[i], fc[i]);
			for (int k = 0; k < 6; ++k) ht[(size_t)i * 7 + k] = rnd();       // independent target particles
		}
		CK(cudaMemcpy(c.raw, rows.data(), sizeof(float) * 7 * n, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(c.A, fa.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(c.B, fb.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(c.C, fc.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(c.tgt, ht.data(), sizeof(float) * 7 * n, cudaMemcpyHostToDevice));
		c.filter = nullptr;
		sweep_f3d<F3DVel>(c, "f3dvel");
		sweep_f3d<F3DDvort>(c, "f3ddvort");
		printf("done\n");
		return 0;
	}
#ifndef PART
#define PART -1
#endif
#ifdef FULL_SWEEP
#define SW(k, POL, REG, nm) if (PART < 0 || PART == (k) % 8) sweep<POL<REG>, cvtx_r1::POL<cvtx_r1::REG>>(c, nm, chunks);
#else
#define SW(k, POL, REG, nm)
#endif
	SW(0, P3DVel, REG_SINGULAR, "vel-singular") SW(1, P3DVel, REG_WINCKELMANS, "vel-winckelmans")
	SW(2, P3DVel, REG_PLANETARY, "vel-planetary") SW(3, P3DVel, REG_GAUSSIAN, "vel-gaussian")
	SW(4, P3DDvort, REG_SINGULAR, "dvort-singular") SW(5, P3DDvort, REG_WINCKELMANS, "dvort-winckelmans")
	SW(6, P3DDvort, REG_PLANETARY, "dvort-planetary") SW(7, P3DDvort, REG_GAUSSIAN, "dvort-gaussian")
	SW(8, P3DVisc, REG_WINCKELMANS, "visc-winckelmans") SW(9, P3DVisc, REG_GAUSSIAN, "visc-gaussian")
	SW(10, P3DVort, REG_WINCKELMANS, "vort-winckelmans") SW(11, P3DVort, REG_PLANETARY, "vort-planetary")
	SW(12, P3DVort, REG_GAUSSIAN, "vort-gaussian") SW(13, P2DVel, REG_SINGULAR, "p2dvel-singular")
	SW(14, P2DVel, REG_WINCKELMANS, "p2dvel-winckelmans") SW(15, P2DVel, REG_PLANETARY, "p2dvel-planetary")
	SW(16, P2DVel, REG_GAUSSIAN, "p2dvel-gaussian") SW(17, P2DVisc, REG_WINCKELMANS, "p2dvisc-winckelmans")
	SW(18, P2DVisc, REG_GAUSSIAN, "p2dvisc-gaussian") SW(19, P3DVelDvort, REG_SINGULAR, "veldvort-singular")
	SW(20, P3DVelDvort, REG_WINCKELMANS, "veldvort-winckelmans") SW(21, P3DVelDvort, REG_PLANETARY, "veldvort-planetary")
	SW(22, P3DVelDvort, REG_GAUSSIAN, "veldvort-gaussian")
	printf("done\n");
	return 0;
}

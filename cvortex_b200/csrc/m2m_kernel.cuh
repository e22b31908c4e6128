// m2m_kernel.cuh -- the all-pairs kernel: every M2M op of the hot path is this
// one template, instantiated with a pair policy from pair_math.cuh.
//
// It replaces the reference's OpenCL scheme (src/nbody.cl:38-73 + the host
// loop src/ocl_P3D.cpp:259-291): there one work-item evaluates ONE pair, a
// 256-wide shared-memory tree adds them up, work-item 0 read-modify-writes the
// target's result, and the host launches one NDRange per 256 sources.  Here
//
//   grid  = (target tiles, source chunks); ONE launch per call
//   block = B threads, each owning T targets held in registers for the
//           whole kernel (positions + per-target attributes)
//   the block streams its source chunk through shared memory in tiles of S
//   packed records, double-buffered with 1-D TMA bulk copies
//   (cp.async.bulk + mbarrier transaction counts) issued by one thread, so the
//   math warps spend no issue slots on loads or address arithmetic;
//   every thread reads the same source record at the same time -> LDS.128
//   broadcasts, 2 per source per warp, amortised over T targets;
//   running sums are FP32 per chain (a tile, or Policy::CHAIN sources) and
//   are flushed into FP64 accumulators, matching the reference CPU path's
//   double accumulation (src/P3D.cpp:237-249) at ~0.1 % extra instructions;
//   ops whose coincident-pair / non-finite guards only ever replace an inf or a
//   NaN (Policy::OPTIMISTIC, pair_math.cuh "GUARDS") run the pair loop without
//   them and check the FP32 running sums once per chain; a chain that comes out
//   non-finite is evaluated again with the guards (bit-identical results, and
//   2-6 ALU-pipe instructions fewer per pair);
//   the epilogue applies Policy::finish() in FP64 and either writes the
//   final floats (one chunk) or FP64 partials that reduce_partials_kernel
//   adds in a fixed order (deterministic, no atomics).
//
// The source-chunk dimension exists for load balance only: 1M targets give 977
// target tiles of 1024, i.e. 3.3 waves on 296 resident blocks (18 % tail
// loss); splitting the sources C ways turns that into 3.3*C waves of work
// units that the hardware block scheduler hands out dynamically.  blockIdx.x
// (fastest) walks the target tiles so that co-resident blocks read the same
// source chunk from L2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "op_table.h"

namespace cvtx {

constexpr int kSrcTile = 256;   // S: packed sources per shared-memory tile (also the padding quantum)

struct M2MArgs {
	const float4 *srcA;        // packed source records, padded to a multiple of kSrcTile
	const float4 *srcB;        // second record (unused when Policy::NSRC4 == 1)
	int n_src_tiles;           // total tiles of kSrcTile sources
	int tiles_per_chunk;       // tiles handled by one blockIdx.y
	const float *tgt;          // raw target rows, Policy::TCOLS floats each
	int n_tgt;
	float *out;                // final result, Policy::NOUT floats per target (gridDim.y == 1)
	double *partial;           // [gridDim.y][n_tgt][NOUT] FP64 partials (gridDim.y > 1)
	PairConsts k;
	int exact_only;            // 1: always evaluate the guarded pair form (never the optimistic one)
};

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP) ----------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	asm volatile(
	    "{\n\t"
	    ".reg .pred p;\n\t"
	    "WAIT_%=:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
	    "@p bra DONE_%=;\n\t"
	    "bra WAIT_%=;\n\t"
	    "DONE_%=:\n\t"
	    "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <class P, int T, int B, int MINB>
__global__ void __launch_bounds__(B, MINB) m2m_kernel(const M2MArgs args)
{
	constexpr int S = kSrcTile;
	constexpr int CHAIN = P::CHAIN ? P::CHAIN : S;
#ifndef CVTX_UNROLL
#define CVTX_UNROLL 8
#endif
	constexpr int UNROLL = CHAIN < CVTX_UNROLL ? CHAIN : CVTX_UNROLL;
	static_assert(S % CHAIN == 0, "a tile must hold whole chains");
	constexpr uint32_t kTileBytes = S * sizeof(float4);

	__shared__ __align__(128) float4 tileA[2][S];
	__shared__ __align__(128) float4 tileB[2][P::NSRC4 == 2 ? S : 1];
	__shared__ __align__(8) uint64_t full[2];

	const int tid = threadIdx.x;
	const int tile0 = blockIdx.y * args.tiles_per_chunk;
	const int ntile = min(args.tiles_per_chunk, args.n_src_tiles - tile0);
	const float4 *gA = args.srcA + (size_t)tile0 * S;
	const float4 *gB = args.srcB + (size_t)tile0 * S;

	if (tid == 0) {
		mbar_init(&full[0], 1);
		mbar_init(&full[1], 1);
		mbar_fence_init();
	}
	__syncthreads();
	if (tid == 0 && ntile > 0) {
		mbar_expect_tx(&full[0], kTileBytes * P::NSRC4);
		bulk_g2s(tileA[0], gA, kTileBytes, &full[0]);
		if (P::NSRC4 == 2) bulk_g2s(tileB[0], gB, kTileBytes, &full[0]);
	}

	// ---- this thread's T targets, strided by B so a warp touches contiguous rows.
	// Two targets share one Vec<2> (packed FP32x2 lanes) when T is even.
	constexpr int W = (T % 2 == 0) ? 2 : 1;
	constexpr int NV = T / W;
	const long base = (long)blockIdx.x * (B * T) + tid;
	Vec<W> tg[NV][P::NTGT];
	double dacc[T][P::NACC];      // (moving these to shared memory to raise occupancy was measured: no gain)
#pragma unroll
	for (int t = 0; t < T; ++t) {
		long i = base + (long)t * B;
		i = i < args.n_tgt ? i : (long)args.n_tgt - 1;          // clamp: tail threads redo the last target, never store
		float one[P::NTGT];
		P::load_target(args.tgt + i * P::TCOLS, one);
#pragma unroll
		for (int c = 0; c < P::NTGT; ++c) tg[t / W][c].set(t % W, one[c]);
#pragma unroll
		for (int c = 0; c < P::NACC; ++c) dacc[t][c] = 0.0;
	}

	// A warp none of whose target slots is real (most of the block in a few-target call) only
	// keeps the tile hand-over in step.  Slot (t = 0, lane 0) is the warp's lowest target index.
	// Only the T = 1 geometry, the one the planner gives few-target calls, carries the test: in
	// the large-problem geometries it would cost the Gaussian kernels 0.8 % for a tail that is
	// one block in a thousand.
	const bool idle_warp = T == 1 && (long)blockIdx.x * (B * T) + (tid & ~31) >= (long)args.n_tgt;
	const bool optimistic = P::OPTIMISTIC && !args.exact_only;

	for (int it = 0; it < ntile; ++it) {
		const int buf = it & 1;
		if (tid == 0 && it + 1 < ntile) {                          // prefetch the next tile into the other buffer
			mbar_expect_tx(&full[buf ^ 1], kTileBytes * P::NSRC4);
			bulk_g2s(tileA[buf ^ 1], gA + (size_t)(it + 1) * S, kTileBytes, &full[buf ^ 1]);
			if (P::NSRC4 == 2) bulk_g2s(tileB[buf ^ 1], gB + (size_t)(it + 1) * S, kTileBytes, &full[buf ^ 1]);
		}
		if (!idle_warp) {
			mbar_wait(&full[buf], (it >> 1) & 1);

			const float4 *sA = tileA[buf];
			const float4 *sB = tileB[P::NSRC4 == 2 ? buf : 0];
#pragma unroll 1
			for (int j0 = 0; j0 < S; j0 += CHAIN) {
				Vec<W> acc[NV][P::NACC];
#pragma unroll
				for (int v = 0; v < NV; ++v)
#pragma unroll
					for (int c = 0; c < P::NACC; ++c) acc[v][c] = bc<W>(0.0f);
				bool guarded = true;
				if (P::OPTIMISTIC && optimistic) {
#pragma unroll UNROLL
					for (int j = 0; j < CHAIN; ++j) {
						const float4 a = sA[j0 + j];
						float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
						if (P::NSRC4 == 2) b = sB[j0 + j];
#pragma unroll
						for (int v = 0; v < NV; ++v) P::template pair<W, false>(tg[v], a, b, acc[v], args.k);
					}
					// inf and NaN survive any further addition, so one sum over this thread's running sums
					// tells whether a guard would have fired anywhere in the chain (a sum of finite values
					// that overflows only costs a needless second evaluation)
					Vec<W> chk = acc[0][0];
#pragma unroll
					for (int v = 0; v < NV; ++v)
#pragma unroll
						for (int c = (v == 0 ? 1 : 0); c < P::NACC; ++c) chk = vadd(chk, acc[v][c]);
					const float s = W == 2 ? chk.lane(0) + chk.lane(1) : chk.lane(0);
					guarded = !(fabsf(s) <= 3.40282346e38f);
					if (guarded) {
#pragma unroll
						for (int v = 0; v < NV; ++v)
#pragma unroll
							for (int c = 0; c < P::NACC; ++c) acc[v][c] = bc<W>(0.0f);
					}
				}
				if (guarded) {
#pragma unroll UNROLL
					for (int j = 0; j < CHAIN; ++j) {
						const float4 a = sA[j0 + j];
						float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
						if (P::NSRC4 == 2) b = sB[j0 + j];
#pragma unroll
						for (int v = 0; v < NV; ++v) P::template pair<W, true>(tg[v], a, b, acc[v], args.k);
					}
				}
#pragma unroll
				for (int t = 0; t < T; ++t)
#pragma unroll
					for (int c = 0; c < P::NACC; ++c) dacc[t][c] += (double)acc[t / W][c].lane(t % W);
			}
		}
		__syncthreads();      // everyone is done with tile[buf] before it is refilled two iterations on
	}

	// ---- epilogue: FP64 finish, then final floats or FP64 partials
#pragma unroll
	for (int t = 0; t < T; ++t) {
		const long i = base + (long)t * B;
		if (i < args.n_tgt) {
			double res[P::NOUT];
			P::finish(args.tgt + i * P::TCOLS, dacc[t], res, args.k);
			if (gridDim.y == 1) {
#pragma unroll
				for (int c = 0; c < P::NOUT; ++c) args.out[i * P::NOUT + c] = (float)res[c];
			} else {
				double *dst = args.partial + ((size_t)blockIdx.y * args.n_tgt + i) * P::NOUT;
#pragma unroll
				for (int c = 0; c < P::NOUT; ++c) dst[c] = res[c];
			}
		}
	}
}

// out[i] = (float) sum_c partial[c][i], chunks added in index order.
__global__ void reduce_partials_kernel(const double *__restrict__ partial, float *__restrict__ out,
                                       long n_vals, int n_chunks)
{
	const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_vals) return;
	double s = 0.0;
	for (int c = 0; c < n_chunks; ++c) s += partial[(size_t)c * n_vals + i];
	out[i] = (float)s;
}

// The same sum for few values and many chunks (a few-target call splits its sources thousands of
// ways): one warp per value, lane l adds chunks l, l + 32, ... in order, then a fixed shuffle tree.
__global__ void reduce_partials_wide_kernel(const double *__restrict__ partial, float *__restrict__ out,
                                            long n_vals, int n_chunks)
{
	const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (i >= n_vals) return;
	double s = 0.0;
	for (int c = lane; c < n_chunks; c += 32) s += partial[(size_t)c * n_vals + i];
	for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
	if (lane == 0) out[i] = (float)s;
}

// Raw rows -> packed float4 records, padded with zero-strength records (pad_source) to n_pad, a
// multiple of kSrcTile.
__global__ void pack_sources_kernel(int kind, int cols, const float *__restrict__ rows, int n, int n_pad,
                                    float4 *__restrict__ A, float4 *__restrict__ Bq)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_pad) return;
	float4 a, b;
	pad_source(kind, a, b);
	if (i < n) {
		float row[7];
		for (int c = 0; c < cols; ++c) row[c] = rows[(size_t)i * cols + c];
		pack_source(kind, row, a, b);
	}
	A[i] = a;
	if (Bq) Bq[i] = b;
}

// ---------------------------------------------------------------------------
// cvtx_F3D_inf_mtrx on the device (reference src/F3D.cpp:204-227, CPU only
// there): out[i * n + j] = u_j(x_i) . dir_i, the dense m x n influence matrix of n filaments on
// m points.  Nothing is reduced, every pair is an output, so the roles flip against m2m_kernel:
// a THREAD owns W filaments (lanes of Vec<W>, registers) so that a warp writes 32 consecutive
// columns of a row (coalesced 128-byte stores), and the POINTS of a row tile are broadcast from
// shared memory.  40 lane-ops + 4 for the projection per element and 4 bytes written: at the
// ~750 G elements/s the FP32 pipe allows that is 3.0 TB/s of stores, so the kernel sits near both
// roofs at once.
template <int W, int B>
__global__ void __launch_bounds__(B) f3d_inf_mtrx_kernel(const float *__restrict__ fil, int n,
                                                          const float *__restrict__ mes, const float *__restrict__ dir,
                                                          int row0, int row1, int rows_per_block,
                                                          float *__restrict__ out /* row `row0` of the matrix */,
                                                          float one /* 1.0f at run time, see cross_rounded() */)
{
	constexpr int TILE = 128;
	__shared__ float4 sx[TILE], sd[TILE];
	const int tid = threadIdx.x;
	const long jbase = (long)blockIdx.x * (B * W) + tid;
	Vec<W> ax, ay, az, bx, by, bz, g;
	bool live[W];
#pragma unroll
	for (int l = 0; l < W; ++l) {
		long j = jbase + (long)l * B;
		live[l] = j < n;
		j = live[l] ? j : (long)n - 1;
		const float *r = fil + j * 7;
		ax.set(l, r[0]); ay.set(l, r[1]); az.set(l, r[2]);
		bx.set(l, r[3]); by.set(l, r[4]); bz.set(l, r[5]);
		g.set(l, r[6] / (4.0f * 3.14159265359f));                       // strength / (4 pi_f), src/F3D.cpp:47
	}
	const int first = row0 + blockIdx.y * rows_per_block;
	const int last = min(first + rows_per_block, row1);
	for (int t0 = first; t0 < last; t0 += TILE) {
		const int cnt = min(TILE, last - t0);
		__syncthreads();
		for (int k = tid; k < cnt; k += B) {
			const float *x = mes + (size_t)(t0 + k) * 3, *d = dir + (size_t)(t0 + k) * 3;
			sx[k] = make_float4(x[0], x[1], x[2], 0.f);
			sd[k] = make_float4(d[0], d[1], d[2], 0.f);
		}
		__syncthreads();
#pragma unroll 4
		for (int k = 0; k < cnt; ++k) {
			const float4 x = sx[k], d = sd[k];
			const Vec<W> px = vsub(x.x, ax), py = vsub(x.y, ay), pz = vsub(x.z, az);      // r1 = x - a
			const Vec<W> qx = vsub(x.x, bx), qy = vsub(x.y, by), qz = vsub(x.z, bz);      // r2 = x - b
			const Vec<W> ox = vsub(px, qx), oy = vsub(py, qy), oz = vsub(pz, qz);         // r0 = r1 - r2
			Vec<W> cx, cy, cz;
			cross_rounded(px, py, pz, qx, qy, qz, one, cx, cy, cz);
			const Vec<W> c2 = vfma(cz, cz, vfma(cy, cy, vmul(cx, cx)));
			const Vec<W> n1 = vfma(pz, pz, vfma(py, py, vmul(px, px)));
			const Vec<W> n2 = vfma(qz, qz, vfma(qy, qy, vmul(qx, qx)));
			const Vec<W> d1 = vfma(pz, oz, vfma(py, oy, vmul(px, ox)));
			const Vec<W> d2 = vfma(qz, oz, vfma(qy, oy, vmul(qx, ox)));
			const Vec<W> t1 = vmul(vrcp(c2), g);
			const Vec<W> t2 = vfms(d1, vrsqrt(n1), vmul(d2, vrsqrt(n2)));
			const Vec<W> proj = vfma(cz, d.z, vfma(cy, d.y, vmul(cx, d.x)));               // c . dir
			const Vec<W> val = vmul(vmul(t1, t2), proj);
			float *row = out + (size_t)(t0 + k - row0) * n;
#pragma unroll
			for (int l = 0; l < W; ++l) {
				const bool ok = (fabsf(t1.lane(l)) <= 3.40282346e38f) && (fabsf(t2.lane(l)) <= 3.40282346e38f);
				if (live[l]) row[jbase + (long)l * B] = ok ? val.lane(l) : 0.0f;
			}
		}
	}
}

}  // namespace cvtx

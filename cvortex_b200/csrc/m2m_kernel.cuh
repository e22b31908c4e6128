// m2m_kernel.cuh -- the all-pairs kernel: every M2M op of the hot path is this
// one template, instantiated with a pair policy from pair_math.cuh.
//
// It replaces the reference's OpenCL scheme (src/nbody.cl:38-73 + the host
// loop src/ocl_P3D.cpp:259-291): there one work-item evaluates ONE pair, a
// 256-wide shared-memory tree adds them up, work-item 0 read-modify-writes the
// target's result, and the host launches one NDRange per 256 sources.  Here
//
//   ONE launch per call.  The work -- (target tile) x (source grain) cells, target-tile-major -- is
//   one linear sequence cut into `gridDim.x` contiguous runs of equal length, gridDim.x = a small
//   multiple (4) of the blocks that are resident at once: every block does the same number of pair
//   evaluations whatever the target count, and the hardware block scheduler always has a next run to
//   hand to an SM whose favoured block has finished (two blocks that share an SM do NOT advance at the
//   same pace: with exactly one resident set the faster one of each SM was measured to finish after 54 %
//   of the launch and its partner ran the rest alone, 1.5 - 3 % slower overall).
//   A block walks its run: B threads, each owning T targets held in registers (positions +
//   per-target attributes) while the block streams source tiles of S packed records through shared
//   memory, double-buffered with 1-D TMA bulk copies (cp.async.bulk + mbarrier transaction counts)
//   issued by one thread, so the math warps spend no issue slots on loads or address arithmetic; the
//   stream simply continues across target-tile boundaries.
//   every thread reads the same source record at the same time -> LDS.128
//   broadcasts, amortised over T targets;
//   running sums are FP32 per chain (a grain: 256 sources, or 32 for small
//   source sets) and are flushed into FP64 accumulators, matching the
//   reference CPU path's double accumulation (src/P3D.cpp:237-249);
//   ops whose coincident-pair guards only ever replace an inf or a NaN
//   (Policy::OPTIMISTIC, pair_math.cuh "GUARDS") run the pair loop without
//   them and check the FP32 running sums once per chain; a chain that comes out
//   non-finite is evaluated again with the guards (bit-identical results);
//   the filament ops (Policy::HYBRID, pair_math.cuh "FILAMENTS") run their fast
//   form over sub-chains of 32 sources and re-evaluate, per target, the
//   sub-chains whose flag fired in the reference's own arithmetic;
//   a run that covers a target tile completely finishes it in registers
//   (Policy::finish() in FP64) and writes the final floats.  Only the tiles a
//   run boundary cuts -- at most two per block -- go through memory: each
//   block writes its FP64 piece, takes a ticket on the tile, and the block that
//   arrives last adds the pieces in run order and writes the result
//   (deterministic, no floating-point atomics, no second kernel, and a few MB
//   of scratch where the first version of this kernel wrote [chunks][m] FP64
//   partials -- 550 MB at 1M x 1M -- for a reduce kernel to read back).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "op_table.h"

namespace cvtx {

constexpr int kSrcTile = 256;   // S: packed sources per shared-memory tile (also the padding quantum)

struct M2MArgs {
	const float4 *srcA;        // packed source records, padded to a multiple of kSrcTile
	const float4 *srcB;        // second record (unused when Policy::NSRC4 == 1)
	const float4 *srcC;        // third record (filaments)
	const float *src_raw;      // the raw source rows (filaments: the reference-arithmetic tier reads them)
	int n_src;                 // real sources
	int n_src_tiles;           // tiles of kSrcTile sources
	int grain;                 // sources per grain = FP32 chain length: 256, or 32 for small source sets
	long long total_grains;    // (target tiles) x (n_src_tiles x 256 / grain)
	const float *tgt;          // raw target rows, Policy::TCOLS floats each
	int n_tgt;
	float *out;                // final result, Policy::NOUT floats per target
	double *pieces;            // [2 x gridDim.x][T][B][NOUT] FP64 pieces of the target tiles a run boundary cuts
	int *tickets;              // [target tiles], zero between launches
	const int *f3d_mode;       // filaments: which fast form (pair_math.cuh f3d_pick_mode), decided while packing
	PairConsts k;
	int exact_only;            // 1: always evaluate the guarded pair form (never the optimistic one)
	int direct;                // 1: no packed copy exists; blocks pack the raw rows of a tile straight into shared memory
	int defer_finish;          // 1: cut target tiles are summed by finish_pieces_kernel after this launch, not by their last block
	unsigned long long *block_times;   // diagnostics (tools/kernel_ab): per block {SM id, start ns, end ns}, or null
	// cvtx_P3D_M2M_vort only: *sparse_gate = (target tile, source tile) pairs whose boxes meet; at or below sparse_max the
	// call belongs to sparse_tiles_kernel and this kernel returns at once (null: no such gate)
	const unsigned long long *sparse_gate;
	unsigned long long sparse_max;
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

// The run of grains block b of R owns: [run_begin(b), run_begin(b + 1)).
__host__ __device__ __forceinline__ long long run_begin(long long b, long long R, long long total) {
	return (long long)(((unsigned __int128)b * (unsigned __int128)total) / (unsigned __int128)R);
}
// The block whose run holds grain x.
__host__ __device__ __forceinline__ long long run_of(long long x, long long R, long long total) {
	return (long long)((((unsigned __int128)(x + 1)) * (unsigned __int128)R - 1) / (unsigned __int128)total);
}

// The filament ops' slow tier: sub-chain [j0, j0 + F3D_SUB) of the source order for ONE target, in the
// reference's own arithmetic (Policy::exact), from the raw rows.  Kept out of line: it runs for one pair
// in a million and must not cost the pair loop registers.
template <class P>
__device__ __noinline__ void exact_subchain(const float *__restrict__ src_raw, long j0, long j1, const float *__restrict__ tgt_row, float *sums)
{
	float acc[P::NACC];
#pragma unroll
	for (int c = 0; c < P::NACC; ++c) acc[c] = 0.0f;
	for (long j = j0; j < j1; ++j) P::exact(src_raw + j * 7, tgt_row, acc);
#pragma unroll
	for (int c = 0; c < P::NACC; ++c) sums[c] = acc[c];
}

// VW: lanes per Vec (pair_math.cuh LANES): 2 = one packed pair at a time, 4 / 8 = two / four packed pairs
// evaluated stage by stage (more independent instructions in flight, more registers).
// OPT (bit set): 1 = the FP64 accumulators live in shared memory instead of registers (T x NACC x 2 registers
// handed back to the pair loop's schedule; one LDS.64 + DADD + STS.64 per sum and chain);
// 2 = the pair loops are written `#pragma unroll 8` over the chain instead of an explicit 8-wide body;
// 4 = filament ops: the sub-chains a target's flag rejects are only marked (and their fast sums dropped) inside the
// sub-chain loop; the reference-arithmetic re-evaluation -- an out-of-line call -- happens once per chain, after it.
enum { M2M_SMEM_ACC = 1, M2M_PLAIN_LOOP = 2, M2M_DEFER_EXACT = 4 };
// dynamic shared memory a launch of m2m_kernel<P, T, B, .., OPT> needs (beyond 48 KB in total the launcher has to
// raise cudaFuncAttributeMaxDynamicSharedMemorySize first)
template <class P, int T, int B, int OPT> constexpr size_t m2m_smem_bytes() { return (OPT & M2M_SMEM_ACC) ? sizeof(double) * T * P::NACC * B : 0; }
template <class P, int T, int B, int MINB, int VW = 2, int OPT = 0, int GRAIN = 0>
__global__ void __launch_bounds__(B, MINB) m2m_kernel(const M2MArgs args)
{
	constexpr int S = kSrcTile;
#ifndef CVTX_UNROLL
#define CVTX_UNROLL 8
#endif
	constexpr int UNROLL = CVTX_UNROLL;
	constexpr uint32_t kTileBytes = S * sizeof(float4);
	constexpr int NR = P::NSRC4;
	if (args.sparse_gate && *args.sparse_gate <= args.sparse_max) return;      // sparse_tiles_kernel has done this call

	__shared__ __align__(128) float4 tile[2][NR][S];
	__shared__ __align__(8) uint64_t full[2];
	__shared__ int s_ticket;
	constexpr bool SACC = (OPT & M2M_SMEM_ACC) != 0;
	extern __shared__ double s_acc[];                                    // SACC: [t][c][thread] (dynamic, m2m_smem_bytes), conflict-free

	const int tid = threadIdx.x;
	const int G = GRAIN ? GRAIN : args.grain;                            // GRAIN != 0: compile-time chain length
	const int gps = S / G;                                              // grains per source tile
	const int gpt = args.n_src_tiles * gps;                             // grains per target tile
	// this block's run, as (target tile, grain within the tile, grains left): 32-bit state in the loop
	int tt, gs, left;
	bool from_start;                                                    // the run entered tile tt at the tile's first grain
	{
		const long long g0 = run_begin(blockIdx.x, gridDim.x, args.total_grains), g1 = run_begin(blockIdx.x + 1, gridDim.x, args.total_grains);
		tt = (int)(g0 / gpt); gs = (int)(g0 % gpt); left = (int)(g1 - g0);
		from_start = gs == 0;
	}

	if (tid == 0) {
		mbar_init(&full[0], 1);
		mbar_init(&full[1], 1);
		mbar_fence_init();
	}
	__syncthreads();

	// Small source sets and the small-tile geometries (GRAIN == 0 instances) of the particle ops take their
	// sources straight from the caller's rows: every thread packs a record or two of the step's tile into
	// shared memory.  A tile is then used by few blocks (few targets) or the whole call is a few microseconds
	// (10k x 10k), and a separate pack kernel -- a launch, a write and a read of the packed copy, a dependency --
	// costs more than the two loads per thread it saves here.
	constexpr bool CAN_DIRECT = GRAIN == 0 && !P::HYBRID;
	const bool direct = CAN_DIRECT && args.direct;
	auto fill = [&](int src_tile, int buf) {                            // all threads; followed by __syncthreads()
		constexpr int KIND = P::NSRC4 == 1 ? SRC_P2D : SRC_P3D, COLS = P::NSRC4 == 1 ? 4 : 7;
		for (int i = tid; i < S; i += B) {
			const long j = (long)src_tile * S + i;
			f4 a, b, c;
			pad_source(KIND, a, b, c);
			if (j < args.n_src) {
				float row[COLS];
#pragma unroll
				for (int q = 0; q < COLS; ++q) row[q] = __ldg(args.src_raw + j * COLS + q);
				pack_source(KIND, row, a, b, c);
			}
			tile[buf][0][i] = a;
			if (NR >= 2) tile[buf][NR >= 2 ? 1 : 0][i] = b;
		}
	};
	auto fetch = [&](int src_tile, int buf) {                           // thread 0 only
		const size_t off = (size_t)src_tile * S;
		mbar_expect_tx(&full[buf], kTileBytes * NR);
		bulk_g2s(tile[buf][0], args.srcA + off, kTileBytes, &full[buf]);
		if (NR >= 2) bulk_g2s(tile[buf][NR >= 2 ? 1 : 0], args.srcB + off, kTileBytes, &full[buf]);
		if (NR >= 3) bulk_g2s(tile[buf][NR >= 3 ? 2 : 0], args.srcC + off, kTileBytes, &full[buf]);
	};
	if (args.block_times && tid == 0) {
		unsigned smid; unsigned long long t;
		asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
		args.block_times[3 * blockIdx.x] = smid; args.block_times[3 * blockIdx.x + 1] = t;
	}
	if (!direct && tid == 0 && left > 0) fetch(gs / gps, 0);

	// Two targets share one Vec<2> (packed FP32x2 lanes) when T is even.
	constexpr int W = (T % 2 == 0) ? (T < VW ? T : VW) : 1;
	constexpr int NV = T / W;
	Vec<W> tg[NV][P::NTGT];
	double dacc[SACC ? 1 : T][P::NACC];
	const bool optimistic = P::OPTIMISTIC && !args.exact_only;
	int f3d_mode = F3D_WIDE;
	if (P::HYBRID) f3d_mode = *args.f3d_mode;

	bool fresh = true;                                                  // the next step starts a target tile
	bool idle_warp = false;
	int it = 0;
	while (left > 0) {
		const int buf = it & 1;
		// this step: grains [gs, gs + n) of tile tt, all inside one source tile
		int n = gps - gs % gps;
		n = n < left ? n : left;
		if (CAN_DIRECT && direct) {
			fill(gs / gps, buf);
			__syncthreads();
		} else if (left > n) {                                          // prefetch the next step's source tile into the other buffer
			const int gs_next = gs + n == gpt ? 0 : gs + n;
			if (tid == 0) fetch(gs_next / gps, buf ^ 1);
		}
		if (fresh) {
			// ---- this thread's T targets of the new tile, strided by B so a warp touches contiguous rows
			fresh = false;
			const long base = (long)tt * (B * T) + tid;
#pragma unroll
			for (int t = 0; t < T; ++t) {
				long i = base + (long)t * B;
				i = i < args.n_tgt ? i : (long)args.n_tgt - 1;      // clamp: tail threads redo the last target, never store
				float one[P::NTGT];
				P::load_target(args.tgt + i * P::TCOLS, one);
#pragma unroll
				for (int c = 0; c < P::NTGT; ++c) tg[t / W][c].set(t % W, one[c]);
#pragma unroll
				for (int c = 0; c < P::NACC; ++c) {
					if (SACC) s_acc[(t * P::NACC + c) * B + tid] = 0.0;
					else dacc[t][c] = 0.0;
				}
			}
			// A warp none of whose target slots is real (most of the block in a few-target call) only
			// keeps the tile hand-over in step.  Only the T = 1 geometry, the one the planner gives
			// few-target calls, carries the test.
			idle_warp = T == 1 && (long)tt * (B * T) + (tid & ~31) >= (long)args.n_tgt;
		}
		const int lo = (gs % gps) * G, hi = lo + n * G;                 // this step's sources within the tile
		if (!idle_warp) {
			if (!direct) mbar_wait(&full[buf], (it >> 1) & 1);
			const float4 *sA = tile[buf][0];
			const float4 *sB = tile[buf][NR >= 2 ? 1 : 0];
			const float4 *sC = tile[buf][NR >= 3 ? 2 : 0];
#pragma unroll 1
			for (int j0 = lo; j0 < hi; j0 += G) {
				Vec<W> acc[NV][P::NACC];
#pragma unroll
				for (int v = 0; v < NV; ++v)
#pragma unroll
					for (int c = 0; c < P::NACC; ++c) acc[v][c] = bc<W>(0.0f);
				if constexpr (P::HYBRID) {
					unsigned long long redo_mask = 0;                      // bit t * 8 + s: target slot t, sub-chain s of this chain
#pragma unroll 1
					for (int s0 = j0; s0 < j0 + G; s0 += F3D_SUB) {
						Vec<W> sub[NV][P::NACC], flag[NV];
#pragma unroll
						for (int v = 0; v < NV; ++v) {
							flag[v] = bc<W>(3.0e38f);
#pragma unroll
							for (int c = 0; c < P::NACC; ++c) sub[v][c] = bc<W>(0.0f);
						}
						if (f3d_mode == F3D_NEW) {
#pragma unroll 1
							for (int j = 0; j < F3D_SUB; j += UNROLL) {
#pragma unroll
								for (int u = 0; u < UNROLL; ++u) {
									const float4 a = sA[s0 + j + u], b = sB[s0 + j + u], c = sC[s0 + j + u];
#pragma unroll
									for (int v = 0; v < NV; ++v) P::template fast<W, F3D_NEW>(tg[v], a, b, c, sub[v], flag[v], args.k);
								}
							}
						} else {
#pragma unroll 1
							for (int j = 0; j < F3D_SUB; j += UNROLL) {
#pragma unroll
								for (int u = 0; u < UNROLL; ++u) {
									const float4 a = sA[s0 + j + u], b = sB[s0 + j + u], c = sC[s0 + j + u];
#pragma unroll
									for (int v = 0; v < NV; ++v) P::template fast<W, F3D_WIDE>(tg[v], a, b, c, sub[v], flag[v], args.k);
								}
							}
						}
						bool redo = false;
#pragma unroll
						for (int t = 0; t < T; ++t) redo |= !(flag[t / W].lane(t % W) > 0.0f);
						if ((OPT & M2M_DEFER_EXACT) && redo) {
							// mark (target slot, sub-chain) and drop the fast sums of the rejected slots
							const int sidx = (s0 - j0) / F3D_SUB;
#pragma unroll
							for (int t = 0; t < T; ++t) {
								if (!(flag[t / W].lane(t % W) > 0.0f)) {
									redo_mask |= 1ull << (t * 8 + sidx);
#pragma unroll
									for (int c = 0; c < P::NACC; ++c) sub[t / W][c].set(t % W, 0.0f);
								}
							}
							redo = false;
						}
						if (redo) {
							const long js = (long)(gs / gps) * S + s0;
							const long je = js + F3D_SUB < (long)args.n_src ? js + F3D_SUB : (long)args.n_src;
#pragma unroll
							for (int t = 0; t < T; ++t) {
								if (!(flag[t / W].lane(t % W) > 0.0f)) {
									long i = (long)tt * (B * T) + tid + (long)t * B;
									i = i < args.n_tgt ? i : (long)args.n_tgt - 1;
									float e[P::NACC];
									exact_subchain<P>(args.src_raw, js, je, args.tgt + i * P::TCOLS, e);
#pragma unroll
									for (int c = 0; c < P::NACC; ++c) sub[t / W][c].set(t % W, e[c]);
								}
							}
						}
#pragma unroll
						for (int v = 0; v < NV; ++v)
#pragma unroll
							for (int c = 0; c < P::NACC; ++c) acc[v][c] = vadd(acc[v][c], sub[v][c]);
					}
					if ((OPT & M2M_DEFER_EXACT) && redo_mask) {
						// the rejected sub-chains of this chain in the reference's own arithmetic, in source order
#pragma unroll 1
						for (int t = 0; t < T; ++t) {
							const unsigned bits = (unsigned)(redo_mask >> (t * 8)) & 0xffu;
							if (!bits) continue;
							long i = (long)tt * (B * T) + tid + (long)t * B;
							i = i < args.n_tgt ? i : (long)args.n_tgt - 1;
							float add[P::NACC];
#pragma unroll
							for (int c = 0; c < P::NACC; ++c) add[c] = 0.0f;
							for (int sidx = 0; sidx < G / F3D_SUB; ++sidx) {
								if (!((bits >> sidx) & 1u)) continue;
								const long js = (long)(gs / gps) * S + j0 + sidx * F3D_SUB;
								const long je = js + F3D_SUB < (long)args.n_src ? js + F3D_SUB : (long)args.n_src;
								float e[P::NACC];
								exact_subchain<P>(args.src_raw, js, je, args.tgt + i * P::TCOLS, e);
#pragma unroll
								for (int c = 0; c < P::NACC; ++c) add[c] = __fadd_rn(add[c], e[c]);
							}
							// (dynamic t: one select per slot instead of a register-indexed write)
#pragma unroll
							for (int tt2 = 0; tt2 < T; ++tt2)
								if (tt2 == t) {
#pragma unroll
									for (int c = 0; c < P::NACC; ++c) acc[tt2 / W][c].set(tt2 % W, __fadd_rn(acc[tt2 / W][c].lane(tt2 % W), add[c]));
								}
						}
					}
				} else {
					bool guarded = true;
					if (P::OPTIMISTIC && optimistic) {
						if (OPT & M2M_PLAIN_LOOP) {
#pragma unroll 8
							for (int j = 0; j < G; ++j) {
								const float4 a = sA[j0 + j];
								float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
								if (NR == 2) b = sB[j0 + j];
#pragma unroll
								for (int v = 0; v < NV; ++v) P::template pair<W, false>(tg[v], a, b, acc[v], args.k);
							}
						} else {
#pragma unroll 1
							for (int j = 0; j < G; j += UNROLL) {
#pragma unroll
								for (int u = 0; u < UNROLL; ++u) {
									const float4 a = sA[j0 + j + u];
									float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
									if (NR == 2) b = sB[j0 + j + u];
#pragma unroll
									for (int v = 0; v < NV; ++v) P::template pair<W, false>(tg[v], a, b, acc[v], args.k);
								}
							}
						}
						// inf and NaN survive any further addition, so one sum over this thread's running sums
						// tells whether a guard would have fired anywhere in the chain (a sum of finite values
						// that overflows only costs a needless second evaluation)
						Vec<W> chk = acc[0][0];
#pragma unroll
						for (int v = 0; v < NV; ++v)
#pragma unroll
							for (int c = (v == 0 ? 1 : 0); c < P::NACC; ++c) chk = vadd(chk, acc[v][c]);
						float s = chk.lane(0);
#pragma unroll
						for (int l = 1; l < W; ++l) s += chk.lane(l);
						guarded = !(fabsf(s) <= 3.40282346e38f);
						if (guarded) {
#pragma unroll
							for (int v = 0; v < NV; ++v)
#pragma unroll
								for (int c = 0; c < P::NACC; ++c) acc[v][c] = bc<W>(0.0f);
						}
					}
					if (guarded) {
						if (OPT & M2M_PLAIN_LOOP) {
#pragma unroll 8
							for (int j = 0; j < G; ++j) {
								const float4 a = sA[j0 + j];
								float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
								if (NR == 2) b = sB[j0 + j];
#pragma unroll
								for (int v = 0; v < NV; ++v) P::template pair<W, true>(tg[v], a, b, acc[v], args.k);
							}
						} else {
#pragma unroll 1
							for (int j = 0; j < G; j += UNROLL) {
#pragma unroll
								for (int u = 0; u < UNROLL; ++u) {
									const float4 a = sA[j0 + j + u];
									float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
									if (NR == 2) b = sB[j0 + j + u];
#pragma unroll
									for (int v = 0; v < NV; ++v) P::template pair<W, true>(tg[v], a, b, acc[v], args.k);
								}
							}
						}
					}
				}
#pragma unroll
				for (int t = 0; t < T; ++t)
#pragma unroll
					for (int c = 0; c < P::NACC; ++c) {
						if (SACC) s_acc[(t * P::NACC + c) * B + tid] += (double)acc[t / W][c].lane(t % W);
						else dacc[t][c] += (double)acc[t / W][c].lane(t % W);
					}
			}
		}
		__syncthreads();      // everyone is done with tile[buf] before it is refilled two steps on
		gs += n;
		left -= n;
		++it;

		if (left == 0 || gs == gpt) {
			// ---- leaving target tile tt: FP64 finish, then final floats or an FP64 piece
			const long base = (long)tt * (B * T) + tid;
			auto sums_of = [&](int t, double *d) {                          // this thread's FP64 sums of target slot t
#pragma unroll
				for (int c = 0; c < P::NACC; ++c) d[c] = SACC ? s_acc[(t * P::NACC + c) * B + tid] : dacc[SACC ? 0 : t][c];
			};
			if (from_start && gs == gpt) {
#pragma unroll
				for (int t = 0; t < T; ++t) {
					const long i = base + (long)t * B;
					if (i < args.n_tgt) {
						double res[P::NOUT], d[P::NACC];
						sums_of(t, d);
						P::finish(args.tgt + i * P::TCOLS, d, res, args.k);
#pragma unroll
						for (int c = 0; c < P::NOUT; ++c) args.out[i * P::NOUT + c] = (float)res[c];
					}
				}
			} else {
				const long long R = gridDim.x;
				const long long g_begin = run_begin(blockIdx.x, R, args.total_grains);
				// slot 2b: the tile the block's run starts in; 2b + 1: the tile its run ends in, if that is another one
				const long long slot = 2 * (long long)blockIdx.x + (g_begin / gpt == tt ? 0 : 1);
				double *mine = args.pieces + (size_t)slot * (T * B * P::NOUT);
#pragma unroll
				for (int t = 0; t < T; ++t) {
					const long i = base + (long)t * B;
					double res[P::NOUT], d[P::NACC];
					sums_of(t, d);
					if (i < args.n_tgt) P::finish(args.tgt + i * P::TCOLS, d, res, args.k);
					else for (int c = 0; c < P::NOUT; ++c) res[c] = 0.0;
#pragma unroll
					for (int c = 0; c < P::NOUT; ++c) __stcg(mine + ((size_t)t * B + tid) * P::NOUT + c, res[c]);
				}
				const long long b_first = run_of((long long)tt * gpt, R, args.total_grains);
				const long long b_last = run_of((long long)(tt + 1) * gpt - 1, R, args.total_grains);
				// a tile cut into very many pieces (few targets, many sources) is left to finish_pieces_kernel:
				// a warp per value there, against one thread per value walking every piece here
				if (args.defer_finish) { if (tid == 0) s_ticket = -1; }
				else {
					__threadfence();
					__syncthreads();
					if (tid == 0) s_ticket = atomicAdd(args.tickets + tt, 1);
				}
				__syncthreads();
				if (s_ticket == (int)(b_last - b_first)) {              // last to arrive: add the pieces in run order
					__threadfence();
					// T x NOUT values per thread, one piece per run that touched the tile: the loads of a piece (and of the next)
					// are all in flight before the first add waits, the adds stay in run order
					constexpr int TB = T > 4 ? 4 : T;
#pragma unroll 1
					for (int t0 = 0; t0 < T; t0 += TB) {
						double sum[TB][P::NOUT];
#pragma unroll
						for (int t = 0; t < TB; ++t)
#pragma unroll
							for (int c = 0; c < P::NOUT; ++c) sum[t][c] = 0.0;
						// (only the first run can have entered the tile from the one before it: every later run starts inside it)
						const long long first_sl = 2 * b_first + (run_begin(b_first, R, args.total_grains) / gpt == tt ? 0 : 1);
#pragma unroll 2
						for (long long bb = b_first; bb <= b_last; ++bb) {
							const long long sl = bb == b_first ? first_sl : 2 * bb;
							const double *pc = args.pieces + (size_t)sl * (T * B * P::NOUT) + ((size_t)t0 * B + tid) * P::NOUT;
#pragma unroll
							for (int t = 0; t < TB; ++t)
#pragma unroll
								for (int c = 0; c < P::NOUT; ++c) sum[t][c] += __ldcg(pc + (size_t)t * B * P::NOUT + c);
						}
#pragma unroll
						for (int t = 0; t < TB; ++t) {
							const long i = base + (long)(t0 + t) * B;
							if (i < args.n_tgt) {
#pragma unroll
								for (int c = 0; c < P::NOUT; ++c) args.out[i * P::NOUT + c] = (float)sum[t][c];
							}
						}
					}
					if (tid == 0) args.tickets[tt] = 0;                  // leave the ticket counter ready for the next launch
				}
				__syncthreads();
			}
			if (gs == gpt) { gs = 0; ++tt; }
			from_start = true;
			fresh = true;
		}
	}
	if (args.block_times && tid == 0) {
		unsigned long long t;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
		args.block_times[3 * blockIdx.x + 2] = t;
	}
}

// ---------------------------------------------------------------------------
// Box-cutoff ops (cvtx_P3D_M2M_vort: only sources inside the 5-sigma cube around a target count, reference
// src/P3D.cpp:298-322) on particles whose ORDER is spatially coherent -- what cvtx_P3D_redistribute_on_grid returns
// (ascending Morton code) and what cvtx_P3D_pedrizzetti_relaxation is called on.  Then a tile of 256 consecutive
// sources is a small cube, a tile of 1024 consecutive targets another, and all but a few per cent of the
// (target tile, source tile) pairs cannot hold a single pair inside the cutoff.  target_tile_masks_kernel
// (aux_kernels.cuh) marks the pairs whose boxes meet; here ONE block per target tile streams only the marked source
// tiles through the same TMA double buffer and the same guarded pair loop, one FP32 chain per tile flushed into
// FP64 in source order -- the chains m2m_kernel forms, minus chains that are exact zeros: the same bits.  The
// call goes here when the marked pairs are at most sparse_max (30 % of all), to m2m_kernel otherwise; both
// kernels are launched and read the count, the one that is not concerned returns at once.
struct SparseArgs {
	const float4 *srcA, *srcB;
	int n_src_tiles;
	const float *tgt; int n_tgt;
	float *out;
	const unsigned *mask; int words;           // [target tiles][words]: bit s of a row = source tile s is marked
	const unsigned long long *gate; unsigned long long sparse_max;
	PairConsts k;
};

template <class P, int T, int B>
__global__ void __launch_bounds__(B, 2) sparse_tiles_kernel(const SparseArgs args)
{
	constexpr int S = kSrcTile, UNROLL = 8, W = 2, NV = T / W;
	constexpr uint32_t kTileBytes = S * sizeof(float4);
	static_assert(P::NSRC4 == 2 && !P::HYBRID && !P::OPTIMISTIC && T % 2 == 0, "box-cutoff particle policies");
	if (*args.gate > args.sparse_max) return;                            // dense: m2m_kernel does this call
	__shared__ __align__(128) float4 tile[2][2][S];
	__shared__ __align__(8) uint64_t full[2];
	__shared__ int s_tile[2];
	const int tid = threadIdx.x, tt = blockIdx.x;
	const unsigned *row = args.mask + (size_t)tt * args.words;
	auto next_marked = [&](int from) {                                   // first marked source tile >= from, or -1 (thread 0)
		for (int w = from >> 5; w < args.words; ++w) {
			unsigned bits = __ldg(row + w);
			if (w == (from >> 5)) bits &= ~0u << (from & 31);
			if (bits) { const int s = w * 32 + __ffs(bits) - 1; return s < args.n_src_tiles ? s : -1; }
		}
		return -1;
	};
	auto fetch = [&](int src_tile, int buf) {
		const size_t off = (size_t)src_tile * S;
		mbar_expect_tx(&full[buf], kTileBytes * 2);
		bulk_g2s(tile[buf][0], args.srcA + off, kTileBytes, &full[buf]);
		bulk_g2s(tile[buf][1], args.srcB + off, kTileBytes, &full[buf]);
	};
	if (tid == 0) {
		mbar_init(&full[0], 1);
		mbar_init(&full[1], 1);
		mbar_fence_init();
		s_tile[0] = next_marked(0);
		if (s_tile[0] >= 0) fetch(s_tile[0], 0);
	}
	__syncthreads();

	Vec<W> tg[NV][P::NTGT];
	double dacc[T][P::NACC];
	const long base = (long)tt * (B * T) + tid;
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
	for (int it = 0;; ++it) {
		const int buf = it & 1;
		const int cur = s_tile[buf];
		if (cur < 0) break;
		if (tid == 0) {                                                  // the marked tile after this one goes into the other buffer
			const int nxt = next_marked(cur + 1);
			s_tile[buf ^ 1] = nxt;
			if (nxt >= 0) fetch(nxt, buf ^ 1);
		}
		mbar_wait(&full[buf], (it >> 1) & 1);
		const float4 *sA = tile[buf][0], *sB = tile[buf][1];
		Vec<W> acc[NV][P::NACC];
#pragma unroll
		for (int v = 0; v < NV; ++v)
#pragma unroll
			for (int c = 0; c < P::NACC; ++c) acc[v][c] = bc<W>(0.0f);
#pragma unroll 1
		for (int j = 0; j < S; j += UNROLL) {
#pragma unroll
			for (int u = 0; u < UNROLL; ++u) {
				const float4 a = sA[j + u], b = sB[j + u];
#pragma unroll
				for (int v = 0; v < NV; ++v) P::template pair<W, true>(tg[v], a, b, acc[v], args.k);
			}
		}
#pragma unroll
		for (int t = 0; t < T; ++t)
#pragma unroll
			for (int c = 0; c < P::NACC; ++c) dacc[t][c] += (double)acc[t / W][c].lane(t % W);
		__syncthreads();                                                  // tile[buf] and s_tile[buf] are free again; s_tile[buf ^ 1] is visible
	}
#pragma unroll
	for (int t = 0; t < T; ++t) {
		const long i = base + (long)t * B;
		if (i < args.n_tgt) {
			double res[P::NOUT];
			P::finish(args.tgt + i * P::TCOLS, dacc[t], res, args.k);
#pragma unroll
			for (int c = 0; c < P::NOUT; ++c) args.out[i * P::NOUT + c] = (float)res[c];
		}
	}
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

// aux_kernels.cuh -- the small kernels around the pair kernel (included by device_api.cu only: they are not
// templates, so they must live in one translation unit): the deferred ordered finish, source packing, the
// per-call choice of the filament form.
#pragma once
#include "m2m_kernel.cuh"

namespace cvtx {

// The ordered finish of the target tiles that more than one run touched, as a kernel of its own (M2MArgs::defer_finish):
// one WARP per output value; lane l adds pieces l, l + 32, ... in run order, then a fixed shuffle tree -- the same sum on
// every run and every device.  Used when a tile is cut into more pieces than one thread should walk (few targets against
// many sources: 1M sources on one target are 1184 pieces).
__global__ void __launch_bounds__(256) finish_pieces_kernel(const double *__restrict__ pieces, float *__restrict__ out, int n_tgt, int nout,
                                                            int slots /* T x B */, long long gpt, long long total_grains, long long R)
{
	const long long tiles_t = total_grains / gpt;
	const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;       // value index over (tile, slot, component)
	const int lane = threadIdx.x & 31;
	const long long per_tile = (long long)slots * nout;
	if (w >= tiles_t * per_tile) return;
	const long long tt = w / per_tile;
	const int v = (int)(w - tt * per_tile);                                              // slot * nout + component, the piece layout
	const int slot = v / nout, c = v - slot * nout;
	// target index of this slot: slot = t * B + thread, target = tile * slots + thread + t * B = tile * slots + slot
	const long long i = tt * slots + slot;
	if (i >= n_tgt) return;
	const long long b_first = run_of(tt * gpt, R, total_grains), b_last = run_of((tt + 1) * gpt - 1, R, total_grains);
	if (b_first == b_last) return;                                                       // one run covered the tile and wrote it
	// (only the first run can have entered the tile from the one before it: every later run starts inside it --
	// one 128-bit division per warp instead of one per piece)
	const long long first_sl = 2 * b_first + (run_begin(b_first, R, total_grains) / gpt == tt ? 0 : 1);
	double s = 0.0;
#pragma unroll 4
	for (long long bb = b_first + lane; bb <= b_last; bb += 32) {
		const long long sl = bb == b_first ? first_sl : 2 * bb;
		s += __ldcg(pieces + (size_t)sl * per_tile + v);
	}
	for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
	if (lane == 0) out[i * nout + c] = (float)s;
}

// Raw rows -> packed float4 records, padded with zero-strength records (pad_source) to n_pad, a
// multiple of kSrcTile.  One block packs one tile.  For filaments each block also leaves the
// statistics f3d_pick_mode() wants (sum of l^3, longest l, bounding box of the end points) in
// stats[blockIdx.x]; f3d_mode_kernel combines them in block order.
struct F3DStats { double sum_len3; float max_len; float lo[3], hi[3]; };

__global__ void __launch_bounds__(kSrcTile) pack_sources_kernel(int kind, int cols, const float *__restrict__ rows, int n, int n_pad,
                                                                float4 *__restrict__ A, float4 *__restrict__ Bq, float4 *__restrict__ Cq,
                                                                F3DStats *__restrict__ stats)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	float4 a, b, c;
	pad_source(kind, a, b, c);
	const bool real = i < n;
	if (real) {
		float row[7];
		for (int k = 0; k < cols; ++k) row[k] = rows[(size_t)i * cols + k];
		pack_source(kind, row, a, b, c);
	}
	if (i < n_pad) {
		A[i] = a;
		if (Bq) Bq[i] = b;
		if (Cq) Cq[i] = c;
	}
	if (stats) {
		__shared__ F3DStats part[kSrcTile / 32];
		const float len = real ? sqrtf(c.w) : 0.0f;
		double l3 = (double)len * len * len;
		float mx = len;
		float lo[3], hi[3];
		const float big = 3.0e38f;
		lo[0] = real ? fminf(a.x, b.x) : big; lo[1] = real ? fminf(a.y, b.y) : big; lo[2] = real ? fminf(a.z, b.z) : big;
		hi[0] = real ? fmaxf(a.x, b.x) : -big; hi[1] = real ? fmaxf(a.y, b.y) : -big; hi[2] = real ? fmaxf(a.z, b.z) : -big;
		for (int o = 16; o > 0; o >>= 1) {                      // fixed shuffle tree: the same sum on every run
			l3 += __shfl_down_sync(0xffffffffu, l3, o);
			mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
			for (int d = 0; d < 3; ++d) {
				lo[d] = fminf(lo[d], __shfl_down_sync(0xffffffffu, lo[d], o));
				hi[d] = fmaxf(hi[d], __shfl_down_sync(0xffffffffu, hi[d], o));
			}
		}
		if ((threadIdx.x & 31) == 0) {
			F3DStats &w = part[threadIdx.x >> 5];
			w.sum_len3 = l3; w.max_len = mx;
			for (int d = 0; d < 3; ++d) { w.lo[d] = lo[d]; w.hi[d] = hi[d]; }
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			F3DStats r = part[0];
			for (int k = 1; k < kSrcTile / 32; ++k) {
				r.sum_len3 += part[k].sum_len3; r.max_len = fmaxf(r.max_len, part[k].max_len);
				for (int d = 0; d < 3; ++d) { r.lo[d] = fminf(r.lo[d], part[k].lo[d]); r.hi[d] = fmaxf(r.hi[d], part[k].hi[d]); }
			}
			stats[blockIdx.x] = r;
		}
	}
}

// mode[0] = f3d_pick_mode over all filaments: one block, thread t takes blocks t, t + 256, ... in order,
// then a fixed tree -- the same sums on every run and every device.  `force` >= 0 pins the mode.
__device__ __forceinline__ void f3d_stats_merge(F3DStats &r, const F3DStats &o) {
	r.sum_len3 += o.sum_len3; r.max_len = fmaxf(r.max_len, o.max_len);
	for (int d = 0; d < 3; ++d) { r.lo[d] = fminf(r.lo[d], o.lo[d]); r.hi[d] = fmaxf(r.hi[d], o.hi[d]); }
}
__global__ void __launch_bounds__(256) f3d_mode_kernel(const F3DStats *__restrict__ stats, int n_blocks, int n, int force, int *__restrict__ mode)
{
	__shared__ F3DStats sh[256];
	F3DStats r;
	r.sum_len3 = 0.0; r.max_len = 0.0f;
	for (int d = 0; d < 3; ++d) { r.lo[d] = 3.0e38f; r.hi[d] = -3.0e38f; }
	for (int k = threadIdx.x; k < n_blocks; k += 256) f3d_stats_merge(r, stats[k]);
	sh[threadIdx.x] = r;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o) f3d_stats_merge(sh[threadIdx.x], sh[threadIdx.x + o]);
		__syncthreads();
	}
	if (threadIdx.x == 0) mode[0] = force >= 0 ? force : f3d_pick_mode(sh[0].sum_len3, (double)n, sh[0].lo, sh[0].hi, sh[0].max_len);
}

}  // namespace cvtx

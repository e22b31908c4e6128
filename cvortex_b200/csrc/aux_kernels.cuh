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

// ---- small calls: the host rows are READ by the device instead of being copied to it ------------------------------
// Two byte ranges (source rows, target rows) from the pinned staging area into device memory, 16 bytes per load, every
// byte read once.  A 10k x 10k call moves 400 KB: two copy-engine transfers cost it ~20 us each in engine start-up and
// the hand-over to the compute engine, a kernel reading through the mapped pointer starts like any other launch and
// runs at the link's rate.  Sizes are multiples of 4 (rows of floats), the buffers 256-byte aligned.
__global__ void __launch_bounds__(256) upload_rows_kernel(const void *__restrict__ a_src, void *__restrict__ a_dst, size_t a_bytes,
                                                          const void *__restrict__ b_src, void *__restrict__ b_dst, size_t b_bytes)
{
	// ONE 16-byte unit per thread: every read of the call is on the link at once (a loop of dependent load -> store
	// rounds pays the link's ~2.5 us round trip once per round)
	const size_t a16 = a_bytes / 16, b16 = b_bytes / 16;
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < a16) ((uint4 *)a_dst)[i] = __ldcs((const uint4 *)a_src + i);
	else if (i < a16 + b16) ((uint4 *)b_dst)[i - a16] = __ldcs((const uint4 *)b_src + (i - a16));
	else {
		const size_t k = i - a16 - b16;                                    // the last few threads: the ranges' tails, word by word
		const size_t ta = (a_bytes % 16) / 4, tb = (b_bytes % 16) / 4;
		if (k < ta) ((unsigned *)a_dst)[a16 * 4 + k] = ((const unsigned *)a_src)[a16 * 4 + k];
		else if (k < ta + tb) ((unsigned *)b_dst)[b16 * 4 + (k - ta)] = ((const unsigned *)b_src)[b16 * 4 + (k - ta)];
	}
}

// ---- box-cutoff ops on spatially coherent orders (m2m_kernel.cuh, sparse_tiles_kernel) ----------------------
// Bounding box of the REAL records of every packed source tile (NaN coordinates are ignored: such a source is outside
// every cutoff cube, as in the reference).  box[tile] = {lo x, y, z, hi x, y, z}; an empty tile gets an inverted box.
__global__ void __launch_bounds__(kSrcTile) source_tile_boxes_kernel(const float4 *__restrict__ A, int n_src, float *__restrict__ box)
{
	__shared__ float part[kSrcTile / 32][6];
	const int i = blockIdx.x * kSrcTile + threadIdx.x;
	const float big = 3.0e38f;
	float v[6] = {big, big, big, -big, -big, -big};
	if (i < n_src) {
		const float4 a = A[i];
		v[0] = v[3] = a.x; v[1] = v[4] = a.y; v[2] = v[5] = a.z;
		for (int d = 0; d < 3; ++d) if (!(v[d] == v[d])) { v[d] = big; v[3 + d] = -big; }
	}
	for (int o = 16; o > 0; o >>= 1)
		for (int d = 0; d < 3; ++d) {
			v[d] = fminf(v[d], __shfl_down_sync(0xffffffffu, v[d], o));
			v[3 + d] = fmaxf(v[3 + d], __shfl_down_sync(0xffffffffu, v[3 + d], o));
		}
	if ((threadIdx.x & 31) == 0) for (int d = 0; d < 6; ++d) part[threadIdx.x >> 5][d] = v[d];
	__syncthreads();
	if (threadIdx.x < 6) {
		float r = part[0][threadIdx.x];
		for (int k = 1; k < kSrcTile / 32; ++k) r = threadIdx.x < 3 ? fminf(r, part[k][threadIdx.x]) : fmaxf(r, part[k][threadIdx.x]);
		box[(size_t)blockIdx.x * 6 + threadIdx.x] = r;
	}
}

// One block per target tile of `slots` consecutive targets: its bounding box, then for every source tile whether the
// two boxes come closer than the cutoff on all three axes (in double: never drops a tile that holds a pair with
// |d| < cutoff); bit s of mask[tile][s / 32].  *active counts the marked pairs over all target tiles.
__global__ void __launch_bounds__(256) target_tile_masks_kernel(const float *__restrict__ tgt, int tcols, int n_tgt, int slots,
                                                                const float *__restrict__ src_box, int n_src_tiles, float cutoff,
                                                                unsigned *__restrict__ mask, int words, unsigned long long *__restrict__ active)
{
	__shared__ float part[8][6];
	__shared__ float tb[6];
	__shared__ unsigned s_count[8];
	const int tid = threadIdx.x;
	const long first = (long)blockIdx.x * slots;
	const float big = 3.0e38f;
	float v[6] = {big, big, big, -big, -big, -big};
	for (long i = first + tid; i < first + slots && i < n_tgt; i += 256)
		for (int d = 0; d < 3; ++d) {
			const float x = tgt[i * tcols + d];
			if (x == x) { v[d] = fminf(v[d], x); v[3 + d] = fmaxf(v[3 + d], x); }
		}
	for (int o = 16; o > 0; o >>= 1)
		for (int d = 0; d < 3; ++d) {
			v[d] = fminf(v[d], __shfl_down_sync(0xffffffffu, v[d], o));
			v[3 + d] = fmaxf(v[3 + d], __shfl_down_sync(0xffffffffu, v[3 + d], o));
		}
	if ((tid & 31) == 0) for (int d = 0; d < 6; ++d) part[tid >> 5][d] = v[d];
	__syncthreads();
	if (tid < 6) {
		float r = part[0][tid];
		for (int k = 1; k < 8; ++k) r = tid < 3 ? fminf(r, part[k][tid]) : fmaxf(r, part[k][tid]);
		tb[tid] = r;
	}
	__syncthreads();
	const double c = (double)cutoff * (1.0 + 1e-6);
	unsigned mine = 0;
	for (int w0 = 0; w0 < words * 32; w0 += 256) {
		const int s = w0 + tid;
		bool meet = false;
		if (s < n_src_tiles && cutoff > 0.0f) {
			const float *b = src_box + (size_t)s * 6;
			meet = true;
			for (int d = 0; d < 3; ++d)
				if ((double)b[d] - (double)tb[3 + d] > c || (double)tb[d] - (double)b[3 + d] > c) meet = false;
		}
		const unsigned bits = __ballot_sync(0xffffffffu, meet);
		if ((tid & 31) == 0 && (s >> 5) < words) { mask[(size_t)blockIdx.x * words + (s >> 5)] = bits; mine += __popc(bits); }
	}
	if ((tid & 31) == 0) s_count[tid >> 5] = mine;
	__syncthreads();
	if (tid == 0) {
		unsigned long long n = 0;
		for (int k = 0; k < 8; ++k) n += s_count[k];
		if (n) atomicAdd(active, n);
	}
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

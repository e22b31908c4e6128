// remesh_device.cu -- particle redistribution onto a regular grid, device stage.
//
// Replaces the tree build of the reference's cvtx_P3D_redistribute_on_grid /
// cvtx_P2D_redistribute_on_grid (src/P3D.cpp:552-589, src/P2D.cpp:325-362): there, every
// thread inserts its particles' (2R+1)^D shares into a private oct/quadtree one key at a
// time and the trees are merged serially (src/GridParticleOcttree.cpp:76-135,216-280).
// Here the particles are first put in cell order (cell = nearest node; the order shares are
// summed in, remesh_math.h), then one of two routes builds the nodes.  Dense route, when the
// grid is small and populated (the usual remeshing case): one thread per grid node, in Morton
// order, walks the particles of the (2R+1)^D cells around it -- 2R+1 contiguous runs per
// x-row -- and sums their shares; no share is ever written to memory.  Sort route, for sparse
// or very fine grids: the shares are produced per particle, sorted by the Morton code of
// their node and summed per node:
//
//   spread_count   particle -> number of non-zero shares            (28 B read / particle)
//   [scan]         offsets of each particle's run of shares         (CUB)
//   spread_emit    particle -> (Morton code, share) records, one warp per particle,
//                  coalesced                                   (K + 4 + 4 COMPS B / share)
//   [radix sort]   (code, record index) pairs, only the code bits the grid uses (CUB)
//   [run lengths]  distinct codes = nodes, records per node         (CUB)
//   [scan]         first record of each node                        (CUB)
//   node_sums      node -> FP64 sum of its shares in record order, rounded to FP32
//
// K, the width of a code, is 4 bytes when the grid fits (2^10 nodes per axis in 3-D, 2^16
// in 2-D -- a million particles at two per cell need 2^7), else 8.  The sort is stable and
// node_sums walks a node's records in order, so the result does not depend on scheduling:
// the same input gives the same bits on every run, and the same bits as the host stage of
// remesh.cpp.  The stage is HBM-bound; the bytes above are its algorithmic traffic
// (DESIGN.md section 9).
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>
#include <omp.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "remesh.h"
#include "runtime.h"

namespace cvtx {
namespace remesh {
namespace {

constexpr int kBlock = 256;

template <int D> struct Layout;
template <> struct Layout<3> { static constexpr int ROW = 7, COMPS = 3; };   // cvtx_P3D
template <> struct Layout<2> { static constexpr int ROW = 4, COMPS = 1; };   // cvtx_P2D

template <int D>
__global__ void __launch_bounds__(kBlock) spread_count(const float *__restrict__ rows, long n, Grid g, uint32_t *__restrict__ count) {
	const long i = (long)blockIdx.x * kBlock + threadIdx.x;
	if (i > n) return;
	if (i == n) { count[n] = 0; return; }          // so that the exclusive scan ends with the total
	float row[Layout<D>::ROW];
	for (int c = 0; c < Layout<D>::ROW; ++c) row[c] = rows[i * Layout<D>::ROW + c];
	count[i] = (uint32_t)spread_particle<D>(row, g, [](uint64_t, const float *) {});
}

// One warp per particle: lane l of round r owns stencil entry 32 r + l (the reference's
// order, x offset outermost), so a particle's records leave the warp as contiguous,
// coalesced runs; a ballot ranks the non-zero shares.  The D (2R+1) per-axis weights and
// per-axis Morton pieces are computed once, by the first lanes, and shuffled to the
// entries that use them.  R (the stencil half-width) is a template parameter so that the
// entry -> (ix, iy, iz) split is constant division.
template <int D, int R, class K>
__global__ void __launch_bounds__(kBlock) spread_emit(const float *__restrict__ rows, long n, Grid g, const uint32_t *__restrict__ offset,
                                                      K *__restrict__ code, uint32_t *__restrict__ record, float *__restrict__ share) {
	constexpr int ROW = Layout<D>::ROW, COMPS = Layout<D>::COMPS, S = 2 * R + 1, ENTRIES = D == 3 ? S * S * S : S * S;
	const long i = ((long)blockIdx.x * kBlock + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (i >= n) return;
	float row[ROW];
	for (int c = 0; c < ROW; ++c) row[c] = rows[i * ROW + c];
	float w_lane = 0.f;
	K m_lane = 0;
	if (lane < D * S) {
		const int a = lane / S, o = lane - a * S;
		const float x = a == 0 ? row[0] : (a == 1 ? row[1] : row[D - 1]);
		const uint32_t k = (D == 3 ? node_index_3d(x, g.origin[a], g.rh) : node_index_2d(x, g.origin[a], g.rh)) + (uint32_t)(o - R);
		w_lane = weight(g.kind, cell_distance(x, k, g, a));
		m_lane = (K)((D == 3 ? spread3(k) : spread2(k)) << a);
	}
	uint32_t at = offset[i];
#pragma unroll
	for (int base = 0; base < ENTRIES; base += 32) {
		const int e = base + lane;
		const bool valid = e < ENTRIES;
		const int ec = valid ? e : 0;
		const int ix = D == 3 ? ec / (S * S) : ec / S, iy = D == 3 ? (ec / S) % S : ec % S, iz = D == 3 ? ec % S : 0;
		float f = rm_mul(__shfl_sync(0xffffffffu, w_lane, ix), __shfl_sync(0xffffffffu, w_lane, S + iy));
		K m = __shfl_sync(0xffffffffu, m_lane, ix) | __shfl_sync(0xffffffffu, m_lane, S + iy);
		if (D == 3) {
			f = rm_mul(f, __shfl_sync(0xffffffffu, w_lane, 2 * S + iz));
			m |= __shfl_sync(0xffffffffu, m_lane, 2 * S + iz);
		}
		float s[COMPS];
		bool nz = false;
		for (int c = 0; c < COMPS; ++c) { s[c] = rm_mul(row[D + c], f); nz = nz || s[c] != 0.f; }
		nz = nz && valid;
		const uint32_t votes = __ballot_sync(0xffffffffu, nz);
		if (nz) {
			const uint32_t pos = at + __popc(votes & ((1u << lane) - 1u));
			code[pos] = m;
			record[pos] = pos;
			for (int c = 0; c < COMPS; ++c) share[(size_t)pos * COMPS + c] = s[c];
		}
		at += __popc(votes);
	}
}

template <int COMPS>
__global__ void __launch_bounds__(kBlock) node_sums(const uint32_t *__restrict__ first, const uint32_t *__restrict__ record,
                                                    const float *__restrict__ share, uint32_t n_nodes, float *__restrict__ strength) {
	const uint32_t k = blockIdx.x * kBlock + threadIdx.x;
	if (k >= n_nodes) return;
	double acc[COMPS];
	for (int c = 0; c < COMPS; ++c) acc[c] = 0.0;
	const uint32_t lo = first[k], hi = first[k + 1];
#pragma unroll 4
	for (uint32_t j = lo; j < hi; ++j) {
		const size_t r = record[j];
		for (int c = 0; c < COMPS; ++c) acc[c] += (double)share[r * COMPS + c];
	}
	for (int c = 0; c < COMPS; ++c) strength[(size_t)k * COMPS + c] = (float)acc[c];
}

// ---- the summation order: particles by cell, then by the caller's index --------------------
template <int D>
__global__ void __launch_bounds__(kBlock) cell_keys(const float *__restrict__ rows, long n, Grid g, uint64_t *__restrict__ key, uint32_t *__restrict__ at) {
	const long i = (long)blockIdx.x * kBlock + threadIdx.x;
	if (i >= n) return;
	float row[D];
	for (int a = 0; a < D; ++a) row[a] = rows[i * Layout<D>::ROW + a];
	key[i] = cell_of<D>(row, g);
	at[i] = (uint32_t)i;
}

template <int D>
__global__ void __launch_bounds__(kBlock) reorder_rows(const float *__restrict__ rows, const uint32_t *__restrict__ at, long n, float *__restrict__ out) {
	const long j = (long)blockIdx.x * kBlock + threadIdx.x;
	if (j >= n) return;
	const size_t from = (size_t)at[j] * Layout<D>::ROW;
	for (int c = 0; c < Layout<D>::ROW; ++c) out[(size_t)j * Layout<D>::ROW + c] = rows[from + c];
}

// ---- dense route: one thread per grid node gathers its shares --------------------------------
__global__ void __launch_bounds__(kBlock) count_cells(const uint64_t *__restrict__ key, long n, uint32_t *__restrict__ count) {
	const long i = (long)blockIdx.x * kBlock + threadIdx.x;
	if (i < n) atomicAdd(&count[key[i]], 1u);
}

// Thread t owns the node whose Morton code is t.  The particles that can reach it are those
// of the (2R+1)^D cells around it; with cells numbered x fastest and particles sorted by
// cell, each row of 2R+1 cells along x is ONE contiguous run of particles, walked in order --
// the same summation order as the sort route and the host stage.  A node exists when at
// least one share is non-zero (the reference does not insert all-zero shares).
template <int D, int KIND>
__global__ void __launch_bounds__(kBlock) gather_nodes(const float4 *__restrict__ rows /* 2 float4 per particle */, const uint32_t *__restrict__ cell_start,
                                                      Grid g, uint32_t domain, uint32_t *__restrict__ exists, float *__restrict__ dense) {
	constexpr int COMPS = Layout<D>::COMPS, R = kHalfWidth[KIND];
	const uint32_t t = blockIdx.x * kBlock + threadIdx.x;
	if (t > domain) return;
	if (t == domain) { exists[t] = 0u; return; }          // so that the exclusive scan ends with the count
	uint32_t node[3] = {0, 0, 0};
	float at[3] = {0.f, 0.f, 0.f};                         // the node's coordinates
	bool inside = true;
	for (int a = 0; a < D; ++a) {
		node[a] = (uint32_t)(D == 3 ? compact3((uint64_t)t >> a) : compact2((uint64_t)t >> a));
		at[a] = node_coord(node[a], g.origin[a], g.h);
		inside = inside && node[a] <= g.top[a];
	}
	if (!inside) { exists[t] = 0u; return; }
	const uint64_t nx = (uint64_t)g.top[0] + 1, ny = (uint64_t)g.top[1] + 1;
	const uint32_t x0 = node[0] >= (uint32_t)R ? node[0] - R : 0u, x1 = node[0] + R <= g.top[0] ? node[0] + R : g.top[0];
	const uint32_t y0 = node[1] >= (uint32_t)R ? node[1] - R : 0u, y1 = node[1] + R <= g.top[1] ? node[1] + R : g.top[1];
	const uint32_t z0 = D == 3 ? (node[2] >= (uint32_t)R ? node[2] - R : 0u) : 0u, z1 = D == 3 ? (node[2] + R <= g.top[2] ? node[2] + R : g.top[2]) : 0u;
	double acc[COMPS];
	for (int c = 0; c < COMPS; ++c) acc[c] = 0.0;
	bool any = false;
	for (uint32_t cz = z0; cz <= z1; ++cz)
		for (uint32_t cy = y0; cy <= y1; ++cy) {
			const uint64_t line = nx * (cy + ny * cz);
			const uint32_t lo = cell_start[line + x0], hi = cell_start[line + x1 + 1];
			for (uint32_t j = lo; j < hi; ++j) {
				const float4 p = rows[2 * (size_t)j], q = rows[2 * (size_t)j + 1];      // x y z w0 | w1 w2 . .   (2-D: x y gamma .)
				// same operations as cell_distance() / spread_particle(), node coordinate hoisted
				float f = rm_mul(weight(KIND, fabsf(rm_mul(rm_sub(p.x, at[0]), g.rh))), weight(KIND, fabsf(rm_mul(rm_sub(p.y, at[1]), g.rh))));
				float s[3] = {0.f, 0.f, 0.f};
				if (D == 3) {
					f = rm_mul(f, weight(KIND, fabsf(rm_mul(rm_sub(p.z, at[2]), g.rh))));
					s[0] = rm_mul(p.w, f);
					s[1] = rm_mul(q.x, f);
					s[2] = rm_mul(q.y, f);
				} else {
					s[0] = rm_mul(p.z, f);
				}
				if (s[0] != 0.f || s[1] != 0.f || s[2] != 0.f) {
					any = true;
					for (int c = 0; c < COMPS; ++c) acc[c] += (double)s[c];
				}
			}
		}
	exists[t] = any ? 1u : 0u;
	if (any) for (int c = 0; c < COMPS; ++c) dense[(size_t)t * COMPS + c] = (float)acc[c];
}

// Sorted rows padded to two float4 (the dense route reads them with two 16-byte loads).
template <int D>
__global__ void __launch_bounds__(kBlock) reorder_rows_padded(const float *__restrict__ rows, const uint32_t *__restrict__ at, long n, float4 *__restrict__ out) {
	const long j = (long)blockIdx.x * kBlock + threadIdx.x;
	if (j >= n) return;
	const float *r = rows + (size_t)at[j] * Layout<D>::ROW;
	if (D == 3) {
		out[2 * j] = make_float4(r[0], r[1], r[2], r[3]);
		out[2 * j + 1] = make_float4(r[4], r[5], r[6], 0.f);
	} else {
		out[2 * j] = make_float4(r[0], r[1], r[2], r[3]);
		out[2 * j + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
	}
}

template <int D, int KIND>
void launch_gather(const float4 *rows, const uint32_t *cell_start, const Grid &g, uint32_t domain, uint32_t *exists, float *dense, cudaStream_t st) {
	gather_nodes<D, KIND><<<(unsigned)(((size_t)domain + 1 + kBlock - 1) / kBlock), kBlock, 0, st>>>(rows, cell_start, g, domain, exists, dense);
}

template <int COMPS, class K>
__global__ void __launch_bounds__(kBlock) collect_nodes(const uint32_t *__restrict__ exists, const uint32_t *__restrict__ position, const float *__restrict__ dense,
                                                       uint32_t domain, K *__restrict__ node_code, float *__restrict__ sums) {
	const uint32_t t = blockIdx.x * kBlock + threadIdx.x;
	if (t >= domain || !exists[t]) return;
	const uint32_t at = position[t];
	node_code[at] = (K)t;
	for (int c = 0; c < COMPS; ++c) sums[(size_t)at * COMPS + c] = dense[(size_t)t * COMPS + c];
}

inline unsigned blocks_for(size_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

// Pinned staging -> the caller's array, in parallel 1 MiB pieces.
void fetch_bytes(void *dst, const void *src, size_t bytes) {
	const size_t piece = 1 << 20;
	const long pieces = (long)((bytes + piece - 1) / piece);
#pragma omp parallel for schedule(static) num_threads(4) if (pieces > 4)
	for (long i = 0; i < pieces; ++i) {
		const size_t lo = (size_t)i * piece;
		std::memcpy((char *)dst + lo, (const char *)src + lo, lo + piece <= bytes ? piece : bytes - lo);
	}
}

enum { ROWS, COUNT, OFFSET, CODE_A, CODE_B, REC_A, REC_B, SHARE, TEMP, FIRST, SUMS,        // node build
       PARTIAL, STRENGTH, KEEP, POSITION, CODE_C, SUMS_C, HISTOGRAM,                       // device-resident pruning
       CELL_A, CELL_B, ORDER_A, ORDER_B, SORTED_ROWS, CELL_START, EXISTS, DENSE_POS, DENSE,     // cell order, dense route
       RELAX_POINTS, RELAX_FIELD, OUT_ROWS };                                              // relaxation, host-array result

// Node build on rows that are already on the device.  On return (stream synchronised)
// node_code[0..n_nodes) and sums[0..n_nodes * COMPS) hold the nodes in ascending code order.
template <int D, class K>
int build_nodes(Device *d, cudaStream_t st, const float *rows_in, long n, const Grid &g, int bits, K **node_code_out, float **sums_out, uint32_t *n_nodes_out) {
	constexpr int COMPS = Layout<D>::COMPS;
	Buffer *b = d->remesh;
	*n_nodes_out = 0;
	size_t temp = 0;

	// particles by cell, then by the caller's index: the order shares are summed in
	double cells = 1.0;
	for (int a = 0; a < D; ++a) cells *= (double)g.top[a] + 1.0;
	const bool dense = dense_route(D, g, bits, n);
	int cell_bits = 1;
	while (cell_bits < 64 && std::ldexp(1.0, cell_bits) < cells) ++cell_bits;
	CUDA_TRY(b[CELL_A].reserve(sizeof(uint64_t) * (size_t)n));
	CUDA_TRY(b[CELL_B].reserve(sizeof(uint64_t) * (size_t)n));
	CUDA_TRY(b[ORDER_A].reserve(sizeof(uint32_t) * (size_t)n));
	CUDA_TRY(b[ORDER_B].reserve(sizeof(uint32_t) * (size_t)n));
	CUDA_TRY(b[SORTED_ROWS].reserve(sizeof(float) * 8 * (size_t)n));
	uint64_t *cell_a = (uint64_t *)b[CELL_A].p, *cell_b = (uint64_t *)b[CELL_B].p;
	uint32_t *order_a = (uint32_t *)b[ORDER_A].p, *order_b = (uint32_t *)b[ORDER_B].p;
	float *rows = (float *)b[SORTED_ROWS].p;
	cell_keys<D><<<blocks_for((size_t)n), kBlock, 0, st>>>(rows_in, n, g, cell_a, order_a);
	CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp, cell_a, cell_b, order_a, order_b, (int)n, 0, cell_bits, st));
	CUDA_TRY(b[TEMP].reserve(temp));
	CUDA_TRY(cub::DeviceRadixSort::SortPairs(b[TEMP].p, temp, cell_a, cell_b, order_a, order_b, (int)n, 0, cell_bits, st));
	if (dense) reorder_rows_padded<D><<<blocks_for((size_t)n), kBlock, 0, st>>>(rows_in, order_b, n, (float4 *)rows);
	else reorder_rows<D><<<blocks_for((size_t)n), kBlock, 0, st>>>(rows_in, order_b, n, rows);
	count_launches(2);

	if (dense) {
		const uint32_t n_cells = (uint32_t)cells, domain = 1u << bits;
		CUDA_TRY(b[CELL_START].reserve(sizeof(uint32_t) * ((size_t)n_cells + 2)));
		CUDA_TRY(b[EXISTS].reserve(sizeof(uint32_t) * ((size_t)domain + 1)));
		CUDA_TRY(b[DENSE_POS].reserve(sizeof(uint32_t) * ((size_t)domain + 1)));
		CUDA_TRY(b[DENSE].reserve(sizeof(float) * COMPS * (size_t)domain));
		uint32_t *cell_start = (uint32_t *)b[CELL_START].p, *exists = (uint32_t *)b[EXISTS].p, *position = (uint32_t *)b[DENSE_POS].p;
		float *dense_sums = (float *)b[DENSE].p;
		// particles per cell -> first particle of each cell (count[] is scanned in place, one slot more than cells)
		CUDA_TRY(cudaMemsetAsync(cell_start, 0, sizeof(uint32_t) * ((size_t)n_cells + 1), st));
		count_cells<<<blocks_for((size_t)n), kBlock, 0, st>>>(cell_b, n, cell_start);
		CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, temp, cell_start, cell_start, (int)n_cells + 1, st));
		CUDA_TRY(b[TEMP].reserve(temp));
		CUDA_TRY(cub::DeviceScan::ExclusiveSum(b[TEMP].p, temp, cell_start, cell_start, (int)n_cells + 1, st));
		const float4 *padded = (const float4 *)rows;
		switch (g.kind) {
		case K_LAMBDA0: launch_gather<D, K_LAMBDA0>(padded, cell_start, g, domain, exists, dense_sums, st); break;
		case K_LAMBDA1: launch_gather<D, K_LAMBDA1>(padded, cell_start, g, domain, exists, dense_sums, st); break;
		case K_LAMBDA2: launch_gather<D, K_LAMBDA2>(padded, cell_start, g, domain, exists, dense_sums, st); break;
		case K_LAMBDA3: launch_gather<D, K_LAMBDA3>(padded, cell_start, g, domain, exists, dense_sums, st); break;
		default: launch_gather<D, K_M4P>(padded, cell_start, g, domain, exists, dense_sums, st); break;
		}
		CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, temp, exists, position, (int)domain + 1, st));
		CUDA_TRY(b[TEMP].reserve(temp));
		CUDA_TRY(cub::DeviceScan::ExclusiveSum(b[TEMP].p, temp, exists, position, (int)domain + 1, st));
		uint32_t n_nodes = 0;
		CUDA_TRY(cudaMemcpyAsync(&n_nodes, position + domain, sizeof(n_nodes), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		count_launches(2);
		if (n_nodes == 0) return CVTX_B200_OK;
		CUDA_TRY(b[CODE_A].reserve(sizeof(K) * (size_t)n_nodes));
		CUDA_TRY(b[SUMS].reserve(sizeof(float) * COMPS * (size_t)n_nodes));
		collect_nodes<COMPS, K><<<blocks_for(domain), kBlock, 0, st>>>(exists, position, dense_sums, domain, (K *)b[CODE_A].p, (float *)b[SUMS].p);
		count_launches(1);
		CUDA_TRY(cudaGetLastError());
		CUDA_TRY(cudaStreamSynchronize(st));
		*node_code_out = (K *)b[CODE_A].p;
		*sums_out = (float *)b[SUMS].p;
		*n_nodes_out = n_nodes;
		return CVTX_B200_OK;
	}

	// sort route: every share becomes a record
	CUDA_TRY(b[COUNT].reserve(sizeof(uint32_t) * (size_t)(n + 1)));
	CUDA_TRY(b[OFFSET].reserve(sizeof(uint32_t) * (size_t)(n + 1)));
	uint32_t *count = (uint32_t *)b[COUNT].p, *offset = (uint32_t *)b[OFFSET].p;

	spread_count<D><<<blocks_for((size_t)n + 1), kBlock, 0, st>>>(rows, n, g, count);
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, temp, count, offset, (int)(n + 1), st));
	CUDA_TRY(b[TEMP].reserve(temp));
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(b[TEMP].p, temp, count, offset, (int)(n + 1), st));
	uint32_t total = 0;
	CUDA_TRY(cudaMemcpyAsync(&total, offset + n, sizeof(total), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	count_launches(1);
	if (total == 0) return CVTX_B200_OK;

	CUDA_TRY(b[CODE_A].reserve(sizeof(K) * (size_t)total));
	CUDA_TRY(b[CODE_B].reserve(sizeof(K) * (size_t)total));
	CUDA_TRY(b[REC_A].reserve(sizeof(uint32_t) * ((size_t)total + 1)));
	CUDA_TRY(b[REC_B].reserve(sizeof(uint32_t) * (size_t)total));
	CUDA_TRY(b[SHARE].reserve(sizeof(float) * COMPS * (size_t)total));
	K *code_a = (K *)b[CODE_A].p, *code_b = (K *)b[CODE_B].p;
	uint32_t *rec_a = (uint32_t *)b[REC_A].p, *rec_b = (uint32_t *)b[REC_B].p;
	float *share = (float *)b[SHARE].p;
	if (g.half == 1) spread_emit<D, 1, K><<<blocks_for((size_t)n * 32), kBlock, 0, st>>>(rows, n, g, offset, code_a, rec_a, share);
	else spread_emit<D, 2, K><<<blocks_for((size_t)n * 32), kBlock, 0, st>>>(rows, n, g, offset, code_a, rec_a, share);

	CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp, code_a, code_b, rec_a, rec_b, (int)total, 0, bits, st));
	CUDA_TRY(b[TEMP].reserve(temp));
	CUDA_TRY(cub::DeviceRadixSort::SortPairs(b[TEMP].p, temp, code_a, code_b, rec_a, rec_b, (int)total, 0, bits, st));

	// code_a / rec_a are free again: they take the distinct codes and their run lengths
	K *node_code = code_a;
	uint32_t *run_length = rec_a, *n_runs = count;       // count[] is no longer needed either
	CUDA_TRY(cub::DeviceRunLengthEncode::Encode(nullptr, temp, code_b, node_code, run_length, n_runs, (int)total, st));
	CUDA_TRY(b[TEMP].reserve(temp));
	CUDA_TRY(cub::DeviceRunLengthEncode::Encode(b[TEMP].p, temp, code_b, node_code, run_length, n_runs, (int)total, st));
	uint32_t n_nodes = 0;
	CUDA_TRY(cudaMemcpyAsync(&n_nodes, n_runs, sizeof(n_nodes), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));

	CUDA_TRY(b[FIRST].reserve(sizeof(uint32_t) * ((size_t)n_nodes + 1)));
	CUDA_TRY(b[SUMS].reserve(sizeof(float) * COMPS * (size_t)n_nodes));
	uint32_t *first = (uint32_t *)b[FIRST].p;
	float *sums = (float *)b[SUMS].p;
	CUDA_TRY(cudaMemsetAsync(run_length + n_nodes, 0, sizeof(uint32_t), st));   // so that the scan ends with `total`
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, temp, run_length, first, (int)n_nodes + 1, st));
	CUDA_TRY(b[TEMP].reserve(temp));
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(b[TEMP].p, temp, run_length, first, (int)n_nodes + 1, st));
	node_sums<COMPS><<<blocks_for(n_nodes), kBlock, 0, st>>>(first, rec_b, share, n_nodes, sums);
	count_launches(2);
	CUDA_TRY(cudaGetLastError());
	*node_code_out = node_code;
	*sums_out = sums;
	*n_nodes_out = n_nodes;
	return CVTX_B200_OK;
}

// ---- device-resident redistribution: bounds and pruning on the device ---------------------
// Block-level partial results are combined on the host in block order, so a result never
// depends on scheduling.  FP64 sums are formed in a different (tree) order than the host
// stage's, which can only matter when a rounded mean falls within an ulp of a threshold.

struct Bounds { float lo[3], hi[3]; double sum[3]; };
constexpr int kRowsPerBlock = kBlock;      // one row per thread: the canonical sum order of remesh.h

template <class T, class Op>
__device__ T block_reduce(T v, Op op, T *scratch /* kBlock / 32 */) {
	for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_down_sync(0xffffffffu, v, o));
	if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
	__syncthreads();
	T r = scratch[0];
	if (threadIdx.x == 0) for (int w = 1; w < kBlock / 32; ++w) r = op(r, scratch[w]);
	__syncthreads();
	return r;      // valid in thread 0
}

template <int D>
__global__ void __launch_bounds__(kBlock) bounds_kernel(const float *__restrict__ rows, long n, Bounds *__restrict__ partial) {
	__shared__ double scratch_d[kBlock / 32];
	__shared__ float scratch_f[kBlock / 32];
	const long base = (long)blockIdx.x * kRowsPerBlock;
	Bounds q;
	for (int a = 0; a < 3; ++a) { q.lo[a] = 3.4e38f; q.hi[a] = -3.4e38f; q.sum[a] = 0.0; }
	for (long i = base + threadIdx.x; i < base + kRowsPerBlock && i < n; i += kBlock)
		for (int a = 0; a < D; ++a) {
			const float x = rows[i * Layout<D>::ROW + a];
			q.lo[a] = fminf(q.lo[a], x);
			q.hi[a] = fmaxf(q.hi[a], x);
			q.sum[a] += (double)x;
		}
	for (int a = 0; a < D; ++a) {
		const float lo = block_reduce(q.lo[a], [](float x, float y) { return fminf(x, y); }, scratch_f);
		const float hi = block_reduce(q.hi[a], [](float x, float y) { return fmaxf(x, y); }, scratch_f);
		const double sum = block_reduce(q.sum[a], [](double x, double y) { return x + y; }, scratch_d);
		if (threadIdx.x == 0) { partial[blockIdx.x].lo[a] = lo; partial[blockIdx.x].hi[a] = hi; partial[blockIdx.x].sum[a] = sum; }
	}
}

// |w| per node in the host stage's FP32 order; per-block FP64 sums of them, min and max.
struct StrengthPart { double sum; float lo, hi; };
template <int COMPS>
__global__ void __launch_bounds__(kBlock) strengths_kernel(const float *__restrict__ w, uint32_t n, float *__restrict__ strength, StrengthPart *__restrict__ partial) {
	__shared__ double scratch_d[kBlock / 32];
	__shared__ float scratch_f[kBlock / 32];
	const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
	float m = 0.f;
	if (i < n) {
		if (COMPS == 1) m = fabsf(w[i]);
		else m = __fsqrt_rn(rm_add(rm_add(rm_mul(w[(size_t)i * 3], w[(size_t)i * 3]), rm_mul(w[(size_t)i * 3 + 1], w[(size_t)i * 3 + 1])),
		                          rm_mul(w[(size_t)i * 3 + 2], w[(size_t)i * 3 + 2])));
		strength[i] = m;
	}
	const double sum = block_reduce((double)m, [](double x, double y) { return x + y; }, scratch_d);
	const float lo = block_reduce(i < n ? m : 3.4e38f, [](float x, float y) { return fminf(x, y); }, scratch_f);
	const float hi = block_reduce(i < n ? m : -3.4e38f, [](float x, float y) { return fmaxf(x, y); }, scratch_f);
	if (threadIdx.x == 0) { partial[blockIdx.x].sum = sum; partial[blockIdx.x].lo = lo; partial[blockIdx.x].hi = hi; }
}

// keep[i] = strength[i] > cut && i < index_limit; per-block FP64 sums of the dropped vorticity.
template <int COMPS>
__global__ void __launch_bounds__(kBlock) keep_kernel(const float *__restrict__ strength, const float *__restrict__ w, uint32_t n, float cut,
                                                      uint32_t index_limit, uint32_t *__restrict__ keep, double *__restrict__ lost /* [blocks][COMPS] */) {
	__shared__ double scratch_d[kBlock / 32];
	const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
	const bool in = i < n, stays = in && strength[i] > cut && i < index_limit;
	if (in) keep[i] = stays ? 1u : 0u;
	if (i == n) keep[n] = 0u;
	for (int c = 0; c < COMPS; ++c) {
		const double s = block_reduce(in && !stays ? (double)w[(size_t)i * COMPS + c] : 0.0, [](double x, double y) { return x + y; }, scratch_d);
		if (threadIdx.x == 0) lost[(size_t)blockIdx.x * COMPS + c] = s;
	}
}

struct Each { float v[3]; };

// Survivors, with their share of the dropped vorticity, to new node arrays (second-cut route).
template <int COMPS, class K>
__global__ void __launch_bounds__(kBlock) compact_kernel(const K *__restrict__ code, const float *__restrict__ w, const uint32_t *__restrict__ keep,
                                                         const uint32_t *__restrict__ position, uint32_t n, Each each, K *__restrict__ code_out,
                                                         float *__restrict__ w_out) {
	const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
	if (i >= n || !keep[i]) return;
	const uint32_t at = position[i];
	code_out[at] = code[i];
	for (int c = 0; c < COMPS; ++c) w_out[(size_t)at * COMPS + c] = rm_add(w[(size_t)i * COMPS + c], each.v[c]);
}

// Survivors as cvtx_P3D / cvtx_P2D rows.
template <int D, class K>
__global__ void __launch_bounds__(kBlock) write_rows_kernel(const K *__restrict__ code, const float *__restrict__ w, const uint32_t *__restrict__ keep,
                                                            const uint32_t *__restrict__ position, uint32_t n, Each each, Grid g, float size,
                                                            float *__restrict__ out) {
	constexpr int ROW = Layout<D>::ROW, COMPS = Layout<D>::COMPS;
	const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
	if (i >= n || !keep[i]) return;
	float *p = out + (size_t)position[i] * ROW;
	const uint64_t m = code[i];
	if (D == 3) {
		p[0] = node_coord((uint32_t)compact3(m), g.origin[0], g.h);
		p[1] = node_coord((uint32_t)compact3(m >> 1), g.origin[1], g.h);
		p[2] = node_coord((uint32_t)compact3(m >> 2), g.origin[2], g.h);
	} else {
		p[0] = node_coord((uint32_t)compact2(m), g.origin[0], g.h);
		p[1] = node_coord((uint32_t)compact2(m >> 1), g.origin[1], g.h);
	}
	for (int c = 0; c < COMPS; ++c) p[D + c] = rm_add(w[(size_t)i * COMPS + c], each.v[c]);
	p[D + COMPS] = size;
}

__global__ void __launch_bounds__(kBlock) histogram_kernel(const float *__restrict__ strength, uint32_t n, double lo, double range, int *__restrict__ count) {
	__shared__ int local[kCutBins];
	for (int b = threadIdx.x; b < kCutBins; b += kBlock) local[b] = 0;
	__syncthreads();
	for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
		const double t = floor(__ddiv_rn(__dmul_rn((double)(kCutBins - 1), __dsub_rn((double)strength[i], lo)), range));
		atomicAdd(&local[cut_bin(t)], 1);
	}
	__syncthreads();
	for (int b = threadIdx.x; b < kCutBins; b += kBlock) if (local[b]) atomicAdd(&count[b], local[b]);
}

// One cut on the device: strengths (optionally), keep flags, positions.  Fills each[] and kept.
template <int COMPS>
int cut_on_device(Device *d, cudaStream_t st, const float *w, uint32_t n, float cut, uint32_t index_limit, Each *each, uint32_t *kept) {
	Buffer *b = d->remesh;
	const unsigned blocks = blocks_for(n + 1);
	CUDA_TRY(b[KEEP].reserve(sizeof(uint32_t) * ((size_t)n + 1)));
	CUDA_TRY(b[POSITION].reserve(sizeof(uint32_t) * ((size_t)n + 1)));
	CUDA_TRY(b[PARTIAL].reserve(sizeof(double) * 3 * (size_t)blocks + sizeof(StrengthPart) * blocks + sizeof(Bounds) * blocks));
	uint32_t *keep = (uint32_t *)b[KEEP].p, *position = (uint32_t *)b[POSITION].p;
	double *lost = (double *)b[PARTIAL].p;
	keep_kernel<COMPS><<<blocks, kBlock, 0, st>>>((const float *)b[STRENGTH].p, w, n, cut, index_limit, keep, lost);
	size_t temp = 0;
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, temp, keep, position, (int)n + 1, st));
	CUDA_TRY(b[TEMP].reserve(temp));
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(b[TEMP].p, temp, keep, position, (int)n + 1, st));
	std::vector<double> host_lost((size_t)blocks * COMPS);
	CUDA_TRY(cudaMemcpyAsync(host_lost.data(), lost, sizeof(double) * host_lost.size(), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaMemcpyAsync(kept, position + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	count_launches(1);
	for (int c = 0; c < 3; ++c) each->v[c] = 0.f;
	for (int c = 0; c < COMPS; ++c) {
		double total = 0.0;
		for (unsigned k = 0; k < blocks; ++k) total += host_lost[(size_t)k * COMPS + c];
		each->v[c] = (float)total / (float)*kept;
	}
	return CVTX_B200_OK;
}

template <int COMPS>
int strengths_on_device(Device *d, cudaStream_t st, const float *w, uint32_t n, double *total, float *lo, float *hi) {
	Buffer *b = d->remesh;
	const unsigned blocks = blocks_for(n);
	CUDA_TRY(b[STRENGTH].reserve(sizeof(float) * (size_t)n));
	CUDA_TRY(b[PARTIAL].reserve(sizeof(double) * 3 * ((size_t)blocks + 1) + sizeof(StrengthPart) * blocks + sizeof(Bounds) * blocks));
	StrengthPart *part = (StrengthPart *)b[PARTIAL].p;
	strengths_kernel<COMPS><<<blocks, kBlock, 0, st>>>(w, n, (float *)b[STRENGTH].p, part);
	std::vector<StrengthPart> host((size_t)blocks);
	CUDA_TRY(cudaMemcpyAsync(host.data(), part, sizeof(StrengthPart) * blocks, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	count_launches(1);
	*total = 0.0;
	*lo = host[0].lo;
	*hi = host[0].hi;
	for (const StrengthPart &p : host) { *total += p.sum; *lo = fminf(*lo, p.lo); *hi = fmaxf(*hi, p.hi); }
	return CVTX_B200_OK;
}

// out: the caller's device array (capacity max_out rows), or -- own_out set -- an internal
// buffer sized to the result, returned through *own_out (the host-array entry points, which
// copy it back).  want_rows = false asks for the count only.  *nodes_out: nodes before pruning.
template <int D, class K>
int redistribute_resident(Device *d, cudaStream_t st, const float *rows, long n, const Grid &g, int bits, float negligible, float *out,
                          float **own_out, bool want_rows, int max_out, int *n_out, size_t *nodes_out) {
	constexpr int ROW = Layout<D>::ROW, COMPS = Layout<D>::COMPS;
	Buffer *b = d->remesh;
	K *code = nullptr;
	float *w = nullptr;
	uint32_t n_nodes = 0;
	*n_out = 0;
	if (int rc = build_nodes<D, K>(d, st, rows, n, g, bits, &code, &w, &n_nodes)) return rc;
	if (nodes_out) *nodes_out = n_nodes;
	if (n_nodes == 0) return CVTX_B200_OK;

	double total = 0.0;
	float lo = 0.f, hi = 0.f;
	if (int rc = strengths_on_device<COMPS>(d, st, w, n_nodes, &total, &lo, &hi)) return rc;
	Each each;
	uint32_t kept = 0;
	if (int rc = cut_on_device<COMPS>(d, st, w, n_nodes, (float)(total / (double)n_nodes) * negligible, n_nodes, &each, &kept)) return rc;
	*n_out = (int)kept;
	if (!want_rows || kept == 0) return CVTX_B200_OK;

	if ((long)kept > (long)max_out) {
		// too many for the caller's array: materialise the survivors, find the strength that fits
		// (same search as the host stage, histograms on the device), cut again
		CUDA_TRY(b[CODE_C].reserve(sizeof(K) * (size_t)kept));
		CUDA_TRY(b[SUMS_C].reserve(sizeof(float) * COMPS * (size_t)kept));
		K *code_c = (K *)b[CODE_C].p;
		float *w_c = (float *)b[SUMS_C].p;
		compact_kernel<COMPS, K><<<blocks_for(n_nodes), kBlock, 0, st>>>(code, w, (const uint32_t *)b[KEEP].p, (const uint32_t *)b[POSITION].p,
		                                                                 n_nodes, each, code_c, w_c);
		count_launches(1);
		code = code_c;
		w = w_c;
		n_nodes = kept;
		if (int rc = strengths_on_device<COMPS>(d, st, w, n_nodes, &total, &lo, &hi)) return rc;
		CUDA_TRY(b[HISTOGRAM].reserve(sizeof(int) * kCutBins));
		int *count_dev = (int *)b[HISTOGRAM].p;
		int failed = CVTX_B200_OK;
		const float cut2 = strength_cut_with(
		    (int)n_nodes, max_out, [&](float *fmin, float *fmax) { *fmin = lo; *fmax = hi; },
		    [&](double from, double range, int *count) {
			    cudaError_t e = cudaMemsetAsync(count_dev, 0, sizeof(int) * kCutBins, st);
			    histogram_kernel<<<296, kBlock, 0, st>>>((const float *)b[STRENGTH].p, n_nodes, from, range, count_dev);
			    if (e == cudaSuccess) e = cudaMemcpyAsync(count, count_dev, sizeof(int) * kCutBins, cudaMemcpyDeviceToHost, st);
			    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
			    count_launches(1);
			    if (e != cudaSuccess) {
				    failed = fail(CVTX_B200_ERR_CUDA, std::string("strength histogram: ") + cudaGetErrorString(e));
				    for (int i = 0; i < kCutBins; ++i) count[i] = i == kCutBins - 1 ? (int)n_nodes : 0;   // ends the search
			    }
		    });
		if (failed) return failed;
		if (int rc = cut_on_device<COMPS>(d, st, w, n_nodes, cut2, (uint32_t)max_out, &each, &kept)) return rc;
		*n_out = (int)kept;
		if (kept == 0) return CVTX_B200_OK;
	}
	if (own_out) {
		CUDA_TRY(b[OUT_ROWS].reserve(sizeof(float) * ROW * (size_t)kept));
		out = *own_out = (float *)b[OUT_ROWS].p;
	}
	const float size = D == 3 ? g.h * g.h * g.h : g.h * g.h;
	write_rows_kernel<D, K><<<blocks_for(n_nodes), kBlock, 0, st>>>(code, w, (const uint32_t *)b[KEEP].p, (const uint32_t *)b[POSITION].p, n_nodes,
	                                                                each, g, size, out);
	count_launches(1);
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaStreamSynchronize(st));
	return CVTX_B200_OK;
}

}  // namespace

int device_redistribute_from_host(int device, int dim, int kind, float h, const void *const *particles, long n, float negligible,
                                  void *out_rows, int max_out, int *n_out, size_t *n_nodes) {
	if (dim != 2 && dim != 3) return fail(CVTX_B200_ERR_ARGUMENT, "dimension must be 2 or 3");
	if (kind < 0 || kind >= K_COUNT) return fail(CVTX_B200_ERR_ARGUMENT, "unknown redistribution function");
	if (n <= 0 || n > 0x7ffffff0L || !n_out) return fail(CVTX_B200_ERR_ARGUMENT, "bad particle count");
	*n_out = 0;
	if (n_nodes) *n_nodes = 0;
	cudaStream_t st = nullptr;
	if (int rc = device_stream(device, &st)) return rc;
	Device *d = get_device(device);
	const int row_floats = dim == 3 ? 7 : 4;
	static const bool trace = [] { const char *e = std::getenv("CVTX_B200_TRACE"); return e && e[0] == '1'; }();

	// gather the particles into the pinned staging area and place the grid, in one pass
	HostStage &hs = host_stage();
	std::lock_guard<std::mutex> stage_lock(hs.mu);
	CUDA_TRY(hs.src.reserve(sizeof(float) * row_floats * (size_t)n));
	const double t0 = omp_get_wtime();
	uint32_t max_index = 0;
	const Grid g = place_grid(dim, kind, kHalfWidth[kind], h, particles, (float *)hs.src.p, n, row_floats, &max_index);
	const int bits = code_bits(dim, max_index);
	if (bits < 0) return fail(CVTX_B200_ERR_ARGUMENT, "grid too large for the node codes (more than 2^21 nodes per axis in 3-D, 2^31 in 2-D)");
	if (too_many_shares(dim, g, bits, n)) return fail(CVTX_B200_ERR_ARGUMENT, "too many particles on a sparse grid for one redistribution call (the sort route holds 2^31 shares; the dense route has no limit)");

	DeviceLock device_lock(d->mu);
	CUDA_TRY(cudaSetDevice(device));
	const double t1 = omp_get_wtime();
	CUDA_TRY(d->remesh[ROWS].reserve(sizeof(float) * row_floats * (size_t)n));
	float *rows = (float *)d->remesh[ROWS].p, *result = nullptr;
	CUDA_TRY(cudaMemcpyAsync(rows, hs.src.p, sizeof(float) * row_floats * (size_t)n, cudaMemcpyHostToDevice, st));
	const bool want = out_rows != nullptr;
	int rc;
	if (bits <= 32)
		rc = dim == 3 ? redistribute_resident<3, uint32_t>(d, st, rows, n, g, bits, negligible, nullptr, &result, want, max_out, n_out, n_nodes)
		              : redistribute_resident<2, uint32_t>(d, st, rows, n, g, bits, negligible, nullptr, &result, want, max_out, n_out, n_nodes);
	else
		rc = dim == 3 ? redistribute_resident<3, uint64_t>(d, st, rows, n, g, bits, negligible, nullptr, &result, want, max_out, n_out, n_nodes)
		              : redistribute_resident<2, uint64_t>(d, st, rows, n, g, bits, negligible, nullptr, &result, want, max_out, n_out, n_nodes);
	if (rc) return rc;
	const double t2 = omp_get_wtime();
	if (want && *n_out > 0) {
		// the created particles: pinned staging first (a pageable destination would be bounced by the driver)
		const size_t bytes = sizeof(float) * row_floats * (size_t)*n_out;
		CUDA_TRY(hs.out.reserve(bytes));
		CUDA_TRY(cudaMemcpyAsync(hs.out.p, result, bytes, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		fetch_bytes(out_rows, hs.out.p, bytes);
	}
	if (trace) std::fprintf(stderr, "cvortex trace:   gather + grid %.3f ms, H2D + kernels %.3f ms, D2H + copy out %.3f ms (%d-bit codes)\n", (t1 - t0) * 1e3,
	                        (t2 - t1) * 1e3, (omp_get_wtime() - t2) * 1e3, bits <= 32 ? 32 : 64);
	return CVTX_B200_OK;
}

int device_redistribute(int device, void *stream, int dim, int kind, const float *rows_dev, long n, float h, float negligible, float *out_dev,
                        int max_out, int *n_out) {
	if (dim != 2 && dim != 3) return fail(CVTX_B200_ERR_ARGUMENT, "dimension must be 2 or 3");
	if (kind < 0 || kind >= K_COUNT) return fail(CVTX_B200_ERR_ARGUMENT, "unknown redistribution function");
	if (n < 0 || n > 0x7ffffff0L || !n_out || !(h > 0.f) || max_out < 0) return fail(CVTX_B200_ERR_ARGUMENT, "bad count, spacing or null n_out");
	*n_out = 0;
	if (n == 0) return CVTX_B200_OK;
	if (!rows_dev) return fail(CVTX_B200_ERR_ARGUMENT, "null particle rows");
	cudaStream_t own = nullptr;
	if (int rc = device_stream(device, &own)) return rc;
	cudaStream_t st = stream ? (cudaStream_t)stream : own;
	Device *d = get_device(device);
	DeviceLock device_lock(d->mu);
	CUDA_TRY(cudaSetDevice(device));

	// bounds of the particle set -> grid placement (same formulas as the host-array route)
	const unsigned blocks = (unsigned)((n + kRowsPerBlock - 1) / kRowsPerBlock);
	CUDA_TRY(d->remesh[PARTIAL].reserve(sizeof(Bounds) * blocks));
	Bounds *partial = (Bounds *)d->remesh[PARTIAL].p;
	if (dim == 3) bounds_kernel<3><<<blocks, kBlock, 0, st>>>(rows_dev, n, partial);
	else bounds_kernel<2><<<blocks, kBlock, 0, st>>>(rows_dev, n, partial);
	std::vector<Bounds> host(blocks);
	CUDA_TRY(cudaMemcpyAsync(host.data(), partial, sizeof(Bounds) * blocks, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	count_launches(1);
	float lo[3], hi[3];
	double sum[3] = {0, 0, 0};
	for (int a = 0; a < 3; ++a) { lo[a] = host[0].lo[a]; hi[a] = host[0].hi[a]; }
	for (const Bounds &q : host)
		for (int a = 0; a < dim; ++a) { lo[a] = fminf(lo[a], q.lo[a]); hi[a] = fmaxf(hi[a], q.hi[a]); sum[a] += q.sum[a]; }
	uint32_t max_index = 0;
	const Grid g = grid_from_bounds(dim, kind, kHalfWidth[kind], h, lo, hi, sum, n, &max_index);
	const int bits = code_bits(dim, max_index);
	if (bits < 0) return fail(CVTX_B200_ERR_ARGUMENT, "grid too large for the node codes (more than 2^21 nodes per axis in 3-D, 2^31 in 2-D)");
	if (too_many_shares(dim, g, bits, n)) return fail(CVTX_B200_ERR_ARGUMENT, "too many particles on a sparse grid for one redistribution call (the sort route holds 2^31 shares; the dense route has no limit)");

	const bool want = out_dev != nullptr;
	if (bits <= 32)
		return dim == 3 ? redistribute_resident<3, uint32_t>(d, st, rows_dev, n, g, bits, negligible, out_dev, nullptr, want, max_out, n_out, nullptr)
		                : redistribute_resident<2, uint32_t>(d, st, rows_dev, n, g, bits, negligible, out_dev, nullptr, want, max_out, n_out, nullptr);
	return dim == 3 ? redistribute_resident<3, uint64_t>(d, st, rows_dev, n, g, bits, negligible, out_dev, nullptr, want, max_out, n_out, nullptr)
	                : redistribute_resident<2, uint64_t>(d, st, rows_dev, n, g, bits, negligible, out_dev, nullptr, want, max_out, n_out, nullptr);
}

}  // namespace remesh
}  // namespace cvtx

// ---- Pedrizzetti relaxation on device-resident particles -----------------------------------
namespace cvtx {
namespace remesh {
namespace {

__global__ void __launch_bounds__(kBlock) particle_positions(const float *__restrict__ rows, int n, float *__restrict__ points) {
	const int i = blockIdx.x * kBlock + threadIdx.x;
	if (i >= n) return;
	for (int a = 0; a < 3; ++a) points[(size_t)i * 3 + a] = rows[(size_t)i * 7 + a];
}

// alpha <- (1 - f dt) alpha + f dt |alpha| omega / |omega|, each operation a single rounded FP32
// op in the reference's order (src/P3D.cpp:687-699); zero where the field is zero.
__global__ void __launch_bounds__(kBlock) relax_blend(float *__restrict__ rows, const float *__restrict__ field, int n, float fdt) {
	const int i = blockIdx.x * kBlock + threadIdx.x;
	if (i >= n) return;
	float *a = rows + (size_t)i * 7 + 3;
	const float *w = field + (size_t)i * 3;
	const float wn = __fsqrt_rn(rm_add(rm_add(rm_mul(w[0], w[0]), rm_mul(w[1], w[1])), rm_mul(w[2], w[2])));
	const float an = __fsqrt_rn(rm_add(rm_add(rm_mul(a[0], a[0]), rm_mul(a[1], a[1])), rm_mul(a[2], a[2])));
	const float keep = rm_sub(1.f, fdt), pull = rm_mul(__fdiv_rn(an, wn), fdt);
	for (int c = 0; c < 3; ++c) {
		const float r = rm_add(rm_mul(a[c], keep), rm_mul(w[c], pull));
		a[c] = wn != 0.f ? r : 0.f;
	}
}

}  // namespace
}  // namespace remesh
}  // namespace cvtx

// ---- thin C ABI (include/cvtx_b200.h) ------------------------------------------------------
extern "C" CVTX_B200_API int cvtx_b200_pedrizzetti_relaxation(int reg, int device, void *stream, float *rows_dev, int n, float fdt, float sigma) {
	using namespace cvtx;
	using namespace cvtx::remesh;
	if (n < 0) return fail(CVTX_B200_ERR_ARGUMENT, "negative count");
	if (n == 0) return CVTX_B200_OK;
	if (!rows_dev) return fail(CVTX_B200_ERR_ARGUMENT, "null particle rows");
	cudaStream_t own = nullptr;
	if (int rc = device_stream(device, &own)) return rc;
	cudaStream_t st = stream ? (cudaStream_t)stream : own;
	Device *d = get_device(device);
	// points / field are per-device scratch: locked for the whole call (the lock is recursive, cvtx_b200_m2m takes it
	// too), and the stream waits for whatever an earlier call on another stream still has in flight on the arena
	DeviceGuard restore;
	DeviceLock device_lock(d->mu);
	CUDA_TRY(cudaSetDevice(device));
	CUDA_TRY(cudaStreamWaitEvent(st, d->arena_idle, 0));
	CUDA_TRY(d->remesh[RELAX_POINTS].reserve(sizeof(float) * 3 * (size_t)n));
	CUDA_TRY(d->remesh[RELAX_FIELD].reserve(sizeof(float) * 3 * (size_t)n));
	float *points = (float *)d->remesh[RELAX_POINTS].p, *field = (float *)d->remesh[RELAX_FIELD].p;
	particle_positions<<<blocks_for((size_t)n), kBlock, 0, st>>>(rows_dev, n, points);
	CUDA_TRY(cudaGetLastError());
	// the vorticity field the particles induce at their own positions: the all-pairs kernel
	if (int rc = cvtx_b200_m2m(CVTX_B200_P3D_VORT, reg, device, st, rows_dev, n, points, n, field, sigma, 0.f)) return rc;
	CUDA_TRY(cudaSetDevice(device));
	relax_blend<<<blocks_for((size_t)n), kBlock, 0, st>>>(rows_dev, field, n, fdt);
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaEventRecord(d->arena_idle, st));
	count_launches(2);
	return CVTX_B200_OK;
}

extern "C" CVTX_B200_API int cvtx_b200_redistribute(int dim, int kind, int device, void *stream, const float *rows_dev, int n,
                                                    float grid_density, float negligible_vort, float *out_dev, int max_out, int *n_out) {
	return cvtx::remesh::device_redistribute(device, stream, dim, kind, rows_dev, n, grid_density, negligible_vort, out_dev, max_out, n_out);
}

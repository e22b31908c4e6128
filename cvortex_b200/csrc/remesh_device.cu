// remesh_device.cu -- particle redistribution onto a regular grid, device stage.
//
// Replaces the tree build of the reference's cvtx_P3D_redistribute_on_grid /
// cvtx_P2D_redistribute_on_grid (src/P3D.cpp:552-589, src/P2D.cpp:325-362): there, every
// thread inserts its particles' (2R+1)^D shares into a private oct/quadtree one key at a
// time and the trees are merged serially (src/GridParticleOcttree.cpp:76-135,216-280).
// Here the same shares are produced per particle, sorted by the Morton code of their node
// and summed per node:
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
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

inline unsigned blocks_for(size_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

// Pinned staging -> the caller's vectors, in parallel pieces; codes are widened to 64
// bits on the way when the device used 32.
template <class K>
void fetch_codes(uint64_t *dst, const K *src, size_t n) {
	const size_t piece = 1 << 18;
	const long pieces = (long)((n + piece - 1) / piece);
#pragma omp parallel for schedule(static) num_threads(4) if (pieces > 4)
	for (long p = 0; p < pieces; ++p) {
		const size_t lo = (size_t)p * piece, hi = lo + piece < n ? lo + piece : n;
		for (size_t i = lo; i < hi; ++i) dst[i] = (uint64_t)src[i];
	}
}
void fetch_bytes(void *dst, const void *src, size_t bytes) {
	const size_t piece = 1 << 20;
	const long pieces = (long)((bytes + piece - 1) / piece);
#pragma omp parallel for schedule(static) num_threads(4) if (pieces > 4)
	for (long i = 0; i < pieces; ++i) {
		const size_t lo = (size_t)i * piece;
		std::memcpy((char *)dst + lo, (const char *)src + lo, lo + piece <= bytes ? piece : bytes - lo);
	}
}

template <int D, class K>
int run(Device *d, cudaStream_t st, const float *rows_host, long n, const Grid &g, int bits, NodeSet *nodes) {
	constexpr int ROW = Layout<D>::ROW, COMPS = Layout<D>::COMPS;
	Buffer *b = d->remesh;
	enum { ROWS, COUNT, OFFSET, CODE_A, CODE_B, REC_A, REC_B, SHARE, TEMP, FIRST, SUMS };
	CUDA_TRY(b[ROWS].reserve(sizeof(float) * ROW * (size_t)n));
	CUDA_TRY(b[COUNT].reserve(sizeof(uint32_t) * (size_t)(n + 1)));
	CUDA_TRY(b[OFFSET].reserve(sizeof(uint32_t) * (size_t)(n + 1)));
	float *rows = (float *)b[ROWS].p;
	uint32_t *count = (uint32_t *)b[COUNT].p, *offset = (uint32_t *)b[OFFSET].p;
	CUDA_TRY(cudaMemcpyAsync(rows, rows_host, sizeof(float) * ROW * (size_t)n, cudaMemcpyHostToDevice, st));

	spread_count<D><<<blocks_for((size_t)n + 1), kBlock, 0, st>>>(rows, n, g, count);
	size_t temp = 0;
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, temp, count, offset, (int)(n + 1), st));
	CUDA_TRY(b[TEMP].reserve(temp));
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(b[TEMP].p, temp, count, offset, (int)(n + 1), st));
	uint32_t total = 0;
	CUDA_TRY(cudaMemcpyAsync(&total, offset + n, sizeof(total), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	count_launches(1);
	nodes->code.clear();
	nodes->strength.clear();
	if (total == 0) return CVTX_B200_OK;
	if (total > 0x7fffffffu) return fail(CVTX_B200_ERR_ARGUMENT, "redistribution creates more than 2^31 particle-node shares");

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

	// results: pinned staging first (a pageable destination would be bounced by the driver)
	HostStage &hs = host_stage();
	const size_t code_bytes = sizeof(K) * (size_t)n_nodes, sum_bytes = sizeof(float) * COMPS * (size_t)n_nodes;
	const size_t sums_at = (code_bytes + 15) & ~(size_t)15;
	CUDA_TRY(hs.out.reserve(sums_at + sum_bytes));
	CUDA_TRY(cudaMemcpyAsync(hs.out.p, node_code, code_bytes, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaMemcpyAsync((char *)hs.out.p + sums_at, sums, sum_bytes, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	nodes->code.resize(n_nodes);
	nodes->strength.resize((size_t)n_nodes * COMPS);
	fetch_codes<K>(nodes->code.data(), (const K *)hs.out.p, n_nodes);
	fetch_bytes(nodes->strength.data(), (const char *)hs.out.p + sums_at, sum_bytes);
	return CVTX_B200_OK;
}

}  // namespace

int device_nodes(int device, int dim, int kind, float h, const void *const *particles, long n, Grid *grid, NodeSet *nodes) {
	if (dim != 2 && dim != 3) return fail(CVTX_B200_ERR_ARGUMENT, "dimension must be 2 or 3");
	if (kind < 0 || kind >= K_COUNT) return fail(CVTX_B200_ERR_ARGUMENT, "unknown redistribution function");
	if (n <= 0 || n > 0x7ffffff0L) return fail(CVTX_B200_ERR_ARGUMENT, "bad particle count");
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
	const float *rows = (const float *)hs.src.p;
	uint32_t max_index = 0;
	*grid = place_grid(dim, kind, kHalfWidth[kind], h, particles, (float *)hs.src.p, n, row_floats, &max_index);
	const int bits = code_bits(dim, max_index);
	if (bits < 0) return fail(CVTX_B200_ERR_ARGUMENT, "grid too large for the node codes (more than 2^21 nodes per axis in 3-D, 2^31 in 2-D)");

	std::lock_guard<std::mutex> device_lock(d->mu);
	CUDA_TRY(cudaSetDevice(device));
	const double t1 = omp_get_wtime();
	int rc;
	if (bits <= 32) rc = dim == 3 ? run<3, uint32_t>(d, st, rows, n, *grid, bits, nodes) : run<2, uint32_t>(d, st, rows, n, *grid, bits, nodes);
	else rc = dim == 3 ? run<3, uint64_t>(d, st, rows, n, *grid, bits, nodes) : run<2, uint64_t>(d, st, rows, n, *grid, bits, nodes);
	if (trace) std::fprintf(stderr, "cvortex trace:   gather + grid %.3f ms, H2D + kernels + D2H %.3f ms (%d-bit codes)\n", (t1 - t0) * 1e3,
	                        (omp_get_wtime() - t1) * 1e3, bits <= 32 ? 32 : 64);
	return rc;
}

}  // namespace remesh
}  // namespace cvtx

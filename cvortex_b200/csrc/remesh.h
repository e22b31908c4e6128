// remesh.h -- internal interface of the grid redistribution: the node set both the
// device stage (remesh_device.cu) and the host stage (remesh.cpp) produce, and the grid
// placement they share.  No CUDA types.
#pragma once
#include <cstdint>
#include <vector>
#include "remesh_math.h"

namespace cvtx {
namespace remesh {

// Grid nodes that received vorticity, in ascending Morton order of their indices, with
// the summed strength (3 components per node in 3-D, 1 in 2-D).
struct NodeSet {
	std::vector<uint64_t> code;
	std::vector<float> strength;
};

// Where the grid sits for a given particle set (reference src/P3D.cpp:538-550,
// src/P2D.cpp:312-323): a node coincides with the mean position, and the origin is
// pushed at least `half` + 5 cells below the lowest particle.  `rows` are the gathered
// particle structs (row_floats floats each, coordinates first) -- gathered here, in the same
// pass, when `particles` (the caller's array of pointers) is not null; `kind` is -1 and `half` the
// caller's (int)roundf(radius) for a user-defined interpolant.  max_index receives the
// largest node index any stencil can touch.
Grid place_grid(int dim, int kind, int half, float h, const void *const *particles, float *rows, long n, int row_floats,
                uint32_t *max_index);

// The same placement from bounds already reduced (lo / hi per axis, FP64 coordinate sums).
Grid grid_from_bounds(int dim, int kind, int half, float h, const float *lo, const float *hi, const double *sum, long n, uint32_t *max_index);

// Bits of Morton code needed for indices <= max_index; -1 when the grid is larger than
// the codes can hold (2^21 nodes per axis in 3-D, 2^31 in 2-D).
int code_bits(int dim, uint32_t max_index);

// Which of the two node-building routes a call takes (both stages, device and host): the dense
// route visits every grid node and gathers its shares from the particles of the cells around
// it; it needs a grid small enough to enumerate and to index (2^27 codes, 2^26 cells),
// populated enough that most of its nodes exist, and not crowded (the reference's 2-D
// benchmark puts 150 particles in a cell: few nodes, long walks).  Otherwise every share is
// written out and sorted by node.  Both give the same bits.
bool dense_route(int dim, const Grid &g, int bits, long n);
bool too_many_shares(int dim, const Grid &g, int bits, long n);     // the sort route's limit (2^31 shares); the dense route has none

// The host-array entry points on `device`: particles (array of pointers to cvtx_P3D / cvtx_P2D)
// in, created particles out (out_rows may be null: count only).  Node build and pruning both
// run on the device; the result is the host stage's, bit for bit.  Returns a cvtx_b200_status.
int device_redistribute_from_host(int device, int dim, int kind, float h, const void *const *particles, long n, float negligible,
                                  void *out_rows, int max_out, int *n_out, size_t *n_nodes);

// The order in which the pruning stage forms its FP64 sums (mean strength, dropped vorticity),
// on the device (block_reduce in remesh_device.cu: a shuffle tree per warp, warps in order,
// blocks in order) and, through this function, on the host: elements in blocks of 256; inside a
// block eight 32-element trees (a[l] += a[l + o] for o = 16, 8, 4, 2, 1), their results added in
// order; block sums added in block order.  Elements past n count as 0.  One order everywhere
// is what lets the device prune for the host-array entry points and still return the host
// stage's bits.
template <class Get>
double canonical_sum(long n, Get &&get) {
	const long blocks = (n + 255) / 256;
	std::vector<double> part((size_t)blocks);
#pragma omp parallel for schedule(static) if (blocks > 64)
	for (long b = 0; b < blocks; ++b) {
		double block = 0.0;
		for (int w = 0; w < 8; ++w) {
			double a[32];
			for (int l = 0; l < 32; ++l) {
				const long i = b * 256 + w * 32 + l;
				a[l] = i < n ? (double)get(i) : 0.0;
			}
			for (int o = 16; o > 0; o >>= 1)
				for (int l = 0; l < o; ++l) a[l] += a[l + o];
			block = w == 0 ? a[0] : block + a[0];
		}
		part[(size_t)b] = block;
	}
	double total = 0.0;
	for (long b = 0; b < blocks; ++b) total += part[(size_t)b];
	return total;
}

// Device-resident redistribution (the additive cvtx_b200_redistribute of cvtx_b200.h): rows_dev
// in, out_dev (capacity max_out rows, may be null = count only) out, both on `device`.
int device_redistribute(int device, void *stream, int dim, int kind, const float *rows_dev, long n, float h, float negligible,
                        float *out_dev, int max_out, int *n_out);

// The strength above which about `wanted` of the n particles remain: repeated 1024-bin
// histograms of [min, max], zooming into the bin where the count from the top crosses
// `wanted` (reference src/redistribution_helper_funcs.cpp:32-91, same arithmetic).  The
// data stay with the caller: minmax(&min, &max) and histogram(lo, range, counts[1024]) --
// bin = cut_bin(floor(1023 (s - lo) / range)) in FP64 (remesh_math.h) -- are supplied by the
// host stage (a loop) and by the device stage (kernels), so both take the same decisions.
template <class MinMax, class Histogram>
float strength_cut_with(int n, int wanted, MinMax &&minmax, Histogram &&histogram) {
	float fmin = 0.f, fmax = 0.f;
	minmax(&fmin, &fmax);
	double lo = fmin, hi = fmax;
	if (n < wanted) return (float)(hi * 1.05);
	std::vector<float> edge(kCutBins);
	std::vector<int> count(kCutBins);
	int k = 0;
	// every round narrows [lo, hi] to one of 1024 bins, so FP32 strengths are separated after 4
	// rounds at most; the cap only guards against NaN input, where no comparison ever succeeds
	for (int round = 0; round < 64; ++round) {
		const double range = (hi - lo) * 1.05;
		for (int i = 0; i < kCutBins; ++i) edge[i] = (float)(lo + i * range / (float)(kCutBins - 1));
		histogram(lo, range, count.data());
		int above = count[kCutBins - 1];
		for (int i = kCutBins - 2; i >= 0; --i) {
			hi = edge[i + 1];
			lo = edge[i];
			above += count[i];
			count[i] = above;
			if (above > wanted) { k = i + 1; break; }
		}
		if (k < 1) break;      // the count never crossed `wanted` (NaN strengths): nothing to refine
		const float miss = (float)(wanted - count[k]) / (float)wanted;
		if (lo == hi || count[k] == count[k - 1] || (double)(miss < 0.f ? -miss : miss) < 0.01f * 0.6) break;
	}
	return edge[k];
}

}  // namespace remesh
}  // namespace cvtx

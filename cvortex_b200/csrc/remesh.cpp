// remesh.cpp -- the cvtx_* entry points that sit either side of the all-pairs sums in a
// vortex-particle time step: redistribution onto a regular grid (2-D and 3-D) and
// Pedrizzetti relaxation.
//
// Replaces, in the reference:
//   * src/RedistFunc.cpp:36-96            the five interpolants and their constructors;
//   * src/P3D.cpp:509-665, src/P2D.cpp:283-436   cvtx_P3D/P2D_redistribute_on_grid and the
//     strength-threshold pruning behind them (src/redistribution_helper_funcs.cpp:32-91,
//     src/array_methods.cpp farray_info / minmax / mean);
//   * src/P3D.cpp:667-707                 cvtx_P3D_pedrizzetti_relaxation.
// (On g++ the reference's own tree build is an unfinished stub -- src/UIntKey96.h:243-244
// `assert(false); /* TO DO. */` -- so under Linux these entry points only work here.)
//
// A redistribution call is two stages.  Stage A turns the particles into the set of grid
// nodes that receive vorticity, with each node's summed strength: on the first enabled
// accelerator when the caller passed one of the five built-in interpolants
// (remesh_device.cu), else on the host (host_nodes_dense / host_nodes below: the same two
// routes, same arithmetic, same bits).
// Stage B (prune below) is the reference's post-processing on that node set -- it has to
// end in the caller's host array anyway: drop nodes weaker than negligible_vort x the mean
// strength, hand the dropped vorticity back evenly, and if the caller's array is still too
// small find the strength threshold that fits it.
//
// Relaxation is one cvtx_P3D_M2M_vort call (GPU when enabled) and a per-particle blend.
#include <omp.h>
#include <parallel/algorithm>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/cvtx_b200.h"
#include "export.h"
#include "host_hooks.h"
#include "remesh.h"

using namespace cvtx;
using namespace cvtx::remesh;

// ---- the interpolants as plain functions (what cvtx_RedistFunc::func points at) -------
namespace {
float f_lambda0(float U) { return weight(K_LAMBDA0, U); }
float f_lambda1(float U) { return weight(K_LAMBDA1, U); }
float f_lambda2(float U) { return weight(K_LAMBDA2, U); }
float f_lambda3(float U) { return weight(K_LAMBDA3, U); }
float f_m4p(float U) { return weight(K_M4P, U); }
float (*const kBuiltin[K_COUNT])(float) = {f_lambda0, f_lambda1, f_lambda2, f_lambda3, f_m4p};

cvtx_RedistFunc builtin(int kind) {
	cvtx_RedistFunc r;
	r.func = kBuiltin[kind];
	r.radius = kRadius[kind];
	return r;
}
// Which built-in a caller's cvtx_RedistFunc is, or -1 for a user-defined one.
int builtin_kind(const cvtx_RedistFunc *r) {
	for (int k = 0; k < K_COUNT; ++k)
		if (r->func == kBuiltin[k] && r->radius == kRadius[k]) return k;
	return -1;
}
}  // namespace

namespace {
constexpr long kPiece = 1 << 16;      // rows per unit of host-side parallel work
// Host threads for the O(n) passes.  Not left to OMP_NUM_THREADS: launchers such as
// torchrun export OMP_NUM_THREADS=1.
int host_threads() {
	static const int n = [] { int p = omp_get_num_procs(); return p > 8 ? 8 : (p < 1 ? 1 : p); }();
	return n;
}
}  // namespace

// ---- grid placement (shared with the device stage) -----------------------------------
Grid cvtx::remesh::place_grid(int dim, int kind, int half, float h, const void *const *particles, float *rows, long n, int row_floats,
                              uint32_t *max_index) {
	// gather (when asked) and bounds in fixed-size pieces, in parallel; the FP64 coordinate sums
	// in the canonical order of remesh.h, which is also what the device-resident entry point's
	// bounds kernel produces -- the grid then sits at the same place on every route
	struct Piece { float lo[3], hi[3]; };
	const long pieces = (n + kPiece - 1) / kPiece;
	std::vector<Piece> piece((size_t)pieces);
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (pieces > 1)
	for (long p = 0; p < pieces; ++p) {
		Piece q;
		const long b = p * kPiece, e = b + kPiece < n ? b + kPiece : n;
		if (particles)   // gather this piece through the caller's pointer array first
			for (long i = b; i < e; ++i) std::memcpy(rows + i * row_floats, particles[i], sizeof(float) * (size_t)row_floats);
		for (int a = 0; a < 3; ++a) q.lo[a] = q.hi[a] = a < dim ? rows[b * row_floats + a] : 0.f;
		for (long i = b; i < e; ++i)
			for (int a = 0; a < dim; ++a) {
				const float x = rows[i * row_floats + a];
				q.lo[a] = q.lo[a] > x ? x : q.lo[a];
				q.hi[a] = q.hi[a] < x ? x : q.hi[a];
			}
		piece[(size_t)p] = q;
	}
	float lo[3], hi[3];
	double sum[3] = {0, 0, 0};
	for (int a = 0; a < 3; ++a) { lo[a] = piece[0].lo[a]; hi[a] = piece[0].hi[a]; }
	for (const Piece &q : piece)
		for (int a = 0; a < dim; ++a) {
			lo[a] = lo[a] > q.lo[a] ? q.lo[a] : lo[a];
			hi[a] = hi[a] < q.hi[a] ? q.hi[a] : hi[a];
		}
	for (int a = 0; a < dim; ++a) sum[a] = canonical_sum(n, [&](long i) { return rows[i * row_floats + a]; });
	return grid_from_bounds(dim, kind, half, h, lo, hi, sum, n, max_index);
}

Grid cvtx::remesh::grid_from_bounds(int dim, int kind, int half, float h, const float *lo, const float *hi, const double *sum, long n,
                                    uint32_t *max_index) {
	Grid g;
	g.h = h;
	g.rh = 1.f / h;
	g.kind = kind;
	g.half = half;
	uint32_t top = 0;
	for (int a = 0; a < 3; ++a) { g.origin[a] = 0.f; g.top[a] = 0; }
	for (int a = 0; a < dim; ++a) {
		const float mean = (float)(sum[a] / (double)n);
		const float corner = lo[a] - 1.f * ((float)g.half * h);
		const float cells = roundf((mean - corner) / h) + 5.f;
		g.origin[a] = mean - cells * h;
		const uint32_t k = (dim == 3 ? node_index_3d(hi[a], g.origin[a], g.rh) : node_index_2d(hi[a], g.origin[a], g.rh)) + (uint32_t)g.half;
		g.top[a] = k;
		top = k > top ? k : top;
	}
	if (max_index) *max_index = top;
	return g;
}

// The sort route materialises every share ((2 half + 1)^D per particle) and indexes them with 32 bits; the
// dense route never writes one.  Only the former has a particle limit per call.
bool cvtx::remesh::too_many_shares(int dim, const Grid &g, int bits, long n) {
	if (dense_route(dim, g, bits, n)) return false;
	const double stencil = dim == 3 ? (2.0 * g.half + 1) * (2.0 * g.half + 1) * (2.0 * g.half + 1) : (2.0 * g.half + 1) * (2.0 * g.half + 1);
	return (double)n * stencil > 2147483647.0;
}

bool cvtx::remesh::dense_route(int dim, const Grid &g, int bits, long n) {
	double cells = 1.0;
	for (int a = 0; a < dim; ++a) cells *= (double)g.top[a] + 1.0;
	const double stencil = dim == 3 ? (2.0 * g.half + 1) * (2.0 * g.half + 1) * (2.0 * g.half + 1) : (2.0 * g.half + 1) * (2.0 * g.half + 1);
	const bool fits = bits > 0 && bits <= 27 && cells <= 67108864.0;
	// CVTX_B200_REMESH_ROUTE=sort|dense pins the route (benchmarks, tests); dense still needs a grid that fits
	static const char *pin = std::getenv("CVTX_B200_REMESH_ROUTE");
	if (pin && !std::strcmp(pin, "sort")) return false;
	if (pin && !std::strcmp(pin, "dense")) return fits;
	// populated enough that most nodes exist, not so crowded that a node walks thousands of particles
	return fits && (double)n * stencil >= 0.25 * cells && (double)n <= 16.0 * cells;
}

int cvtx::remesh::code_bits(int dim, uint32_t max_index) {
	int b = 0;
	while (b < 32 && (max_index >> b) != 0) ++b;
	if (b < 1) b = 1;
	if (dim == 3 ? b > kBits3D : b > 31) return -1;
	return dim * b;
}

namespace {

// ---- stage A on the host --------------------------------------------------------------
// Same shares, same summation order, same FP64 sums as remesh_device.cu.
// The particle order every stage emits shares in: by cell (remesh_math.h cell_of), then by
// the caller's index.
template <int D>
std::vector<uint32_t> cell_order(const float *rows, long n, const Grid &g) {
	constexpr int ROW = D == 3 ? 7 : 4;
	struct Key { uint64_t cell; uint32_t at; };
	std::vector<Key> key((size_t)n);
#pragma omp parallel for schedule(static)
	for (long i = 0; i < n; ++i) key[(size_t)i] = {cell_of<D>(rows + i * ROW, g), (uint32_t)i};
	__gnu_parallel::stable_sort(key.begin(), key.end(), [](const Key &a, const Key &b) { return a.cell < b.cell; });
	std::vector<uint32_t> order((size_t)n);
#pragma omp parallel for schedule(static)
	for (long i = 0; i < n; ++i) order[(size_t)i] = key[(size_t)i].at;
	return order;
}

template <int D>
void host_nodes(const float *rows, long n, const Grid &g, NodeSet *nodes) {
	constexpr int ROW = D == 3 ? 7 : 4, COMPS = D == 3 ? 3 : 1;
	const std::vector<uint32_t> order = cell_order<D>(rows, n, g);
	std::vector<uint32_t> offset((size_t)n + 1, 0);
#pragma omp parallel for schedule(static)
	for (long i = 0; i < n; ++i) offset[i + 1] = (uint32_t)spread_particle<D>(rows + (size_t)order[(size_t)i] * ROW, g, [](uint64_t, const float *) {});
	for (long i = 0; i < n; ++i) offset[i + 1] += offset[i];
	const size_t total = offset[n];
	struct Record { uint64_t code; uint32_t at; };
	std::vector<Record> rec(total);
	std::vector<float> share(total * COMPS);
#pragma omp parallel for schedule(static)
	for (long i = 0; i < n; ++i) {
		uint32_t at = offset[i];
		spread_particle<D>(rows + (size_t)order[(size_t)i] * ROW, g, [&](uint64_t m, const float *s) {
			rec[at] = {m, at};
			for (int c = 0; c < COMPS; ++c) share[(size_t)at * COMPS + c] = s[c];
			++at;
		});
	}
	__gnu_parallel::stable_sort(rec.begin(), rec.end(), [](const Record &a, const Record &b) { return a.code < b.code; });
	nodes->code.clear();
	nodes->strength.clear();
	for (size_t j = 0; j < total;) {
		double acc[COMPS] = {};
		size_t e = j;
		for (; e < total && rec[e].code == rec[j].code; ++e)
			for (int c = 0; c < COMPS; ++c) acc[c] += (double)share[(size_t)rec[e].at * COMPS + c];
		nodes->code.push_back(rec[j].code);
		for (int c = 0; c < COMPS; ++c) nodes->strength.push_back((float)acc[c]);
		j = e;
	}
}

// Stage A on the host, dense route: what remesh_device.cu's gather_nodes does, one grid node
// per loop iteration instead of per thread.  Nodes are visited in Morton order; a node's shares
// come from the particles of the (2R+1)^D cells around it, walked in cell order -- with cells
// numbered x fastest each x-row of cells is one contiguous run of the cell-sorted particles.
// Same shares and same summation order as host_nodes() above, so the same bits, without ever
// materialising a share.
template <int D, int KIND>
void host_nodes_dense(const float *rows, long n, const Grid &g, int bits, NodeSet *nodes) {
	constexpr int ROW = D == 3 ? 7 : 4, COMPS = D == 3 ? 3 : 1, R = kHalfWidth[KIND];
	const std::vector<uint32_t> order = cell_order<D>(rows, n, g);
	std::vector<float> sorted((size_t)n * ROW);                 // the particles, physically in cell order
#pragma omp parallel for schedule(static) num_threads(host_threads())
	for (long j = 0; j < n; ++j) std::memcpy(&sorted[(size_t)j * ROW], rows + (size_t)order[(size_t)j] * ROW, sizeof(float) * ROW);
	uint64_t n_cells = 1;
	for (int a = 0; a < D; ++a) n_cells *= (uint64_t)g.top[a] + 1;
	std::vector<uint32_t> cell_start(n_cells + 1, 0);
	for (long j = 0; j < n; ++j) ++cell_start[cell_of<D>(&sorted[(size_t)j * ROW], g) + 1];
	for (uint64_t c = 0; c < n_cells; ++c) cell_start[c + 1] += cell_start[c];

	// Morton codes in runs of 4096 (a 16^3 brick in 3-D), handed out dynamically: bricks outside
	// the grid cost nothing, bricks inside differ in how many particles they hold
	constexpr uint64_t kBrick = 4096;
	const uint64_t domain = (uint64_t)1 << bits, nx = (uint64_t)g.top[0] + 1, ny = (uint64_t)g.top[1] + 1;
	const long pieces = (long)((domain + kBrick - 1) / kBrick);
	std::vector<NodeSet> found((size_t)pieces);
#pragma omp parallel for schedule(dynamic, 1) num_threads(host_threads())
	for (long p = 0; p < pieces; ++p) {
		NodeSet &mine = found[(size_t)p];
		const uint64_t t_end = (uint64_t)(p + 1) * kBrick < domain ? (uint64_t)(p + 1) * kBrick : domain;
		for (uint64_t t = (uint64_t)p * kBrick; t < t_end; ++t) {
			uint32_t node[3] = {0, 0, 0};
			float at[3] = {0.f, 0.f, 0.f};                      // the node's coordinates
			bool inside = true;
			for (int a = 0; a < D; ++a) {
				node[a] = (uint32_t)(D == 3 ? compact3(t >> a) : compact2(t >> a));
				at[a] = node_coord(node[a], g.origin[a], g.h);
				inside = inside && node[a] <= g.top[a];
			}
			if (!inside) continue;
			uint32_t lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
			for (int a = 0; a < D; ++a) {
				lo[a] = node[a] >= (uint32_t)R ? node[a] - (uint32_t)R : 0u;
				hi[a] = node[a] + (uint32_t)R <= g.top[a] ? node[a] + (uint32_t)R : g.top[a];
			}
			double acc[COMPS] = {};
			bool any = false;
			for (uint32_t cz = lo[2]; cz <= hi[2]; ++cz)
				for (uint32_t cy = lo[1]; cy <= hi[1]; ++cy) {
					const uint64_t line = nx * (cy + ny * cz);
					for (uint32_t j = cell_start[line + lo[0]]; j < cell_start[line + hi[0] + 1]; ++j) {
						const float *row = &sorted[(size_t)j * ROW];
						// cell_distance() with the node coordinate hoisted: same operations
						float f = weight(KIND, std::fabs((row[0] - at[0]) * g.rh)) * weight(KIND, std::fabs((row[1] - at[1]) * g.rh));
						if (D == 3) f = f * weight(KIND, std::fabs((row[2] - at[2]) * g.rh));
						float s[COMPS];
						bool nz = false;
						for (int c = 0; c < COMPS; ++c) { s[c] = row[D + c] * f; nz = nz || s[c] != 0.f; }
						if (nz) {
							any = true;
							for (int c = 0; c < COMPS; ++c) acc[c] += (double)s[c];
						}
					}
				}
			if (any) {
				mine.code.push_back(t);
				for (int c = 0; c < COMPS; ++c) mine.strength.push_back((float)acc[c]);
			}
		}
	}
	nodes->code.clear();
	nodes->strength.clear();
	for (const NodeSet &piece : found) {
		nodes->code.insert(nodes->code.end(), piece.code.begin(), piece.code.end());
		nodes->strength.insert(nodes->strength.end(), piece.strength.begin(), piece.strength.end());
	}
}

// User-defined interpolant: the weights come from the caller's function pointer, one call
// per axis and stencil offset, so this cannot share spread_particle().  Records are
// produced in the same order and summed the same way.
template <int D>
void host_nodes_user(const float *rows, long n, Grid g, const cvtx_RedistFunc *rf, NodeSet *nodes) {
	constexpr int ROW = D == 3 ? 7 : 4, COMPS = D == 3 ? 3 : 1;
	const int R = g.half, S = 2 * R + 1;
	struct Record { uint64_t code; uint32_t at; };
	std::vector<Record> rec;
	std::vector<float> share;
	std::vector<float> w((size_t)D * S);
	for (long i = 0; i < n; ++i) {
		const float *row = rows + i * ROW;
		uint32_t k0[3] = {0, 0, 0};
		for (int a = 0; a < D; ++a) {
			k0[a] = D == 3 ? node_index_3d(row[a], g.origin[a], g.rh) : node_index_2d(row[a], g.origin[a], g.rh);
			for (int o = 0; o < S; ++o) w[(size_t)a * S + o] = rf->func(cell_distance(row[a], k0[a] + (uint32_t)(o - R), g, a));
		}
		for (int a = 0; a < S; ++a)
			for (int b = 0; b < S; ++b)
				for (int c = 0; c < (D == 3 ? S : 1); ++c) {
					float f = w[a] * w[(size_t)S + b];
					if (D == 3) f = f * w[(size_t)2 * S + c];
					float s[COMPS];
					bool any = false;
					for (int q = 0; q < COMPS; ++q) { s[q] = row[D + q] * f; any = any || s[q] != 0.f; }
					if (!any) continue;
					const uint64_t m = D == 3 ? morton3(k0[0] + (uint32_t)(a - R), k0[1] + (uint32_t)(b - R), k0[2] + (uint32_t)(c - R))
					                          : morton2(k0[0] + (uint32_t)(a - R), k0[1] + (uint32_t)(b - R));
					rec.push_back({m, (uint32_t)rec.size()});
					for (int q = 0; q < COMPS; ++q) share.push_back(s[q]);
				}
	}
	std::stable_sort(rec.begin(), rec.end(), [](const Record &a, const Record &b) { return a.code < b.code; });
	nodes->code.clear();
	nodes->strength.clear();
	for (size_t j = 0; j < rec.size();) {
		double acc[COMPS] = {};
		size_t e = j;
		for (; e < rec.size() && rec[e].code == rec[j].code; ++e)
			for (int c = 0; c < COMPS; ++c) acc[c] += (double)share[(size_t)rec[e].at * COMPS + c];
		nodes->code.push_back(rec[j].code);
		for (int c = 0; c < COMPS; ++c) nodes->strength.push_back((float)acc[c]);
		j = e;
	}
}

// ---- stage B: pruning (reference src/P3D.cpp:590-665) ----------------------------------

// The strength above which about `wanted` of the n particles remain: repeated 1024-bin
// histograms of [min, max], zooming into the bin where the count from the top crosses
// `wanted` (reference src/redistribution_helper_funcs.cpp:32-91, same arithmetic).
float strength_cut(const std::vector<float> &s, int n, int wanted) {
	return strength_cut_with(
	    n, wanted,
	    [&](float *fmin, float *fmax) {
		    *fmin = *fmax = n > 0 ? s[0] : 0.f;
		    for (int i = 0; i < n; ++i) { *fmin = *fmin < s[i] ? *fmin : s[i]; *fmax = *fmax > s[i] ? *fmax : s[i]; }
	    },
	    [&](double lo, double range, int *count) {
		    for (int i = 0; i < kCutBins; ++i) count[i] = 0;
		    for (int i = 0; i < n; ++i) {
			    ++count[cut_bin(std::floor((double)(kCutBins - 1) * (s[i] - lo) / range))];
		    }
	    });
}

// Which nodes survive a cut: node i stays when strength[i] > cut and i < index_limit; the
// vorticity of the others is handed back evenly to the survivors (reference
// src/P3D.cpp:636-665).  The node list is walked in fixed pieces, in parallel; the FP64 sums
// follow the canonical order of remesh.h, so the result does not depend on the thread count
// and equals the device's.
template <int COMPS>
struct Cut {
	float cut;
	long index_limit, kept;
	float each[COMPS];            // what every survivor receives
	std::vector<long> start;      // output position of each piece's first survivor
	bool keeps(float strength, long i) const { return strength > cut && i < index_limit; }
};

template <int COMPS>
Cut<COMPS> plan_cut(const NodeSet &nodes, const std::vector<float> &strength, long n, float cut, long index_limit) {
	Cut<COMPS> c;
	c.cut = cut;
	c.index_limit = index_limit;
	const long pieces = (n + kPiece - 1) / kPiece;
	std::vector<long> tally((size_t)pieces);
	const float *w = nodes.strength.data();
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (pieces > 1)
	for (long p = 0; p < pieces; ++p) {
		long kept = 0;
		const long b = p * kPiece, e = b + kPiece < n ? b + kPiece : n;
		for (long i = b; i < e; ++i) kept += c.keeps(strength[(size_t)i], i) ? 1 : 0;
		tally[(size_t)p] = kept;
	}
	c.kept = 0;
	c.start.resize((size_t)pieces);
	for (long p = 0; p < pieces; ++p) {
		c.start[(size_t)p] = c.kept;
		c.kept += tally[(size_t)p];
	}
	// the dropped vorticity, summed in the canonical order over n + 1 slots (the device's keep
	// kernel covers one slot past the last node)
	for (int q = 0; q < COMPS; ++q) {
		const double lost = canonical_sum(n + 1, [&](long i) {
			return i < n && !c.keeps(strength[(size_t)i], i) ? w[(size_t)i * COMPS + q] : 0.f;
		});
		c.each[q] = (float)lost / (float)c.kept;
	}
	return c;
}

// Hand every survivor of `c`, in order, to sink(position, code, adjusted strength).
template <int COMPS, class Sink>
void apply_cut(const Cut<COMPS> &c, const NodeSet &nodes, const std::vector<float> &strength, long n, Sink &&sink) {
	const long pieces = (n + kPiece - 1) / kPiece;
	const float *w = nodes.strength.data();
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (pieces > 1)
	for (long p = 0; p < pieces; ++p) {
		const long b = p * kPiece, e = b + kPiece < n ? b + kPiece : n;
		long at = c.start[(size_t)p];
		for (long i = b; i < e; ++i)
			if (c.keeps(strength[(size_t)i], i)) {
				float adjusted[COMPS];
				for (int q = 0; q < COMPS; ++q) adjusted[q] = w[(size_t)i * COMPS + q] + c.each[q];
				sink(at++, nodes.code[(size_t)i], adjusted);
			}
	}
}

// |w| per node and, in FP64 and the canonical order (remesh.h), their sum.
template <int COMPS>
double magnitudes(const std::vector<float> &w, int n, std::vector<float> &out) {
	out.resize((size_t)n);
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (n > kPiece)
	for (long i = 0; i < (long)n; ++i) {
		float m;
		if (COMPS == 1) m = std::fabs(w[(size_t)i]);
		else m = std::sqrt(w[(size_t)i * 3] * w[(size_t)i * 3] + w[(size_t)i * 3 + 1] * w[(size_t)i * 3 + 1] + w[(size_t)i * 3 + 2] * w[(size_t)i * 3 + 2]);
		out[(size_t)i] = m;
	}
	return canonical_sum(n, [&](long i) { return out[(size_t)i]; });
}

template <int D, class Particle>
int prune_and_write(NodeSet &nodes, const Grid &g, Particle *out, int max_out, float negligible) {
	constexpr int COMPS = D == 3 ? 3 : 1;
	long n = (long)nodes.code.size();
	if (n == 0) return 0;
	std::vector<float> strength;
	const double total = magnitudes<COMPS>(nodes.strength, (int)n, strength);
	Cut<COMPS> c = plan_cut<COMPS>(nodes, strength, n, (float)(total / (double)n) * negligible, n);
	if (!out) return (int)c.kept;
	if (c.kept > max_out) {
		// the caller's array is too small: materialise the survivors, find the strength that
		// fits, cut again.  The reference also drops every node at index >= max_out whatever
		// its strength (src/P3D.cpp:649: `i < max_keepable` tests the input index); kept as is.
		NodeSet kept;
		kept.code.resize((size_t)c.kept);
		kept.strength.resize((size_t)c.kept * COMPS);
		apply_cut<COMPS>(c, nodes, strength, n, [&](long at, uint64_t code, const float *w) {
			kept.code[(size_t)at] = code;
			for (int q = 0; q < COMPS; ++q) kept.strength[(size_t)at * COMPS + q] = w[q];
		});
		nodes.code.swap(kept.code);
		nodes.strength.swap(kept.strength);
		n = c.kept;
		magnitudes<COMPS>(nodes.strength, (int)n, strength);
		c = plan_cut<COMPS>(nodes, strength, n, strength_cut(strength, (int)n, max_out), max_out);
	}
	const float size = D == 3 ? g.h * g.h * g.h : g.h * g.h;
	apply_cut<COMPS>(c, nodes, strength, n, [&](long at, uint64_t m, const float *w) {
		float *p = (float *)&out[at];
		if (D == 3) {
			p[0] = node_coord((uint32_t)compact3(m), g.origin[0], g.h);
			p[1] = node_coord((uint32_t)compact3(m >> 1), g.origin[1], g.h);
			p[2] = node_coord((uint32_t)compact3(m >> 2), g.origin[2], g.h);
			p[3] = w[0];
			p[4] = w[COMPS > 1 ? 1 : 0];
			p[5] = w[COMPS > 2 ? 2 : 0];
			p[6] = size;
		} else {
			p[0] = node_coord((uint32_t)compact2(m), g.origin[0], g.h);
			p[1] = node_coord((uint32_t)compact2(m >> 1), g.origin[1], g.h);
			p[2] = w[0];
			p[3] = size;
		}
	});
	return (int)c.kept;
}

// CVTX_B200_TRACE=1: one stderr line per call with the wall time of each stage.
bool trace_on() {
	static const int on = [] { const char *e = std::getenv("CVTX_B200_TRACE"); return e && e[0] == '1' ? 1 : 0; }();
	return on != 0;
}

template <int D, class Particle>
int redistribute(const char *entry, const Particle **in, int n_in, Particle *out, int max_out, const cvtx_RedistFunc *rf, float h, float negligible) {
	constexpr int ROW = D == 3 ? 7 : 4;
	static_assert(sizeof(Particle) == sizeof(float) * ROW, "particle layout");
	if (n_in <= 0 || !in || !rf || !(h > 0.f)) return 0;
	NodeSet nodes;
	Grid g;
	const double t0 = omp_get_wtime();
	const int kind = builtin_kind(rf);
	const std::vector<int> devs = enabled_accelerators();
	if (kind >= 0 && !devs.empty()) {
		note_dispatch(1, 1);
		int created = 0;
		size_t n_nodes = 0;
		const int rc = device_redistribute_from_host(devs[0], D, kind, h, (const void *const *)in, n_in, negligible, out, max_out, &created, &n_nodes);
		if (rc == CVTX_B200_ERR_ARGUMENT) {
			// a limit of this implementation (grid or share count), not a device failure: say so and create nothing
			std::fprintf(stderr, "cvortex: %s: %s; no particles created.\n", entry, cvtx_b200_last_error());
			return 0;
		}
		if (rc != CVTX_B200_OK) gpu_failure(entry, rc);
		if (trace_on())
			std::fprintf(stderr, "cvortex trace: %s n=%d -> %zu nodes -> %d particles; %.3f ms (device: nodes + pruning)\n", entry, n_in, n_nodes,
			             created, (omp_get_wtime() - t0) * 1e3);
		return created;
	} else {
		note_dispatch(0, 0);
		std::vector<float> rows((size_t)n_in * ROW);
		uint32_t max_index = 0;
		const int half = kind >= 0 ? kHalfWidth[kind] : (int)roundf(rf->radius);
		g = place_grid(D, kind, half, h, (const void *const *)in, rows.data(), n_in, ROW, &max_index);
		if (code_bits(D, max_index) < 0) {
			std::fprintf(stderr, "cvortex: %s: grid too large for the node codes (more than 2^21 nodes per axis in 3-D, 2^31 in 2-D); no particles created.\n", entry);
			return 0;
		}
		const int bits = code_bits(D, max_index);
		const double stencil = D == 3 ? (2.0 * half + 1) * (2.0 * half + 1) * (2.0 * half + 1) : (2.0 * half + 1) * (2.0 * half + 1);
		if (too_many_shares(D, g, bits, n_in) || (kind < 0 && (double)n_in * stencil > 2147483647.0)) {
			std::fprintf(stderr, "cvortex: %s: %d particles on a sparse grid exceed the 2^31 shares one call can sort; no particles created.\n", entry, n_in);
			return 0;
		}
		if (kind < 0) host_nodes_user<D>(rows.data(), n_in, g, rf, &nodes);
		else if (dense_route(D, g, bits, n_in)) {
			switch (kind) {
			case K_LAMBDA0: host_nodes_dense<D, K_LAMBDA0>(rows.data(), n_in, g, bits, &nodes); break;
			case K_LAMBDA1: host_nodes_dense<D, K_LAMBDA1>(rows.data(), n_in, g, bits, &nodes); break;
			case K_LAMBDA2: host_nodes_dense<D, K_LAMBDA2>(rows.data(), n_in, g, bits, &nodes); break;
			case K_LAMBDA3: host_nodes_dense<D, K_LAMBDA3>(rows.data(), n_in, g, bits, &nodes); break;
			default: host_nodes_dense<D, K_M4P>(rows.data(), n_in, g, bits, &nodes); break;
			}
		}
		else host_nodes<D>(rows.data(), n_in, g, &nodes);
	}
	const double t1 = omp_get_wtime();
	const size_t n_nodes = nodes.code.size();
	const int kept = prune_and_write<D>(nodes, g, out, max_out, negligible);
	if (trace_on())
		std::fprintf(stderr, "cvortex trace: %s n=%d -> %zu nodes -> %d particles; nodes %.3f ms (host), prune+write %.3f ms\n", entry, n_in,
		             n_nodes, kept, (t1 - t0) * 1e3, (omp_get_wtime() - t1) * 1e3);
	return kept;
}

}  // namespace

extern "C" {

CVTX_API const cvtx_RedistFunc cvtx_RedistFunc_lambda0(void) { return builtin(K_LAMBDA0); }
CVTX_API const cvtx_RedistFunc cvtx_RedistFunc_lambda1(void) { return builtin(K_LAMBDA1); }
CVTX_API const cvtx_RedistFunc cvtx_RedistFunc_lambda2(void) { return builtin(K_LAMBDA2); }
CVTX_API const cvtx_RedistFunc cvtx_RedistFunc_lambda3(void) { return builtin(K_LAMBDA3); }
CVTX_API const cvtx_RedistFunc cvtx_RedistFunc_m4p(void) { return builtin(K_M4P); }

CVTX_API int cvtx_P3D_redistribute_on_grid(const cvtx_P3D **input_array_start, const int n_input_particles, cvtx_P3D *output_particles,
                                           int max_output_particles, const cvtx_RedistFunc *redistributor, float grid_density,
                                           float negligible_vort) {
	return redistribute<3>("cvtx_P3D_redistribute_on_grid", input_array_start, n_input_particles, output_particles,
	                       max_output_particles, redistributor, grid_density, negligible_vort);
}

CVTX_API int cvtx_P2D_redistribute_on_grid(const cvtx_P2D **input_array_start, const int n_input_particles, cvtx_P2D *output_particles,
                                           int max_output_particles, const cvtx_RedistFunc *redistributor, float grid_density,
                                           float negligible_vort) {
	return redistribute<2>("cvtx_P2D_redistribute_on_grid", input_array_start, n_input_particles, output_particles,
	                       max_output_particles, redistributor, grid_density, negligible_vort);
}

// alpha_new = (1 - f dt) alpha + f dt |alpha| omega(x)/|omega(x)|, omega = the vorticity field
// the particles themselves induce at x (reference src/P3D.cpp:667-707); a particle sitting in
// zero field loses its vorticity, as there.
CVTX_API void cvtx_P3D_pedrizzetti_relaxation(cvtx_P3D **input_array_start, const int n_input_particles, float fdt,
                                              const cvtx_VortFunc *kernel, float regularisation_radius) {
	const int n = n_input_particles;
	if (n <= 0) return;
	std::vector<bsv_V3f> where((size_t)n), field((size_t)n);
#pragma omp parallel for schedule(static)
	for (int i = 0; i < n; ++i) where[i] = input_array_start[i]->coord;
	cvtx_P3D_M2M_vort((const cvtx_P3D **)input_array_start, n, where.data(), n, field.data(), kernel, regularisation_radius);
	const float keep = 1.f - fdt;
#pragma omp parallel for schedule(static)
	for (int i = 0; i < n; ++i) {
		const bsv_V3f a = input_array_start[i]->vorticity, w = field[i];
		const float wn = bsv_V3f_abs(w);
		const float pull = bsv_V3f_abs(a) / wn * fdt;
		bsv_V3f r = bsv_V3f_plus(bsv_V3f_mult(a, keep), bsv_V3f_mult(w, pull));
		if (!(wn != 0.f)) r = bsv_V3f_zero();
		input_array_start[i]->vorticity = r;
	}
}

}  // extern "C"

// remesh_math.h -- the per-particle arithmetic of particle redistribution onto a
// regular grid, shared by the CUDA kernels (remesh_device.cu) and the host path
// (remesh.cpp) so that both produce the same bits.
//
// Follows the reference's cvtx_P3D_redistribute_on_grid / cvtx_P2D_redistribute_on_grid
// (src/P3D.cpp:509-634, src/P2D.cpp:283-405) and what they call:
//   * the five interpolants of src/RedistFunc.cpp:36-96 (Lambda_0..Lambda_3 and M4');
//   * nearest grid node / node position of src/UIntKey96.cpp:88-107,131-138 and
//     src/UIntKey64.cpp:79-96,116-122 (the 2-D key is formed in FP64 there);
//   * the output order: the reference flattens its oct/quadtree depth first with the
//     child index x + 2y (+ 4z) per bit level (src/GridParticleOcttree.cpp:137-180,
//     src/UIntKey96.h:209-219) -- i.e. ascending Morton code, x in the lowest bit.
//
// Every product and difference is a single correctly rounded FP32 operation in the
// reference's order (rm_mul / rm_sub keep nvcc from contracting them into FMAs), so a
// node receives bit-identical contributions on every path; only the order in which a
// node's contributions are summed is ours (FP64, ascending contribution index).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define RM_HD __host__ __device__ __forceinline__
#else
#define RM_HD inline
#endif

namespace cvtx {
namespace remesh {

enum Kind { K_LAMBDA0 = 0, K_LAMBDA1, K_LAMBDA2, K_LAMBDA3, K_M4P, K_COUNT };

// Support radii in cells (src/RedistFunc.cpp:43,55,67,80,93); the stencil half-width is
// (int)roundf(radius), src/P3D.cpp:542.
constexpr float kRadius[K_COUNT] = {0.5f, 1.0f, 1.5f, 2.0f, 2.0f};
constexpr int kHalfWidth[K_COUNT] = {1, 1, 2, 2, 2};
constexpr int kMaxHalfWidth = 2;
constexpr int kMaxStencil1D = 2 * kMaxHalfWidth + 1;

RM_HD float rm_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
	return __fmul_rn(a, b);
#else
	return a * b;
#endif
}
RM_HD float rm_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
	return __fsub_rn(a, b);
#else
	return a - b;
#endif
}
RM_HD float rm_add(float a, float b) {
#if defined(__CUDA_ARCH__)
	return __fadd_rn(a, b);
#else
	return a + b;
#endif
}

// The 1-D interpolant at distance U >= 0 (in cells).
RM_HD float weight(int kind, float U) {
	switch (kind) {
	case K_LAMBDA0:  // nearest grid point
		return U < 0.5f ? 1.f : 0.f;
	case K_LAMBDA1:  // linear
		return U <= 1.f ? rm_sub(1.f, U) : 0.f;
	case K_LAMBDA2:
		if (U < 0.5f) return rm_sub(1.f, rm_mul(U, U));
		if (U < 1.5f) return rm_mul(rm_mul(0.5f, rm_sub(1.f, U)), rm_sub(2.f, U));
		return 0.f;
	case K_LAMBDA3:
		if (U < 1.f) return rm_mul(rm_mul(0.5f, rm_sub(1.f, rm_mul(U, U))), rm_sub(2.f, U));
		if (U < 2.f) return rm_mul(rm_mul(rm_mul(1.f / 6.f, rm_sub(1.f, U)), rm_sub(2.f, U)), rm_sub(3.f, U));
		return 0.f;
	default:         // K_M4P
		if (U < 1.f) return rm_add(rm_sub(1.f, rm_mul(rm_mul(2.5f, U), U)), rm_mul(rm_mul(rm_mul(1.5f, U), U), U));
		if (U < 2.f) return rm_mul(rm_mul(rm_mul(0.5f, rm_sub(1.f, U)), rm_sub(2.f, U)), rm_sub(2.f, U));
		return 0.f;
	}
}

// The grid: node k sits at origin + h * k, per axis.
struct Grid {
	float h, rh;        // spacing and its FP32 reciprocal
	float origin[3];
	int kind, half;     // interpolant and stencil half-width
	uint32_t top[3];    // largest node index, per axis, that any particle's stencil touches (0 on unused axes)
};

// Nearest node index along one axis.
RM_HD uint32_t node_index_3d(float x, float origin, float rh) {   // src/UIntKey96.cpp:99-106
	return (uint32_t)roundf(rm_mul(rm_sub(x, origin), rh));
}
RM_HD uint32_t node_index_2d(float x, float origin, float rh) {   // src/UIntKey64.cpp:86-94: FP64, then roundf
#if defined(__CUDA_ARCH__)
	const double t = __dmul_rn(__dsub_rn((double)x, (double)origin), (double)rh);
#else
	const double t = ((double)x - (double)origin) * (double)rh;
#endif
	return (uint32_t)roundf((float)t);
}
RM_HD float node_coord(uint32_t k, float origin, float h) {       // src/UIntKey96.cpp:131-138
	return rm_add(origin, rm_mul(h, (float)k));
}
// Distance, in cells, from a particle coordinate to node k.
RM_HD float cell_distance(float x, uint32_t k, const Grid &g, int axis) {
	return fabsf(rm_mul(rm_sub(x, node_coord(k, g.origin[axis], g.h)), g.rh));
}

// A particle's cell = its nearest node.  Cells are numbered x fastest over the box
// [0, top]^D.  A node's shares are summed cell by cell in this order and, within a cell, in
// the order the caller gave the particles -- the one summation order every stage follows
// (device sort route, device dense route, host), which is what makes their results
// bit-identical.
template <int D>
RM_HD uint64_t cell_of(const float *row, const Grid &g) {
	uint64_t lin = 0;
	for (int a = D - 1; a >= 0; --a) {
		const uint32_t k = D == 3 ? node_index_3d(row[a], g.origin[a], g.rh) : node_index_2d(row[a], g.origin[a], g.rh);
		lin = lin * ((uint64_t)g.top[a] + 1) + k;
	}
	return lin;
}

// Bin of the strength-threshold histogram (reference src/redistribution_helper_funcs.cpp:56-61):
// t = floor(1023 (s - lo) / range) is converted with a C cast there, clamped to [0, 1023]
// afterwards.  Once the search has zoomed in, strong particles give t far beyond INT_MAX; on
// x86-64 the cast then yields INT_MIN and they are counted in bin 0, not in the top bin -- which
// steers the search, so it is reproduced here explicitly (also for NaN) on host and device alike.
constexpr int kCutBins = 1024;
RM_HD int cut_bin(double t) {
	if (!(t >= -2147483648.0 && t < 2147483648.0)) return 0;
	const int b = (int)t;
	return b < 0 ? 0 : (b >= kCutBins ? kCutBins - 1 : b);
}

// Morton codes: bit b of x lands at D*b, of y at D*b+1, of z at D*b+2.
constexpr int kBits3D = 21;   // 63-bit code
RM_HD uint64_t spread3(uint64_t v) {   // 21 bits -> every third bit
	v &= 0x1fffffull;
	v = (v | v << 32) & 0x1f00000000ffffull;
	v = (v | v << 16) & 0x1f0000ff0000ffull;
	v = (v | v << 8) & 0x100f00f00f00f00full;
	v = (v | v << 4) & 0x10c30c30c30c30c3ull;
	v = (v | v << 2) & 0x1249249249249249ull;
	return v;
}
RM_HD uint64_t compact3(uint64_t v) {
	v &= 0x1249249249249249ull;
	v = (v ^ (v >> 2)) & 0x10c30c30c30c30c3ull;
	v = (v ^ (v >> 4)) & 0x100f00f00f00f00full;
	v = (v ^ (v >> 8)) & 0x1f0000ff0000ffull;
	v = (v ^ (v >> 16)) & 0x1f00000000ffffull;
	v = (v ^ (v >> 32)) & 0x1fffffull;
	return v;
}
RM_HD uint64_t spread2(uint64_t v) {   // 32 bits -> every second bit
	v &= 0xffffffffull;
	v = (v | v << 16) & 0x0000ffff0000ffffull;
	v = (v | v << 8) & 0x00ff00ff00ff00ffull;
	v = (v | v << 4) & 0x0f0f0f0f0f0f0f0full;
	v = (v | v << 2) & 0x3333333333333333ull;
	v = (v | v << 1) & 0x5555555555555555ull;
	return v;
}
RM_HD uint64_t compact2(uint64_t v) {
	v &= 0x5555555555555555ull;
	v = (v ^ (v >> 1)) & 0x3333333333333333ull;
	v = (v ^ (v >> 2)) & 0x0f0f0f0f0f0f0f0full;
	v = (v ^ (v >> 4)) & 0x00ff00ff00ff00ffull;
	v = (v ^ (v >> 8)) & 0x0000ffff0000ffffull;
	v = (v ^ (v >> 16)) & 0xffffffffull;
	return v;
}
RM_HD uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) { return spread3(x) | spread3(y) << 1 | spread3(z) << 2; }
RM_HD uint64_t morton2(uint32_t x, uint32_t y) { return spread2(x) | spread2(y) << 1; }

// One particle's contributions, in the reference's stencil order (x offset outermost,
// src/UIntKey96.cpp:52-65).  `emit(code, s)` is called for every node that receives a
// non-zero share (all-zero shares are not inserted, src/GridParticleOcttree.cpp:89);
// returns how many were emitted.  D = 3: row = x y z wx wy wz vol, s has 3 components;
// D = 2: row = x y gamma area, s has 1.
template <int D, class Emit>
RM_HD int spread_particle(const float *row, const Grid &g, Emit &&emit) {
	const int R = g.half, S = 2 * R + 1;
	uint32_t k0[D];
	float w[D][kMaxStencil1D];
	for (int a = 0; a < D; ++a) {
		k0[a] = D == 3 ? node_index_3d(row[a], g.origin[a], g.rh) : node_index_2d(row[a], g.origin[a], g.rh);
		for (int o = 0; o < S; ++o) w[a][o] = weight(g.kind, cell_distance(row[a], k0[a] + (uint32_t)(o - R), g, a));
	}
	int n = 0;
	if (D == 3) {
		const float wx = row[3], wy = row[4], wz = row[5];
		for (int i = 0; i < S; ++i)
			for (int j = 0; j < S; ++j) {
				const float fij = rm_mul(w[0][i], w[1][j]);
				for (int k = 0; k < S; ++k) {
					const float f = rm_mul(fij, w[D - 1][k]);
					const float s[3] = {rm_mul(wx, f), rm_mul(wy, f), rm_mul(wz, f)};
					if (s[0] == 0.f && s[1] == 0.f && s[2] == 0.f) continue;
					emit(morton3(k0[0] + (uint32_t)(i - R), k0[1] + (uint32_t)(j - R), k0[D - 1] + (uint32_t)(k - R)), s);
					++n;
				}
			}
	} else {
		const float gam = row[2];
		for (int i = 0; i < S; ++i)
			for (int j = 0; j < S; ++j) {
				const float s[1] = {rm_mul(gam, rm_mul(w[0][i], w[1][j]))};
				if (s[0] == 0.f) continue;
				emit(morton2(k0[0] + (uint32_t)(i - R), k0[1] + (uint32_t)(j - R)), s);
				++n;
			}
	}
	return n;
}

}  // namespace remesh
}  // namespace cvtx

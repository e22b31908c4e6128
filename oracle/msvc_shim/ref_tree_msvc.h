/*
 * ref_tree_msvc.h -- TEST INFRASTRUCTURE ONLY (force-included by oracle/Makefile for two
 * reference translation units; never part of the product).
 *
 * The reference's grid trees (src/GridParticleOcttree.cpp, src/GridParticleQuadtree.cpp) rest on
 * UIntKey96/UIntKey64::matching_leading_bits (src/UIntKey96.h:239-266, src/UIntKey64.h:217-237),
 * whose __GNUC__ branch is an unfinished stub -- `assert(false); / * TO DO. * /` followed by
 * arithmetic that is not a leading-bit count -- so cvtx_P3D/P2D_redistribute_on_grid abort (or,
 * with NDEBUG, build a scrambled tree) under g++.  Only the _MSC_VER branch is functional.
 *
 * To run the reference's OWN redistribution as an oracle, those two units are compiled, from the
 * sources where they lie and unmodified, with the compiler presenting itself the way MSVC does:
 * every system header they use is read first (still as GNU), then __GNUC__ is hidden, _MSC_VER is
 * set and the one MSVC intrinsic the branch needs is supplied.  `unsigned long` is 32 bits on
 * MSVC and the reference passes the address of a uint32_t, so the stand-in stores 32 bits.
 */
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stack>
#include <string>
#include <tuple>
#include <vector>
#include <bsv/bsv.h>
#include <bsv/bsv_V2f.h>
#include <bsv/bsv_V3f.h>

static inline unsigned char _BitScanReverse(unsigned long *index, unsigned long mask) {
	const uint32_t m = (uint32_t)mask;
	if (m == 0) return 0;
	*(uint32_t *)index = 31u - (uint32_t)__builtin_clz(m);
	return 1;
}

#undef __GNUC__
#define _MSC_VER 1929

#ifndef CVTX_B200_BSV_COMPAT_H
#define CVTX_B200_BSV_COMPAT_H
/*
 * bsv/bsv.h -- ABI-compatible stand-in for the hjabird/bsv vector header.
 *
 * The cvortex public header does `#include <bsv/bsv.h>` (reference
 * include/cvortex/libcvtx.h:50) and the reference pulls the real package in
 * with find_package(bsv CONFIG REQUIRED) (reference CMakeLists.txt:69-71); it
 * is NOT vendored in the reference tree and no version is pinned there.  This
 * file was written from the way the reference *uses* bsv, not from bsv's
 * sources:
 *   - members are reached as `.x[i]`                (reference src/P3D.cpp:245)
 *   - values are brace-initialised `{a, b, c}`      (reference src/P3D.cpp:249)
 *   - `abs` is the Euclidean norm                   (reference src/P3D.cpp:64)
 *   - `isequal` is exact component-wise ==, so NaN != NaN
 *                                                   (reference src/F3D.cpp:81)
 * Layout contract relied on by the cvtx_* C ABI (and by Julia's ccall):
 *   sizeof(bsv_V3f) == 12, sizeof(bsv_V2f) == 8, alignment 4; the double
 *   variants are 24 / 16 bytes, alignment 8.
 *
 * Everything is `static inline`, plain C99 and valid C++.
 */
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float  x[3]; } bsv_V3f;
typedef struct { float  x[2]; } bsv_V2f;
typedef struct { double x[3]; } bsv_V3d;
typedef struct { double x[2]; } bsv_V2d;

/* ---- generators: one macro per arity keeps the four types consistent ---- */
#define BSV_COMPAT_DEFINE_3(V, S, SQRT)                                        \
	static inline V V##_zero(void) {                                           \
		V r; r.x[0] = 0; r.x[1] = 0; r.x[2] = 0; return r; }                   \
	static inline V V##_plus(const V a, const V b) {                           \
		V r; r.x[0] = a.x[0] + b.x[0]; r.x[1] = a.x[1] + b.x[1];               \
		r.x[2] = a.x[2] + b.x[2]; return r; }                                  \
	static inline V V##_minus(const V a, const V b) {                          \
		V r; r.x[0] = a.x[0] - b.x[0]; r.x[1] = a.x[1] - b.x[1];               \
		r.x[2] = a.x[2] - b.x[2]; return r; }                                  \
	static inline V V##_mult(const V a, const S s) {                           \
		V r; r.x[0] = a.x[0] * s; r.x[1] = a.x[1] * s; r.x[2] = a.x[2] * s;    \
		return r; }                                                            \
	static inline V V##_div(const V a, const S s) {                            \
		V r; r.x[0] = a.x[0] / s; r.x[1] = a.x[1] / s; r.x[2] = a.x[2] / s;    \
		return r; }                                                            \
	static inline S V##_dot(const V a, const V b) {                            \
		return a.x[0] * b.x[0] + a.x[1] * b.x[1] + a.x[2] * b.x[2]; }          \
	static inline V V##_cross(const V a, const V b) {                          \
		V r; r.x[0] = a.x[1] * b.x[2] - a.x[2] * b.x[1];                       \
		r.x[1] = a.x[2] * b.x[0] - a.x[0] * b.x[2];                            \
		r.x[2] = a.x[0] * b.x[1] - a.x[1] * b.x[0]; return r; }                \
	static inline S V##_abs(const V a) {                                       \
		return SQRT(a.x[0] * a.x[0] + a.x[1] * a.x[1] + a.x[2] * a.x[2]); }    \
	static inline int V##_isequal(const V a, const V b) {                      \
		return a.x[0] == b.x[0] && a.x[1] == b.x[1] && a.x[2] == b.x[2]; }

#define BSV_COMPAT_DEFINE_2(V, S, SQRT)                                        \
	static inline V V##_zero(void) {                                           \
		V r; r.x[0] = 0; r.x[1] = 0; return r; }                               \
	static inline V V##_plus(const V a, const V b) {                           \
		V r; r.x[0] = a.x[0] + b.x[0]; r.x[1] = a.x[1] + b.x[1]; return r; }   \
	static inline V V##_minus(const V a, const V b) {                          \
		V r; r.x[0] = a.x[0] - b.x[0]; r.x[1] = a.x[1] - b.x[1]; return r; }   \
	static inline V V##_mult(const V a, const S s) {                           \
		V r; r.x[0] = a.x[0] * s; r.x[1] = a.x[1] * s; return r; }             \
	static inline V V##_div(const V a, const S s) {                            \
		V r; r.x[0] = a.x[0] / s; r.x[1] = a.x[1] / s; return r; }             \
	static inline S V##_dot(const V a, const V b) {                            \
		return a.x[0] * b.x[0] + a.x[1] * b.x[1]; }                            \
	static inline S V##_abs(const V a) {                                       \
		return SQRT(a.x[0] * a.x[0] + a.x[1] * a.x[1]); }                      \
	static inline int V##_isequal(const V a, const V b) {                      \
		return a.x[0] == b.x[0] && a.x[1] == b.x[1]; }

BSV_COMPAT_DEFINE_3(bsv_V3f, float, sqrtf)
BSV_COMPAT_DEFINE_3(bsv_V3d, double, sqrt)
BSV_COMPAT_DEFINE_2(bsv_V2f, float, sqrtf)
BSV_COMPAT_DEFINE_2(bsv_V2d, double, sqrt)

#undef BSV_COMPAT_DEFINE_3
#undef BSV_COMPAT_DEFINE_2

/* ---- precision conversions ---- */
static inline bsv_V3d bsv_V3f_toV3d(const bsv_V3f a) {
	bsv_V3d r; r.x[0] = a.x[0]; r.x[1] = a.x[1]; r.x[2] = a.x[2]; return r; }
static inline bsv_V3f bsv_V3d_toV3f(const bsv_V3d a) {
	bsv_V3f r; r.x[0] = (float)a.x[0]; r.x[1] = (float)a.x[1];
	r.x[2] = (float)a.x[2]; return r; }
static inline bsv_V2d bsv_V2f_toV2d(const bsv_V2f a) {
	bsv_V2d r; r.x[0] = a.x[0]; r.x[1] = a.x[1]; return r; }
static inline bsv_V2f bsv_V2d_toV2f(const bsv_V2d a) {
	bsv_V2f r; r.x[0] = (float)a.x[0]; r.x[1] = (float)a.x[1]; return r; }

#ifdef __cplusplus
}
#endif
#endif /* CVTX_B200_BSV_COMPAT_H */

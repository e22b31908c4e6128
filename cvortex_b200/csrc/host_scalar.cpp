// host_scalar.cpp -- the scalar host entry points of the cvtx_* ABI:
// single pair (S2S), one-on-many (S2M), many-on-one (M2S), the dense filament
// influence matrix, and host_m2m_*(), the all-pairs loops that run ONLY when
// the caller asked for the host: every accelerator disabled through
// cvtx_accelerator_disable() (the reference's documented CPU/GPU switch,
// reference src/accelerators.cpp:107-118 and bench/benchP3D.c:165-167) or a
// user-defined cvtx_VortFunc whose maths exists only as host function
// pointers (empty cl_kernel_name_ext, reference src/P3D.cpp:355).  They are
// never a fallback for a failed GPU call -- host_api.cu aborts on those.
//
// Semantics follow the reference's CPU path: FP32 pair arithmetic through the
// cvtx_VortFunc pointers, one double accumulator per output component over the
// sources, coincident points contribute zero
//   3D particles src/P3D.cpp:51-322, 2D src/P2D.cpp:49-250, filaments src/F3D.cpp:34-227.
// The pair formulas are written once, as small templates over a tiny vector
// type, and every S2M / M2S / M2M wrapper is generated from them.
#include <cmath>
#include <cassert>
#include "export.h"
#include "host_scalar.h"
#include "host_hooks.h"
#include "op_table.h"

namespace {

constexpr float kPi = 3.14159265359f;          // CVTX_PI_F (reference src/P3D.cpp:47)

struct V3 { float x, y, z; };
inline V3 ld(const bsv_V3f &v) { return {v.x[0], v.x[1], v.x[2]}; }
inline bsv_V3f st(V3 v) { bsv_V3f r; r.x[0] = v.x; r.x[1] = v.y; r.x[2] = v.z; return r; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline bool same(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool has_nan(V3 a) { return a.x != a.x || a.y != a.y || a.z != a.z; }
const V3 kZero = {0.f, 0.f, 0.f};

// ---- pair formulas --------------------------------------------------------
// velocity of particle p at x, without the 1/4pi (src/P3D.cpp:51-72)
inline V3 p3d_vel(const cvtx_P3D *p, V3 x, const cvtx_VortFunc *k, float recip_sigma) {
	const V3 c = ld(p->coord);
	if (same(c, x)) return kZero;
	const V3 rad = x - c;
	const float r = norm(rad);
	const float scale = -k->g_3D(r * recip_sigma) * std::pow(r, -3.f);
	return cross(rad, ld(p->vorticity)) * scale;
}

// vortex stretching of q by p (src/P3D.cpp:86-114)
inline V3 p3d_dvort(const cvtx_P3D *p, const cvtx_P3D *q, const cvtx_VortFunc *k, float sigma) {
	const V3 cp = ld(p->coord), cq = ld(q->coord);
	if (same(cp, cq)) return kZero;
	const V3 rad = cq - cp;
	const float r = norm(rad), rho = std::fabs(r / sigma);
	float g, zeta;
	k->combined_3D(rho, &g, &zeta);
	const V3 c = cross(ld(q->vorticity), ld(p->vorticity));
	const float rho3 = rho * rho * rho;
	const V3 first = (c * g) / rho3;
	const V3 second = rad * ((-1.f / (r * r)) * ((3 * g) / rho3 - zeta) * dot(rad, c));
	return (first + second) * (1.f / (4.f * kPi * std::pow(sigma, 3.f)));
}

// viscous exchange on q from p (src/P3D.cpp:116-144)
inline V3 p3d_visc(const cvtx_P3D *p, const cvtx_P3D *q, const cvtx_VortFunc *k, float sigma, float nu) {
	assert(k->eta_3D != nullptr);
	const V3 cp = ld(p->coord), cq = ld(q->coord);
	if (same(cp, cq)) return kZero;
	const float rho = std::fabs(norm(cp - cq) / sigma);
	const V3 diff = ld(p->vorticity) * q->volume + ld(q->vorticity) * (-1 * p->volume);
	return (diff * k->eta_3D(rho)) * (2 * nu / std::pow(sigma, 2.f));
}

// vorticity field of p at x (src/P3D.cpp:146-161)
inline V3 p3d_vort(const cvtx_P3D *p, V3 x, const cvtx_VortFunc *k, float sigma) {
	const float r = norm(ld(p->coord) - x);
	const float coeff = k->zeta_3D(r / sigma) / (4.f * kPi * sigma * sigma * sigma);
	return ld(p->vorticity) * coeff;
}

// filament on a point (src/F3D.cpp:34-54)
inline V3 f3d_vel(const cvtx_F3D *f, V3 x) {
	const V3 r1 = x - ld(f->start), r2 = x - ld(f->end), r0 = r1 - r2, c = cross(r1, r2);
	const float t1 = f->strength / (4 * kPi * std::pow(norm(c), 2.f));
	const float t2 = dot(r1, r0) / norm(r1) - dot(r2, r0) / norm(r2);
	const float big = 3.40282346e38f;
	return (std::fabs(t1) <= big && std::fabs(t2) <= big) ? c * (t1 * t2) : kZero;
}

// filament on a particle's vorticity (src/F3D.cpp:56-85)
inline V3 f3d_dvort(const cvtx_F3D *f, const cvtx_P3D *q) {
	const V3 x = ld(q->coord), w = ld(q->vorticity);
	const V3 r1 = x - ld(f->start), r2 = x - ld(f->end), r0 = r1 - r2;
	const float t1 = f->strength / (4 * kPi);
	const V3 t211 = r0 / (-std::pow(norm(cross(r1, r0)), 2.f));
	const float t212 = dot(r0, r1) / norm(r1) + (-dot(r0, r2) / norm(r2));
	const float t221 = 3.0f / norm(r0);
	const float nx = norm(cross(r0, r1));
	const float t222 = nx / norm(r1) + (-nx / norm(r2));
	const V3 A = t211 * (t1 * t212);
	const float B = t221 * t1 * t222;
	const V3 ret = w * B + cross(A, w);
	return (has_nan(ret) || t222 != t222 || t212 != t212) ? kZero : ret;
}

// 2D velocity without the 1/2pi (src/P2D.cpp:49-69)
inline void p2d_vel(const cvtx_P2D *p, bsv_V2f x, const cvtx_VortFunc *k, float recip_sigma, float *ux, float *uy) {
	if (p->coord.x[0] == x.x[0] && p->coord.x[1] == x.x[1]) { *ux = 0.f; *uy = 0.f; return; }
	const float rx = x.x[0] - p->coord.x[0], ry = x.x[1] - p->coord.x[1];
	const float r = std::sqrt(rx * rx + ry * ry);
	const float g = k->g_2D(r * recip_sigma);
	*ux = ry * p->vorticity * g / (r * r);
	*uy = -rx * p->vorticity * g / (r * r);
}

// 2D viscous exchange (src/P2D.cpp:167-194)
inline float p2d_visc(const cvtx_P2D *p, const cvtx_P2D *q, const cvtx_VortFunc *k, float sigma, float nu) {
	assert(k->eta_2D != nullptr);
	if (p->coord.x[0] == q->coord.x[0] && p->coord.x[1] == q->coord.x[1]) return 0.f;
	const float rx = p->coord.x[0] - q->coord.x[0], ry = p->coord.x[1] - q->coord.x[1];
	const float rho = std::fabs(std::sqrt(rx * rx + ry * ry) / sigma);
	const float diff = p->vorticity * q->area + (-q->vorticity * p->area);
	return diff * k->eta_2D(rho) * (2 * nu / std::pow(sigma, 2.f));
}

// ---- many-on-one sums: FP32 terms into double accumulators -----------------
struct Sum3 {
	double x = 0, y = 0, z = 0;
	void add(V3 v) { x += v.x; y += v.y; z += v.z; }
	V3 value() const { return {(float)x, (float)y, (float)z}; }
};

inline V3 m2s_p3d_vel(const cvtx_P3D **a, int n, V3 x, const cvtx_VortFunc *k, float sigma) {
	const float rs = 1.f / std::fabs(sigma);
	Sum3 s;
	for (long i = 0; i < n; ++i) s.add(p3d_vel(a[i], x, k, rs));
	return s.value() * (1.f / (4.f * kPi));
}
inline V3 m2s_p3d_dvort(const cvtx_P3D **a, int n, const cvtx_P3D *q, const cvtx_VortFunc *k, float sigma) {
	Sum3 s;
	for (long i = 0; i < n; ++i) s.add(p3d_dvort(a[i], q, k, sigma));
	return s.value();
}
inline V3 m2s_p3d_visc(const cvtx_P3D **a, int n, const cvtx_P3D *q, const cvtx_VortFunc *k, float sigma, float nu) {
	Sum3 s;
	for (long i = 0; i < n; ++i) s.add(p3d_visc(a[i], q, k, sigma, nu));
	return s.value();
}
// FP32 running sum restricted to the 5-sigma box (src/P3D.cpp:298-322)
inline V3 m2s_p3d_vort(const cvtx_P3D **a, int n, V3 x, const cvtx_VortFunc *k, float sigma) {
	const float cutoff = 5.f * sigma, rs = 1 / sigma;
	V3 sum = kZero;
	for (long i = 0; i < n; ++i) {
		const V3 rad = ld(a[i]->coord) - x;
		if (std::fabs(rad.x) < cutoff && std::fabs(rad.y) < cutoff && std::fabs(rad.z) < cutoff)
			sum = ld(a[i]->vorticity) * k->zeta_3D(norm(rad) * rs) + sum;
	}
	return sum / (4.f * kPi * sigma * sigma * sigma);
}
inline V3 m2s_f3d_vel(const cvtx_F3D **a, int n, V3 x) {
	Sum3 s;
	for (long i = 0; i < n; ++i) s.add(f3d_vel(a[i], x));
	return s.value();
}
inline V3 m2s_f3d_dvort(const cvtx_F3D **a, int n, const cvtx_P3D *q) {
	Sum3 s;
	for (long i = 0; i < n; ++i) s.add(f3d_dvort(a[i], q));
	return s.value();
}
inline bsv_V2f m2s_p2d_vel(const cvtx_P2D **a, int n, bsv_V2f x, const cvtx_VortFunc *k, float sigma) {
	const float rs = 1.f / std::fabs(sigma);
	double sx = 0, sy = 0;
	for (long i = 0; i < n; ++i) {
		float ux, uy;
		p2d_vel(a[i], x, k, rs, &ux, &uy);
		sx += ux; sy += uy;
	}
	const float scale = 1.f / (2.f * std::acos(-1.f));
	bsv_V2f r; r.x[0] = (float)sx * scale; r.x[1] = (float)sy * scale;
	return r;
}
inline float m2s_p2d_visc(const cvtx_P2D **a, int n, const cvtx_P2D *q, const cvtx_VortFunc *k, float sigma, float nu) {
	double s = 0;
	for (long i = 0; i < n; ++i) s += (double)p2d_visc(a[i], q, k, sigma, nu);
	return (float)s;
}

}  // namespace

// ---- all-pairs on the host, only on explicit request (see file header) ------
namespace cvtx {
void host_m2m_p3d_vel(const cvtx_P3D **a, int n, const bsv_V3f *x, int m, bsv_V3f *out, const cvtx_VortFunc *k, float s) {
#pragma omp parallel for schedule(static)
	for (long i = 0; i < m; ++i) out[i] = st(m2s_p3d_vel(a, n, ld(x[i]), k, s));
}
void host_m2m_p3d_dvort(const cvtx_P3D **a, int n, const cvtx_P3D **q, int m, bsv_V3f *out, const cvtx_VortFunc *k, float s) {
#pragma omp parallel for schedule(static)
	for (long i = 0; i < m; ++i) out[i] = st(m2s_p3d_dvort(a, n, q[i], k, s));
}
void host_m2m_p3d_visc(const cvtx_P3D **a, int n, const cvtx_P3D **q, int m, bsv_V3f *out, const cvtx_VortFunc *k, float s, float nu) {
#pragma omp parallel for schedule(static)
	for (long i = 0; i < m; ++i) out[i] = st(m2s_p3d_visc(a, n, q[i], k, s, nu));
}
void host_m2m_p3d_vort(const cvtx_P3D **a, int n, const bsv_V3f *x, int m, bsv_V3f *out, const cvtx_VortFunc *k, float s) {
#pragma omp parallel for schedule(guided)
	for (long i = 0; i < m; ++i) out[i] = st(m2s_p3d_vort(a, n, ld(x[i]), k, s));
}
void host_m2m_p2d_vel(const cvtx_P2D **a, int n, const bsv_V2f *x, int m, bsv_V2f *out, const cvtx_VortFunc *k, float s) {
#pragma omp parallel for schedule(static)
	for (long i = 0; i < m; ++i) out[i] = m2s_p2d_vel(a, n, x[i], k, s);
}
void host_m2m_p2d_visc(const cvtx_P2D **a, int n, const cvtx_P2D **q, int m, float *out, const cvtx_VortFunc *k, float s, float nu) {
#pragma omp parallel for schedule(static)
	for (long i = 0; i < m; ++i) out[i] = m2s_p2d_visc(a, n, q[i], k, s, nu);
}
void host_m2m_f3d_vel(const cvtx_F3D **a, int n, const bsv_V3f *x, int m, bsv_V3f *out) {
#pragma omp parallel for schedule(static)
	for (long i = 0; i < m; ++i) out[i] = st(m2s_f3d_vel(a, n, ld(x[i])));
}
void host_m2m_f3d_dvort(const cvtx_F3D **a, int n, const cvtx_P3D **q, int m, bsv_V3f *out) {
#pragma omp parallel for schedule(static)
	for (long i = 0; i < m; ++i) out[i] = st(m2s_f3d_dvort(a, n, q[i]));
}
/* result_matrix[i * num_filaments + j] = u_j(x_i) . dir_i  (reference src/F3D.cpp:204-227) */
void host_f3d_inf_mtrx(const cvtx_F3D **a, int n, const bsv_V3f *x, const bsv_V3f *dir, int m, float *out) {
#pragma omp parallel for schedule(static)
	for (int i = 0; i < m; ++i)
		for (int j = 0; j < n; ++j)
			out[(long)i * n + j] = dot(f3d_vel(a[j], ld(x[i])), ld(dir[i]));
}
}  // namespace cvtx

// ---- exported scalar entry points -------------------------------------------
using cvtx::gpu_m2s;
using cvtx::note_dispatch;
using cvtx::OP_P3D_VEL; using cvtx::OP_P3D_DVORT; using cvtx::OP_P3D_VISC; using cvtx::OP_P3D_VORT;
using cvtx::OP_P2D_VEL; using cvtx::OP_P2D_VISC; using cvtx::OP_F3D_VEL; using cvtx::OP_F3D_DVORT;

extern "C" {

CVTX_API bsv_V3f cvtx_P3D_S2S_vel(const cvtx_P3D *self, const bsv_V3f mes_point, const cvtx_VortFunc *kernel, float regularisation_radius) {
	return st(p3d_vel(self, ld(mes_point), kernel, 1.f / std::fabs(regularisation_radius)) * (1.f / (4.f * kPi)));
}
CVTX_API bsv_V3f cvtx_P3D_S2S_dvort(const cvtx_P3D *self, const cvtx_P3D *induced_particle, const cvtx_VortFunc *kernel, float regularisation_radius) {
	return st(p3d_dvort(self, induced_particle, kernel, regularisation_radius));
}
CVTX_API bsv_V3f cvtx_P3D_S2S_visc_dvort(const cvtx_P3D *self, const cvtx_P3D *induced_particle, const cvtx_VortFunc *kernel, float regularisation_radius, float kinematic_visc) {
	return st(p3d_visc(self, induced_particle, kernel, regularisation_radius, kinematic_visc));
}
CVTX_API bsv_V3f cvtx_P3D_S2S_vort(const cvtx_P3D *self, const bsv_V3f mes_point, const cvtx_VortFunc *kernel, float regularisation_radius) {
	return st(p3d_vort(self, ld(mes_point), kernel, regularisation_radius));
}

CVTX_API void cvtx_P3D_S2M_vel(const cvtx_P3D *self, const bsv_V3f *mes_start, const int num_mes, bsv_V3f *result_array, const cvtx_VortFunc *kernel, float regularisation_radius) {
#pragma omp parallel for
	for (int i = 0; i < num_mes; ++i) result_array[i] = cvtx_P3D_S2S_vel(self, mes_start[i], kernel, regularisation_radius);
}
CVTX_API void cvtx_P3D_S2M_dvort(const cvtx_P3D *self, const cvtx_P3D **induced_start, const int num_induced, bsv_V3f *result_array, const cvtx_VortFunc *kernel, float regularisation_radius) {
#pragma omp parallel for
	for (int i = 0; i < num_induced; ++i) result_array[i] = cvtx_P3D_S2S_dvort(self, induced_start[i], kernel, regularisation_radius);
}
CVTX_API void cvtx_P3D_S2M_visc_dvort(const cvtx_P3D *self, const cvtx_P3D **induced_start, const int num_induced, bsv_V3f *result_array, const cvtx_VortFunc *kernel, float regularisation_radius, float kinematic_visc) {
#pragma omp parallel for
	for (int i = 0; i < num_induced; ++i) result_array[i] = cvtx_P3D_S2S_visc_dvort(self, induced_start[i], kernel, regularisation_radius, kinematic_visc);
}
CVTX_API void cvtx_P3D_S2M_vort(const cvtx_P3D *self, const bsv_V3f *mes_start, const int num_mes, bsv_V3f *result_array, const cvtx_VortFunc *kernel, float regularisation_radius) {
#pragma omp parallel for
	for (int i = 0; i < num_mes; ++i) result_array[i] = cvtx_P3D_S2S_vort(self, mes_start[i], kernel, regularisation_radius);
}

CVTX_API bsv_V3f cvtx_P3D_M2S_vel(const cvtx_P3D **array_start, const int num_particles, const bsv_V3f mes_point, const cvtx_VortFunc *kernel, float regularisation_radius) {
	bsv_V3f r;
	if (gpu_m2s("cvtx_P3D_M2S_vel", OP_P3D_VEL, kernel->cl_kernel_name_ext, (const void *const *)array_start, num_particles, &mes_point, r.x, regularisation_radius, 0.f)) return r;
	note_dispatch(0, 0);
	return st(m2s_p3d_vel(array_start, num_particles, ld(mes_point), kernel, regularisation_radius));
}
CVTX_API bsv_V3f cvtx_P3D_M2S_dvort(const cvtx_P3D **array_start, const int num_particles, const cvtx_P3D *induced_particle, const cvtx_VortFunc *kernel, float regularisation_radius) {
	bsv_V3f r;
	if (gpu_m2s("cvtx_P3D_M2S_dvort", OP_P3D_DVORT, kernel->cl_kernel_name_ext, (const void *const *)array_start, num_particles, induced_particle, r.x, regularisation_radius, 0.f)) return r;
	note_dispatch(0, 0);
	return st(m2s_p3d_dvort(array_start, num_particles, induced_particle, kernel, regularisation_radius));
}
CVTX_API bsv_V3f cvtx_P3D_M2S_visc_dvort(const cvtx_P3D **array_start, const int num_particles, const cvtx_P3D *induced_particle, const cvtx_VortFunc *kernel, float regularisation_radius, float kinematic_visc) {
	bsv_V3f r;
	if (gpu_m2s("cvtx_P3D_M2S_visc_dvort", OP_P3D_VISC, kernel->cl_kernel_name_ext, (const void *const *)array_start, num_particles, induced_particle, r.x, regularisation_radius, kinematic_visc)) return r;
	note_dispatch(0, 0);
	return st(m2s_p3d_visc(array_start, num_particles, induced_particle, kernel, regularisation_radius, kinematic_visc));
}
CVTX_API bsv_V3f cvtx_P3D_M2S_vort(const cvtx_P3D **array_start, const int num_particles, const bsv_V3f mes_point, const cvtx_VortFunc *kernel, float regularisation_radius) {
	bsv_V3f r;
	if (gpu_m2s("cvtx_P3D_M2S_vort", OP_P3D_VORT, kernel->cl_kernel_name_ext, (const void *const *)array_start, num_particles, &mes_point, r.x, regularisation_radius, 0.f)) return r;
	note_dispatch(0, 0);
	return st(m2s_p3d_vort(array_start, num_particles, ld(mes_point), kernel, regularisation_radius));
}

CVTX_API bsv_V3f cvtx_F3D_S2S_vel(const cvtx_F3D *self, const bsv_V3f mes_point) { return st(f3d_vel(self, ld(mes_point))); }
CVTX_API bsv_V3f cvtx_F3D_S2S_dvort(const cvtx_F3D *self, const cvtx_P3D *induced_particle) { return st(f3d_dvort(self, induced_particle)); }
CVTX_API bsv_V3f cvtx_F3D_M2S_vel(const cvtx_F3D **array_start, const int num_filaments, const bsv_V3f mes_point) {
	bsv_V3f r;
	if (gpu_m2s("cvtx_F3D_M2S_vel", OP_F3D_VEL, nullptr, (const void *const *)array_start, num_filaments, &mes_point, r.x, 0.f, 0.f)) return r;
	note_dispatch(0, 0);
	return st(m2s_f3d_vel(array_start, num_filaments, ld(mes_point)));
}
CVTX_API bsv_V3f cvtx_F3D_M2S_dvort(const cvtx_F3D **array_start, const int num_filaments, const cvtx_P3D *induced_particle) {
	bsv_V3f r;
	if (gpu_m2s("cvtx_F3D_M2S_dvort", OP_F3D_DVORT, nullptr, (const void *const *)array_start, num_filaments, induced_particle, r.x, 0.f, 0.f)) return r;
	note_dispatch(0, 0);
	return st(m2s_f3d_dvort(array_start, num_filaments, induced_particle));
}
CVTX_API bsv_V2f cvtx_P2D_S2S_vel(const cvtx_P2D *self, const bsv_V2f mes_point, const cvtx_VortFunc *kernel, float regularisation_radius) {
	float ux, uy;
	p2d_vel(self, mes_point, kernel, 1.f / std::fabs(regularisation_radius), &ux, &uy);
	const float scale = 1.f / (2.f * std::acos(-1.f));
	bsv_V2f r; r.x[0] = ux * scale; r.x[1] = uy * scale;
	return r;
}
CVTX_API void cvtx_P2D_S2M_vel(const cvtx_P2D *self, const bsv_V2f *mes_start, const int num_mes, bsv_V2f *result_array, const cvtx_VortFunc *kernel, float regularisation_radius) {
#pragma omp parallel for
	for (int i = 0; i < num_mes; ++i) result_array[i] = cvtx_P2D_S2S_vel(self, mes_start[i], kernel, regularisation_radius);
}
CVTX_API bsv_V2f cvtx_P2D_M2S_vel(const cvtx_P2D **array_start, const int num_particles, const bsv_V2f mes_point, const cvtx_VortFunc *kernel, float regularisation_radius) {
	bsv_V2f r;
	if (gpu_m2s("cvtx_P2D_M2S_vel", OP_P2D_VEL, kernel->cl_kernel_name_ext, (const void *const *)array_start, num_particles, &mes_point, r.x, regularisation_radius, 0.f)) return r;
	note_dispatch(0, 0);
	return m2s_p2d_vel(array_start, num_particles, mes_point, kernel, regularisation_radius);
}
CVTX_API float cvtx_P2D_S2S_visc_dvort(const cvtx_P2D *self, const cvtx_P2D *induced_particle, const cvtx_VortFunc *kernel, float regularisation_radius, float kinematic_visc) {
	return p2d_visc(self, induced_particle, kernel, regularisation_radius, kinematic_visc);
}
CVTX_API void cvtx_P2D_S2M_visc_dvort(const cvtx_P2D *self, const cvtx_P2D **induced_start, const int num_induced, float *result_array, const cvtx_VortFunc *kernel, float regularisation_radius, float kinematic_visc) {
#pragma omp parallel for
	for (int i = 0; i < num_induced; ++i) result_array[i] = cvtx_P2D_S2S_visc_dvort(self, induced_start[i], kernel, regularisation_radius, kinematic_visc);
}
CVTX_API float cvtx_P2D_M2S_visc_dvort(const cvtx_P2D **array_start, const int num_particles, const cvtx_P2D *induced_particle, const cvtx_VortFunc *kernel, float regularisation_radius, float kinematic_visc) {
	float r;
	if (gpu_m2s("cvtx_P2D_M2S_visc_dvort", OP_P2D_VISC, kernel->cl_kernel_name_ext, (const void *const *)array_start, num_particles, induced_particle, &r, regularisation_radius, kinematic_visc)) return r;
	note_dispatch(0, 0);
	return m2s_p2d_visc(array_start, num_particles, induced_particle, kernel, regularisation_radius, kinematic_visc);
}

}  // extern "C"

// vortfunc.cpp -- the four built-in regularisations as host function tables.
//
// ABI: cvtx_VortFunc_{singular,winckelmans,planetary,gaussian}() return the
// 80-byte struct by value (reference include/cvortex/libcvtx.h:86-94,113-116;
// constructors at reference src/VortFunc.cpp:201-251).  The function pointers
// serve the scalar host entry points (single pair, one-to-many, many-to-one)
// and user code that calls them directly; the all-pairs GPU kernels never call
// them -- they key on cl_kernel_name_ext and use the compile-time policies of
// pair_math.cuh.  Formulas: reference src/VortFunc.cpp:64-199, restated here
// per regularisation as one small struct of static functions.
#include <cmath>
#include <cstdio>
#include <cstring>
#include "../../include/cvortex/libcvtx.h"
#include "export.h"

namespace {

constexpr float kSqrt2OverPi = 0.7978845608028654f;
constexpr float kRecipSqrt2 = 0.7071067811865475f;

// A regularisation without a viscous kernel: say so once, then act inviscid
// (what the reference's warn_bad_eta_fn does, src/VortFunc.cpp:52-62).
float no_eta(float) {
	static bool told = false;
	if (!told) {
		told = true;
		std::fprintf(stderr, "cvortex: this regularisation has no eta function; "
		                     "viscous terms are taken as zero.\n");
	}
	return 0.f;
}

struct Singular {
	static float g3(float) { return 1.f; }
	static float g2(float) { return 1.f; }
	static float zeta3(float) { return 0.f; }
};

struct Winckelmans {
	static float g3(float rho) {                       // (rho^2 + 5/2) rho^3 (rho^2+1)^-5/2
		const float q = rho * rho;
		return (q + 2.5f) * rho * q * std::pow(q + 1.f, -2.5f);
	}
	static float g2(float rho) {                       // (rho^4 + 2 rho^2)/(rho^2+1)^2
		const float q = rho * rho;
		return (q * q + q * 2.f) / (q * q + 2.f * q + 1.f);
	}
	static float zeta3(float rho) { return 7.5f * std::pow(rho * rho + 1.f, -3.5f); }
	static float eta3(float rho) { return 52.5f * std::pow(rho * rho + 1.f, -4.5f); }
	static float eta2(float rho) {                     // 24 exp(4/a^3)/a^4 as the reference codes it
		const float a = rho * rho + 1.f;
		const float ia4 = (1.f / (a * a)) * (1.f / (a * a));
		return 24.f * std::exp(4.f * a * ia4) * ia4;
	}
};

struct Planetary {
	static float g3(float rho) { return rho < 1.f ? rho * rho * rho : 1.f; }
	static float g2(float rho) { return rho < 1.f ? rho * rho : 1.f; }
	static float zeta3(float rho) { return rho < 1.f ? 3.f : 0.f; }
};

struct Gaussian {
	static float g3(float rho) {                       // erf(rho/sqrt2) - rho sqrt(2/pi) exp(-rho^2/2)
		if (rho > 6.f) return 1.f;
		// erf by Abramowitz & Stegun 7.1.26, the reference's choice
		static const float a[5] = {0.254829592f, -0.284496736f, 1.421413741f, -1.453152027f, 1.061405429f};
		const float z = rho * kRecipSqrt2;
		const float t = 1.f / (1.f + 0.3275911f * z);
		const float t2 = t * t, t3 = t2 * t, t4 = t2 * t2, t5 = t3 * t2;
		const float e = std::exp(-z * z);
		const float erf_as = 1.f - (a[0] * t + a[1] * t2 + a[2] * t3 + a[3] * t4 + a[4] * t5) * e;
		return erf_as - rho * kSqrt2OverPi * e;
	}
	static float g2(float rho) { return 1.f - std::exp(-rho * rho * 0.5f); }
	static float zeta3(float rho) { return kSqrt2OverPi * std::exp(-rho * rho * 0.5f); }
	static float eta2(float rho) { return std::exp(-rho * rho * 0.5f); }
};

template <class R> void combined(float rho, float *g, float *zeta) { *g = R::g3(rho); *zeta = R::zeta3(rho); }

template <class R>
cvtx_VortFunc table(const char *key, float (*eta3)(float), float (*eta2)(float)) {
	cvtx_VortFunc vf;
	std::memset(&vf, 0, sizeof(vf));
	vf.g_3D = &R::g3;
	vf.g_2D = &R::g2;
	vf.zeta_3D = &R::zeta3;
	vf.combined_3D = &combined<R>;
	vf.eta_3D = eta3;
	vf.eta_2D = eta2;
	std::strncpy(vf.cl_kernel_name_ext, key, sizeof(vf.cl_kernel_name_ext) - 1);
	return vf;
}

}  // namespace

extern "C" {

CVTX_API const cvtx_VortFunc cvtx_VortFunc_singular(void) { return table<Singular>("singular", &no_eta, &no_eta); }
CVTX_API const cvtx_VortFunc cvtx_VortFunc_winckelmans(void) {
	return table<Winckelmans>("winckelmans", &Winckelmans::eta3, &Winckelmans::eta2);
}
CVTX_API const cvtx_VortFunc cvtx_VortFunc_planetary(void) { return table<Planetary>("planetary", &no_eta, &no_eta); }
/* eta_3D of the Gaussian is its zeta (reference src/VortFunc.cpp:245). */
CVTX_API const cvtx_VortFunc cvtx_VortFunc_gaussian(void) {
	return table<Gaussian>("gaussian", &Gaussian::zeta3, &Gaussian::eta2);
}

}  // extern "C"

// out_of_scope.cpp -- cvtx_* symbols outside the all-pairs hot path.
//
// The particle redistribution / relaxation subsystem of the reference
// (src/P3D.cpp:509-707, src/P2D.cpp:283-436, src/RedistFunc.cpp, the grid
// oct/quadtrees and keys) is CPU-only there, is not an all-pairs sum, and is
// marked OUT OF SCOPE for this backend (SURVEY.md section 2 rows 14-18 and
// section 8f rank 4).  The symbols are kept so that programs written against
// libcvtx.h still link; calling one says what is missing and aborts instead of
// returning something wrong.
#include <cstdio>
#include <cstdlib>
#include "export.h"

namespace {
[[noreturn]] void not_in_this_build(const char *what) {
	std::fprintf(stderr,
	             "cvortex (B200 build): %s is outside the all-pairs hot path this library implements; "
	             "link the reference CPU library for redistribution / relaxation.\n", what);
	std::abort();
}
float redist_unavailable(float) { not_in_this_build("cvtx_RedistFunc::func"); }
cvtx_RedistFunc redist(float radius) { cvtx_RedistFunc r; r.func = &redist_unavailable; r.radius = radius; return r; }
}  // namespace

extern "C" {
/* radii as in reference src/RedistFunc.cpp (0.5, 1, 1.5, 2, 2) */
CVTX_API const cvtx_RedistFunc cvtx_RedistFunc_lambda0(void) { return redist(0.5f); }
CVTX_API const cvtx_RedistFunc cvtx_RedistFunc_lambda1(void) { return redist(1.f); }
CVTX_API const cvtx_RedistFunc cvtx_RedistFunc_lambda2(void) { return redist(1.5f); }
CVTX_API const cvtx_RedistFunc cvtx_RedistFunc_lambda3(void) { return redist(2.f); }
CVTX_API const cvtx_RedistFunc cvtx_RedistFunc_m4p(void) { return redist(2.f); }

CVTX_API int cvtx_P3D_redistribute_on_grid(const cvtx_P3D **, const int, cvtx_P3D *, int, const cvtx_RedistFunc *, float, float) {
	not_in_this_build("cvtx_P3D_redistribute_on_grid");
}
CVTX_API void cvtx_P3D_pedrizzetti_relaxation(cvtx_P3D **, const int, float, const cvtx_VortFunc *, float) {
	not_in_this_build("cvtx_P3D_pedrizzetti_relaxation");
}
CVTX_API int cvtx_P2D_redistribute_on_grid(const cvtx_P2D **, const int, cvtx_P2D *, int, const cvtx_RedistFunc *, float, float) {
	not_in_this_build("cvtx_P2D_redistribute_on_grid");
}
}  // extern "C"

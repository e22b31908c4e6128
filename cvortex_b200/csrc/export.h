// export.h -- symbol visibility and the ABI layout checks of libcvortex.so.
// The library is built with -fvisibility=hidden; exactly the 52 cvtx_* symbols
// of include/cvortex/libcvtx.h and the cvtx_b200_* symbols of
// include/cvtx_b200.h are exported, unmangled.
#pragma once
#include <cstddef>
#include "../../include/cvortex/libcvtx.h"

#define CVTX_API __attribute__((visibility("default")))

// Layouts existing callers (C, Julia ccall) depend on -- SURVEY.md Appendix C.
static_assert(sizeof(bsv_V3f) == 12 && sizeof(bsv_V2f) == 8, "bsv vector layout");
static_assert(sizeof(cvtx_P3D) == 28 && sizeof(cvtx_F3D) == 28 && sizeof(cvtx_P2D) == 16, "particle layout");
static_assert(sizeof(cvtx_VortFunc) == 80 && offsetof(cvtx_VortFunc, cl_kernel_name_ext) == 48, "cvtx_VortFunc layout (LP64)");

// host_hooks.h -- pieces of host_api.cu that the other cvtx_* entry points use
// (no CUDA types, so plain C++ translation units can include it).
#pragma once
#include <cstddef>
#include <vector>

namespace cvtx {

std::vector<int> enabled_accelerators();                     // devices the caller has switched on
void note_dispatch(int on_gpu, int n_devices);               // feeds cvtx_b200_last_dispatch()
void gather_rows(void *dst, const void *const *ptrs, long n, size_t row_bytes);
void copy_parallel(void *dst, const void *src, size_t bytes);   // memcpy in 1 MB pieces over the gather threads
// One-target (M2S) calls: the reference evaluates them in a SERIAL loop over the sources
// (src/P3D.cpp:230-322, src/P2D.cpp:99-119,214-231, src/F3D.cpp:87-128).  From kM2SMinSources sources up,
// with an accelerator enabled and a built-in regularisation, they are the all-pairs kernel with one
// target (few-target geometry, sources packed inside the kernel).  Returns false when the call is not
// the GPU's to take; `result` receives the output row.  reg_name: cvtx_VortFunc::cl_kernel_name_ext,
// NULL for the filament ops.  tgt_row: a point (P3D vel / vort, F3D vel: 3 floats; P2D vel: 2) or the
// induced particle struct itself.
bool gpu_m2s(const char *entry, int op, const char *reg_name, const void *const *src_ptrs, int n_src,
             const void *tgt_row, float *result, float sigma, float nu);
[[noreturn]] void gpu_failure(const char *entry, int rc);    // message + abort: never a silent CPU substitute

}  // namespace cvtx

// host_hooks.h -- pieces of host_api.cu that the other cvtx_* entry points use
// (no CUDA types, so plain C++ translation units can include it).
#pragma once
#include <cstddef>
#include <vector>

namespace cvtx {

std::vector<int> enabled_accelerators();                     // devices the caller has switched on
void note_dispatch(int on_gpu, int n_devices);               // feeds cvtx_b200_last_dispatch()
void gather_rows(void *dst, const void *const *ptrs, long n, size_t row_bytes);
[[noreturn]] void gpu_failure(const char *entry, int rc);    // message + abort: never a silent CPU substitute

}  // namespace cvtx

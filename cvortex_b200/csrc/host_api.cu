// host_api.cu -- the cvtx_* all-pairs entry points and accelerator control of
// the public ABI (include/cvortex/libcvtx.h), on top of the staged
// multi-device runner of device_api.cu.
//
// Replaces, in the reference:
//   * the dispatch blocks of src/P3D.cpp:343-366,387-410,432-456,476-498,
//     src/P2D.cpp:141-162,252-276, src/F3D.cpp:162-202 ("small, or no kernel
//     name, or the GPU call failed -> OpenMP loop, else OpenCL");
//   * src/accelerators.cpp:39-118 + src/opencl_acc.cpp:55-258 (init/finalise,
//     device list, enable/disable, names).
// Dispatch rule here: if at least one accelerator is enabled and the
// cvtx_VortFunc names one of the four built-in kernels, the call runs on the
// GPU(s) at ANY size and a failure aborts with a message; the host loops of
// host_scalar.cpp run only when the caller disabled every accelerator or
// supplied a user-defined regularisation.  There is no silent substitution
// in either direction.
//
// A call does: gather the array-of-pointers input into the pinned staging area
// (parallel 28/16-byte row copies) -> H2D -> pack + pair kernels -> D2H ->
// result_array.  With several accelerators enabled the targets are split into
// contiguous shards, one per device, sources replicated.
#include <cuda_runtime.h>
#include <omp.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include "export.h"
#include "host_scalar.h"
#include "op_table.h"
#include "runtime.h"

using namespace cvtx;

namespace {

std::mutex g_mu;
bool g_init = false;
std::vector<char> g_enabled;
std::string g_info;
std::atomic<int> g_last_dispatch{-1};       // 1 = GPU, 0 = host loops, -1 = none yet
std::atomic<int> g_last_devices{0};

// Below this many pairs one device finishes before a second could be fed
// (about 2 ms of kernel time), so sharding is skipped.
constexpr double kShardMinPairs = 2.0e9;

std::vector<int> enabled_devices() {
	std::lock_guard<std::mutex> lk(g_mu);
	std::vector<int> v;
	for (size_t i = 0; i < g_enabled.size(); ++i) if (g_enabled[i]) v.push_back((int)i);
	return v;
}
}  // namespace
std::vector<int> cvtx::enabled_accelerators() { return enabled_devices(); }
void cvtx::note_dispatch(int on_gpu, int n_devices) { g_last_dispatch = on_gpu; if (on_gpu) g_last_devices = n_devices; }
namespace {

void build_info(int n_dev) {
	int rt = 0, drv = 0;
	cudaRuntimeGetVersion(&rt);
	cudaDriverGetVersion(&drv);
	char buf[256];
	g_info = "cvortex version: 0.3.8 (B200 CUDA backend)\n";
#if defined(__CUDACC_VER_MAJOR__)
	std::snprintf(buf, sizeof(buf), "compiler: nvcc %d.%d / GCC %d.%d.%d\n", __CUDACC_VER_MAJOR__, __CUDACC_VER_MINOR__,
	              __GNUC__, __GNUC_MINOR__, __GNUC_PATCHLEVEL__);
	g_info += buf;
#endif
	g_info += "using OpenMP: TRUE\nusing OpenCL: FALSE\n";
	std::snprintf(buf, sizeof(buf), "using CUDA: TRUE (runtime %d, driver %d, sm_100a kernels)\naccelerators: %d\n", rt, drv, n_dev);
	g_info += buf;
	for (int i = 0; i < n_dev; ++i) {
		std::snprintf(buf, sizeof(buf), "  [%d] %s, %d SMs\n", i, cvtx_b200_device_name(i), cvtx_b200_device_sm_count(i));
		g_info += buf;
	}
}

}  // namespace
[[noreturn]] void cvtx::gpu_failure(const char *entry, int rc) {
	std::fprintf(stderr, "cvortex: %s failed on the GPU path (status %d): %s\n"
	                     "cvortex: refusing to substitute a CPU result; aborting.\n",
	             entry, rc, cvtx_b200_last_error());
	std::abort();
}
namespace {

// Threads for the host-side gather.  Not left to OMP_NUM_THREADS: launchers such as
// torchrun export OMP_NUM_THREADS=1, which would serialise a 1M-pointer chase.
[[maybe_unused]] int gather_threads() {
	static int n = 0;
	if (n == 0) {
		const char *env = std::getenv("CVTX_B200_GATHER_THREADS");
		n = env ? std::atoi(env) : 0;
		if (n <= 0) { n = omp_get_num_procs(); if (n > 8) n = 8; }
		if (n < 1) n = 1;
	}
	return n;
}

// Rows [lo, hi) through their pointers into contiguous memory.  Hosts usually keep their particles in ONE array and
// hand over pointers into it (the reference's own benchmark does, bench/bencharraysetup.c:43-58): consecutive
// pointers that are a row apart are copied as one block -- one compare per row instead of one copy call per row,
// 24 -> 10 us on the 10k rows of a small call -- and a lone row by a fixed-size copy the compiler inlines.
template <size_t ROW> void gather_span(char *out, const void *const *ptrs, long lo, long hi) {
	long i = lo;
	while (i < hi) {
		const char *base = (const char *)ptrs[i];
		long j = i + 1;
		while (j < hi && (const char *)ptrs[j] == base + (size_t)(j - i) * ROW) ++j;
		if (j == i + 1) std::memcpy(out + (size_t)i * ROW, base, ROW);
		else std::memcpy(out + (size_t)i * ROW, base, (size_t)(j - i) * ROW);
		i = j;
	}
}
void gather_span_any(char *out, const void *const *ptrs, long lo, long hi, size_t row_bytes) {
	switch (row_bytes) {
	case 28: gather_span<28>(out, ptrs, lo, hi); break;      // cvtx_P3D, cvtx_F3D
	case 16: gather_span<16>(out, ptrs, lo, hi); break;      // cvtx_P2D
	case 12: gather_span<12>(out, ptrs, lo, hi); break;      // cvtx_Vec3f
	case 8:  gather_span<8>(out, ptrs, lo, hi); break;       // cvtx_Vec2f
	default: for (long i = lo; i < hi; ++i) std::memcpy(out + (size_t)i * row_bytes, ptrs[i], row_bytes);
	}
}
}  // namespace
// Copy n rows of `row_bytes` through an array of pointers into contiguous memory.
void cvtx::gather_rows(void *dst, const void *const *ptrs, long n, size_t row_bytes) {
	char *out = (char *)dst;
	const int threads = n > 32768 ? gather_threads() : 1;
	if (threads <= 1) { gather_span_any(out, ptrs, 0, n, row_bytes); return; }
#pragma omp parallel for schedule(static) num_threads(threads)
	for (int t = 0; t < threads; ++t) gather_span_any(out, ptrs, n * t / threads, n * (t + 1) / threads, row_bytes);
}
namespace {
void copy_rows(void *dst, const void *src, long n, size_t row_bytes) {
	const size_t total = (size_t)n * row_bytes, piece = 1 << 20;
	const long pieces = (long)((total + piece - 1) / piece);
#pragma omp parallel for schedule(static) num_threads(gather_threads()) if (pieces > 8)
	for (long i = 0; i < pieces; ++i) {
		const size_t lo = (size_t)i * piece, len = lo + piece <= total ? piece : total - lo;
		std::memcpy((char *)dst + lo, (const char *)src + lo, len);
	}
}

}  // namespace
void cvtx::copy_parallel(void *dst, const void *src, size_t bytes) { copy_rows(dst, src, 1, bytes); }
namespace {

// What a failure on the GPU route does.  Default: message + abort -- a CPU result is never substituted
// silently.  CVTX_B200_ON_FAILURE=host (read once) is the opt-in for hosts that would rather lose speed than the
// session (the reference's behaviour, src/P3D.cpp:353-364, but never silent): the failure is reported on stderr
// every time and the caller's host loop runs.
bool failure_goes_to_host(const char *entry, int rc) {
	static const bool to_host = [] { const char *e = std::getenv("CVTX_B200_ON_FAILURE"); return e && !std::strcmp(e, "host"); }();
	if (!to_host) return false;
	std::fprintf(stderr, "cvortex: %s failed on the GPU path (status %d): %s\n"
	                     "cvortex: CVTX_B200_ON_FAILURE=host -> running the HOST loops for this call.\n",
	             entry, rc, cvtx_b200_last_error());
	cudaGetLastError();
	g_last_dispatch = 0;
	return true;
}

// The GPU route of every M2M entry point.  Returns false when the call is not
// the GPU's to take (nothing enabled / user-defined regularisation), in which
// case the caller runs the host loops; aborts on a GPU failure.
bool gpu_m2m(const char *entry, int op, const cvtx_VortFunc *kernel,
             const void *const *src_ptrs, int n_src,
             const void *tgt_flat, const void *const *tgt_ptrs, int n_tgt,
             void *result, float sigma, float nu)
{
	int reg = REG_SINGULAR;
	if (!op_is_filament(op)) {
		reg = reg_from_name(kernel->cl_kernel_name_ext);
		if (reg < 0 || !op_supported(op, reg)) return false;
	}
	std::vector<int> devs = enabled_devices();
	if (devs.empty()) return false;
	g_last_dispatch = 1;
	if (n_tgt <= 0) return true;
	if ((double)n_src * (double)n_tgt < kShardMinPairs) devs.resize(1);
	g_last_devices = (int)devs.size();

	int tcols = 0, scols = 0;
	cvtx_b200_op_info(op, reg, &scols, &tcols, nullptr, nullptr, nullptr);
	const size_t srow = sizeof(float) * scols, trow = sizeof(float) * tcols;
	HostStage &hs = host_stage();
	DeviceGuard restore;
	std::lock_guard<std::mutex> lk(hs.mu);
	cudaError_t e = cudaSetDevice(devs[0]);
	if (e == cudaSuccess) e = hs.src.reserve(srow * (size_t)(n_src > 0 ? n_src : 0));
	if (e == cudaSuccess) e = hs.tgt.reserve(trow * (size_t)n_tgt);
	if (e != cudaSuccess) {
		fail(CVTX_B200_ERR_CUDA, std::string("pinned staging: ") + cudaGetErrorString(e));
		if (failure_goes_to_host(entry, CVTX_B200_ERR_CUDA)) return false;
		gpu_failure(entry, CVTX_B200_ERR_CUDA);
	}
	if (n_src > 0) gather_rows(hs.src.p, src_ptrs, n_src, srow);
	if (tgt_ptrs) gather_rows(hs.tgt.p, tgt_ptrs, n_tgt, trow);
	else copy_rows(hs.tgt.p, tgt_flat, n_tgt, trow);
	const int rc = run_staged(op, reg, devs, n_src > 0 ? n_src : 0, n_tgt, (float *)result, sigma, nu, nullptr, nullptr);
	if (rc != CVTX_B200_OK) {
		if (failure_goes_to_host(entry, rc)) return false;
		gpu_failure(entry, rc);
	}
	return true;
}

}  // namespace

// Below this many sources a one-target call stays in the host loop: the GPU route costs ~0.1 ms of
// gather + copies + launch + sync whatever the size, the serial host loop ~50 ns per source.
constexpr int kM2SMinSources = 4096;

bool cvtx::gpu_m2s(const char *entry, int op, const char *reg_name, const void *const *src_ptrs, int n_src,
                   const void *tgt_row, float *result, float sigma, float nu)
{
	if (n_src < kM2SMinSources) return false;
	cvtx_VortFunc k;
	std::memset(&k, 0, sizeof k);
	if (reg_name) std::strncpy(k.cl_kernel_name_ext, reg_name, sizeof(k.cl_kernel_name_ext) - 1);
	return gpu_m2m(entry, op, reg_name ? &k : nullptr, src_ptrs, n_src, tgt_row, nullptr, 1, result, sigma, nu);
}

extern "C" {

// ---- diagnostics (declared in cvtx_b200.h) -------------------------------------
CVTX_API int cvtx_b200_last_dispatch(void) { return g_last_dispatch.load(); }
CVTX_API int cvtx_b200_last_devices_used(void) { return g_last_devices.load(); }

// ---- library / accelerator control ---------------------------------------------
CVTX_API void cvtx_initialise() {
	std::lock_guard<std::mutex> lk(g_mu);
	if (g_init) return;                                  // idempotent, like reference src/opencl_acc.cpp:55-67
	int n = cvtx_b200_device_count();
	if (n < 0) {
		std::fprintf(stderr, "cvortex: CUDA initialisation failed: %s\n", cvtx_b200_last_error());
		n = 0;
	}
	if (n == 0) {
		// Same ABI behaviour as the reference without an OpenCL device (zero accelerators, host
		// loops), but never silently: say it, and let deployments that must not run on the host
		// turn it into a hard error.
		std::fprintf(stderr, "cvortex: no CUDA accelerator found; cvtx_*_M2M_* calls will run the host loops.\n");
		const char *req = std::getenv("CVTX_B200_REQUIRE_GPU");
		if (req && req[0] == '1') { std::fprintf(stderr, "cvortex: CVTX_B200_REQUIRE_GPU=1 -> aborting.\n"); std::abort(); }
	}
	g_enabled.assign((size_t)n, 0);
	// Default: accelerator 0 only, as the reference does (src/opencl_acc.cpp:192-201).
	// CVTX_B200_ENABLE=all turns every device on for callers that cannot be
	// changed to call cvtx_accelerator_enable() themselves.
	const char *env = std::getenv("CVTX_B200_ENABLE");
	if (n > 0) g_enabled[0] = 1;
	if (env && !std::strcmp(env, "all")) g_enabled.assign((size_t)n, 1);
	build_info(n);
	g_init = true;
}

CVTX_API void cvtx_finalise() {
	{
		std::lock_guard<std::mutex> lk(g_mu);
		if (!g_init) return;
		g_enabled.clear();
		g_info.clear();
		g_init = false;
	}
	cvtx_b200_release();
}

CVTX_API const char *cvtx_information() { return g_info.c_str(); }

CVTX_API int cvtx_num_accelerators() {
	std::lock_guard<std::mutex> lk(g_mu);
	return (int)g_enabled.size();
}

CVTX_API int cvtx_num_enabled_accelerators() {
	std::lock_guard<std::mutex> lk(g_mu);
	int c = 0;
	for (char e : g_enabled) c += e ? 1 : 0;
	return c;
}

CVTX_API const char *cvtx_accelerator_name(int accelerator_id) {
	{
		std::lock_guard<std::mutex> lk(g_mu);
		if (accelerator_id < 0 || accelerator_id >= (int)g_enabled.size()) return nullptr;
	}
	return cvtx_b200_device_name(accelerator_id);
}

CVTX_API int cvtx_accelerator_enabled(int accelerator_id) {
	std::lock_guard<std::mutex> lk(g_mu);
	if (accelerator_id < 0 || accelerator_id >= (int)g_enabled.size()) return 0;
	return g_enabled[accelerator_id] ? 1 : 0;
}

CVTX_API void cvtx_accelerator_enable(int accelerator_id) {
	std::lock_guard<std::mutex> lk(g_mu);
	if (accelerator_id >= 0 && accelerator_id < (int)g_enabled.size()) g_enabled[accelerator_id] = 1;
}

CVTX_API void cvtx_accelerator_disable(int accelerator_id) {
	std::lock_guard<std::mutex> lk(g_mu);
	if (accelerator_id >= 0 && accelerator_id < (int)g_enabled.size()) g_enabled[accelerator_id] = 0;
}

// ---- the hot path --------------------------------------------------------------
#define PTRS(p) ((const void *const *)(p))

CVTX_API void cvtx_P3D_M2M_vel(const cvtx_P3D **array_start, const int num_particles, const bsv_V3f *mes_start,
                               const int num_mes, bsv_V3f *result_array, const cvtx_VortFunc *kernel, float regularisation_radius) {
	if (gpu_m2m("cvtx_P3D_M2M_vel", OP_P3D_VEL, kernel, PTRS(array_start), num_particles, mes_start, nullptr, num_mes,
	            result_array, regularisation_radius, 0.f)) return;
	g_last_dispatch = 0;
	host_m2m_p3d_vel(array_start, num_particles, mes_start, num_mes, result_array, kernel, regularisation_radius);
}

CVTX_API void cvtx_P3D_M2M_dvort(const cvtx_P3D **array_start, const int num_particles, const cvtx_P3D **induced_start,
                                 const int num_induced, bsv_V3f *result_array, const cvtx_VortFunc *kernel, float regularisation_radius) {
	if (gpu_m2m("cvtx_P3D_M2M_dvort", OP_P3D_DVORT, kernel, PTRS(array_start), num_particles, nullptr, PTRS(induced_start),
	            num_induced, result_array, regularisation_radius, 0.f)) return;
	g_last_dispatch = 0;
	host_m2m_p3d_dvort(array_start, num_particles, induced_start, num_induced, result_array, kernel, regularisation_radius);
}

CVTX_API void cvtx_P3D_M2M_visc_dvort(const cvtx_P3D **array_start, const int num_particles, const cvtx_P3D **induced_start,
                                      const int num_induced, bsv_V3f *result_array, const cvtx_VortFunc *kernel,
                                      float regularisation_radius, float kinematic_visc) {
	if (gpu_m2m("cvtx_P3D_M2M_visc_dvort", OP_P3D_VISC, kernel, PTRS(array_start), num_particles, nullptr, PTRS(induced_start),
	            num_induced, result_array, regularisation_radius, kinematic_visc)) return;
	g_last_dispatch = 0;
	host_m2m_p3d_visc(array_start, num_particles, induced_start, num_induced, result_array, kernel, regularisation_radius, kinematic_visc);
}

CVTX_API void cvtx_P3D_M2M_vort(const cvtx_P3D **array_start, const int num_particles, const bsv_V3f *mes_start,
                                const int num_mes, bsv_V3f *result_array, const cvtx_VortFunc *kernel, float regularisation_radius) {
	if (gpu_m2m("cvtx_P3D_M2M_vort", OP_P3D_VORT, kernel, PTRS(array_start), num_particles, mes_start, nullptr, num_mes,
	            result_array, regularisation_radius, 0.f)) return;
	g_last_dispatch = 0;
	host_m2m_p3d_vort(array_start, num_particles, mes_start, num_mes, result_array, kernel, regularisation_radius);
}

CVTX_API void cvtx_P2D_M2M_vel(const cvtx_P2D **array_start, const int num_particles, const bsv_V2f *mes_start,
                               const int num_mes, bsv_V2f *result_array, const cvtx_VortFunc *kernel, float regularisation_radius) {
	if (gpu_m2m("cvtx_P2D_M2M_vel", OP_P2D_VEL, kernel, PTRS(array_start), num_particles, mes_start, nullptr, num_mes,
	            result_array, regularisation_radius, 0.f)) return;
	g_last_dispatch = 0;
	host_m2m_p2d_vel(array_start, num_particles, mes_start, num_mes, result_array, kernel, regularisation_radius);
}

CVTX_API void cvtx_P2D_M2M_visc_dvort(const cvtx_P2D **array_start, const int num_particles, const cvtx_P2D **induced_start,
                                      const int num_induced, float *result_array, const cvtx_VortFunc *kernel,
                                      float regularisation_radius, float kinematic_visc) {
	if (gpu_m2m("cvtx_P2D_M2M_visc_dvort", OP_P2D_VISC, kernel, PTRS(array_start), num_particles, nullptr, PTRS(induced_start),
	            num_induced, result_array, regularisation_radius, kinematic_visc)) return;
	g_last_dispatch = 0;
	host_m2m_p2d_visc(array_start, num_particles, induced_start, num_induced, result_array, kernel, regularisation_radius, kinematic_visc);
}

CVTX_API void cvtx_F3D_M2M_vel(const cvtx_F3D **array_start, const int num_filaments, const bsv_V3f *mes_start,
                               const int num_mes, bsv_V3f *result_array) {
	if (gpu_m2m("cvtx_F3D_M2M_vel", OP_F3D_VEL, nullptr, PTRS(array_start), num_filaments, mes_start, nullptr, num_mes,
	            result_array, 0.f, 0.f)) return;
	g_last_dispatch = 0;
	host_m2m_f3d_vel(array_start, num_filaments, mes_start, num_mes, result_array);
}

CVTX_API void cvtx_F3D_M2M_dvort(const cvtx_F3D **array_start, const int num_filaments, const cvtx_P3D **induced_start,
                                 const int num_induced, bsv_V3f *result_array) {
	if (gpu_m2m("cvtx_F3D_M2M_dvort", OP_F3D_DVORT, nullptr, PTRS(array_start), num_filaments, nullptr, PTRS(induced_start),
	            num_induced, result_array, 0.f, 0.f)) return;
	g_last_dispatch = 0;
	host_m2m_f3d_dvort(array_start, num_filaments, induced_start, num_induced, result_array);
}

// Dense influence matrix.  GPU route: filaments gathered and uploaded once,
// then the matrix is produced in row slabs that fit the pinned result buffer and streamed back --
// the m x n floats leaving over PCIe are what bounds this call, not the kernel.
CVTX_API void cvtx_F3D_inf_mtrx(const cvtx_F3D **array_start, const int num_filaments, const bsv_V3f *mes_start,
                                const bsv_V3f *dir_start, const int num_mes, float *result_matrix) {
	std::vector<int> devs = enabled_devices();
	if (devs.empty()) {
		g_last_dispatch = 0;
		host_f3d_inf_mtrx(array_start, num_filaments, mes_start, dir_start, num_mes, result_matrix);
		return;
	}
	g_last_dispatch = 1;
	g_last_devices = 1;
	if (num_filaments <= 0 || num_mes <= 0) return;
	const int dev = devs[0];
	Device *d = get_device(dev);
	HostStage &hs = host_stage();
	DeviceGuard restore;
	std::lock_guard<std::mutex> lk(hs.mu);
	const size_t fb = sizeof(cvtx_F3D) * (size_t)num_filaments, pb = sizeof(bsv_V3f) * (size_t)num_mes;
	const size_t row_bytes = sizeof(float) * (size_t)num_filaments;
	long slab_rows = (long)((size_t)(128u << 20) / row_bytes);
	if (slab_rows < 1) slab_rows = 1;
	if (slab_rows > num_mes) slab_rows = num_mes;
	int rc = CVTX_B200_OK;
	auto step = [&](cudaError_t e, const char *what) {
		if (e != cudaSuccess && rc == CVTX_B200_OK) rc = fail(CVTX_B200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
		return e == cudaSuccess;
	};
	cudaStream_t st = nullptr;
	bool ok = true;
	if ((rc = device_stream(dev, &st)) != CVTX_B200_OK) ok = false;
	ok = ok && step(hs.src.reserve(fb), "pinned filaments") && step(hs.tgt.reserve(2 * pb), "pinned points")
	     && step(hs.out.reserve(row_bytes * (size_t)slab_rows), "pinned result slab");
	if (ok) {
		DeviceLock dl(d->mu);
		ok = step(d->d_src.reserve(fb), "device filaments") && step(d->d_tgt.reserve(2 * pb), "device points")
		     && step(d->d_out.reserve(row_bytes * (size_t)slab_rows), "device result slab");
	}
	if (ok) {
		gather_rows(hs.src.p, PTRS(array_start), num_filaments, sizeof(cvtx_F3D));
		std::memcpy(hs.tgt.p, mes_start, pb);
		std::memcpy((char *)hs.tgt.p + pb, dir_start, pb);
		ok = step(cudaMemcpyAsync(d->d_src.p, hs.src.p, fb, cudaMemcpyHostToDevice, st), "H2D filaments")
		     && step(cudaMemcpyAsync(d->d_tgt.p, hs.tgt.p, 2 * pb, cudaMemcpyHostToDevice, st), "H2D points");
	}
	for (long r0 = 0; ok && r0 < num_mes; r0 += slab_rows) {
		const long rows = r0 + slab_rows <= num_mes ? slab_rows : num_mes - r0;
		const float *mes_d = (const float *)d->d_tgt.p + 3 * r0;
		const float *dir_d = (const float *)((const char *)d->d_tgt.p + pb) + 3 * r0;
		const int krc = cvtx_b200_f3d_inf_mtrx(dev, st, (const float *)d->d_src.p, num_filaments, mes_d, dir_d, (int)rows,
		                                       (float *)d->d_out.p);
		if (krc != CVTX_B200_OK) { rc = krc; ok = false; break; }
		ok = step(cudaMemcpyAsync(hs.out.p, d->d_out.p, row_bytes * (size_t)rows, cudaMemcpyDeviceToHost, st), "D2H slab")
		     && step(cudaStreamSynchronize(st), "sync");
		if (ok) copy_rows(result_matrix + (size_t)r0 * num_filaments, hs.out.p, rows, row_bytes);
	}
	if (!ok) gpu_failure("cvtx_F3D_inf_mtrx", rc);
}

}  // extern "C"

// device_api.cu -- the thin C ABI over the sm_100a kernels (include/cvtx_b200.h).
//
// Takes the place of the reference's OpenCL host layer
//   src/ocl_P3D.cpp:150-725, src/ocl_P2D.cpp:107-631, src/ocl_F3D.cpp:72-646
// (pack to cl_float3 -> one NDRange per 256 sources chained by events ->
// blocking read -> scale) and of the per-device state it pulls from
//   src/opencl_acc.cpp:203-239, src/OclPlatformState.cpp:87-213.
// One call = pack kernel + ONE pair kernel (+ a reduce kernel when the source
// set was split for load balance), all asynchronous on the caller's stream.
// No CPU path lives here: errors are returned, never papered over.
#include <cuda_runtime.h>
#include <chrono>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../include/cvtx_b200.h"
#include "aux_kernels.cuh"
#include "kernel_table.h"
#include "runtime.h"

using namespace cvtx;

static_assert((int)CVTX_B200_P3D_VEL == OP_P3D_VEL && (int)CVTX_B200_F3D_DVORT == OP_F3D_DVORT && (int)CVTX_B200_P3D_VEL_DVORT == OP_P3D_VEL_DVORT, "op ids");
static_assert((int)CVTX_B200_SINGULAR == REG_SINGULAR && (int)CVTX_B200_GAUSSIAN == REG_GAUSSIAN, "reg ids");

// ---- shared runtime pieces (declared in runtime.h) ----------------------------
namespace {
thread_local std::string g_err;
std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_force_T{0}, g_force_chunks{0};
std::atomic<int> g_guard_mode{-1};
std::atomic<int> g_sparse_route{1};
std::atomic<int> g_f3d_mode{-2};                              // -2 = not set: CVTX_B200_F3D_MODE decides; -1 auto, 0 new, 1 wide                           // -1 = not set: CVTX_B200_GUARDED decides, read once
std::mutex g_devices_mu;
std::vector<Device *> g_devices;
int g_device_count = -2;                                       // -2 = not probed yet
}  // namespace

int cvtx::fail(int code, const std::string &msg) { g_err = msg; return code; }
void cvtx::count_launches(unsigned long long n) { g_launches += n; }

cudaError_t cvtx::Buffer::reserve(size_t bytes) {
	if (bytes <= cap) return cudaSuccess;
	release();
	const size_t want = bytes + bytes / 4 + 4096;              // grow-only with headroom
	cudaError_t e = pinned ? cudaHostAlloc(&p, want, cudaHostAllocPortable) : cudaMalloc(&p, want);
	if (e == cudaSuccess) cap = want; else p = nullptr;
	return e;
}

void cvtx::Buffer::release() {
	if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); }
	p = nullptr; cap = 0;
}

cvtx::HostStage &cvtx::host_stage() {
	static HostStage hs;
	hs.src.pinned = hs.tgt.pinned = hs.out.pinned = true;
	return hs;
}

namespace {

int probe_devices() {
	std::lock_guard<std::mutex> lk(g_devices_mu);
	if (g_device_count != -2) return g_device_count;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) {
		g_err = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
		cudaGetLastError();
		// no driver / no device is "zero accelerators", anything else is an error
		g_device_count = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? 0 : -1;
		return g_device_count;
	}
	for (int i = 0; i < n; ++i) {
		Device *d = new Device();
		if (cudaGetDeviceProperties(&d->prop, i) != cudaSuccess) std::memset(&d->prop, 0, sizeof(d->prop));
		g_devices.push_back(d);
	}
	g_device_count = n;
	return n;
}

int ensure_ready(Device *d) {                   // caller holds d->mu and has done cudaSetDevice
	if (d->ready) return CVTX_B200_OK;
	CUDA_TRY(cudaEventCreateWithFlags(&d->arena_idle, cudaEventDisableTiming));
	CUDA_TRY(cudaEventCreate(&d->k_start));
	CUDA_TRY(cudaEventCreate(&d->k_stop));
	CUDA_TRY(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
	d->ready = true;
	return CVTX_B200_OK;
}

// ---- launch planning ----------------------------------------------------------
// Source sets below this many tiles are walked in grains (= FP32 chains) of 32 sources instead of
// 256, so that a 10k x 10k call still cuts into enough equal runs for every resident block.  A
// property of the SOURCES alone: every shard of a multi-GPU call rounds alike.
constexpr int kSmallSourceTiles = 64;
// The grid is this many times the blocks that are resident at once when the problem is big enough:
// two blocks that share an SM do not advance at the same pace (the warp schedulers favour one of them:
// with exactly one resident set the favoured block of every SM was measured to finish after 54 % of the
// launch and its partner ran the rest alone, profiles/kernel_ab_r2.txt), so the hardware block scheduler
// has to have a next run to hand out.  Runs stay equal; 3 ... 8 measured the same, 16 worse.
constexpr int kRunsPerSlot = 4;
constexpr int kMinRunGrains = 24;
constexpr int kInKernelFinishPieces = 48;
constexpr size_t kZeroCopyResultBytes = 512u << 10;      // results up to this size are stored to pinned host memory by the kernel
constexpr size_t kKernelUploadBytes = 4u << 20;          // inputs up to this size are read from pinned host memory by a kernel (upload_rows_kernel)

struct Plan { KernelChoice k; int grain, grid; long long tiles_t, total_grains; };

// Pick the block geometry (targets per thread) and the grid.  The work -- (target tile) x (grain) cells
// -- is cut into `grid` equal runs.  Every (geometry, grid) candidate is priced in microseconds:
//   waves x ( prologue + pairs of a run / what a block gets of its SM )  +  the ordered finish of cut tiles
// where a block's share of the SM depends on how many blocks are resident with it (4 warps alone keep the
// FP32 pipe ~60 % busy, 8 warps 93 %, 16 and more 100 %) and a single resident set of co-resident blocks pays for the
// uneven pace of the two (+30 %: the slower one finishes alone).  Large problems come out at the op's
// preferred geometry with kRunsPerSlot runs per resident slot; small ones (10k x 10k: 100 us) at ONE run per
// SM in a small-tile geometry, which is what an exhaustive sweep finds too (profiles/plan_sweep_r2.txt).
struct Planner {
	int device, n_src, n_tgt, sm_count, op, reg; Plan plan;
	template <class P> void run() {
		const int n_src_tiles = (n_src + kSrcTile - 1) / kSrcTile;
		static const double cand_eff[4] = {1.0, 0.985, 0.96, 0.79};      // profiles/sweep_ops_r1.txt, ubench_r1.txt
		const int fT = g_force_T.load(), fG = g_force_chunks.load();
		const int pref = P::PREF_T;
		const int first = pref >= 8 ? 0 : (pref >= 4 ? 1 : (pref >= 2 ? 2 : 3));
		const int grain = n_src_tiles < kSmallSourceTiles ? 32 : kSrcTile;
		// pairs per microsecond one SM sustains on this op at 85 % of its FP32 issue rate (1.965 GHz x 128 lanes)
		const double sm_rate = 0.85 * 128.0 * 1965.0 / (double)P::LANE_OPS;
		double best = 1e300;
		plan = Plan();
		for (int v = 0; v < 4; ++v) {
			const KernelChoice k = kernel_choice(op, reg, v, grain == kSrcTile, device);
			if (fT ? k.T != fT : v < first) continue;
			const long long slots = (long long)k.B * k.T;
			const long long tiles_t = ((long long)n_tgt + slots - 1) / slots;
			const long long total = tiles_t * n_src_tiles * (kSrcTile / grain);
			const long long resident = (long long)sm_count * k.occ;
			long long mult = total / (resident * kMinRunGrains);
			mult = mult < 1 ? 1 : (mult > kRunsPerSlot ? kRunsPerSlot : mult);
			// a step = one source tile against one target tile.  However few of a block's target slots are real, the
			// warps that have one walk the tile's 256 sources one after the other and the tile has to arrive first:
			// ~7 us for one target per thread (profiles/few_sweep_r2.txt), more with more targets per thread
			const long long gps = kSrcTile / grain;
			const double step_floor_us = 2.5 + 4.5 * (double)k.T * (double)P::LANE_OPS / 21.0;
			const long long cand_grid[5] = {resident * mult, (long long)sm_count, 2LL * sm_count, 4LL * sm_count, 8LL * sm_count};
			for (int gi = 0; gi < 5; ++gi) {
				long long grid = cand_grid[gi];
				if (fG > 0) grid = fG;
				if (grid > total) grid = total;
				if (grid < 1) grid = 1;
				if (gi > 0 && (fG > 0 || grid >= cand_grid[0])) continue;
				const long long per_block = (total + grid - 1) / grid;
				const long long waves = (grid + resident - 1) / resident;
				long long together = (grid + sm_count - 1) / sm_count;              // blocks that share an SM
				if (together > k.occ) together = k.occ;
				// FP32-pipe utilisation by resident warps per SM (measured on the T = 4 / T = 2 geometries run with one
				// block per SM, profiles/sweep_ops_r2.txt), relative to the geometry at its own full occupancy, which
				// is what cand_eff and the TUNE constants were measured at
				auto by_warps = [](double w) { return w >= 16.0 ? 1.0 : (w >= 12.0 ? 0.97 : (w >= 8.0 ? 0.93 : (w >= 4.0 ? 0.6 : 0.15 * w))); };
				const double util = by_warps((double)together * k.B / 32.0) / by_warps((double)k.occ * k.B / 32.0);
				const double uneven = (together >= 2 && waves == 1) ? 1.3 : 1.0;
				double run_us = (double)per_block * grain * (double)slots * (double)together * uneven / (sm_rate * cand_eff[v] * util);
				const double steps = (double)((per_block + gps - 1) / gps);
				if (run_us < steps * step_floor_us) run_us = steps * step_floor_us;
				const double pieces = (double)(grid + tiles_t - 1) / (double)tiles_t;      // runs that touch a cut tile
				const double finish_us = grid <= tiles_t ? 0.0 : (pieces > kInKernelFinishPieces ? 12.0 : 2.0 + 0.4 * pieces);
				// + what every run costs whatever it does (block launch and retirement, its FP64 piece): 30 ns
				const double cost = (double)waves * (3.0 + run_us) + finish_us + 0.03 * (double)grid;
				if (cost < best) {
					best = cost;
					plan.k = k; plan.grain = grain; plan.grid = (int)grid; plan.tiles_t = tiles_t; plan.total_grains = total;
				}
			}
		}
	}
};

// Chains shorter than this many tiles per target are evaluated in the guarded form straight away:
// a self-interaction call re-evaluates one chain per target, which is noise among thousands of
// chains and half the work among two.
constexpr int kMinTilesOptimistic = 16;

// CVTX_B200_TRACE=1: every staged call prints where its wall-clock time went (stderr, microseconds)
bool trace_enabled() {
	static const bool on = [] { const char *e = getenv("CVTX_B200_TRACE"); return e && e[0] == '1'; }();
	return on;
}
double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// CVTX_B200_SPARSE=0 switches the sparse-tile route of the box-cutoff op off (tests, A/B measurements)
bool sparse_route_enabled() {
	static const bool on = [] { const char *e = getenv("CVTX_B200_SPARSE"); return !(e && e[0] == '0'); }();
	return on && g_sparse_route.load() != 0;
}

// 0 = by size (the default), 1 = guarded form only, 2 = optimistic form at any size
int guard_mode() {
	int v = g_guard_mode.load();
	if (v < 0) {
		const char *e = getenv("CVTX_B200_GUARDED");
		v = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 0;
		g_guard_mode = v;
	}
	return v;
}

// ---- run-time pipe peaks (cvtx_b200_measure_peak): the roofline denominators, measured ----
// Packed FP32 FMA (FFMA2): 8 independent chains per thread, two lane-ops per instruction slot.
__global__ void __launch_bounds__(256) peak_ffma2_kernel(float *out, int iters, float a, float b) {
	float2 x[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 16; ++r)
#pragma unroll
			for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], make_float2(a, a), make_float2(b, b));
	}
	float s = 0.f;
#pragma unroll
	for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
	if (s == 123.456f) out[0] = s;
}
__global__ void __launch_bounds__(256) peak_mufu_kernel(float *out, int iters) {
	float x[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) x[i] = 1.0f + threadIdx.x * 1e-3f + i;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 16; ++r)
#pragma unroll
			for (int i = 0; i < 8; ++i) x[i] = mufu_rsqrt(x[i]);
	}
	float s = 0.f;
#pragma unroll
	for (int i = 0; i < 8; ++i) s += x[i];
	if (s == 123.456f) out[0] = s;
}

// -1 = decided per call from the filaments (f3d_pick_mode); 0 / 1 pin the fast form (tests, experiments)
int f3d_mode_override() {
	int v = g_f3d_mode.load();
	if (v == -2) {
		const char *e = getenv("CVTX_B200_F3D_MODE");
		v = (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : -1;
		g_f3d_mode = v;
	}
	return v;
}

struct Info {
	int lane, sfu, tcols, nout;
	template <class P> void run() { lane = P::LANE_OPS; sfu = P::SFU_OPS; tcols = P::TCOLS; nout = P::NOUT; }
};

struct ConstsOf {
	float sigma, nu; PairConsts k;
	template <class P> void run() { k = P::make_consts(sigma, nu); }
};

}  // namespace

cvtx::Device *cvtx::get_device(int device) {
	const int n = probe_devices();
	if (device < 0 || device >= n) return nullptr;
	return g_devices[device];
}

int cvtx::device_stream(int device, cudaStream_t *stream) {
	Device *d = get_device(device);
	if (!d) return fail(CVTX_B200_ERR_ARGUMENT, "no such CUDA device");
	DeviceGuard restore;
	DeviceLock lk(d->mu);
	CUDA_TRY(cudaSetDevice(device));
	if (int rc = ensure_ready(d)) return rc;
	*stream = d->stream;
	return CVTX_B200_OK;
}

// =============================================================================
extern "C" {

int cvtx_b200_device_count(void) { return probe_devices(); }

const char *cvtx_b200_device_name(int device) {
	Device *d = get_device(device);
	return d ? d->prop.name : nullptr;
}

int cvtx_b200_device_sm_count(int device) {
	Device *d = get_device(device);
	return d ? d->prop.multiProcessorCount : -1;
}

int cvtx_b200_device_clock_khz(int device) {
	if (!get_device(device)) return -1;
	int khz = 0;
	if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device) != cudaSuccess) return -1;
	return khz;
}

void cvtx_b200_release(void) {
	release_exchange();
	DeviceGuard restore;
	{
		std::lock_guard<std::mutex> lk(g_devices_mu);
		for (size_t i = 0; i < g_devices.size(); ++i) {
			Device *d = g_devices[i];
			DeviceLock dl(d->mu);
			if (!d->ready) continue;
			cudaSetDevice((int)i);
			cudaDeviceSynchronize();
			Buffer *all[] = {&d->packedA, &d->packedB, &d->packedC, &d->pieces, &d->tickets, &d->aux, &d->d_src, &d->d_tgt, &d->d_out};
			for (Buffer *b : all) b->release();
			for (Buffer &b : d->remesh) b.release();
			cudaEventDestroy(d->arena_idle); cudaEventDestroy(d->k_start); cudaEventDestroy(d->k_stop);
			cudaStreamDestroy(d->stream);
			d->ready = false; d->timed = false;
		}
	}
	HostStage &hs = host_stage();
	std::lock_guard<std::mutex> hl(hs.mu);
	hs.src.release(); hs.tgt.release(); hs.out.release();
}

int cvtx_b200_op_info(int op, int reg, int *src_cols_, int *tgt_cols, int *out_cols, int *lane_ops, int *sfu_ops) {
	Info q = {};
	if (!dispatch_op(op, reg, q)) return fail(CVTX_B200_ERR_UNSUPPORTED, "no kernel for this (op, regularisation)");
	if (src_cols_) *src_cols_ = src_cols(op);
	if (tgt_cols) *tgt_cols = q.tcols;
	if (out_cols) *out_cols = q.nout;
	if (lane_ops) *lane_ops = q.lane;
	if (sfu_ops) *sfu_ops = q.sfu;
	return CVTX_B200_OK;
}

int cvtx_b200_plan(int op, int device, int n_src, int n_tgt, int *block, int *tpt, int *grid_x, int *grid_y) {
	Device *d = get_device(device);
	Info q = {};
	if (!d || n_src < 0 || n_tgt < 0) return fail(CVTX_B200_ERR_ARGUMENT, "bad device or counts");
	if (!dispatch_op(op, op_is_filament(op) ? 0 : REG_WINCKELMANS, q)) return fail(CVTX_B200_ERR_UNSUPPORTED, "bad op");
	Planner pl = {device, n_src, n_tgt, d->prop.multiProcessorCount, op, op_is_filament(op) ? 0 : REG_WINCKELMANS, {}};
	{
		DeviceGuard restore;
		DeviceLock lk(d->mu);
		CUDA_TRY(cudaSetDevice(device));
		dispatch_op(op, op_is_filament(op) ? 0 : REG_WINCKELMANS, pl);
	}
	const Plan &p = pl.plan;
	if (block) *block = p.k.B;
	if (tpt) *tpt = p.k.T;
	if (grid_x) *grid_x = p.grid;
	if (grid_y) *grid_y = p.grain;
	return CVTX_B200_OK;
}

unsigned long long cvtx_b200_kernel_launches(void) { return g_launches.load(); }

void cvtx_b200_tune(int force_T, int force_chunks) { g_force_T = force_T; g_force_chunks = force_chunks; }

void cvtx_b200_f3d_mode(int mode) { g_f3d_mode = (mode == 0 || mode == 1) ? mode : -1; }

void cvtx_b200_guarded_only(int mode) { g_guard_mode = (mode == 1 || mode == 2) ? mode : 0; }

void cvtx_b200_sparse_route(int on) { g_sparse_route = on ? 1 : 0; }

const char *cvtx_b200_last_error(void) { return g_err.c_str(); }

float cvtx_b200_last_pair_kernel_ms(int device) {
	Device *d = get_device(device);
	if (!d) return -1.f;
	DeviceGuard restore;
	DeviceLock lk(d->mu);
	if (!d->timed) return -1.f;
	float ms = -1.f;
	if (cudaSetDevice(device) != cudaSuccess) return -1.f;
	if (cudaEventSynchronize(d->k_stop) != cudaSuccess) return -1.f;
	if (cudaEventElapsedTime(&ms, d->k_start, d->k_stop) != cudaSuccess) return -1.f;
	return ms;
}

int cvtx_b200_measure_peak(int device, int what, double *ops_per_second)
{
	g_err.clear();
	Device *d = get_device(device);
	if (!d || !ops_per_second || (what != 0 && what != 1)) return fail(CVTX_B200_ERR_ARGUMENT, "bad device, selector or pointer");
	DeviceGuard restore;
	DeviceLock lk(d->mu);
	CUDA_TRY(cudaSetDevice(device));
	if (int rc = ensure_ready(d)) return rc;
	float *sink = nullptr;
	CUDA_TRY(cudaMalloc(&sink, sizeof(float)));
	cudaEvent_t e0, e1;
	CUDA_TRY(cudaEventCreate(&e0));
	CUDA_TRY(cudaEventCreate(&e1));
	const int blocks = d->prop.multiProcessorCount * 8, threads = 256;
	const int iters = what == 0 ? 4096 : 1024;
	double best = 0.0;
	cudaError_t err = cudaSuccess;
	for (int rep = 0; rep < 4 && err == cudaSuccess; ++rep) {          // first pass warms up
		cudaEventRecord(e0, d->stream);
		if (what == 0) peak_ffma2_kernel<<<blocks, threads, 0, d->stream>>>(sink, iters, 0.999f, 1e-3f);
		else peak_mufu_kernel<<<blocks, threads, 0, d->stream>>>(sink, iters);
		cudaEventRecord(e1, d->stream);
		err = cudaEventSynchronize(e1);
		float ms = 0.f;
		if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
		// per thread and iteration: 16 x 8 instructions; an FFMA2 is two lane-ops
		const double ops = (double)blocks * threads * iters * 16.0 * 8.0 * (what == 0 ? 2.0 : 1.0);
		if (rep > 0 && ms > 0.f && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
	g_launches += 4;
	CUDA_TRY(err);
	*ops_per_second = best;
	return CVTX_B200_OK;
}

int cvtx_b200_m2m(int op, int reg, int device, void *stream_, const float *src, int n_src,
                  const float *tgt, int n_tgt, float *out, float sigma, float nu)
{
	g_err.clear();
	Device *d = get_device(device);
	if (!d) return fail(CVTX_B200_ERR_ARGUMENT, "no such CUDA device");
	if (n_src < 0 || n_tgt < 0) return fail(CVTX_B200_ERR_ARGUMENT, "negative count");
	if (n_src > 2147483647 - 2 * kSrcTile) return fail(CVTX_B200_ERR_ARGUMENT, "too many sources for 32-bit tile indexing");
	if (op_is_filament(op)) reg = REG_SINGULAR;
	Info q = {};
	if (!dispatch_op(op, reg, q)) return fail(CVTX_B200_ERR_UNSUPPORTED, "no kernel for this (op, regularisation)");
	if (n_tgt == 0) return CVTX_B200_OK;
	if (!tgt || !out || (n_src > 0 && !src)) return fail(CVTX_B200_ERR_ARGUMENT, "null pointer");

	cudaStream_t st = (cudaStream_t)stream_;
	DeviceGuard restore;
	DeviceLock lk(d->mu);
	CUDA_TRY(cudaSetDevice(device));
	if (int rc = ensure_ready(d)) return rc;
	if (n_src == 0) {                                  // no sources: the sums are empty
		CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n_tgt * q.nout, st));
		return CVTX_B200_OK;
	}

	Planner pl = {device, n_src, n_tgt, d->prop.multiProcessorCount, op, reg, {}};
	dispatch_op(op, reg, pl);
	const Plan &plan = pl.plan;
	const int n_src_tiles = (n_src + kSrcTile - 1) / kSrcTile;
	const int n_pad = n_src_tiles * kSrcTile;
	const int records = src_records(op);
	if ((plan.total_grains + plan.grid - 1) / plan.grid > 2147483647LL || plan.tiles_t > 2147483647LL)
		return fail(CVTX_B200_ERR_ARGUMENT, "problem too large for one launch (more than 2^31 grains per block)");
	const bool filament = op_is_filament(op);

	// the arena may still be in use by an earlier call on another stream
	CUDA_TRY(cudaStreamWaitEvent(st, d->arena_idle, 0));
	const size_t need_packed = plan.k.can_direct ? 0 : (size_t)n_pad * sizeof(float4);     // no packed copy in direct mode
	const size_t need_pieces = sizeof(double) * 2 * (size_t)plan.grid * plan.k.T * plan.k.B * q.nout;
	const size_t need_tickets = sizeof(int) * (size_t)plan.tiles_t;
	// cvtx_P3D_M2M_vort on enough targets and a packed source set: the sparse-tile route is tried (m2m_kernel.cuh)
	constexpr int kSparseSlots = 8 * 128;                                // sparse_tiles_kernel<P, 8, 128>: targets per tile
	const long long sparse_tiles_t = ((long long)n_tgt + kSparseSlots - 1) / kSparseSlots;
	const bool try_sparse = op == OP_P3D_VORT && !plan.k.can_direct && sparse_tiles_t >= 2LL * d->prop.multiProcessorCount
	                        && g_force_T.load() == 0 && g_force_chunks.load() == 0 && sparse_route_enabled();
	const int sparse_words = (n_src_tiles + 31) / 32;
	const size_t need_aux = filament ? sizeof(F3DStats) * (size_t)n_src_tiles + 64
	                      : (try_sparse ? 64 + sizeof(float) * 6 * (size_t)n_src_tiles + sizeof(unsigned) * (size_t)sparse_tiles_t * sparse_words : 0);
	if (need_packed > d->packedA.cap || (records >= 2 && need_packed > d->packedB.cap) || (records >= 3 && need_packed > d->packedC.cap)
	    || need_pieces > d->pieces.cap || need_tickets > d->tickets.cap || need_aux > d->aux.cap) {
		CUDA_TRY(cudaDeviceSynchronize());             // growing frees memory earlier launches may still read
		CUDA_TRY(d->packedA.reserve(need_packed));
		if (records >= 2) CUDA_TRY(d->packedB.reserve(need_packed));
		if (records >= 3) CUDA_TRY(d->packedC.reserve(need_packed));
		CUDA_TRY(d->pieces.reserve(need_pieces));
		CUDA_TRY(d->aux.reserve(need_aux));
		if (need_tickets > d->tickets.cap) {
			CUDA_TRY(d->tickets.reserve(need_tickets));
			CUDA_TRY(cudaMemsetAsync(d->tickets.p, 0, d->tickets.cap, st));      // the kernel leaves them at zero
		}
	}

	// aux: [0, 64) the filament mode word, then one F3DStats per packed tile
	int *mode_word = (int *)d->aux.p;
	F3DStats *stats = filament ? (F3DStats *)((char *)d->aux.p + 64) : nullptr;
	unsigned long long launched = 1;
	const bool direct = plan.k.can_direct;             // the pair kernel packs the raw rows itself (m2m_kernel.cuh)
	if (!direct) {
		pack_sources_kernel<<<n_src_tiles, kSrcTile, 0, st>>>(src_kind(op), src_cols(op), src, n_src, n_pad, (float4 *)d->packedA.p,
		                                                        records >= 2 ? (float4 *)d->packedB.p : nullptr,
		                                                        records >= 3 ? (float4 *)d->packedC.p : nullptr, stats);
		CUDA_TRY(cudaGetLastError());
		++launched;
	}
	if (filament) {
		f3d_mode_kernel<<<1, 256, 0, st>>>(stats, n_src_tiles, n_src, f3d_mode_override(), mode_word);
		CUDA_TRY(cudaGetLastError());
		++launched;
	}

	ConstsOf ck = {sigma, nu, {}};
	dispatch_op(op, reg, ck);
	M2MArgs args = {};
	args.srcA = (const float4 *)d->packedA.p;
	args.srcB = records >= 2 ? (const float4 *)d->packedB.p : nullptr;
	args.srcC = records >= 3 ? (const float4 *)d->packedC.p : nullptr;
	args.src_raw = src;
	args.n_src = n_src;
	args.n_src_tiles = n_src_tiles;
	args.grain = plan.grain;
	args.total_grains = plan.total_grains;
	args.tgt = tgt;
	args.n_tgt = n_tgt;
	args.out = out;
	args.pieces = (double *)d->pieces.p;
	args.tickets = (int *)d->tickets.p;
	args.f3d_mode = mode_word;
	args.k = ck.k;
	const int gm = guard_mode();
	args.exact_only = (gm == 1 || (gm == 0 && n_src_tiles < kMinTilesOptimistic)) ? 1 : 0;
	args.direct = direct ? 1 : 0;
	// pieces per cut target tile ~ runs per tile; beyond this the ordered finish is a kernel of its own
	const long long runs_per_tile = (plan.grid + plan.tiles_t - 1) / plan.tiles_t;
	const bool defer = runs_per_tile > kInKernelFinishPieces;
	args.defer_finish = defer ? 1 : 0;
	CUDA_TRY(cudaEventRecord(d->k_start, st));
	if (try_sparse) {
		// aux: [0, 8) the count of (target tile, source tile) pairs whose boxes meet, [64, ..) source tile boxes, then the masks
		unsigned long long *gate = (unsigned long long *)d->aux.p;
		float *boxes = (float *)((char *)d->aux.p + 64);
		unsigned *mask = (unsigned *)(boxes + 6 * (size_t)n_src_tiles);
		CUDA_TRY(cudaMemsetAsync(gate, 0, sizeof(unsigned long long), st));
		source_tile_boxes_kernel<<<n_src_tiles, kSrcTile, 0, st>>>((const float4 *)d->packedA.p, n_src, boxes);
		target_tile_masks_kernel<<<(unsigned)sparse_tiles_t, 256, 0, st>>>(tgt, q.tcols, n_tgt, kSparseSlots, boxes, n_src_tiles, ck.k.c3,
		                                                                    mask, sparse_words, gate);
		CUDA_TRY(cudaGetLastError());
		SparseArgs sa = {};
		sa.srcA = (const float4 *)d->packedA.p; sa.srcB = (const float4 *)d->packedB.p; sa.n_src_tiles = n_src_tiles;
		sa.tgt = tgt; sa.n_tgt = n_tgt; sa.out = out; sa.mask = mask; sa.words = sparse_words; sa.gate = gate;
		sa.sparse_max = (unsigned long long)(0.3 * (double)sparse_tiles_t * (double)n_src_tiles);
		sa.k = ck.k;
		void *sparams[1] = {&sa};
		CUDA_TRY(cudaLaunchKernel(vort_sparse_fn(reg), dim3((unsigned)sparse_tiles_t), dim3(128), sparams, 0, st));
		args.sparse_gate = gate;
		args.sparse_max = sa.sparse_max;
		launched += 3;
	}
	void *params[1] = {&args};
	CUDA_TRY(cudaLaunchKernel(plan.k.fn, dim3(plan.grid), dim3(plan.k.B), params, plan.k.smem, st));
	if (defer) {
		const long long gpt = plan.total_grains / plan.tiles_t;
		const long long warps = plan.tiles_t * (long long)plan.k.T * plan.k.B * q.nout;
		finish_pieces_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>((const double *)d->pieces.p, out, n_tgt, q.nout,
		                                                                            plan.k.T * plan.k.B, gpt, plan.total_grains, plan.grid);
		CUDA_TRY(cudaGetLastError());
		++launched;
	}
	CUDA_TRY(cudaEventRecord(d->k_stop, st));
	d->timed = true;
	CUDA_TRY(cudaEventRecord(d->arena_idle, st));
	g_launches += launched;
	return CVTX_B200_OK;
}

int cvtx_b200_f3d_inf_mtrx(int device, void *stream_, const float *fil, int n_fil, const float *mes,
                           const float *dir, int n_mes, float *out)
{
	g_err.clear();
	Device *d = get_device(device);
	if (!d) return fail(CVTX_B200_ERR_ARGUMENT, "no such CUDA device");
	if (n_fil < 0 || n_mes < 0) return fail(CVTX_B200_ERR_ARGUMENT, "negative count");
	if (n_fil == 0 || n_mes == 0) return CVTX_B200_OK;
	if (!fil || !mes || !dir || !out) return fail(CVTX_B200_ERR_ARGUMENT, "null pointer");
	DeviceGuard restore;
	DeviceLock lk(d->mu);
	CUDA_TRY(cudaSetDevice(device));
	if (int rc = ensure_ready(d)) return rc;
	constexpr int B = 256, W = 2;
	const int gx = (n_fil + B * W - 1) / (B * W);
	// rows per block: enough blocks to fill the chip several times over, at least one tile of rows
	long want_blocks = 16L * d->prop.multiProcessorCount;
	int gy = (int)((want_blocks + gx - 1) / gx);
	if (gy > (n_mes + 127) / 128) gy = (n_mes + 127) / 128;
	if (gy < 1) gy = 1;
	if (gy > 65535) gy = 65535;
	int rows_per_block = (n_mes + gy - 1) / gy;
	rows_per_block = (rows_per_block + 127) / 128 * 128;
	gy = (n_mes + rows_per_block - 1) / rows_per_block;
	f3d_inf_mtrx_kernel<W, B><<<dim3(gx, gy), B, 0, (cudaStream_t)stream_>>>(fil, n_fil, mes, dir, 0, n_mes, rows_per_block, out, 1.0f);
	CUDA_TRY(cudaGetLastError());
	g_launches += 1;
	return CVTX_B200_OK;
}

int cvtx_b200_m2m_host(int op, int reg, int device, const float *src, int n_src, const float *tgt, int n_tgt,
                       float *out, float sigma, float nu, size_t *h2d_bytes, size_t *d2h_bytes)
{
	g_err.clear();
	if (h2d_bytes) *h2d_bytes = 0;
	if (d2h_bytes) *d2h_bytes = 0;
	if (!get_device(device)) return fail(CVTX_B200_ERR_ARGUMENT, "no such CUDA device");
	if (n_src < 0 || n_tgt < 0) return fail(CVTX_B200_ERR_ARGUMENT, "negative count");
	if (op_is_filament(op)) reg = REG_SINGULAR;
	Info q = {};
	if (!dispatch_op(op, reg, q)) return fail(CVTX_B200_ERR_UNSUPPORTED, "no kernel for this (op, regularisation)");
	if (n_tgt == 0) return CVTX_B200_OK;
	if (!tgt || !out || (n_src > 0 && !src)) return fail(CVTX_B200_ERR_ARGUMENT, "null pointer");
	const size_t sb = sizeof(float) * (size_t)n_src * src_cols(op);
	const size_t tb = sizeof(float) * (size_t)n_tgt * q.tcols;
	HostStage &hs = host_stage();
	DeviceGuard restore;
	std::lock_guard<std::mutex> lk(hs.mu);
	CUDA_TRY(cudaSetDevice(device));
	CUDA_TRY(hs.src.reserve(sb));
	CUDA_TRY(hs.tgt.reserve(tb));
	const double t0 = trace_enabled() ? now_us() : 0.0;
	if (sb) std::memcpy(hs.src.p, src, sb);
	std::memcpy(hs.tgt.p, tgt, tb);
	if (trace_enabled()) std::fprintf(stderr, "cvtx trace: rows into the staging area %.1f us (%zu bytes)\n", now_us() - t0, sb + tb);
	return run_staged(op, reg, std::vector<int>(1, device), n_src, n_tgt, out, sigma, nu, h2d_bytes, d2h_bytes);
}

}  // extern "C"

// =============================================================================
// Staged multi-device runner: one host thread drives every device asynchronously.  The source
// set crosses PCIe once in total: device g receives rows [g n / G, (g + 1) n / G) straight into
// their place in its full-size buffer, the devices then all-gather the shards between themselves
// (exchange.cu: NCCL over NVLink), and each runs pack + pair kernels on its own target shard and
// returns that shard's result.  Targets are embarrassingly parallel -- each output is an independent
// sum over all sources (reference src/P3D.cpp:335-339) -- so the results need no collective.
int cvtx::run_staged(int op, int reg, const std::vector<int> &devices_in, int n_src, int n_tgt,
                     float *out, float sigma, float nu, size_t *h2d_bytes, size_t *d2h_bytes)
{
	Info q = {};
	if (op_is_filament(op)) reg = REG_SINGULAR;
	if (!dispatch_op(op, reg, q)) return fail(CVTX_B200_ERR_UNSUPPORTED, "no kernel for this (op, regularisation)");
	if (devices_in.empty()) return fail(CVTX_B200_ERR_ARGUMENT, "no device given");
	DeviceGuard restore;
	HostStage &hs = host_stage();
	// a device without a target has nothing to do: use the first min(G, n_tgt) devices
	std::vector<int> devices(devices_in.begin(), devices_in.begin() + (n_tgt < (int)devices_in.size() ? (n_tgt > 0 ? n_tgt : 1) : (int)devices_in.size()));
	const int G = (int)devices.size();
	const size_t srow = sizeof(float) * src_cols(op), trow = sizeof(float) * q.tcols, orow = sizeof(float) * q.nout;
	const bool trace = trace_enabled();
	const double t_in = trace ? now_us() : 0.0;
	double t_up = 0.0, t_launched = 0.0, t_done = 0.0;
	CUDA_TRY(cudaSetDevice(devices[0]));
	CUDA_TRY(hs.out.reserve(orow * (size_t)n_tgt));
	std::vector<cudaStream_t> streams(G, nullptr);
	std::vector<const void *> shard(G, nullptr);
	std::vector<void *> full(G, nullptr);
	std::vector<long> soff(G + 1, 0);
	for (int g = 0; g <= G; ++g) soff[g] = (long)n_src * g / G;
	size_t up = 0, down = 0;
	for (int g = 0; g < G; ++g) {
		const long lo = (long)n_tgt * g / G, hi = (long)n_tgt * (g + 1) / G;
		Device *d = get_device(devices[g]);
		if (!d) return fail(CVTX_B200_ERR_ARGUMENT, "no such CUDA device");
		{
			DeviceLock lk(d->mu);
			CUDA_TRY(cudaSetDevice(devices[g]));
			if (int rc = ensure_ready(d)) return rc;
			streams[g] = d->stream;
			CUDA_TRY(d->d_src.reserve(srow * (size_t)n_src));
			CUDA_TRY(d->d_tgt.reserve(trow * (size_t)(hi - lo)));
			CUDA_TRY(d->d_out.reserve(orow * (size_t)(hi - lo)));
		}
		full[g] = d->d_src.p;
		shard[g] = (const char *)d->d_src.p + srow * (size_t)soff[g];
		const size_t sb = srow * (size_t)(soff[g + 1] - soff[g]);
		const size_t tb = trow * (size_t)(hi - lo);
		if (G == 1 && sb + tb <= kKernelUploadBytes) {
			// one device, little data: one launch that reads both ranges through the mapped staging pointers
			const size_t units = sb / 16 + tb / 16 + 8;                        // + the tails of both ranges
			const unsigned blocks = (unsigned)((units + 255) / 256);
			upload_rows_kernel<<<blocks, 256, 0, streams[g]>>>(hs.src.p, (void *)shard[g], sb, hs.tgt.p, d->d_tgt.p, tb);
			CUDA_TRY(cudaGetLastError());
			count_launches(1);
		} else {
			if (sb) CUDA_TRY(cudaMemcpyAsync((void *)shard[g], (const char *)hs.src.p + srow * (size_t)soff[g], sb, cudaMemcpyHostToDevice, streams[g]));
			if (tb) CUDA_TRY(cudaMemcpyAsync(d->d_tgt.p, (const char *)hs.tgt.p + trow * lo, tb, cudaMemcpyHostToDevice, streams[g]));
		}
		up += sb + tb;
	}
	if (int rc = all_gather_rows(devices, streams, shard, soff, full, srow)) return rc;
	if (trace) { if (getenv("CVTX_B200_TRACE_SYNC")) cudaStreamSynchronize(streams[0]); t_up = now_us(); }
	// A caller whose result array is page-locked itself (cudaHostAlloc / cudaHostRegister) gets the result written
	// there directly -- by the kernel or by the D2H copy -- and no second copy on the host (10 us of a 10k x 10k call).
	char *landing = (char *)hs.out.p;
	bool direct_out = false;
	{
		cudaPointerAttributes pa = {};
		if (G == 1 && cudaPointerGetAttributes(&pa, out) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer == (void *)out) {
			landing = (char *)out;
			direct_out = true;
		} else cudaGetLastError();
	}
	for (int g = 0; g < G; ++g) {
		const long lo = (long)n_tgt * g / G, hi = (long)n_tgt * (g + 1) / G;
		if (hi == lo) continue;
		Device *d = get_device(devices[g]);
		// small results are written by the kernel straight into page-locked host memory (unified addressing: the
		// host pointer is valid on the device): the D2H copy of a 10k x 10k call costs as much as a tenth of its kernel
		const bool zero_copy = orow * (size_t)(hi - lo) <= kZeroCopyResultBytes;
		float *result = zero_copy ? (float *)(landing + orow * lo) : (float *)d->d_out.p;
		if (int rc = cvtx_b200_m2m(op, reg, devices[g], streams[g], (const float *)d->d_src.p, n_src,
		                           (const float *)d->d_tgt.p, (int)(hi - lo), result, sigma, nu))
			return rc;
		CUDA_TRY(cudaSetDevice(devices[g]));
		if (!zero_copy) CUDA_TRY(cudaMemcpyAsync(landing + orow * lo, d->d_out.p, orow * (size_t)(hi - lo), cudaMemcpyDeviceToHost, streams[g]));
		down += orow * (size_t)(hi - lo);
	}
	if (trace) t_launched = now_us();
	for (int g = 0; g < G; ++g) {
		const long lo = (long)n_tgt * g / G, hi = (long)n_tgt * (g + 1) / G;
		CUDA_TRY(cudaSetDevice(devices[g]));
		CUDA_TRY(cudaStreamSynchronize(streams[g]));
		if (trace && g == G - 1) t_done = now_us();
		if (hi > lo && !direct_out) copy_parallel((char *)out + orow * lo, (const char *)hs.out.p + orow * lo, orow * (size_t)(hi - lo));
	}
	if (trace)
		std::fprintf(stderr, "cvtx trace: %d x %d on %d device(s): uploads enqueued%s %.1f us, launches %.1f, wait %.1f, result copy %.1f; pair kernel %.1f us\n",
		             n_src, n_tgt, G, getenv("CVTX_B200_TRACE_SYNC") ? " and done" : "", t_up - t_in, t_launched - t_up, t_done - t_launched, now_us() - t_done,
		             1e3 * (double)cvtx_b200_last_pair_kernel_ms(devices[0]));
	if (h2d_bytes) *h2d_bytes = up;
	if (d2h_bytes) *d2h_bytes = down;
	return CVTX_B200_OK;
}

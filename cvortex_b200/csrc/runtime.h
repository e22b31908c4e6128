// runtime.h -- internal to libcvortex.so: per-device state, the library-wide
// pinned staging area and the staged multi-device runner shared by
// device_api.cu (thin C ABI) and host_api.cu (the cvtx_* entry points).
//
// Replaces the reference's global `ocl_state` (src/opencl_acc.cpp:42-49) and
// the per-call clCreateBuffer / clEnqueueWriteBuffer traffic of its host
// wrappers (src/ocl_P3D.cpp:187-222, :260-275): buffers here are grow-only and
// live for the life of the library.
#pragma once
#include <cuda_runtime.h>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/cvtx_b200.h"
#include "host_hooks.h"

namespace cvtx {

struct Buffer {
	void *p = nullptr; size_t cap = 0; bool pinned = false;
	cudaError_t reserve(size_t bytes);
	void release();
};

// Per-device state.  `mu` is recursive: an entry point that keeps its scratch buffers locked for the whole call
// (relaxation: points / field) runs cvtx_b200_m2m, which takes the lock itself, under it.
struct Device {
	std::recursive_mutex mu;
	bool ready = false;
	cudaDeviceProp prop;
	// arena of cvtx_b200_m2m
	Buffer packedA, packedB, packedC, pieces, tickets, aux;
	cudaEvent_t arena_idle = nullptr, k_start = nullptr, k_stop = nullptr;
	bool timed = false;
	// raw rows of the staged (host-array) path
	cudaStream_t stream = nullptr;
	Buffer d_src, d_tgt, d_out;
	// work arrays of the grid redistribution (remesh_device.cu)
	Buffer remesh[32];
};

using DeviceLock = std::lock_guard<std::recursive_mutex>;

// Library-wide pinned (portable) staging: sources, targets and results of ONE
// all-pairs call at a time.  Hold `mu` from the first write into src/tgt until
// run_staged() has returned.
struct HostStage {
	std::mutex mu;
	Buffer src, tgt, out;
};
HostStage &host_stage();

int fail(int code, const std::string &msg);
void count_launches(unsigned long long n);     // feeds cvtx_b200_kernel_launches()
Device *get_device(int device);
// Make `device` current, create its stream / events on first use, hand back its private stream.
int device_stream(int device, cudaStream_t *stream);

// Run (op, reg) with n_src source rows in host_stage().src and n_tgt target
// rows in host_stage().tgt.  Targets are split into contiguous shards over
// `devices`; the sources are uploaded ONCE in total -- device g receives rows
// [g n / G, (g + 1) n / G) -- and all-gathered between the devices
// (all_gather_rows); each shard's result lands in host_stage().out and is
// copied to `out`.  Synchronous.  Caller holds host_stage().mu.
int run_staged(int op, int reg, const std::vector<int> &devices, int n_src, int n_tgt,
               float *out, float sigma, float nu, size_t *h2d_bytes, size_t *d2h_bytes);

// exchange.cu: the all-gather of raw source rows over NCCL / NVLink (see there).
int all_gather_rows(const std::vector<int> &devices, const std::vector<cudaStream_t> &streams,
                    const std::vector<const void *> &shard, const std::vector<long> &row_off,
                    const std::vector<void *> &full, size_t row_bytes);
void release_exchange();

// Restores the calling thread's current device when it goes out of scope: every entry point
// switches devices, and a host application that shares the thread with its own CUDA or torch
// code must not find itself on another GPU afterwards.
struct DeviceGuard {
	int prev = -1;
	DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
	~DeviceGuard() { int cur = -1; if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev); }
	DeviceGuard(const DeviceGuard &) = delete;
	DeviceGuard &operator=(const DeviceGuard &) = delete;
};

}  // namespace cvtx

#define CUDA_TRY(expr)                                                                                    \
	do {                                                                                                  \
		cudaError_t e_ = (expr);                                                                          \
		if (e_ != cudaSuccess)                                                                            \
			return cvtx::fail(CVTX_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
	} while (0)

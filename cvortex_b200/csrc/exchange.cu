// exchange.cu -- the one exchange step of the multi-GPU path: the sources arrive
// sharded over the enabled devices and every device needs all of them; the results need no
// collective (each output is an independent sum over all sources, reference src/P3D.cpp:335-339).
//
// The reference has nothing here: its OpenCL path runs on ONE device ("currently only one
// accelerator is used", /root/reference README.md:149-150, src/ocl_P3D.cpp:51-52).
//
// One process drives all devices.  The exchange is an all-gather of raw source rows over NCCL
// (NVLink 5 / NVSwitch): ncclCommInitAll over the device list, re-created lazily when the caller
// changes the enabled set (cvtx_accelerator_enable / disable), and one grouped ncclBroadcast per
// shard so that shards of unequal length need no padding.  libnccl.so.2 is opened at run time:
// a C or Julia host that never enables a second device does not need NCCL installed.  If it cannot
// be opened the same rows travel as peer-to-peer copies (cudaMemcpyPeerAsync, also NVLink); the
// route taken is reported by cvtx_b200_exchange_backend(), never hidden.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include "op_table.h"
#include "runtime.h"

using namespace cvtx;

namespace {

struct NcclApi {
	void *handle = nullptr;
	bool tried = false;
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	ncclResult_t (*GetVersion)(int *) = nullptr;
	std::string describe;
};

std::mutex g_ex_mu;
NcclApi g_nccl;
std::vector<int> g_comm_devices;
std::vector<ncclComm_t> g_comms;
std::vector<cudaEvent_t> g_ready;          // peer route: "shard of device g is in place"
std::string g_backend = "none yet";

bool load_nccl() {
	if (g_nccl.tried) return g_nccl.handle != nullptr;
	g_nccl.tried = true;
	const char *off = std::getenv("CVTX_B200_EXCHANGE");
	if (off && !std::strcmp(off, "peer")) { g_nccl.describe = "NCCL switched off by CVTX_B200_EXCHANGE=peer"; return false; }
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
	if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
	if (!h) { g_nccl.describe = std::string("libnccl.so.2 not loadable: ") + dlerror(); return false; }
#define CVTX_SYM(field, name) g_nccl.field = (decltype(g_nccl.field))dlsym(h, name)
	CVTX_SYM(CommInitAll, "ncclCommInitAll");
	CVTX_SYM(CommDestroy, "ncclCommDestroy");
	CVTX_SYM(GroupStart, "ncclGroupStart");
	CVTX_SYM(GroupEnd, "ncclGroupEnd");
	CVTX_SYM(Broadcast, "ncclBroadcast");
	CVTX_SYM(GetErrorString, "ncclGetErrorString");
	CVTX_SYM(GetVersion, "ncclGetVersion");
#undef CVTX_SYM
	if (!g_nccl.CommInitAll || !g_nccl.CommDestroy || !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.Broadcast || !g_nccl.GetErrorString) {
		dlclose(h);
		g_nccl.describe = "libnccl.so.2 lacks a required symbol";
		return false;
	}
	g_nccl.handle = h;
	int v = 0;
	if (g_nccl.GetVersion) g_nccl.GetVersion(&v);
	char buf[64];
	std::snprintf(buf, sizeof buf, "nccl %d.%d.%d", v / 10000, (v / 100) % 100, v % 100);
	g_nccl.describe = buf;
	return true;
}

void drop_comms() {
	for (size_t i = 0; i < g_comms.size(); ++i) if (g_comms[i]) g_nccl.CommDestroy(g_comms[i]);
	g_comms.clear();
	for (size_t i = 0; i < g_ready.size(); ++i) {
		if (g_ready[i]) { cudaSetDevice(g_comm_devices[i]); cudaEventDestroy(g_ready[i]); }
	}
	g_ready.clear();
	g_comm_devices.clear();
}

int nccl_fail(const char *what, ncclResult_t r) {
	return fail(CVTX_B200_ERR_CUDA, std::string(what) + ": " + g_nccl.GetErrorString(r));
}

// Communicators (or, on the peer route, peer access + events) for exactly this device list.
int ensure_exchange(const std::vector<int> &devices, bool *use_nccl) {
	*use_nccl = load_nccl();
	if (devices == g_comm_devices && (g_comms.size() == devices.size() || g_ready.size() == devices.size())) return CVTX_B200_OK;
	drop_comms();
	g_comm_devices = devices;
	const int G = (int)devices.size();
	if (*use_nccl) {
		g_comms.assign(G, nullptr);
		const ncclResult_t r = g_nccl.CommInitAll(g_comms.data(), G, devices.data());
		if (r != ncclSuccess) {
			g_comms.clear();
			// a box whose NCCL cannot come up (no shared memory segment, a container without /dev/shm ...) still has NVLink
			std::fprintf(stderr, "cvortex: ncclCommInitAll over %d devices failed (%s); using peer-to-peer copies.\n", G, g_nccl.GetErrorString(r));
			g_nccl.describe += std::string(" (ncclCommInitAll failed: ") + g_nccl.GetErrorString(r) + ")";
			g_nccl.handle = nullptr;
			*use_nccl = false;
		}
	}
	if (!*use_nccl) {
		g_ready.assign(G, nullptr);
		for (int a = 0; a < G; ++a) {
			CUDA_TRY(cudaSetDevice(devices[a]));
			CUDA_TRY(cudaEventCreateWithFlags(&g_ready[a], cudaEventDisableTiming));
			for (int b = 0; b < G; ++b) {
				if (a == b) continue;
				int can = 0;
				CUDA_TRY(cudaDeviceCanAccessPeer(&can, devices[a], devices[b]));
				if (can) {
					const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
					if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_TRY(e);
					cudaGetLastError();
				}
			}
		}
	}
	g_backend = *use_nccl ? g_nccl.describe + ", ncclCommInitAll + grouped ncclBroadcast per shard"
	                      : "peer-to-peer copies (cudaMemcpyPeerAsync); " + g_nccl.describe;
	return CVTX_B200_OK;
}

}  // namespace

// Every device g of `devices` holds rows [row_off[g], row_off[g + 1]) of the source set, at shard[g]
// (device g memory), and has room for all row_off[G] rows at full[g].  On return (stream-ordered:
// the work is queued on streams[g]) full[g] holds every row in order, on every device.  shard[g] may
// point into full[g] at its own offset (in place) or anywhere else on device g.
int cvtx::all_gather_rows(const std::vector<int> &devices, const std::vector<cudaStream_t> &streams,
                          const std::vector<const void *> &shard, const std::vector<long> &row_off,
                          const std::vector<void *> &full, size_t row_bytes)
{
	const int G = (int)devices.size();
	if (G <= 1) {
		if (G == 1 && shard[0] != full[0] && row_off[1] > 0) {
			CUDA_TRY(cudaSetDevice(devices[0]));
			CUDA_TRY(cudaMemcpyAsync(full[0], shard[0], row_bytes * (size_t)row_off[1], cudaMemcpyDeviceToDevice, streams[0]));
		}
		return CVTX_B200_OK;
	}
	std::lock_guard<std::mutex> lk(g_ex_mu);
	bool use_nccl = false;
	if (int rc = ensure_exchange(devices, &use_nccl)) return rc;
	if (use_nccl) {
		ncclResult_t r = g_nccl.GroupStart();
		if (r != ncclSuccess) return nccl_fail("ncclGroupStart", r);
		for (int root = 0; root < G; ++root) {
			const size_t bytes = row_bytes * (size_t)(row_off[root + 1] - row_off[root]);
			if (bytes == 0) continue;
			for (int g = 0; g < G; ++g) {
				void *recv = (char *)full[g] + row_bytes * (size_t)row_off[root];
				const void *send = g == root ? shard[root] : recv;
				r = g_nccl.Broadcast(send, recv, bytes, ncclChar, root, g_comms[g], streams[g]);
				if (r != ncclSuccess) { g_nccl.GroupEnd(); return nccl_fail("ncclBroadcast", r); }
			}
		}
		r = g_nccl.GroupEnd();
		if (r != ncclSuccess) return nccl_fail("ncclGroupEnd", r);
		return CVTX_B200_OK;
	}
	// peer route: every device pulls the other shards once their owners have them in place
	for (int g = 0; g < G; ++g) {
		CUDA_TRY(cudaSetDevice(devices[g]));
		void *own = (char *)full[g] + row_bytes * (size_t)row_off[g];
		const size_t bytes = row_bytes * (size_t)(row_off[g + 1] - row_off[g]);
		if (shard[g] != own && bytes) CUDA_TRY(cudaMemcpyAsync(own, shard[g], bytes, cudaMemcpyDeviceToDevice, streams[g]));
		CUDA_TRY(cudaEventRecord(g_ready[g], streams[g]));
	}
	for (int g = 0; g < G; ++g) {
		CUDA_TRY(cudaSetDevice(devices[g]));
		for (int k = 1; k < G; ++k) {
			const int root = (g + k) % G;                              // staggered: no two devices start on the same owner
			const size_t bytes = row_bytes * (size_t)(row_off[root + 1] - row_off[root]);
			if (bytes == 0) continue;
			CUDA_TRY(cudaStreamWaitEvent(streams[g], g_ready[root], 0));
			CUDA_TRY(cudaMemcpyPeerAsync((char *)full[g] + row_bytes * (size_t)row_off[root], devices[g],
			                             (const char *)full[root] + row_bytes * (size_t)row_off[root], devices[root], bytes, streams[g]));
		}
	}
	return CVTX_B200_OK;
}

void cvtx::release_exchange() {
	std::lock_guard<std::mutex> lk(g_ex_mu);
	if (g_nccl.handle || !g_ready.empty()) drop_comms();
}

extern "C" {

const char *cvtx_b200_exchange_backend(void) {
	std::lock_guard<std::mutex> lk(g_ex_mu);
	static std::string copy;
	copy = g_backend;
	return copy.c_str();
}

// Sources sharded over several devices, everything resident: device devices[g] holds n_src_shard[g]
// source rows at src_shard_dev[g], n_tgt[g] target rows at tgt_dev[g] and receives out_dev[g].  The
// source set of the call is the concatenation of the shards in list order.
int cvtx_b200_m2m_sharded(int op, int reg, int n_dev, const int *devices, const float *const *src_shard_dev,
                          const int *n_src_shard, const float *const *tgt_dev, const int *n_tgt,
                          float *const *out_dev, float sigma, float nu)
{
	if (n_dev < 1 || !devices || !src_shard_dev || !n_src_shard || !tgt_dev || !n_tgt || !out_dev)
		return fail(CVTX_B200_ERR_ARGUMENT, "null list or no device");
	if (op_is_filament(op)) reg = REG_SINGULAR;
	if (!op_supported(op, reg)) return fail(CVTX_B200_ERR_UNSUPPORTED, "no kernel for this (op, regularisation)");
	const int G = n_dev;
	std::vector<int> devs(devices, devices + G);
	std::vector<long> off(G + 1, 0);
	for (int g = 0; g < G; ++g) {
		if (!get_device(devs[g])) return fail(CVTX_B200_ERR_ARGUMENT, "no such CUDA device");
		for (int h = 0; h < g; ++h) if (devs[h] == devs[g]) return fail(CVTX_B200_ERR_ARGUMENT, "a device is listed twice");
		if (n_src_shard[g] < 0 || n_tgt[g] < 0) return fail(CVTX_B200_ERR_ARGUMENT, "negative count");
		off[g + 1] = off[g] + n_src_shard[g];
	}
	if (off[G] > 2147483647L - 512) return fail(CVTX_B200_ERR_ARGUMENT, "too many sources");
	int prev = 0;
	cudaGetDevice(&prev);
	const size_t srow = sizeof(float) * src_cols(op);
	std::vector<cudaStream_t> streams(G);
	std::vector<const void *> shard(G);
	std::vector<void *> full(G);
	int rc = CVTX_B200_OK;
	for (int g = 0; g < G && rc == CVTX_B200_OK; ++g) {
		rc = device_stream(devs[g], &streams[g]);
		if (rc != CVTX_B200_OK) break;
		Device *d = get_device(devs[g]);
		DeviceLock lk(d->mu);
		const cudaError_t e = d->d_src.reserve(srow * (size_t)off[G]);
		if (e != cudaSuccess) rc = fail(CVTX_B200_ERR_CUDA, std::string("device source buffer: ") + cudaGetErrorString(e));
		shard[g] = src_shard_dev[g];
		full[g] = d->d_src.p;
	}
	if (rc == CVTX_B200_OK) rc = all_gather_rows(devs, streams, shard, off, full, srow);
	for (int g = 0; g < G && rc == CVTX_B200_OK; ++g) {
		if (n_tgt[g] == 0) continue;
		rc = cvtx_b200_m2m(op, reg, devs[g], streams[g], (const float *)full[g], (int)off[G], tgt_dev[g], n_tgt[g], out_dev[g], sigma, nu);
	}
	for (int g = 0; g < G; ++g) {
		if (!streams[g]) continue;
		cudaSetDevice(devs[g]);
		const cudaError_t e = cudaStreamSynchronize(streams[g]);
		if (e != cudaSuccess && rc == CVTX_B200_OK) rc = fail(CVTX_B200_ERR_CUDA, std::string("sharded call: ") + cudaGetErrorString(e));
	}
	cudaSetDevice(prev);
	return rc;
}

}  // extern "C"

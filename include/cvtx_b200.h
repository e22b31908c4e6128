#ifndef CVTX_B200_H
#define CVTX_B200_H
/*
 * cvtx_b200.h -- the thin C ABI of the B200 all-pairs backend.
 *
 * This is the seam the reference has between its front ends and its OpenCL
 * host layer: `int opencl_brute_force_<OBJ>_M2M_<fn>(...)`, 0 = done, -1 =
 * "could not, caller decides" (reference src/ocl_P3D.h:33-117, src/ocl_P2D.h,
 * src/ocl_F3D.h:32-77), plus the device bookkeeping of src/opencl_acc.h:43-95.
 * The public drop-in surface stays the reference's own `cvtx_*` ABI
 * (include/cvortex/libcvtx.h in this repo, same 52 symbols); libcvortex.so
 * exports both.  Everything here is `extern "C"`, plain pointers and sizes.
 *
 * Two levels:
 *   cvtx_b200_m2m()       device pointers in, device pointer out, asynchronous
 *                         on the caller's stream -- for callers that keep their
 *                         particles on the GPU (one process per GPU under
 *                         torch.distributed uses this), and what the kernel-only
 *                         benchmark times.
 *   cvtx_b200_m2m_host()  flat host arrays in/out on one device, synchronous:
 *                         pinned staging + H2D + cvtx_b200_m2m + D2H.
 * The pointer-array entry points (`cvtx_P3D_M2M_vel(const cvtx_P3D **...)`) sit on
 * top of these in host_api.cu: gather, shard targets over the enabled
 * devices, call, scatter.
 *
 * Row formats are the reference's structs verbatim:
 *   P3D particle = 7 floats  x y z wx wy wz vol             (libcvtx.h:53-57)
 *   P2D particle = 4 floats  x y vorticity area             (libcvtx.h:66-70)
 *   F3D filament = 7 floats  ax ay az bx by bz strength     (libcvtx.h:60-63)
 *   bsv_V3f / bsv_V2f points and results = 3 / 2 packed floats
 *
 * There is no CPU fallback anywhere behind this header: a call either runs
 * the CUDA kernels or returns an error code with cvtx_b200_last_error() set.
 */
#include <stddef.h>

#ifndef CVTX_B200_API
# if defined(__GNUC__)
#  define CVTX_B200_API __attribute__((visibility("default")))
# else
#  define CVTX_B200_API
# endif
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Ops: which reference entry point a call implements. */
enum cvtx_b200_op {
	CVTX_B200_P3D_VEL = 0,        /* cvtx_P3D_M2M_vel         src (n,7)  tgt (m,3)  out (m,3) */
	CVTX_B200_P3D_DVORT = 1,      /* cvtx_P3D_M2M_dvort       src (n,7)  tgt (m,7)  out (m,3) */
	CVTX_B200_P3D_VISC_DVORT = 2, /* cvtx_P3D_M2M_visc_dvort  src (n,7)  tgt (m,7)  out (m,3) */
	CVTX_B200_P3D_VORT = 3,       /* cvtx_P3D_M2M_vort        src (n,7)  tgt (m,3)  out (m,3) */
	CVTX_B200_P2D_VEL = 4,        /* cvtx_P2D_M2M_vel         src (n,4)  tgt (m,2)  out (m,2) */
	CVTX_B200_P2D_VISC_DVORT = 5, /* cvtx_P2D_M2M_visc_dvort  src (n,4)  tgt (m,4)  out (m,1) */
	CVTX_B200_F3D_VEL = 6,        /* cvtx_F3D_M2M_vel         src (n,7)  tgt (m,3)  out (m,3) */
	CVTX_B200_F3D_DVORT = 7,      /* cvtx_F3D_M2M_dvort       src (n,7)  tgt (m,7)  out (m,3) */
	/* Additive (no counterpart in libcvtx.h): cvtx_P3D_M2M_vel evaluated AT the induced particles'
	 * positions fused with cvtx_P3D_M2M_dvort on them, one pass over the sources.
	 * out row = {u_x, u_y, u_z, dw_x, dw_y, dw_z}. */
	CVTX_B200_P3D_VEL_DVORT = 8   /*                          src (n,7)  tgt (m,7)  out (m,6) */
};

/* Regularisations: the four values of cvtx_VortFunc::cl_kernel_name_ext the
 * reference accelerates (src/VortFunc.cpp:210,223,236,249).  Ignored by the
 * filament ops.  visc ops accept WINCKELMANS and GAUSSIAN only. */
enum cvtx_b200_reg {
	CVTX_B200_SINGULAR = 0,
	CVTX_B200_WINCKELMANS = 1,
	CVTX_B200_PLANETARY = 2,
	CVTX_B200_GAUSSIAN = 3
};

enum cvtx_b200_status {
	CVTX_B200_OK = 0,
	CVTX_B200_ERR_UNSUPPORTED = -1,   /* (op, reg) has no kernel */
	CVTX_B200_ERR_ARGUMENT = -2,      /* bad device / negative count / null pointer */
	CVTX_B200_ERR_CUDA = -3           /* a CUDA call failed; see cvtx_b200_last_error() */
};

/* ---- devices (replaces reference src/opencl_acc.cpp:55-130, OclDeviceState) ---- */
CVTX_B200_API int cvtx_b200_device_count(void);                       /* CUDA devices visible; <0 = CUDA error */
CVTX_B200_API const char *cvtx_b200_device_name(int device);          /* library-owned string, NULL for a bad index */
CVTX_B200_API int cvtx_b200_device_sm_count(int device);
CVTX_B200_API int cvtx_b200_device_clock_khz(int device);             /* cudaDevAttrClockRate */
CVTX_B200_API void cvtx_b200_release(void);                           /* free every per-device arena / pinned buffer */

/* ---- the hot path ------------------------------------------------------------ */
/* Asynchronous on `stream` (a cudaStream_t, NULL = the legacy default stream)
 * of `device`.  All pointers are device pointers on that device; n_src rows of
 * sources act on n_tgt rows of targets; out receives n_tgt rows and is
 * overwritten (zeros when n_src == 0).  Scratch comes from a per-device
 * arena owned by the library; successive calls on one device are ordered
 * against each other even across different streams. */
CVTX_B200_API int cvtx_b200_m2m(int op, int reg, int device, void *stream,
                  const float *src_dev, int n_src,
                  const float *tgt_dev, int n_tgt,
                  float *out_dev, float sigma, float nu);

/* Synchronous, host arrays, one device.  Bytes moved are reported through the
 * optional out-parameters (NULL to ignore). */
CVTX_B200_API int cvtx_b200_m2m_host(int op, int reg, int device,
                       const float *src, int n_src,
                       const float *tgt, int n_tgt,
                       float *out, float sigma, float nu,
                       size_t *h2d_bytes, size_t *d2h_bytes);

/* Several devices, everything resident, SOURCES SHARDED (BASELINE north_star: "via NCCL all-gather
 * over NVLink when sources arrive sharded").  Device devices[g] holds n_src_shard[g] source
 * rows at src_shard_dev[g] and n_tgt[g] target rows at tgt_dev[g], and receives out_dev[g]; the source
 * set of the call is the concatenation of the shards in list order.  The shards are all-gathered between
 * the devices -- ncclCommInitAll over the list (re-created when the list changes), one grouped
 * ncclBroadcast per shard, so unequal shards need no padding -- and every device then runs the
 * unchanged pair kernel on its own targets.  Results are bit-identical to cvtx_b200_m2m on one device
 * with the concatenated sources.  The caller's buffers must be complete before the call (it runs on the
 * library's own per-device streams) and the call returns when every device has finished.  The reference
 * has no counterpart: its accelerated path uses one device (README.md:149-150). */
CVTX_B200_API int cvtx_b200_m2m_sharded(int op, int reg, int n_devices, const int *devices,
                                        const float *const *src_shard_dev, const int *n_src_shard,
                                        const float *const *tgt_dev, const int *n_tgt,
                                        float *const *out_dev, float sigma, float nu);
/* How source shards travel between devices in this process: "nccl 2.x.y, ..." or "peer-to-peer copies ..."
 * (libnccl.so.2 is opened at run time; CVTX_B200_EXCHANGE=peer switches it off).  Library-owned string. */
CVTX_B200_API const char *cvtx_b200_exchange_backend(void);

/* cvtx_F3D_inf_mtrx on device pointers (libcvtx.h:299-305; CPU-only in the reference,
 * src/F3D.cpp:204-227): out_dev[i * n_fil + j] = u_j(mes_i) . dir_i for n_fil filament rows
 * (7 floats) and n_mes points / directions (3 floats each).  Asynchronous on `stream`. */
CVTX_B200_API int cvtx_b200_f3d_inf_mtrx(int device, void *stream, const float *fil_dev, int n_fil,
                                         const float *mes_dev, const float *dir_dev, int n_mes, float *out_dev);

/* Redistribution onto a regular grid with the particles resident on the device (additive:
 * the reference's cvtx_P3D_redistribute_on_grid / cvtx_P2D_redistribute_on_grid,
 * src/P3D.cpp:509-634 and src/P2D.cpp:283-405, take host pointer arrays and return host
 * particles).  rows_dev: n cvtx_P3D (dim 3, 7 floats) or cvtx_P2D (dim 2, 4 floats) structs;
 * out_dev: room for max_out structs of the same kind, or NULL to ask only for the count
 * that survives negligible_vort, as the reference's NULL idiom does.  Same grid placement,
 * same nodes, same order and same pruning rules as the host-array entry points; *n_out
 * receives the number of particles created.  Runs on `stream` (NULL: the library's own) and
 * returns after it has drained -- the counts steer the host. */
enum cvtx_b200_redist {
	CVTX_B200_LAMBDA0 = 0, CVTX_B200_LAMBDA1 = 1, CVTX_B200_LAMBDA2 = 2, CVTX_B200_LAMBDA3 = 3, CVTX_B200_M4P = 4
};
CVTX_B200_API int cvtx_b200_redistribute(int dim, int kind, int device, void *stream, const float *rows_dev, int n,
                                         float grid_density, float negligible_vort, float *out_dev, int max_out, int *n_out);

/* cvtx_P3D_pedrizzetti_relaxation (libcvtx.h:259-265, reference src/P3D.cpp:667-707) on n
 * cvtx_P3D structs resident on the device, in place: the vorticity field at the particles'
 * own positions comes from the CVTX_B200_P3D_VORT kernel, the blend is the reference's FP32
 * arithmetic.  Asynchronous on `stream` (NULL: the library's own). */
CVTX_B200_API int cvtx_b200_pedrizzetti_relaxation(int reg, int device, void *stream, float *rows_dev, int n,
                                                   float fdt, float sigma);

/* ---- introspection ------------------------------------------------------------ */
/* Shape and roofline metadata of (op, reg): floats per source / target / output
 * row, and the algorithmic FP32 lane-ops and MUFU ops per pair of the kernel's
 * formulation. */
CVTX_B200_API int cvtx_b200_op_info(int op, int reg, int *src_cols, int *tgt_cols, int *out_cols,
                      int *lane_ops_per_pair, int *sfu_ops_per_pair);
/* Run-time pipe peaks of `device`, the measured counterparts of the nominal roofline
 * denominators (SMs x 128 lanes x clock, SMs x 16 x clock): what = 0 -> FP32 lane-ops/s sustained
 * by a packed-FMA (FFMA2) loop, what = 1 -> MUFU ops/s sustained by an rsqrt loop.  Synchronous,
 * a few milliseconds. */
CVTX_B200_API int cvtx_b200_measure_peak(int device, int what, double *ops_per_second);
/* Geometry the planner would use: threads per block, targets per thread, the number of
 * persistent blocks (grid_x) and the sources per grain = FP32 chain (grid_y: 256, or 32 for
 * small source sets). */
CVTX_B200_API int cvtx_b200_plan(int op, int device, int n_src, int n_tgt, int *block, int *tgt_per_thread,
                   int *grid_x, int *grid_y);
/* Kernels launched by this library since load (pack + pair + reduce). */
CVTX_B200_API unsigned long long cvtx_b200_kernel_launches(void);
/* Time the all-pairs kernel alone: cudaEvents recorded on `stream` immediately
 * around the pair-kernel launch of the most recent cvtx_b200_m2m() call on
 * `device`.  Blocks until that kernel finishes.  <0 if nothing was recorded. */
CVTX_B200_API float cvtx_b200_last_pair_kernel_ms(int device);
/* Experiments and tests only: force targets-per-thread (1, 2, 4, 8; 0 = planner) and the
 * number of persistent blocks the work is cut into (0 = planner). */
CVTX_B200_API void cvtx_b200_tune(int force_tgt_per_thread, int force_chunks);
/* Experiments and tests only.  Ops whose coincident-pair / non-finite tests can
 * only ever replace an inf or a NaN run the pair loop without them and evaluate
 * again, with the tests, any 256-source chain whose running sums come out
 * non-finite (same bits either way).  mode = 1 makes every
 * chain take the tested form directly, mode = 2 uses the untested form even for
 * source sets of a few tiles (where the default does not bother), mode = 0
 * restores the default; CVTX_B200_GUARDED=0|1|2 in the environment sets the
 * initial mode. */
CVTX_B200_API void cvtx_b200_guarded_only(int mode);
/* Experiments and tests only.  cvtx_P3D_M2M_vort counts only the sources inside the 5-sigma cube around a target
 * (reference src/P3D.cpp:298-322).  On spatially coherent particle orders -- what the redistribution returns --
 * most (target tile, source tile) pairs cannot hold such a source and are skipped (same bits);
 * on = 0 switches that route off, on = 1 (the default) on.  CVTX_B200_SPARSE=0 in the environment does the same. */
CVTX_B200_API void cvtx_b200_sparse_route(int on);
/* Experiments and tests only.  The filament ops pick their fast pair form per call from the
 * filaments themselves: 0 pins the cancellation-free form for short
 * filaments, 1 the form that selects per pair (long filaments, few filaments), anything else
 * restores the automatic choice.  CVTX_B200_F3D_MODE=0|1
 * in the environment sets the initial value. */
CVTX_B200_API void cvtx_b200_f3d_mode(int mode);
/* Which route the most recent cvtx_*_M2M_* call of the public ABI took:
 * 1 = the CUDA kernels, 0 = the host loops (every accelerator disabled, or a
 * user-defined cvtx_VortFunc), -1 = no call yet; and on how many devices. */
CVTX_B200_API int cvtx_b200_last_dispatch(void);
CVTX_B200_API int cvtx_b200_last_devices_used(void);
/* Thread-local description of the last failure in this thread ("" if none). */
CVTX_B200_API const char *cvtx_b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* CVTX_B200_H */

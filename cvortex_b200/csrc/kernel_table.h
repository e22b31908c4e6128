// kernel_table.h -- the compiled instances of m2m_kernel, looked up at run time.
//
// Every (op, regularisation) policy exists in four block geometries (8 targets per thread x 128 threads,
// 4 x 256, 2 x 256, 1 x 128); the two large ones twice, with the FP32 chain length as a compile-time 256
// (the form ptxas schedules best) and as a run-time value for small source sets
// (those instances also pack raw rows themselves, m2m_kernel.cuh "direct").  Vector width and accumulator
// placement of the large geometries are the policy's measured TUNE constants (pair_math.cuh).  The
// instances are spread over kernels_*.cu so that they compile in parallel; device_api.cu only sees this
// table.
#pragma once
#include <cstddef>

namespace cvtx {

struct KernelChoice {
	const void *fn;       // the __global__ function, for cudaLaunchKernel
	int T, B;             // targets per thread, threads per block
	size_t smem;          // dynamic shared memory per block (FP64 accumulators of the M2M_SMEM_ACC instances)
	int occ;              // blocks of it that are resident on one SM
	bool can_direct;      // packs the caller's raw rows itself: no pack kernel, no packed copy
};

// variant: 0 = 8 x 128, 1 = 4 x 256, 2 = 2 x 256, 3 = 1 x 128.  (op, reg) must be a supported pair (op_table.h).
KernelChoice kernel_choice(int op, int reg, int variant, bool grain256, int device);

// sparse_tiles_kernel<P3DVort<reg>, 8, 128> (m2m_kernel.cuh): the box-cutoff op on spatially coherent particle orders.
const void *vort_sparse_fn(int reg);

}  // namespace cvtx

// kernels_p3d_visc_vort.cu -- instances of m2m_kernel (kernel_inst.cuh); split by op so that the library builds in parallel.
#include "kernel_inst.cuh"

namespace cvtx {
KernelChoice choice_p3d_visc(int reg, int v, bool g, int device) { (void)reg; return choice_by_eta<P3DVisc>(reg, v, g, device); }
KernelChoice choice_p3d_vort(int reg, int v, bool g, int device) { (void)reg; return choice_by_reg<P3DVort>(reg, v, g, device); }
const void *vort_sparse_fn(int reg) {
	switch (reg) {
	case REG_SINGULAR:    return (const void *)sparse_tiles_kernel<P3DVort<REG_SINGULAR>, 8, 128>;
	case REG_WINCKELMANS: return (const void *)sparse_tiles_kernel<P3DVort<REG_WINCKELMANS>, 8, 128>;
	case REG_PLANETARY:   return (const void *)sparse_tiles_kernel<P3DVort<REG_PLANETARY>, 8, 128>;
	default:              return (const void *)sparse_tiles_kernel<P3DVort<REG_GAUSSIAN>, 8, 128>;
	}
}
}  // namespace cvtx

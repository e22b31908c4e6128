// kernels_p3d_vel.cu -- instances of m2m_kernel (kernel_inst.cuh); split by op so that the library builds in parallel.
#include "kernel_inst.cuh"

namespace cvtx {
KernelChoice choice_p3d_vel(int reg, int v, bool g, int device) { (void)reg; return choice_by_reg<P3DVel>(reg, v, g, device); }
}  // namespace cvtx

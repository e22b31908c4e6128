// kernels_p2d.cu -- instances of m2m_kernel (kernel_inst.cuh); split by op so that the library builds in parallel.
#include "kernel_inst.cuh"

namespace cvtx {
KernelChoice choice_p2d_vel(int reg, int v, bool g, int device) { (void)reg; return choice_by_reg<P2DVel>(reg, v, g, device); }
KernelChoice choice_p2d_visc(int reg, int v, bool g, int device) { (void)reg; return choice_by_eta<P2DVisc>(reg, v, g, device); }
}  // namespace cvtx

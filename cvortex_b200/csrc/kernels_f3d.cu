// kernels_f3d.cu -- instances of m2m_kernel (kernel_inst.cuh); split by op so that the library builds in parallel.
#include "kernel_inst.cuh"

namespace cvtx {
KernelChoice choice_f3d_vel(int reg, int v, bool g, int device) { (void)reg; return choice<F3DVel>(v, g, device); }
KernelChoice choice_f3d_dvort(int reg, int v, bool g, int device) { (void)reg; return choice<F3DDvort>(v, g, device); }

KernelChoice kernel_choice(int op, int reg, int variant, bool grain256, int device) {
	switch (op) {
	case OP_P3D_VEL:   return choice_p3d_vel(reg, variant, grain256, device);
	case OP_P3D_DVORT: return choice_p3d_dvort(reg, variant, grain256, device);
	case OP_P3D_VISC:  return choice_p3d_visc(reg, variant, grain256, device);
	case OP_P3D_VORT:  return choice_p3d_vort(reg, variant, grain256, device);
	case OP_P2D_VEL:   return choice_p2d_vel(reg, variant, grain256, device);
	case OP_P2D_VISC:  return choice_p2d_visc(reg, variant, grain256, device);
	case OP_F3D_VEL:   return choice_f3d_vel(reg, variant, grain256, device);
	case OP_F3D_DVORT: return choice_f3d_dvort(reg, variant, grain256, device);
	default:           return choice_p3d_vel_dvort(reg, variant, grain256, device);
	}
}
}  // namespace cvtx

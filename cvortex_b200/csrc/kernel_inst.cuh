// kernel_inst.cuh -- included by kernels_*.cu only: instantiates m2m_kernel for a policy and describes the
// instance to the planner (kernel_table.h).
#pragma once
#include <cuda_runtime.h>
#include "kernel_table.h"
#include "m2m_kernel.cuh"

namespace cvtx {

template <class P, int T, int B, int MINB, int VW, int OPT, int GRAIN>
KernelChoice choice_of(int device) {
	auto kern = m2m_kernel<P, T, B, MINB, VW, OPT, GRAIN>;
	static int occ_cache[64];                          // per device: the attribute below is per device too
	const size_t smem = m2m_smem_bytes<P, T, B, OPT>();
	int &occ = occ_cache[device & 63];
	if (occ == 0) {
		if (smem > 0) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		int o = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, B, smem) != cudaSuccess || o < 1) { cudaGetLastError(); o = MINB; }
		occ = o;
	}
	KernelChoice c = {(const void *)kern, T, B, smem, occ, GRAIN == 0 && !P::HYBRID};
	return c;
}

template <class P> KernelChoice choice(int v, bool grain256, int device) {
	if (v == 0) return grain256 ? choice_of<P, 8, 128, 2, P::VW8, P::OPT8, 256>(device) : choice_of<P, 8, 128, 2, P::VW8, P::OPT8, 0>(device);
	if (v == 1) return grain256 ? choice_of<P, 4, 256, 2, P::VW4, P::OPT4, 256>(device) : choice_of<P, 4, 256, 2, P::VW4, P::OPT4, 0>(device);
	// (the filament ops evaluate their reference-arithmetic tier once per chain in EVERY geometry: where it is added
	// decides the FP32 rounding of a flagged target's chain, and a result may not depend on the geometry)
	constexpr int SMALL = P::HYBRID ? M2M_DEFER_EXACT : 0;
	static_assert(!P::HYBRID || ((P::OPT8 & M2M_DEFER_EXACT) && (P::OPT4 & M2M_DEFER_EXACT)), "filament geometries must agree on the tier order");
	if (v == 2) return choice_of<P, 2, 256, 3, 2, SMALL, 0>(device);
	return choice_of<P, 1, 128, 8, 1, SMALL, 0>(device);
}

template <template <int> class POLICY> KernelChoice choice_by_reg(int reg, int v, bool grain256, int device) {
	switch (reg) {
	case REG_SINGULAR:    return choice<POLICY<REG_SINGULAR>>(v, grain256, device);
	case REG_WINCKELMANS: return choice<POLICY<REG_WINCKELMANS>>(v, grain256, device);
	case REG_PLANETARY:   return choice<POLICY<REG_PLANETARY>>(v, grain256, device);
	default:              return choice<POLICY<REG_GAUSSIAN>>(v, grain256, device);
	}
}
template <template <int> class POLICY> KernelChoice choice_by_eta(int reg, int v, bool grain256, int device) {
	return reg == REG_WINCKELMANS ? choice<POLICY<REG_WINCKELMANS>>(v, grain256, device) : choice<POLICY<REG_GAUSSIAN>>(v, grain256, device);
}

// one function per op, each in its own translation unit
KernelChoice choice_p3d_vel(int reg, int v, bool g, int device);
KernelChoice choice_p3d_dvort(int reg, int v, bool g, int device);
KernelChoice choice_p3d_visc(int reg, int v, bool g, int device);
KernelChoice choice_p3d_vort(int reg, int v, bool g, int device);
KernelChoice choice_p3d_vel_dvort(int reg, int v, bool g, int device);
KernelChoice choice_p2d_vel(int reg, int v, bool g, int device);
KernelChoice choice_p2d_visc(int reg, int v, bool g, int device);
KernelChoice choice_f3d_vel(int reg, int v, bool g, int device);
KernelChoice choice_f3d_dvort(int reg, int v, bool g, int device);

}  // namespace cvtx

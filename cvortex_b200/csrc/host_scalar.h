// host_scalar.h -- host all-pairs loops reachable only on explicit request
// (every accelerator disabled, or a user-defined cvtx_VortFunc); see
// host_scalar.cpp.
#pragma once
#include "../../include/cvortex/libcvtx.h"

namespace cvtx {
void host_m2m_p3d_vel(const cvtx_P3D **a, int n, const bsv_V3f *x, int m, bsv_V3f *out, const cvtx_VortFunc *k, float sigma);
void host_m2m_p3d_dvort(const cvtx_P3D **a, int n, const cvtx_P3D **q, int m, bsv_V3f *out, const cvtx_VortFunc *k, float sigma);
void host_m2m_p3d_visc(const cvtx_P3D **a, int n, const cvtx_P3D **q, int m, bsv_V3f *out, const cvtx_VortFunc *k, float sigma, float nu);
void host_m2m_p3d_vort(const cvtx_P3D **a, int n, const bsv_V3f *x, int m, bsv_V3f *out, const cvtx_VortFunc *k, float sigma);
void host_m2m_p2d_vel(const cvtx_P2D **a, int n, const bsv_V2f *x, int m, bsv_V2f *out, const cvtx_VortFunc *k, float sigma);
void host_m2m_p2d_visc(const cvtx_P2D **a, int n, const cvtx_P2D **q, int m, float *out, const cvtx_VortFunc *k, float sigma, float nu);
void host_m2m_f3d_vel(const cvtx_F3D **a, int n, const bsv_V3f *x, int m, bsv_V3f *out);
void host_m2m_f3d_dvort(const cvtx_F3D **a, int n, const cvtx_P3D **q, int m, bsv_V3f *out);
void host_f3d_inf_mtrx(const cvtx_F3D **a, int n, const bsv_V3f *x, const bsv_V3f *dir, int m, float *out);
}  // namespace cvtx

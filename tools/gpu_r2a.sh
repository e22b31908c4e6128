#!/bin/bash
# Round 2, call A: ncu --set full on the kernels furthest below their roof + baseline sweep.
mkdir -p gpurun_out
for spec in "P3D_M2M_vort gaussian vortg" "P2D_M2M_visc_dvort gaussian p2dviscg" "P2D_M2M_vel planetary p2dvelp" "P3D_M2M_dvort planetary dvortp" "P3D_M2M_visc_dvort winckelmans viscw"; do
  set -- $spec
  echo "== ncu --set full $1 $2"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:m2m_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2a_$3 python tools/prof_one.py $1 $2 262144 2>&1 | tail -2
done
echo "== per-op sweep"; timeout 900 python tools/sweep_ops.py 262144 2>&1 | tee gpurun_out/sweep_ops_r2a.log | tail -25

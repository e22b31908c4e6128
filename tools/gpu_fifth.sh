#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "sampled|passed|failed|Error" | tail -20 | tee gpurun_out/pytest_gpu.log
echo "== sanitizer memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "empty_and_tiny or (kernel_geometries and P3D_M2M_vel) or (bench_regime and F3D_M2M_dvort) or (bench_regime and P2D_M2M_visc)" 2>&1 | tail -6 | tee gpurun_out/sanitizer_memcheck.log
echo "== sanitizer racecheck"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(kernel_geometries and P3D_M2M_vel and 4-3) or (kernel_geometries and P2D_M2M_visc and 2-5)" 2>&1 | tail -6 | tee gpurun_out/sanitizer_racecheck.log
echo "== bench"; timeout 1200 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json | cut -c1-200
echo "== bench velw"; timeout 900 python bench.py --workload p3d_vel_winckelmans_1M --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_velw.json | cut -c1-200
for w in p3d_visc_winckelmans_4M p2d_vel+visc_gaussian_4M f3d_vel+dvort_100k_on_2M; do
  echo "== bench $w"; timeout 1200 python bench.py --workload $w --steps 2 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_$w.json | cut -c1-200
done

#!/bin/bash
# Second GPU visit: full GPU test suite, bench (both arms), launch list, ncu captures.
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -3 | tee gpurun_out/bench_ref.json
echo "== bench"; timeout 1200 python bench.py 2>&1 | tail -3 | tee gpurun_out/bench.json
echo "== bench winckelmans"; timeout 900 python bench.py --workload p3d_vel_winckelmans_1M --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_velw.json
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; tail -5 gpurun_out/launches.csv
for spec in "P3D_M2M_vel winckelmans velw" "P3D_M2M_vel gaussian velg" "P3D_M2M_dvort gaussian dvortg"; do
  set -- $spec
  echo "== ncu full $1 $2"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:m2m_kernel -s 1 -c 1 -f -o gpurun_out/prof_$3 python tools/prof_one.py $1 $2 262144 2>&1 | tail -3
done
ls -la gpurun_out

#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 900 python tools/sweep_ops.py 262144 2>&1 | tee gpurun_out/sweep_ops.log
echo "== config 1 (10k)"; timeout 300 python bench.py --workload p3d_vel_winckelmans_10k --steps 50 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_10k.json | cut -c1-1500
echo "== dram traffic at 1M"; timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:m2m_kernel -c 4 --csv --log-file gpurun_out/traffic_1m.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; tail -12 gpurun_out/traffic_1m.csv | cut -c1-400

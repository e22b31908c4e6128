#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== ubench head"; timeout 600 ./tools/ubench 262144 2>&1 | head -13 | tee gpurun_out/ubench_peaks.log
echo "== small targets"; timeout 300 python tools/small_targets.py 2>&1 | tee gpurun_out/small_targets.log
echo "== bench"; timeout 1200 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['fused_vel_dvort'])"

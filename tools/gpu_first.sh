#!/bin/bash
# First GPU contact: smoke, micro-benchmarks, then the GPU parity suite.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20 | tee gpurun_out/smoke.log
echo "== ubench"; timeout 600 ./tools/ubench 262144 2>&1 | tee gpurun_out/ubench.log | tail -60
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=25 2>&1 | tail -150 | tee gpurun_out/pytest_gpu.log | tail -60

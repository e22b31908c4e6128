#!/bin/bash
# Final short validation of round 1 (about 4 minutes of box time): gpurun --timeout 330 -- 'bash tools/gpu_round1c.sh'
mkdir -p gpurun_out
echo "== smoke";   timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -9 | tee gpurun_out/smoke.log
echo "== pytest";  timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench";   timeout 300 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json | cut -c1-200
echo "== per-op sweep"; timeout 200 python tools/sweep_ops.py 262144 2>&1 | tee gpurun_out/sweep_ops.log | grep -E "F3D|vort "
echo "== ubench_alu"; timeout 100 ./tools/ubench_alu 262144 2>&1 | tee gpurun_out/ubench_alu.log | grep -E "F3D|DIFFER"
echo "== bench f3d"; timeout 200 python bench.py --workload f3d_vel+dvort_100k_on_2M --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_f3d.json | cut -c1-200

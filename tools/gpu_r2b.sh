#!/bin/bash
# Round 2, call B: validation of the persistent pair kernel + filament v2 (smoke, GPU tests, the reference's own tests, both bench arms).
mkdir -p gpurun_out
echo "== smoke";   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -10 | tee gpurun_out/smoke_r2b.log
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/pytest_gpu_r2b.log
echo "== reference all_tests on the GPU path"; timeout 600 ./oracle/_ref/all_tests_b200 > gpurun_out/ref_all_tests_r2b.log 2>&1; grep "^Passed" gpurun_out/ref_all_tests_r2b.log
echo "== bench";   timeout 1200 python bench.py 2> gpurun_out/bench_r2b.err | tail -1 | tee gpurun_out/bench_r2b.json | cut -c1-400; tail -5 gpurun_out/bench_r2b.err
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_ref_r2b.json | cut -c1-300
echo "== per-op sweep"; timeout 900 python tools/sweep_ops.py 262144 2>&1 | tee gpurun_out/sweep_ops_r2b.log | tail -30

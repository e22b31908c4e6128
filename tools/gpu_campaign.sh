#!/bin/bash
# One-GPU validation campaign of round 2 (run through gpurun from the repo root); everything lands in gpurun_out/ and the
# summaries worth keeping are copied to profiles/ afterwards.     gpurun --timeout 3000 -- 'bash tools/gpu_campaign.sh'
mkdir -p gpurun_out
echo "== smoke";   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -9 | tee gpurun_out/smoke_r2.txt
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2.txt 2>&1; tail -2 gpurun_out/pytest_gpu_r2.txt
echo "== reference all_tests on the GPU path"; timeout 600 ./oracle/_ref/all_tests_b200 > gpurun_out/reference_all_tests_on_b200_r2.txt 2>&1; grep "^Passed" gpurun_out/reference_all_tests_on_b200_r2.txt
echo "== the reference's own benchmark program on the GPU path"; timeout 900 bash tools/reference_bench.sh 2>&1 | tail -40 | tee gpurun_out/reference_all_bench_r2.txt | grep -i -E "huge|small" | head -12
echo "== compute-sanitizer memcheck (kernel geometries, few-target path, filaments)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kernel_geometries or empty_and_tiny or one_target or vortex_line or page_locked or pointer_arrays" 2>&1 | tail -6 | tee gpurun_out/sanitizer_memcheck_r2.txt
echo "== compute-sanitizer racecheck"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kernel_geometries and (winckelmans or F3D)" 2>&1 | tail -6 | tee gpurun_out/sanitizer_racecheck_r2.txt
echo "== bench (default flags) and the reference arm"
timeout 600 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -c 300 gpurun_out/bench_r2_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_r2_reference_arm.json 2>> gpurun_out/bench_r2_n1.err; tail -c 300 gpurun_out/bench_r2_reference_arm.json
echo "== small calls"; timeout 200 python tools/latency_breakdown.py 2>&1 | tee gpurun_out/latency_breakdown_final.txt | cut -c1-200

#!/bin/bash
# One-GPU measurement campaign (run through gpurun from the repo root); everything lands in
# gpurun_out/ and the summaries worth keeping are copied to profiles/ afterwards.
#   gpurun --timeout 3000 -- 'bash tools/gpu_campaign.sh'
mkdir -p gpurun_out
echo "== smoke";   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/smoke.log
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== reference all_tests on the GPU path"; ./oracle/_ref/all_tests_b200 > gpurun_out/ref_all_tests.log 2>&1; grep "^Passed" gpurun_out/ref_all_tests.log
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-160
echo "== bench";   timeout 1200 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json | cut -c1-160
echo "== bench headline (vel Winckelmans 1M)"; timeout 900 python bench.py --workload p3d_vel_winckelmans_1M --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_velw.json | cut -c1-160
for w in p3d_vel_winckelmans_10k p3d_visc_winckelmans_4M p2d_vel+visc_gaussian_4M f3d_vel+dvort_100k_on_2M; do
  echo "== bench $w"; timeout 1200 python bench.py --workload $w --steps 3 --warmup 3 2>&1 | tail -1 | tee "gpurun_out/bench_$w.json" | cut -c1-160
done
echo "== per-op sweep"; timeout 900 python tools/sweep_ops.py 262144 2>&1 | tee gpurun_out/sweep_ops.log | tail -25
echo "== ubench";  timeout 600 ./tools/ubench 262144 2>&1 | tee gpurun_out/ubench.log | head -14
echo "== launch list (same command as the bench)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; tail -4 gpurun_out/launches.csv | cut -c1-260
echo "== dram traffic at 1M"; timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:m2m_kernel -c 4 --csv --log-file gpurun_out/traffic_1m.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; tail -6 gpurun_out/traffic_1m.csv | cut -c1-300
for spec in "P3D_M2M_vel winckelmans velw" "P3D_M2M_vel gaussian velg" "P3D_M2M_dvort gaussian dvortg"; do
  set -- $spec
  echo "== ncu --set full $1 $2"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:m2m_kernel -s 1 -c 1 -f -o gpurun_out/prof_$3 python tools/prof_one.py $1 $2 262144 2>&1 | tail -2
done

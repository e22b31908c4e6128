#!/bin/bash
# Short one-GPU validation of the optimistic pair loop (about 7 minutes of box time):
#   gpurun --timeout 560 -- 'bash tools/gpu_round1b.sh'
# Order = priority: a cut-off run still leaves the earlier results in gpurun_out/.
mkdir -p gpurun_out
echo "== smoke";   timeout 240 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -9 | tee gpurun_out/smoke.log
echo "== pytest";  timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
if ! grep -q " passed" gpurun_out/pytest_gpu.log || grep -q "failed" gpurun_out/pytest_gpu.log; then
  echo "== pytest again, guarded form only"; CVTX_B200_GUARDED=1 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_guarded.log
fi
echo "== ubench_alu"; timeout 120 ./tools/ubench_alu 262144 2>&1 | tee gpurun_out/ubench_alu.log
echo "== bench";   timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json | cut -c1-200
echo "== bench, guarded form only (A/B at 1M)"; CVTX_B200_GUARDED=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 2>&1 | tail -1 | tee gpurun_out/bench_guarded.json | cut -c1-200
echo "== per-op sweep"; timeout 300 python tools/sweep_ops.py 262144 2>&1 | tee gpurun_out/sweep_ops.log | tail -25
echo "== per-op sweep, guarded form only"; CVTX_B200_GUARDED=1 timeout 300 python tools/sweep_ops.py 262144 2>&1 | tee gpurun_out/sweep_ops_guarded.log | tail -25
echo "== bench --impl reference"; timeout 300 python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-200
for spec in "P3D_M2M_dvort gaussian dvortg" "P3D_M2M_vel gaussian velg"; do
  set -- $spec
  echo "== ncu --set full $1 $2"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:m2m_kernel -s 1 -c 1 -f -o gpurun_out/prof_$3 python tools/prof_one.py $1 $2 262144 2>&1 | tail -2
done
echo "== launch list (same command as the bench)"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; tail -4 gpurun_out/launches.csv | cut -c1-260
echo "== bench headline (vel Winckelmans 1M)"; timeout 300 python bench.py --workload p3d_vel_winckelmans_1M --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_velw.json | cut -c1-200

#!/bin/bash
# Round 2, call F: profiles.  ncu --set full on the dominant kernels and on the ones furthest below their roof (VERDICT r1
# weak #7: none of them had a capture); reports stay on the box (they exceed what comes back), their summaries return.
mkdir -p gpurun_out /tmp/prof
for spec in "P3D_M2M_dvort gaussian dvortg" "P3D_M2M_vel gaussian velg" "P3D_M2M_vel winckelmans velw" "P3D_M2M_visc_dvort winckelmans viscw" \
            "P3D_M2M_vort gaussian vortg" "P2D_M2M_visc_dvort gaussian p2dviscg" "P2D_M2M_vel planetary p2dvelp" "P3D_M2M_dvort planetary dvortp" "F3D_M2M_vel singular f3dvel"; do
  set -- $spec
  echo "== ncu --set full $1 $2"
  timeout 600 ncu --set full --clock-control none -k regex:m2m_kernel -s 1 -c 1 -f -o /tmp/prof/prof_r2_$3 python tools/prof_one.py $1 $2 262144 2>&1 | tail -1
  python tools/ncu_summary.py /tmp/prof/prof_r2_$3.ncu-rep > gpurun_out/ncu_r2_$3.txt 2>&1
  ncu -i /tmp/prof/prof_r2_$3.ncu-rep --page details --csv 2>/dev/null | grep -i -E "warp cycles per issued|No Eligible|Issue Slots Busy|stall" | cut -c1-300 >> gpurun_out/ncu_r2_$3.txt
done
ls -la /tmp/prof | tail -12
echo "== launch list (same command as the bench)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_under_ncu.log 2>&1; grep -c m2m_kernel gpurun_out/launches_r2.csv
echo "== dram traffic at 1M"; timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:m2m_kernel -c 4 --csv --log-file gpurun_out/traffic_1m_r2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1; tail -12 gpurun_out/traffic_1m_r2.csv | cut -c1-300 | tail -3
du -sh gpurun_out

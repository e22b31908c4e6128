#!/bin/bash
# Round 2, call D: direct-pack small-source path + faster ordered finish; GPU tests, config 1 latency, few-target calls.
mkdir -p gpurun_out
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r2d.log
echo "== bench 10k"; timeout 600 python bench.py --workload p3d_vel_winckelmans_10k --steps 200 --warmup 20 --no-extra 2> gpurun_out/bench10k_r2d.err | tail -1 > gpurun_out/bench10k_r2d.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench10k_r2d.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'kern',d['roofline']['avg_launch_ms'])
PY
echo "== few targets"; timeout 600 python tools/small_targets.py 2>&1 | tail -12 | tee gpurun_out/small_targets_r2d.txt

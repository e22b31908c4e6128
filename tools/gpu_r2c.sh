#!/bin/bash
# Round 2, call C: validation after the kernel retune (vector widths, smem accumulators, compile-time chain, 4 runs per slot).
mkdir -p gpurun_out
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r2c.log
echo "== bench";   timeout 1200 python bench.py 2> gpurun_out/bench_r2c.err | tail -1 > gpurun_out/bench_r2c.json; tail -3 gpurun_out/bench_r2c.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2c.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'parity',d['parity']['rel_l2_vs_ref'])
for k,v in d['roofline']['all_kernels'].items(): print(' ',k,round(v['gpairs_per_s'],1),round(v['frac_fp32'],3))
for k,v in d['extra'].items(): print(' X',k,round(v['value'],1),v['ms_per_step'],v.get('e2e') and round(v['e2e']['value'],1),{kk:(round(vv['gpairs_per_s']),round(vv['frac_fp32'],3)) for kk,vv in v['kernels'].items()})
PY
echo "== per-op sweep"; timeout 900 python tools/sweep_ops.py 262144 2>&1 | tee gpurun_out/sweep_ops_r2c.log | tail -30

#!/bin/bash
# Is the reference's own test program deterministic on the GPU path?  20 runs alone, 20 runs next to another process that keeps the GPU busy.
mkdir -p gpurun_out
echo "== alone"
for i in $(seq 1 20); do ./oracle/_ref/all_tests_b200 > gpurun_out/probe_a_$i.txt 2>&1; grep -o "Passed [0-9]* of 420" gpurun_out/probe_a_$i.txt; done | sort | uniq -c
echo "== next to a busy process"
python tools/sweep_ops.py 131072 > /dev/null 2>&1 &
BG=$!
sleep 8
for i in $(seq 1 20); do ./oracle/_ref/all_tests_b200 > gpurun_out/probe_b_$i.txt 2>&1; grep -o "Passed [0-9]* of 420" gpurun_out/probe_b_$i.txt; done | sort | uniq -c
kill $BG 2>/dev/null; wait $BG 2>/dev/null
for f in gpurun_out/probe_*_*.txt; do if ! grep -q "Passed 382 of 420" $f; then echo "ODD RUN $f"; grep -A2 "Test failed" $f | grep -v "^--" | awk 'NR%3==2' | sort | uniq -c | sort -rn | head -20; break; fi; done

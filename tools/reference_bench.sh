#!/bin/bash
# The reference's OWN benchmark program (bench/*.c, unmodified; built by `make -C oracle ref_bench`)
# run against libcvortex.so (B200) and against the reference's CPU build (oracle/_ref, all host
# cores; its "-gpu" lines run its OpenMP path there because that build has no accelerator).
#   gpurun -- 'bash tools/reference_bench.sh > gpurun_out/reference_bench.txt'
# Output: one line per benchmark, "name  size  min-ms".
FUNCS="vel-winckelmans-gpu vel-gaussian-gpu dvort-winckelmans-gpu dvort-gaussian-gpu viscdvort-winckelmans-gpu viscdvort-gaussian-gpu vort-gaussian-gpu redistribute-m4p redistribute-lambda1"
summ() { awk -F'\t' '/Test name:/{n=$3} /Prob. size:/{s=$3} /Minimum:/{printf "%-44s %9d %14s\n", n, s, $3}'; }
echo "== libcvortex.so on $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1): reference bench/all_bench, min of 3 repeats (ms)"
./oracle/_ref/all_bench_b200 -types P3D P2D -funcs $FUNCS -scales small large huge -repeats 3 2>&1 | summ
echo "== reference CPU build on $(nproc) cores: same program, 1 repeat (ms)"
./oracle/_ref/all_bench_ref -types P3D P2D -funcs $FUNCS -scales small medium -repeats 1 2>&1 | summ
echo "== reference CPU build: redistribution at the sizes above"
./oracle/_ref/all_bench_ref -types P3D P2D -funcs redistribute-m4p redistribute-lambda1 -scales large huge -repeats 1 2>&1 | summ

#!/bin/bash
# tools/profile_gpu.sh — ncu evidence for the bench workload (run on the GPU box through gpurun; 1 GPU).
#   1. launch list (gpu__time_duration per launch) of one bench step            -> gpurun_out/launches_<tag>.csv
#   2. ONE `--set full` capture of five consecutive launches = one whole Lanczos iteration at k ~ 150
#      (operator, project, reduce, update, scale)                                -> gpurun_out/iteration_<tag>.ncu-rep (+ raw csv)
# Numbers printed by bench.py under ncu are never bench values.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --max-iteration ${MAXIT:-192}"
NCU="ncu --clock-control none"
# one full step after the warm-up step: skip the warm-up's launches
timeout 900 $NCU --metrics gpu__time_duration.sum -s ${SKIP:-3900} -c ${COUNT:-1300} --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/launches_$TAG.log 2>&1
timeout 900 $NCU --set full --import-source on -k 'regex:k_project|k_update|k_sell_spmv_dot|k_csr_stream_dot|k_scale_norm|k_reduce' -s ${KSKIP:-750} -c 5 -f -o $OUT/iteration_$TAG $BENCH > $OUT/iteration_$TAG.log 2>&1
ncu -i $OUT/iteration_$TAG.ncu-rep --page raw --csv > $OUT/iteration_${TAG}_raw.csv 2>/dev/null
ls -la $OUT | tail -8

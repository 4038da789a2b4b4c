#!/bin/bash
# tools/profile_gpu.sh — ncu evidence for the bench workload (run on the GPU box through gpurun; 1 GPU).
#   1. launch list (gpu__time_duration per launch) of one bench step  -> gpurun_out/launches_<tag>.csv
#   2. one `--set full` capture per hot kernel at a late iteration      -> gpurun_out/<kernel>_<tag>.ncu-rep (+ raw csv)
# Numbers printed by bench.py under ncu are never bench values.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --max-iteration ${MAXIT:-192}"
NCU="ncu --clock-control none"
# one full step after the warm-up step: skip the warm-up's launches
timeout 900 $NCU --metrics gpu__time_duration.sum -s ${SKIP:-4700} -c ${COUNT:-1300} --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/launches_$TAG.log 2>&1
for K in ${KERNELS:-k_project k_update k_csr_stream_dot k_scale_norm k_combine}; do
  timeout 600 $NCU --set full --import-source on -k regex:$K -s ${KSKIP:-150} -c 1 -f -o $OUT/${K}_$TAG $BENCH > $OUT/${K}_$TAG.log 2>&1
  ncu -i $OUT/${K}_$TAG.ncu-rep --page raw --csv > $OUT/${K}_${TAG}_raw.csv 2>/dev/null
done
ls -la $OUT

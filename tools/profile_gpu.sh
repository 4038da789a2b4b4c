#!/bin/bash
# tools/profile_gpu.sh — ncu evidence for the bench workload (run on the GPU box through gpurun; 1 GPU).
#   1. launch list (gpu__time_duration per launch) of one bench step            -> gpurun_out/launches_<tag>.csv
#   2. `--set full` captures of the operator + fused orthogonalisation launches at three depths k of a Lanczos run
#      (k ~ 20, 100, 180: DRAM traffic vs algorithmic bytes as the basis grows)    -> gpurun_out/iteration_<tag>_k*.ncu-rep
#   3. one capture each of the XXZ block kernel (L = 28, double and complex) and of the DIA kernels
# Numbers printed by bench.py under ncu are never bench values.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --parity-iterations 0 --max-iteration ${MAXIT:-192}"
NCU="ncu --clock-control none"
# per step: 4 runs x (begin: 2-3 launches + 192 x 2 launches + combine ...) ~ 1560 launches; skip the warm-up step
timeout 900 $NCU --metrics gpu__time_duration.sum -s ${SKIP:-1700} -c ${COUNT:-800} --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/launches_$TAG.log 2>&1
for K in 20 100 180; do
  # first Lanczos run of the timed step: launch index = warm-up step (~1600) + 3 (begin) + 2 K
  timeout 900 $NCU --set full --import-source on -k 'regex:k_orth|k_dia_spmv_dot|k_sell_spmv_dot' -s $((800 + K)) -c 2 -f -o $OUT/iteration_${TAG}_k$K $BENCH > $OUT/iteration_${TAG}_k$K.log 2>&1
  ncu -i $OUT/iteration_${TAG}_k$K.ncu-rep --page raw --csv > $OUT/iteration_${TAG}_k${K}_raw.csv 2>/dev/null
  ncu -i $OUT/iteration_${TAG}_k$K.ncu-rep --page details > $OUT/iteration_${TAG}_k${K}_details.txt 2>/dev/null
done
for DT in f64; do
  timeout 300 $NCU --set full --import-source on -k regex:k_xxz_block_apply -s 3 -c 1 -f -o $OUT/xxz28_${TAG} python tools/bench_xxz.py 28 block > $OUT/xxz28_${TAG}.log 2>&1
  ncu -i $OUT/xxz28_${TAG}.ncu-rep --page raw --csv > $OUT/xxz28_${TAG}_raw.csv 2>/dev/null
  ncu -i $OUT/xxz28_${TAG}.ncu-rep --page details > $OUT/xxz28_${TAG}_details.txt 2>/dev/null
done
timeout 300 $NCU --set full -k regex:k_xxz_block_apply -s 26 -c 1 -f -o $OUT/xxz28c_${TAG} python tools/bench_xxz.py 28 block > $OUT/xxz28c_${TAG}.log 2>&1
ncu -i $OUT/xxz28c_${TAG}.ncu-rep --page raw --csv > $OUT/xxz28c_${TAG}_raw.csv 2>/dev/null
timeout 300 $NCU --set full -k regex:k_dia_spmv_dot -s 3 -c 1 -f -o $OUT/dia_${TAG} python tools/bench_spmv.py laplacian > $OUT/dia_${TAG}.log 2>&1
ncu -i $OUT/dia_${TAG}.ncu-rep --page raw --csv > $OUT/dia_${TAG}_raw.csv 2>/dev/null
rm -f $OUT/*_${TAG}*.ncu-rep  # the exported csv / details pages are what travels back (gpurun_out is capped at 64 MiB)
ls -la $OUT | tail -12

#!/bin/bash
# tools/sanitize_gpu.sh — compute-sanitizer passes over a subset of the GPU parity tests (the CUDA analogue of the
# reference CI's ASan + valgrind runs).  Run on the GPU box: gpurun -- bash tools/sanitize_gpu.sh
set -u
OUT=gpurun_out
mkdir -p $OUT
SUBSET=${SUBSET:-'dot_norm or inner_product or schmidt or csr_operator or sell_operator or xxz_matrix_free or xxz_block_kernel or dia_storage or simple_matrix or hermitian or single_element or multiple_eigenpairs_8x8 or exponentiate_real or capacity or gerschgorin or lanczos_on_sell or iteration_by_iteration'}
# (racecheck models neither cooperative launches nor grid barriers polled through global memory: run it with
#  LLZ_FUSED_ORTH=0, i.e. on the separate kernels the fused one is built from)
for TOOL in ${TOOLS:-memcheck racecheck}; do
  if [ $TOOL = racecheck ]; then export LLZ_FUSED_ORTH=0; else unset LLZ_FUSED_ORTH; fi
  timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 86 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -x -q -k "$SUBSET" > $OUT/sanitizer_$TOOL.log 2>&1
  echo "$TOOL exit code: $?" | tee -a $OUT/sanitizer_$TOOL.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/sanitizer_$TOOL.log | tail -4
done

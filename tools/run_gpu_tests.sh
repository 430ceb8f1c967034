#!/bin/bash
# Runs the GPU parity suite in isolated processes (a faulting kernel poisons its CUDA context, not the others)
# and a k-NN timing probe.  Usage on the GPU box: bash tools/run_gpu_tests.sh [groups...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
GROUPS_DEFAULT="features mirror distance knn dither kmeans palquant matcher dropin sliding motion reconstruct tile_classes encode"
GROUPS_RUN="${@:-$GROUPS_DEFAULT}"
: > gpurun_out/summary.txt
for g in $GROUPS_RUN; do
  timeout 900 python -m pytest tests/test_gpu_core.py -m gpu -q -x -k "$g" > gpurun_out/test_$g.log 2>&1
  rc=$?
  echo "$g rc=$rc $(tail -1 gpurun_out/test_$g.log)" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt

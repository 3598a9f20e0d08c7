#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for v in 0 1 2; do LUW_TILE_VARIANT=$v QB_PRECS=1,0 timeout 600 python tests/quickbench_dev.py > gpurun_out/quick_v$v.log 2>&1; cat gpurun_out/quick_v$v.log; done

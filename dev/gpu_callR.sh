#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "full_size" 2>&1 | grep -E "assert|Error|passed|failed" | head -12 | tee gpurun_out/pytest_gpu_R.log

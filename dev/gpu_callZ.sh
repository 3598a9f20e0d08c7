#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cpp" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_Z.log

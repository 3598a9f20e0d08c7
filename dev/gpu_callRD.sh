#!/bin/bash
mkdir -p gpurun_out
LUW_RUN_REFERENCE_DRIVER=1 timeout 240 python -m pytest tests/test_reference_driver.py -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_RD.log
tail -45 gpurun_out/reference_driver.log 2>/dev/null | cut -c1-140
timeout 120 python -m pytest tests -m gpu -x -q -k "cpp" 2>&1 | tail -3 | tee -a gpurun_out/pytest_gpu_RD.log

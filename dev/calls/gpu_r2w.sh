#!/bin/bash
# round 2, call W: binned voxeliser on the device (timing, equality with the all-triangles kernel), full GPU suite after the rebuild
mkdir -p gpurun_out
timeout 120 python dev/voxel_bench.py > gpurun_out/r2w_voxel_bench.json 2> gpurun_out/r2w_voxel_bench.err; echo "voxel bench rc=$?"; cat gpurun_out/r2w_voxel_bench.json; tail -3 gpurun_out/r2w_voxel_bench.err
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/r2w_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2w_pytest_gpu.log

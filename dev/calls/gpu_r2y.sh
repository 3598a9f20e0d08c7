#!/bin/bash
# round 2, call Y: the reworked k_inlet_knn (shared-memory slots, two-level worst search): parity, timing, ncu; ncu of the binned voxeliser; memcheck; C1-sized bench line
mkdir -p gpurun_out
timeout 120 ./baseline/_ref/luw_inlet_parity > gpurun_out/r2y_inlet_parity.log 2>&1; echo "inlet parity rc=$?"; tail -1 gpurun_out/r2y_inlet_parity.log
timeout 120 python -m pytest tests/test_inlet_gpu.py -q > gpurun_out/r2y_pytest_inlet.log 2>&1; echo "pytest inlet rc=$?"; tail -2 gpurun_out/r2y_pytest_inlet.log
timeout 120 python dev/inlet_bench.py > gpurun_out/r2y_inlet_bench.json 2> gpurun_out/r2y_inlet_bench.err; echo "bench rc=$?"; cat gpurun_out/r2y_inlet_bench.json
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_inlet_knn -c 1 -o gpurun_out/r2y_ncu_inlet_knn -f python dev/inlet_bench.py > gpurun_out/r2y_ncu_knn.log 2>&1; echo "ncu knn rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_voxelize_mesh_binned -c 1 -o gpurun_out/r2y_ncu_vox_binned -f python dev/voxel_bench.py > gpurun_out/r2y_ncu_vox.log 2>&1; echo "ncu vox rc=$?"
timeout 100 python bench.py --workload profile256_fp32 --also '' --steps 400 --warmup 40 > gpurun_out/r2y_bench_c1.json 2> gpurun_out/r2y_bench_c1.err; echo "c1 bench rc=$?"; cut -c1-400 gpurun_out/r2y_bench_c1.json; tail -2 gpurun_out/r2y_bench_c1.err
LUW_INLET_PARITY_POSITIONS_ONLY=1 timeout 100 compute-sanitizer --tool memcheck ./baseline/_ref/luw_inlet_parity > gpurun_out/r2y_memcheck_inlet.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2y_memcheck_inlet.log

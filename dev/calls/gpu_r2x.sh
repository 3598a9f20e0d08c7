#!/bin/bash
# round 2, call X: f2 in place (a .luw deck through the reference driver, this repo's boundary mapping against the reference's own in the same binary), ncu of the search kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reference_driver.py -q -x -k wrf_style > gpurun_out/r2x_pytest_nwp.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2x_pytest_nwp.log
grep -h "inlet/outlet\|Flux\|flux\|boundary cells\|Threads used" gpurun_out/reference_driver_nwp_*_ours.log | head -20
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_inlet_knn -c 1 -o gpurun_out/r2x_ncu_inlet_knn -f python dev/inlet_bench.py > gpurun_out/r2x_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2x_ncu.log

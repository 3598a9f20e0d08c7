#!/bin/bash
# round 2, call V: f2 on the device (inlet parity harness, search timings), the reference driver tests with the new apply_* functions linked in
mkdir -p gpurun_out
timeout 200 ./baseline/_ref/luw_inlet_parity > gpurun_out/r2v_inlet_parity.log 2>&1; echo "inlet parity rc=$?"; tail -3 gpurun_out/r2v_inlet_parity.log
timeout 120 python dev/inlet_bench.py > gpurun_out/r2v_inlet_bench.json 2> gpurun_out/r2v_inlet_bench.err; echo "bench rc=$?"; cat gpurun_out/r2v_inlet_bench.json; tail -3 gpurun_out/r2v_inlet_bench.err
timeout 300 python -m pytest tests/test_reference_driver.py -q -x > gpurun_out/r2v_pytest_driver.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2v_pytest_driver.log

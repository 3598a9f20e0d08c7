#!/bin/bash
# round 2, call ZC: the FP32 two-pass fallbacks (V7) and the unrolled k_inlet_knn: targeted tests, parity harness, timings
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_thermal_gpu.py tests/test_inlet_gpu.py -q -x -k "decomposed_equals_single_domain or fp32_two_pass or fp32_les_on_a_narrow or does_not_depend_on_the_tile_variant or inlet or nearest or (thermal and fp32)" -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r2zc_pytest.log; tail -4 gpurun_out/r2zc_pytest.log
timeout 40 ./baseline/_ref/luw_inlet_parity 2>&1 | tail -1
timeout 40 python dev/inlet_bench.py 2>/dev/null | tee gpurun_out/r2zc_inlet_bench.json
timeout 40 python bench.py --workload profile256_fp32 --also '' --steps 400 --warmup 40 --no-cpu --no-e2e --traffic off --sustain 0 2>/dev/null | cut -c1-160

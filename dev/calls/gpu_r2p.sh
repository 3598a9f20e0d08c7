#!/bin/bash
# round 2, call P: TYPE_T no longer forces the general body (thermal workloads), flux-correction parity harness, driver tests with the surface flux correction linked in
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_thermal_gpu.py tests/test_reference_driver.py tests/test_gpu_parity.py -m gpu -q -x -rfE -p no:cacheprovider -k "thermal or reference or flux or driver or variant or zone" 2>&1 | tail -8 > gpurun_out/r2p_pytest.log
tail -4 gpurun_out/r2p_pytest.log
baseline/_ref/luw_flux_parity | tail -3
for w in urban_fp16s_thermal urban_fp16c_thermal; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e --traffic off --also '' --sustain 0 2>> gpurun_out/r2p_sweep.err | tee -a gpurun_out/r2p_bench.txt | cut -c1-120
done
timeout 300 python dev/variant_sweep.py urban_fp16s d 40 10 2>> gpurun_out/r2p_sweep.err | tee -a gpurun_out/r2p_sweep.txt
grep -E "^\|  [0-9]+ " gpurun_out/reference_driver.log | tail -1

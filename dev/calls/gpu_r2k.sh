#!/bin/bash
# round 2, call K: TYPE_E lanes through the packed relaxation (rate 1) instead of the out-of-line scalar equilibrium; parity + sweeps
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_thermal_gpu.py tests/test_cpp_host.py -m gpu -q -x -rfE -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/r2k_pytest.log
tail -4 gpurun_out/r2k_pytest.log
for w in urban_fp16s:d,4 urban_fp16s_uf:d channel512_fp16s:d channel512_fp16c:d,5 urban_fp16s:d; do
  timeout 300 python dev/variant_sweep.py ${w%%:*} ${w##*:} 40 10 2>> gpurun_out/r2k_sweep.err | tee -a gpurun_out/r2k_sweep.txt
done

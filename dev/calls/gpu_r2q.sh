#!/bin/bash
# round 2, call Q: one reduction + precomputed edge / zone masks per tile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_thermal_gpu.py -m gpu -q -x -p no:cacheprovider -k "tiled or bench_instantiation or variant or zone or periodic or two_kernel or decomposed or example" 2>&1 | tail -3 > gpurun_out/r2q_pytest.log; tail -2 gpurun_out/r2q_pytest.log
for w in urban_fp16s:d urban_fp16s_uf:d channel512_fp16s:d channel512_fp16c:d urban_fp16s:d; do
  timeout 300 python dev/variant_sweep.py ${w%%:*} ${w##*:} 40 10 2>> gpurun_out/r2q_sweep.err | tee -a gpurun_out/r2q_sweep.txt
done

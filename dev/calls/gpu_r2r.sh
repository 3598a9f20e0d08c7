#!/bin/bash
# round 2, call R: 256-wide tiles (V10: 256x2, 2 CTAs/SM; V11: 256x4, 1 CTA/SM)
mkdir -p gpurun_out
LUW_TILE_VARIANT=10 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "tiled_strict or tiled_fast or periodic" 2>&1 | tail -2 > gpurun_out/r2r_pytest.log; tail -1 gpurun_out/r2r_pytest.log
for w in urban_fp16s:d,10,11 channel512_fp16s:d,10,11,5 urban_fp16s_uf:d,10; do
  timeout 300 python dev/variant_sweep.py ${w%%:*} ${w##*:} 40 10 2>> gpurun_out/r2r_sweep.err | tee -a gpurun_out/r2r_sweep.txt
done
tail -3 gpurun_out/r2r_sweep.err

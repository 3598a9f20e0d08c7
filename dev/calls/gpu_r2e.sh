#!/bin/bash
# round 2, call E: odd Nx on the tile kernels, deck-driven parity through the reference driver (LUW_DUMP_DIR), sync-every-16 in the drop-in run loop, tile variants V6 / V8 on the urban step
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q -rfE -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
grep -E "MLUPs|Steps/s|normal Steps" gpurun_out/reference_driver.log | tail -5
for w in urban_fp16s:4,5,6,8 urban_fp16s_uf:5,8 channel512_fp16s:6,8 channel512_fp16c:3,8; do
  timeout 300 python dev/variant_sweep.py ${w%%:*} ${w##*:} 40 10 2>> gpurun_out/r2e_sweep.err | tee -a gpurun_out/r2e_sweep.txt
done
LUW_PRODUCER_LAG=0 timeout 300 python dev/variant_sweep.py urban_fp16s 8 40 10 2>> gpurun_out/r2e_sweep.err | tee -a gpurun_out/r2e_sweep.txt

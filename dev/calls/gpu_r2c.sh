#!/bin/bash
# round 2, call C: lean kernel v4 (loop constants from the host, no spills, element loads of x-shifted boxes, L2 prefetch of the strip's next tile), occupancy variant V7
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1200 python -m pytest tests -m gpu -q -rfE -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
for w in urban_fp16s:4,5,7 urban_fp16s_uf:4,5,7 channel512_fp16s:0,5,6,7 channel512_fp16c:3,5,6,7; do
  timeout 300 python dev/variant_sweep.py ${w%%:*} ${w##*:} 40 10 2>> gpurun_out/r2c_sweep.err | tee -a gpurun_out/r2c_sweep.txt
done
echo "--- LUW_LEAN_PREFETCH=0" | tee -a gpurun_out/r2c_sweep.txt
for w in urban_fp16s:5,7 channel512_fp16s:5,6; do
  LUW_LEAN_PREFETCH=0 timeout 300 python dev/variant_sweep.py ${w%%:*} ${w##*:} 40 10 2>> gpurun_out/r2c_sweep.err | tee -a gpurun_out/r2c_sweep.txt
done
for v in 4 5; do
LUW_TILE_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide -s 4 -c 1 -o gpurun_out/r2c_urban_v$v -f python bench.py --workload urban_fp16s --steps 4 --warmup 3 --no-cpu --no-e2e --traffic off --also '' --sustain 0 > gpurun_out/r2c_ncu_v$v.log 2>&1
tail -1 gpurun_out/r2c_ncu_v$v.log
done
LUW_TILE_VARIANT=5 timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python dev/prof_case.py 256 12 8 1 63 1 urban 3 > gpurun_out/r2c_sanitizer_racecheck_v5.log 2>&1
tail -3 gpurun_out/r2c_sanitizer_racecheck_v5.log

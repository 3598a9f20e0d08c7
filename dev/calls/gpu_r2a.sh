#!/bin/bash
# round 2, call A: new parity tests on the round-1 kernels + the lean-loop kernel (V5 / V6), variant sweep, ncu of V5 on the urban step, compute-sanitizer
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_box.txt
rm -f gpurun_out/parity_measured.jsonl
timeout 1200 python -m pytest tests -m gpu -q -rfE -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
for w in urban_fp16s:4,5,6 urban_fp16s_uf:4,5 channel512_fp16s:0,5,6 channel512_fp16c:3,5,6 channel512_fp32:0,5; do
  timeout 300 python dev/variant_sweep.py ${w%%:*} ${w##*:} 40 10 2>> gpurun_out/r2a_sweep.err | tee -a gpurun_out/r2a_sweep.txt
done
LUW_TILE_VARIANT=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide -s 4 -c 1 -o gpurun_out/r2a_urban_v5 -f python bench.py --workload urban_fp16s --steps 4 --warmup 3 --no-cpu --no-e2e --traffic off --also '' --sustain 0 > gpurun_out/r2a_ncu.log 2>&1
tail -2 gpurun_out/r2a_ncu.log
# race / memory checker on small tiled cases (TMA + plain stores into the same arrays, hand-rolled mbarrier protocol)
for tool in memcheck racecheck; do
  for v in 4 5; do
    LUW_TILE_VARIANT=$v timeout 600 compute-sanitizer --tool $tool --print-limit 20 python dev/prof_case.py 256 12 8 1 63 1 urban 3 > gpurun_out/r2a_sanitizer_${tool}_v$v.log 2>&1
    tail -3 gpurun_out/r2a_sanitizer_${tool}_v$v.log
  done
done

#!/bin/bash
# round 2, call B: lean kernel v3 (E cells + zones in the fast body, no per-strip rendezvous, lean producer wait) + regrouped FAST math
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1200 python -m pytest tests -m gpu -q -rfE -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
for w in urban_fp16s:4,5 urban_fp16s_uf:4,5 channel512_fp16s:0,5,6 channel512_fp16c:3,5,6 channel512_fp32:0,5; do
  timeout 300 python dev/variant_sweep.py ${w%%:*} ${w##*:} 40 10 2>> gpurun_out/r2b_sweep.err | tee -a gpurun_out/r2b_sweep.txt
done
LUW_TILE_VARIANT=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide -s 4 -c 1 -o gpurun_out/r2b_urban_v5 -f python bench.py --workload urban_fp16s --steps 4 --warmup 3 --no-cpu --no-e2e --traffic off --also '' --sustain 0 > gpurun_out/r2b_ncu.log 2>&1
tail -2 gpurun_out/r2b_ncu.log

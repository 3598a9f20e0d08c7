#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 3; do for w in urban_fp16s urban_fp16s_nz channel512_fp16s; do
LUW_VERBOSE=1 LUW_TILE_VARIANT=$v timeout 300 python bench.py --no-cpu --no-e2e --steps 60 --warmup 10 --workload $w 2> gpurun_out/err_${v}_$w.log | python -c "import json,sys; d=json.load(sys.stdin); print('v$v $w', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), d['clocks'])" | tee -a gpurun_out/variants_C.txt
grep "luw" gpurun_out/err_${v}_$w.log | head -1
done; done
LUW_TILE_VARIANT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_tile -s 4 -c 1 -o gpurun_out/prof_C_urban_v1 -f python bench.py --workload urban_fp16s --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_C.log 2>&1
tail -2 gpurun_out/ncu_C.log

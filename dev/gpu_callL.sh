#!/bin/bash
mkdir -p gpurun_out
run() { LUW_CUDA_LIB=latticeurbanwind_b200/$1/libluw_cuda.so LUW_TILE_VARIANT=$3 timeout 300 python bench.py --no-cpu --no-e2e --steps 60 --warmup 10 --workload $2 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$1 $2 variant=$3', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3))" | tee -a gpurun_out/misc_L.txt; }
for l in lib lib_r1 lib_r4; do run $l urban_fp16s ""; run $l urban_fp16s 1; run $l urban_fp16s_uf ""; run $l channel512_fp16s ""; run $l channel512_fp16c ""; done

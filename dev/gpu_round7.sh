#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
TRACE_TILES_X=4 LUW_CUDA_LIB=latticeurbanwind_b200/lib_trace/libluw_cuda.so python dev/trace_tiles.py 1 62 urban 2>&1
TRACE_TILES_X=4 LUW_CUDA_LIB=latticeurbanwind_b200/lib_trace/libluw_cuda.so python dev/trace_tiles.py 1 4 channel 2>&1
for w in channel512_fp16s urban_fp16s urban_fp16s_uf; do timeout 300 python bench.py --no-cpu --no-e2e --steps 100 --workload $w | python -c "import json,sys; d=json.load(sys.stdin); print('$w', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3))"; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_tile -s 4 -c 1 -o gpurun_out/prof_urban_fp16s_r7 -f python bench.py --workload urban_fp16s --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_r7.log 2>&1

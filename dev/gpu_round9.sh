#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
TRACE_TILES_X=4 LUW_CUDA_LIB=latticeurbanwind_b200/lib_trace/libluw_cuda.so python dev/trace_tiles.py 1 62 urban 2>&1
TRACE_TILES_X=4 LUW_CUDA_LIB=latticeurbanwind_b200/lib_trace/libluw_cuda.so python dev/trace_tiles.py 1 0 periodic_box 2>&1
for w in channel512_fp16s urban_fp16s; do timeout 300 python bench.py --no-cpu --no-e2e --steps 100 --workload $w | python -c "import json,sys; d=json.load(sys.stdin); print('$w', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3))"; done

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_P.log
run() { if [ -n "$2" ]; then export LUW_TILE_VARIANT=$2; else unset LUW_TILE_VARIANT; fi; timeout 300 python bench.py --no-cpu --no-e2e --steps 60 --warmup 10 --workload $1 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$1 variant=$2', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3))" | tee -a gpurun_out/misc_P.txt; }
run channel512_fp16c ""
run channel512_fp16c 0
run channel512_fp16c 1
run channel512_fp16c 4

#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_Y.log
run() { timeout 300 python bench.py --no-cpu --no-e2e --steps 100 --warmup 10 --workload $1 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$1', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3))" | tee -a gpurun_out/misc_Y.txt; }
run channel512_fp16s; run urban_fp16s; run channel512_fp16c; run urban_fp16s_uf; run channel512_fp32; run channel512_fp16s

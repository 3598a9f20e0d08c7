#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_O.log
for a in fast strict; do timeout 600 python bench.py --no-cpu --steps 40 --warmup 5 --workload urban_fp16s --arith $a 2>gpurun_out/err_O.log | python -c "
import json,sys
d=json.load(sys.stdin); print('$a', round(d['value']), d['roofline']['frac'], d['roofline'].get('kernel_ms_isolated'), d['e2e']['job'].get('vk_inlet'), d['e2e']['job'].get('stats_sample_ms'))"; tail -2 gpurun_out/err_O.log; done

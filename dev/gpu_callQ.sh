#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_Q.log
timeout 600 python bench.py --no-cpu --steps 100 --warmup 10 2>gpurun_out/err_Q.log | python -c "
import json,sys
d=json.load(sys.stdin); print('channel', round(d['value']), d['roofline']['frac'], round(d['roofline']['kernel_ms_isolated'],3), 'e2e', round(d['e2e']['value']), d['gpu_launches'], d['e2e']['job'].get('vk_inlet'))"; tail -2 gpurun_out/err_Q.log
timeout 600 python bench.py --no-cpu --steps 100 --warmup 10 --workload urban_fp16s_uf 2>gpurun_out/err_Q.log | python -c "
import json,sys
d=json.load(sys.stdin); print('urban_uf', round(d['value']), d['roofline']['frac'], 'e2e', round(d['e2e']['value']), d['gpu_launches'])"; tail -2 gpurun_out/err_Q.log

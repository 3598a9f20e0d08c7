#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cell_sets or stats or vk" 2>&1 | grep -E "assert|Error|passed|failed" | head -12 | tee gpurun_out/pytest_gpu_S.log
timeout 600 python bench.py --no-cpu --steps 100 --warmup 10 2>gpurun_out/err_S.log | python -c "
import json,sys
d=json.load(sys.stdin); print('channel', round(d['value']), d['roofline']['frac'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], d['gpu_launches'])"; tail -2 gpurun_out/err_S.log
timeout 600 python bench.py --no-cpu --steps 100 --warmup 10 --workload urban_fp16s 2>gpurun_out/err_S.log | python -c "
import json,sys
d=json.load(sys.stdin); print('urban', round(d['value']), d['roofline']['frac'], 'e2e', round(d['e2e']['value']), d['gpu_launches'])"; tail -2 gpurun_out/err_S.log

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_N.log
timeout 900 python bench.py --also urban_fp16s,urban_fp16s_uf,channel512_fp32,channel512_fp16c > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -3 gpurun_out/bench_r1_final.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r1_final.json'))
print(d['config']['name'], round(d['value']), d['roofline']['frac'], 'e2e', round(d['e2e']['value']), d['e2e']['job'], d['cpu_baseline'])
for a in d['also']: print(a['config']['name'], round(a['value']), round(a['roofline']['frac'],3), 'e2e', round(a['e2e']['value']))
"
timeout 400 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/bench_ref_r1_final.json 2>> gpurun_out/bench_r1_final.err; cut -c1-200 gpurun_out/bench_ref_r1_final.json

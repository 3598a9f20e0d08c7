#!/bin/bash
# round 2, call H (1 GPU): C++ host-API end-to-end leg, 1 G-cell domain without host mirrors, lag only on 5-stage rings (FP32 / V0 back to their figures), the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cpp_host.py -m gpu -q -rfEs -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log
for w in channel512_fp32:d channel512_fp16s:d,0 channel512_fp16c:d; do
  timeout 300 python dev/variant_sweep.py ${w%%:*} ${w##*:} 40 10 2>> gpurun_out/r2h_sweep.err | tee -a gpurun_out/r2h_sweep.txt
done
timeout 900 python bench.py > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench_default.json'))
print('headline', d['config']['name'], round(d['value']), 'frac', round(d['roofline']['frac'],3), 'sustained', d.get('sustained',{}).get('value'), 'e2e', round(d['e2e']['value']), 'cpp', d['e2e'].get('cpp_host'))
for a in d.get('also',[]): print(' also', a['config']['name'], round(a['value']), round(a['roofline']['frac'],3))
print('traffic', d['roofline'].get('traffic'), d['roofline'].get('traffic_source'))
PY

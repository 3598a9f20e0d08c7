#!/bin/bash
# round 2, call ZB: validation of the tree as committed -- smoke(), the full GPU suite, the default bench line, the FP32 LES lines, the f2 search timings
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 330 python -m pytest tests -m gpu -q -rfE -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/r2zb_pytest.log
tail -4 gpurun_out/r2zb_pytest.log
timeout 300 python bench.py > gpurun_out/r2zb_bench_default.json 2> gpurun_out/r2zb_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2zb_bench_default.json'))
print('headline', d['config']['name'], round(d['value']), 'frac', round(d['roofline']['frac'],3), d['roofline']['kernel'], 'sustained', round(d.get('sustained',{}).get('value',0)), 'e2e', round(d['e2e']['value']), 'cpp', round(d['e2e']['cpp_host'].get('value',0)), 'traffic', d['roofline'].get('traffic'), 'clocks', d['clocks'])
for a in d.get('also',[]): print(' also', a['config']['name'], round(a['value']), round(a['roofline']['frac'],3))
PY
timeout 60 python bench.py --workload profile256_fp32 --also '' --steps 400 --warmup 40 --no-cpu > gpurun_out/r2zb_bench_c1_fp32.json 2>/dev/null; cut -c1-180 gpurun_out/r2zb_bench_c1_fp32.json
timeout 90 python bench.py --workload urban512_fp32 --also '' --no-cpu > gpurun_out/r2zb_bench_urban512_fp32.json 2>/dev/null; cut -c1-180 gpurun_out/r2zb_bench_urban512_fp32.json
timeout 60 python dev/inlet_bench.py > gpurun_out/r2zb_inlet_bench.json 2>/dev/null; cat gpurun_out/r2zb_inlet_bench.json

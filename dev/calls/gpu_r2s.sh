#!/bin/bash
# round 2, call S: final validation of the tree as committed -- smoke(), the full GPU suite, the default bench line, the reference arm, the thermal step's launch list
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1800 python -m pytest tests -m gpu -q -rfE -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/r2s_pytest.log
tail -4 gpurun_out/r2s_pytest.log
timeout 900 python bench.py > gpurun_out/r2s_bench_default.json 2> gpurun_out/r2s_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s_bench_default.json'))
print('headline', d['config']['name'], round(d['value']), 'frac', round(d['roofline']['frac'],3), d['roofline']['kernel'], 'sustained', round(d.get('sustained',{}).get('value',0)), 'e2e', round(d['e2e']['value']), 'cpp', round(d['e2e']['cpp_host'].get('value',0)), 'traffic', d['roofline'].get('traffic'), 'clocks', d['clocks'])
for a in d.get('also',[]): print(' also', a['config']['name'], round(a['value']), round(a['roofline']['frac'],3))
PY
timeout 600 python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/r2s_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r2s_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2s_thermal_launches.csv python bench.py --workload urban_fp16s_thermal --steps 3 --warmup 3 --no-cpu --no-e2e --traffic off --also '' --sustain 0 > /dev/null 2>&1
grep -E "k_thermal_g|k_stream_collide" gpurun_out/r2s_thermal_launches.csv | awk -F'","' '{print substr($5,1,60), $NF}' | tail -4

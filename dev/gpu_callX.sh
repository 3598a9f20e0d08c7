#!/bin/bash
# final evidence of the round: full bench line (+ other workloads), reference arm, ncu full captures, launch list
mkdir -p gpurun_out
timeout 900 python bench.py --also urban_fp16s,urban_fp16s_uf,channel512_fp32,channel512_fp16c > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err; tail -3 gpurun_out/bench_r1_final2.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r1_final2.json'))
print(d['config']['name'], round(d['value']), d['roofline']['frac'], d['roofline']['kernel_ms_isolated'], 'e2e', round(d['e2e']['value']), d['e2e']['job'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])
for a in d['also']: print(a['config']['name'], round(a['value']), round(a['roofline']['frac'],3), 'e2e', round(a['e2e']['value']))
"
timeout 400 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/bench_ref_r1_final2.json 2>> gpurun_out/bench_r1_final2.err; cut -c1-160 gpurun_out/bench_ref_r1_final2.json
for W in channel512_fp16s urban_fp16s channel512_fp16c; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_tile -s 4 -c 1 -o gpurun_out/prof_r1c_${W} -f python bench.py --workload $W --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_${W}.log 2>&1
tail -1 gpurun_out/ncu_${W}.log
ncu -i gpurun_out/prof_r1c_${W}.ncu-rep --page raw --csv > gpurun_out/raw_c_${W}.csv 2>/dev/null
ncu -i gpurun_out/prof_r1c_${W}.ncu-rep --page source --csv --print-source sass > gpurun_out/src_c_${W}.csv 2>/dev/null
rm -f gpurun_out/prof_r1c_${W}.ncu-rep  # gpurun copies at most 64 MiB back
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_bench.log 2>&1

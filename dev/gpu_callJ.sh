#!/bin/bash
mkdir -p gpurun_out
for W in channel512_fp16s urban_fp16s; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_tile -s 4 -c 1 -o gpurun_out/prof_r1b_${W} -f python bench.py --workload $W --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_${W}.log 2>&1
tail -1 gpurun_out/ncu_${W}.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_bench.log 2>&1
run() { LUW_NO_TILE=$3 timeout 300 python bench.py --no-cpu --no-e2e --steps 60 --warmup 10 --workload $1 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$1 notile=$3', d['roofline']['kernel'], round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3))" | tee -a gpurun_out/misc_J.txt; }
run channel512_fp16c x 0
run channel512_fp16s x 1
run urban_fp16s x 1
run channel512_fp32 x 0

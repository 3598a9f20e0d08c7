#!/bin/bash
# round 2, call L: FP16C tile variants after the lean-loop changes (is the two-pass lean kernel now ahead of the single-pass one?), ncu of the current urban kernel
mkdir -p gpurun_out
for w in channel512_fp16c:3,5,6 urban_fp16c_uf:3,5 urban_fp16s_uf:5,4; do
  timeout 300 python dev/variant_sweep.py ${w%%:*} ${w##*:} 40 10 2>> gpurun_out/r2l_sweep.err | tee -a gpurun_out/r2l_sweep.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide -s 4 -c 1 -o /tmp/r2l_urban -f python bench.py --workload urban_fp16s --steps 4 --warmup 3 --no-cpu --no-e2e --traffic off --also '' --sustain 0 > gpurun_out/r2l_ncu.log 2>&1
ncu -i /tmp/r2l_urban.ncu-rep --page raw --csv > gpurun_out/r2l_raw.csv 2>/dev/null
ncu -i /tmp/r2l_urban.ncu-rep --page source --csv > gpurun_out/r2l_src.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --workload urban_fp16s --steps 2 --warmup 3 --no-cpu --traffic off --also '' --sustain 0 > gpurun_out/r2l_launches.log 2>&1
tail -2 gpurun_out/r2l_launches.log | cut -c1-200

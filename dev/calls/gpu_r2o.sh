#!/bin/bash
# round 2, call O: V9 (128x8 tiles, 16 consumer warps per producer, 1 CTA/SM), yielding consumer waits, where the thermal step's time goes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "variant or tiled_strict" 2>&1 | tail -3 > gpurun_out/r2o_pytest.log; tail -2 gpurun_out/r2o_pytest.log
timeout 300 python dev/variant_sweep.py urban_fp16s d,9 40 10 2>> gpurun_out/r2o_sweep.err | tee -a gpurun_out/r2o_sweep.txt
LUW_LEAN_YIELD=1 timeout 300 python dev/variant_sweep.py urban_fp16s d,9 40 10 2>> gpurun_out/r2o_sweep.err | tee -a gpurun_out/r2o_sweep.txt
timeout 300 python dev/variant_sweep.py channel512_fp16s d,9,5 40 10 2>> gpurun_out/r2o_sweep.err | tee -a gpurun_out/r2o_sweep.txt
timeout 300 python dev/variant_sweep.py urban_fp16s_uf d,9 40 10 2>> gpurun_out/r2o_sweep.err | tee -a gpurun_out/r2o_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_thermal_g -s 2 -c 1 -o /tmp/r2o_thermal -f python bench.py --workload urban_fp16s_thermal --steps 3 --warmup 3 --no-cpu --no-e2e --traffic off --also '' --sustain 0 > gpurun_out/r2o_ncu.log 2>&1
ncu -i /tmp/r2o_thermal.ncu-rep --page raw --csv > gpurun_out/r2o_thermal_raw.csv 2>/dev/null
ncu -i /tmp/r2o_thermal.ncu-rep --page source --csv > gpurun_out/r2o_thermal_src.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2o_thermal_launches.csv python bench.py --workload urban_fp16s_thermal --steps 3 --warmup 3 --no-cpu --no-e2e --traffic off --also '' --sustain 0 > /dev/null 2>&1
grep -E "k_thermal_g|k_stream_collide" gpurun_out/r2o_thermal_launches.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | cut -c1-160 | tail -6

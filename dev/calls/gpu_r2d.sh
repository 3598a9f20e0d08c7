#!/bin/bash
# round 2, call D: lagged refill in the producer loop (LUW_PRODUCER_LAG), A/B on the tile shapes; one ncu capture exported to CSV on the box (the .ncu-rep stays there: gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -rfE -p no:cacheprovider -k "tiled or bench_instantiation or periodic" 2>&1 | tail -15 > gpurun_out/r2d_pytest.log
tail -3 gpurun_out/r2d_pytest.log
for lag in 0 1; do
  echo "--- LUW_PRODUCER_LAG=$lag" | tee -a gpurun_out/r2d_sweep.txt
  for w in urban_fp16s:4,5,7 channel512_fp16s:0,5,6 channel512_fp16c:3 urban_fp16s_uf:4,5; do
    LUW_PRODUCER_LAG=$lag timeout 300 python dev/variant_sweep.py ${w%%:*} ${w##*:} 40 10 2>> gpurun_out/r2d_sweep.err | tee -a gpurun_out/r2d_sweep.txt
  done
done
echo "--- thermal (one-cell-per-thread kernel)" | tee -a gpurun_out/r2d_sweep.txt
for w in urban_fp16s_thermal urban_fp16c_thermal urban_fp16c_uf; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e --traffic off --also '' --sustain 0 2>> gpurun_out/r2d_sweep.err | tee -a gpurun_out/r2d_sweep.txt
done
LUW_TILE_VARIANT=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide -s 4 -c 1 -o /tmp/r2d_urban_v5 -f python bench.py --workload urban_fp16s --steps 4 --warmup 3 --no-cpu --no-e2e --traffic off --also '' --sustain 0 > gpurun_out/r2d_ncu.log 2>&1
ncu -i /tmp/r2d_urban_v5.ncu-rep --page raw --csv > gpurun_out/r2d_raw.csv 2>/dev/null
ncu -i /tmp/r2d_urban_v5.ncu-rep --page source --csv > gpurun_out/r2d_src.csv 2>/dev/null
LUW_TILE_VARIANT=5 timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python dev/prof_case.py 256 12 8 1 63 1 urban 3 > gpurun_out/r2d_sanitizer_racecheck_v5.log 2>&1
tail -3 gpurun_out/r2d_sanitizer_racecheck_v5.log

#!/bin/bash
# round 2, call G: two-kernel thermal step (tiled momentum kernel + k_thermal_g), element loads in the tile kernel (racecheck V4), full GPU test suite, thermal workloads
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1800 python -m pytest tests -m gpu -q -rfE -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2g_pytest.log
tail -5 gpurun_out/r2g_pytest.log
for w in urban_fp16s_thermal urban_fp16c_thermal urban_fp16c_uf urban_fp16s_uf; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e --traffic off --also '' --sustain 0 2>> gpurun_out/r2g_sweep.err | tee -a gpurun_out/r2g_bench.txt | cut -c1-200
done
LUW_TILE_VARIANT=4 timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python dev/prof_case.py 256 12 8 1 63 1 urban 3 > gpurun_out/r2g_sanitizer_racecheck_v4.log 2>&1
tail -2 gpurun_out/r2g_sanitizer_racecheck_v4.log
LUW_TILE_VARIANT=0 timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python dev/prof_case.py 256 12 8 1 4 0 channel 3 > gpurun_out/r2g_sanitizer_racecheck_v0_strict.log 2>&1
tail -2 gpurun_out/r2g_sanitizer_racecheck_v0_strict.log
timeout 300 python dev/variant_sweep.py channel512_fp16s d,0 40 10 2>> gpurun_out/r2g_sweep.err | tee -a gpurun_out/r2g_sweep.txt
timeout 300 python dev/variant_sweep.py channel512_fp32 d 40 10 2>> gpurun_out/r2g_sweep.err | tee -a gpurun_out/r2g_sweep.txt
timeout 300 python dev/variant_sweep.py channel512_fp16c d 40 10 2>> gpurun_out/r2g_sweep.err | tee -a gpurun_out/r2g_sweep.txt

#!/bin/bash
# round 2, call F: overlapped halo exchange (boundary strips first, second stream) -- two ranks sharing the one GPU over CUDA IPC; kernel parity after the strip-order change; racecheck
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -q -x -rfEs -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r2f_pytest_dist.log
tail -6 gpurun_out/r2f_pytest_dist.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -rfE -p no:cacheprovider -k "tiled or bench_instantiation or periodic or decomposed or variant" 2>&1 | tail -8 > gpurun_out/r2f_pytest.log
tail -3 gpurun_out/r2f_pytest.log
for v in 5 4; do
LUW_TILE_VARIANT=$v timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python dev/prof_case.py 256 12 8 1 63 1 urban 3 > gpurun_out/r2f_sanitizer_racecheck_v$v.log 2>&1
tail -2 gpurun_out/r2f_sanitizer_racecheck_v$v.log
done
LUW_TILE_VARIANT=5 timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python dev/prof_case.py 253 12 8 1 63 1 urban 3 > gpurun_out/r2f_sanitizer_memcheck_v5_oddNx.log 2>&1
tail -2 gpurun_out/r2f_sanitizer_memcheck_v5_oddNx.log
timeout 300 python dev/variant_sweep.py urban_fp16s d,5 40 10 2>> gpurun_out/r2f_sweep.err | tee -a gpurun_out/r2f_sweep.txt

#!/bin/bash
# usage: gpu_prof.sh <workload> [variant]: bench line + ncu full capture of the step kernel
W=${1:-channel512_fp16s}; V=${2:-0}
export LUW_TILE_VARIANT=$V
mkdir -p gpurun_out
timeout 600 python bench.py --workload $W --no-cpu > gpurun_out/bench_${W}_v$V.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench_${W}_v$V.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_tile -s 4 -c 1 -o gpurun_out/prof_${W}_v$V -f python bench.py --workload $W --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_${W}.log 2>&1
tail -2 gpurun_out/ncu_${W}.log

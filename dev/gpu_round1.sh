#!/bin/bash
# First GPU round: parity tests, smoke, bench (headline + urban), ncu launch list + full capture of the step kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
ls /etc/OpenCL/vendors > gpurun_out/opencl_vendors.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_channel512_fp16s.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench_channel512_fp16s.json
timeout 600 python bench.py --workload channel512_fp32 --no-cpu > gpurun_out/bench_channel512_fp32.json 2>> gpurun_out/bench_err.log
timeout 600 python bench.py --workload urban_fp16s --no-cpu > gpurun_out/bench_urban_fp16s.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_urban_fp16s.json
timeout 600 python bench.py --workload urban_fp16s_uf --no-cpu > gpurun_out/bench_urban_fp16s_uf.json 2>> gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_channel512_fp16s.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_tile -s 4 -c 2 -o gpurun_out/prof_channel512_fp16s -f python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e >> gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_tile -s 4 -c 2 -o gpurun_out/prof_urban_fp16s -f python bench.py --workload urban_fp16s --steps 4 --warmup 3 --no-cpu --no-e2e >> gpurun_out/ncu_bench.log 2>&1
for v in 0 1 2; do LUW_TILE_VARIANT=$v QB_PRECS=1,0 timeout 600 python tests/quickbench_dev.py > gpurun_out/quick_v$v.log 2>&1; done
LUW_NO_TILE=1 QB_PRECS=1,0 timeout 600 python tests/quickbench_dev.py > gpurun_out/quick_notile.log 2>&1
tail -3 gpurun_out/quick_v*.log

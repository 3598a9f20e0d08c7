#!/bin/bash
# Round-1 GPU call A: parity tests, smoke, bench (both arms), ncu launch list, ncu --set full captures of the step kernel.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.txt; ls /etc/OpenCL/vendors >> gpurun_out/box.txt 2>&1; nproc >> gpurun_out/box.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 900 python bench.py --also urban_fp16s,urban_fp16s_uf,channel512_fp32 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
timeout 400 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/bench_ref_r1.json 2>> gpurun_out/bench_r1.err; cat gpurun_out/bench_ref_r1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_bench.log 2>&1
for W in channel512_fp16s urban_fp16s; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_tile -s 4 -c 1 -o gpurun_out/prof_r1_${W} -f python bench.py --workload $W --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_${W}.log 2>&1
tail -2 gpurun_out/ncu_${W}.log
done
ls -la gpurun_out

#!/bin/bash
# 2-GPU call: decomposed parity over two devices (peer copies), then the N=2 bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "decomposed or nccl" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_E.log
for w in channel512_fp16s urban_fp16s; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --workload $w > gpurun_out/bench_n2_$w.json 2> gpurun_out/bench_n2_$w.err
echo "exit=$? bytes=$(wc -c < gpurun_out/bench_n2_$w.json)"
grep -v "^\*\|OMP_NUM" gpurun_out/bench_n2_$w.err | tail -15
python -c "import json; d=json.load(open('gpurun_out/bench_n2_$w.json')); print('N2 $w', round(d['value']), round(d['ms_per_step'],3), d['roofline']['kernel_ms'], d['halo'])"
done
timeout 300 python bench.py --no-cpu --no-e2e --steps 100 --warmup 10 --workload channel512_fp16s | python -c "import json,sys; d=json.load(sys.stdin); print('N1 channel', round(d['value']), round(d['ms_per_step'],3))"
timeout 300 python bench.py --no-cpu --no-e2e --steps 100 --warmup 10 --workload urban_fp16s | python -c "import json,sys; d=json.load(sys.stdin); print('N1 urban', round(d['value']), round(d['ms_per_step'],3))"

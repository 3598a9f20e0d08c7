#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "nccl" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_F.log
for tr in ipc nccl; do for w in channel512_fp16s urban_fp16s; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --workload $w --transport $tr > gpurun_out/bench_n2_${tr}_$w.json 2> gpurun_out/bench_n2_${tr}_$w.err
echo "exit=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_n2_${tr}_$w.err | tail -5
python -c "import json; d=json.loads(open('gpurun_out/bench_n2_${tr}_$w.json').read().strip().splitlines()[-1]); print('N2 $tr $w', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['halo'])"
done; done

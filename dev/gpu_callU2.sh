#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "nccl" 2>&1 | tail -2 | tee gpurun_out/pytest_gpu_U2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2_U2.json 2> gpurun_out/bench_n2_U2.err
echo "exit=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_n2_U2.err | tail -3
python -c "import json; d=json.loads(open('gpurun_out/bench_n2_U2.json').read().strip().splitlines()[-1]); print('N2', d['config']['decomposition'], round(d['value']), round(d['ms_per_step'],3), d['halo'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 4 --warmup 3 | cut -c1-120

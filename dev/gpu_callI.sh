#!/bin/bash
mkdir -p gpurun_out
run() { # workload decomp
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 100 --warmup 10 --workload $1 $2 > gpurun_out/bench_n4_I.json 2> gpurun_out/bench_n4_I.err
echo "exit=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_n4_I.err | tail -5
python -c "import json; d=json.loads(open('gpurun_out/bench_n4_I.json').read().strip().splitlines()[-1]); print('N4 $1 $2', d['config']['decomposition'], round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['halo'])" | tee -a gpurun_out/n4_I.txt
}
run channel512_fp16s "--decomp 1,1,4"
cp gpurun_out/bench_n4_I.json gpurun_out/bench_n4_channel_114.json
run channel512_fp16s ""
cp gpurun_out/bench_n4_I.json gpurun_out/bench_n4_channel_default.json
run urban_fp16s ""
cp gpurun_out/bench_n4_I.json gpurun_out/bench_n4_urban_default.json
